#!/bin/sh
# builds a variant of libmrp_b200.so under build/<name>/ with extra nvcc flags
# usage: tools/build_variant.sh name "-DFLAG ..." [files.cu ...]   (default: every .cu is rebuilt)
set -e
name=$1; flags=$2; shift 2
root=$(cd "$(dirname "$0")/.." && pwd)
src=$root/libmultirobotplanning_b200/csrc
out=$root/build/$name
mkdir -p "$out"
files=${*:-$(cd $src && ls *.cu)}
objs=""
for f in $(cd $src && ls *.cu); do
  o=$out/${f%.cu}.o
  case " $files " in
    *" $f "*) /usr/local/cuda/bin/nvcc -O3 -std=c++17 -lineinfo -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC $flags -c $src/$f -o $o ;;
    *) o=$src/${f%.cu}.o ;;
  esac
  objs="$objs $o"
done
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -shared -o $out/libmrp_b200.so $objs $src/widen.o -lcudart -lpthread
echo $out/libmrp_b200.so
