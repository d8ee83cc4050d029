# instances/s of the 1000-instance ECBS batch against lanes and slice (second run: arenas grown)
run() { echo "== $*"; env PROF_RUNS=2 "$@" python tools/prof_ecbs.py 1000 2>/dev/null | head -1; }
run MRP_HOST_LANES=1
run MRP_HOST_LANES=1 MRP_HOST_SLICE=128
run MRP_HOST_LANES=1 MRP_HOST_SLICE=512
run MRP_HOST_LANES=1 MRP_HOST_SLICE=0
run MRP_HOST_LANES=4
run MRP_HOST_LANES=4 MRP_HOST_SLICE=128
run MRP_HOST_LANES=4 MRP_HOST_SLICE=0
run MRP_HOST_LANES=2 MRP_HOST_LANE_SIZE=250
run MRP_HOST_LANES=8 MRP_HOST_LANE_SIZE=62 CUDA_DEVICE_MAX_CONNECTIONS=32
nproc
