N=${N:-2}; G=${G:-2048}
run() { echo "== N=$N G=$G $*"; env "$@" python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 tools/multi_gpu_check.py --goals $G --fields-only 2>&1 | grep "rank 0/" | cut -c1-260; }
run MRP_GATHER_SPARE_SMS=16
run MRP_GATHER_SPARE_SMS=0 MRP_GATHER_CHUNK=4096
run MRP_GATHER_SPARE_SMS=32 MRP_GATHER_CHUNK=232
run MRP_GATHER_SPARE_SMS=16 MRP_GATHER_CHUNK=132
