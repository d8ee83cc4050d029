# strong-scaling step (fields by goal + packed all-gather) against the SMs left to the collective,
# the chunk size and NCCL's channel count: N=4 G=4096 sh tools/mg_sweep.sh
N=${N:-2}; G=${G:-2048}
run() { echo "== N=$N G=$G $*"; env "$@" python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 tools/multi_gpu_check.py --goals $G --fields-only 2>&1 | grep "rank 0/" | cut -c1-260; }
run MRP_GATHER_SPARE_SMS=16
run MRP_GATHER_SPARE_SMS=32
run MRP_GATHER_SPARE_SMS=8
run MRP_GATHER_SPARE_SMS=16 NCCL_MAX_NCHANNELS=8
run MRP_GATHER_SPARE_SMS=16 NCCL_MIN_NCHANNELS=16 NCCL_MAX_NCHANNELS=16
run MRP_GATHER_SPARE_SMS=24 MRP_GATHER_CHUNK=248
run MRP_GATHER_SPARE_SMS=48
run MRP_GATHER_SPARE_SMS=64
run MRP_GATHER_SPARE_SMS=48 NCCL_MIN_NCHANNELS=32
run MRP_GATHER_SPARE_SMS=32 NCCL_MIN_NCHANNELS=32
run MRP_GATHER_SPARE_SMS=32 MRP_GATHER_CHUNK=58
