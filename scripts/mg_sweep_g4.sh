mkdir -p gpurun_out
G=${G:-4}
run() { tag=$1; shift; (env "$@" timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1 --master-port $((29620 + RANDOM % 200)) bench.py --gpus $G --steps 2 --warmup 1 --skip-single --skip-e2e $EXTRA > gpurun_out/r2_g${G}_$tag.json) 2> gpurun_out/r2_g${G}_$tag.err; }
run cta4 RFB_MG_NCCL_MAX_CTAS=4
run cta8 RFB_MG_NCCL_MAX_CTAS=8
