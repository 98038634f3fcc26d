# usage (under gpurun --gpus 8): bash scripts/mg_sweep_g8.sh  -- a few variants of the 8-GPU 32768^2 run, device-resident only
mkdir -p gpurun_out
G=${G:-8}
run() { tag=$1; shift; (env "$@" timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1 --master-port $((29620 + RANDOM % 200)) bench.py --gpus $G --steps 2 --warmup 1 --skip-single --skip-e2e $EXTRA > gpurun_out/r2_g${G}_$tag.json) 2> gpurun_out/r2_g${G}_$tag.err; }
run def
EXTRA="--block 256" run nb256
run m2 RFB_MG_MERGE=2 RFB_MG_KMAX=4096 RFB_MG_PIECE_TILES=280
