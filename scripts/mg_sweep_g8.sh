# usage (under gpurun --gpus G): G=8 bash scripts/mg_sweep_g8.sh  -- the full G-GPU 32768^2 bench line + one device-only variant
mkdir -p gpurun_out
G=${G:-8}
run() { tag=$1; shift; (env "$@" timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1 --master-port $((29620 + RANDOM % 200)) bench.py --gpus $G --steps 3 --warmup 2 $EXTRA > gpurun_out/r2_g${G}_$tag.json) 2> gpurun_out/r2_g${G}_$tag.err; }
EXTRA="--skip-single --skip-e2e --block 256" run nb256rl
run full2
