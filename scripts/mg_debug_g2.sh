mkdir -p gpurun_out
(timeout 500 python -m pytest tests/test_gpu_dist.py -q --timeout 300 -x 2>&1 | tail -15) > gpurun_out/r2_g2_dist3.log
run() { tag=$1; shift; (env "$@" timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $((29620 + RANDOM % 200)) bench.py --gpus 2 --steps 1 --warmup 1 --skip-single --skip-e2e --size 16384 > gpurun_out/r2_g2_dbg_$tag.json) 2> gpurun_out/r2_g2_dbg_$tag.err; }
run def
run s500 RFB_MG_SLICE_US=500
run m1 RFB_MG_MERGE=1
(timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29688 bench.py --gpus 2 --steps 2 --warmup 1 --skip-single --skip-e2e > gpurun_out/r2_g2_rl3.json) 2> gpurun_out/r2_g2_rl3.err
