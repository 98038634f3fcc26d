mkdir -p gpurun_out
run() { tag=$1; shift; (env "$@" timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $((29620 + RANDOM % 200)) bench.py --gpus 2 --steps 2 --warmup 1 --skip-single --skip-e2e > gpurun_out/r2_g2_$tag.json) 2> gpurun_out/r2_g2_$tag.err; }
run m8 RFB_MG_MERGE=8
run m8k8 RFB_MG_MERGE=8 RFB_MG_KMAX=8192 RFB_MG_PIECE_TILES=420
run m16k8 RFB_MG_MERGE=16 RFB_MG_KMAX=8192 RFB_MG_PIECE_TILES=420
