set -x
run() { tag=$1; shift; env "$@" python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 200 --warmup 5 --no-cpu --configs '' > gpurun_out/r2_n2_mc_$tag.json 2> gpurun_out/r2_n2_mc_$tag.err; tail -c 300 gpurun_out/r2_n2_mc_$tag.err; }
run hold16 HISPMV_MC_CTAS=16
run holdce HISPMV_MC_CTAS=-1
run freece HISPMV_MC_CTAS=-1 HISPMV_BENCH_HOLD_EXCHANGE=0
run hold32 HISPMV_MC_CTAS=32
run hold8 HISPMV_MC_CTAS=8
