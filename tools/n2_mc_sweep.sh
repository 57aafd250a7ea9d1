set -x
run() { tag=$1; shift; env "$@" python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 200 --warmup 5 --no-cpu --configs '' > gpurun_out/r2_n2_mc_$tag.json 2> gpurun_out/r2_n2_mc_$tag.err; tail -c 300 gpurun_out/r2_n2_mc_$tag.err; }
run free148x128 HISPMV_MC_CTAS=148 HISPMV_MC_THREADS=128 HISPMV_BENCH_HOLD_EXCHANGE=0
run free74x128 HISPMV_MC_CTAS=74 HISPMV_MC_THREADS=128 HISPMV_BENCH_HOLD_EXCHANGE=0
run free148x64 HISPMV_MC_CTAS=148 HISPMV_MC_THREADS=64 HISPMV_BENCH_HOLD_EXCHANGE=0
run free296x64 HISPMV_MC_CTAS=296 HISPMV_MC_THREADS=64 HISPMV_BENCH_HOLD_EXCHANGE=0
run hold148x128 HISPMV_MC_CTAS=148 HISPMV_MC_THREADS=128
