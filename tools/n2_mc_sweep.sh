set -x
run() { tag=$1; shift; env "$@" python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 100 --warmup 5 --no-cpu --configs '' > gpurun_out/r2_n2_mc_$tag.json 2> gpurun_out/r2_n2_mc_$tag.err; tail -c 600 gpurun_out/r2_n2_mc_$tag.err; }
run 16 HISPMV_MC_CTAS=16
run ce HISPMV_MC_CTAS=-1
run 4 HISPMV_MC_CTAS=4
run 32 HISPMV_MC_CTAS=32
