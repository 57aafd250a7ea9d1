set -x
run() { tag=$1; shift; env "$@" python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --workload c5 --steps 30 --warmup 3 --no-cpu --configs '' ${XARGS} > gpurun_out/r2_n2_c5_$tag.json 2> gpurun_out/r2_n2_c5_$tag.err; tail -c 300 gpurun_out/r2_n2_c5_$tag.err; }
XARGS="--x-exchange nccl" run nccl A=1
XARGS="--x-exchange multicast" run mc16 HISPMV_MC_CTAS=16
XARGS="--x-exchange multicast" run mcce HISPMV_MC_CTAS=-1
XARGS="--x-exchange multicast" run mc8 HISPMV_MC_CTAS=8
