set -x
run() { tag=$1; shift; env $ENVS python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --steps 200 --warmup 5 --no-cpu --configs '' "$@" > gpurun_out/r2_n8_xs_$tag.json 2> gpurun_out/r2_n8_xs_$tag.err; tail -c 200 gpurun_out/r2_n8_xs_$tag.err; }
ENVS="HISPMV_SLICE_PATH=multicast" run mc16
ENVS="HISPMV_SLICE_PATH=multicast HISPMV_MC_CTAS=4" run mc4
ENVS="A=1" run root --x-source root
ENVS="HISPMV_PEER_CTAS=1" run peer1
