set -x
run() { tag=$1; shift; env $ENVS python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 200 --warmup 5 --no-cpu "$@" > gpurun_out/r2_n2_xs_$tag.json 2> gpurun_out/r2_n2_xs_$tag.err; tail -c 300 gpurun_out/r2_n2_xs_$tag.err; }
ENVS="A=1" run peer8 --configs ''
ENVS="HISPMV_PEER_CTAS=4" run peer4 --configs ''
ENVS="HISPMV_PEER_CTAS=16" run peer16 --configs ''
ENVS="HISPMV_SLICE_PATH=multicast" run mc --configs ''
ENVS="A=1" run c5peer --configs '' --workload c5 --steps 30 --x-exchange multicast
python -m pytest tests/test_multi_gpu.py -x -q -m gpu 2>&1 | tail -3
