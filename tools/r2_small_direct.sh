set -x
python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "small or plugin or pyhispmv or linear or model or layer" 2>&1 | tail -4
python tools/chain_one_gpu.py 2>&1 | tail -6
HISPMV_SMALL_DIRECT=0 python tools/chain_one_gpu.py 2>&1 | tail -6
