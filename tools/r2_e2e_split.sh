python -m pytest tests/test_blocked.py tests/test_full_size.py -x -q -m gpu 2>&1 | tail -3
python bench.py --no-cpu --configs "" --steps 50 > gpurun_out/r2_bench_e2e_split.json 2> gpurun_out/r2_bench_e2e_split.err; python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/r2_bench_e2e_split.json') if l.startswith('{')][-1])
print('split', d['value'], d['e2e'], d['e2e_pageable']['value'], d['parity_ok'])
PY
HISPMV_RUN_SPLIT_X=0 python bench.py --no-cpu --configs "" --steps 50 > gpurun_out/r2_bench_e2e_nosplit.json 2> /dev/null; python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/r2_bench_e2e_nosplit.json') if l.startswith('{')][-1])
print('nosplit', d['value'], d['e2e'], d['e2e_pageable']['value'], d['parity_ok'])
PY
