set -x
python -m pytest tests -x -q -m gpu > gpurun_out/r2_pytest_full4.log 2>&1; tail -5 gpurun_out/r2_pytest_full4.log
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2_bench_ref_arm.json 2> gpurun_out/r2_bench_ref_arm.err; head -c 600 gpurun_out/r2_bench_ref_arm.json
python bench.py > gpurun_out/r2_bench_n1_c.json 2> gpurun_out/r2_bench_n1_c.err; tail -c 1500 gpurun_out/r2_bench_n1_c.err; head -c 1200 gpurun_out/r2_bench_n1_c.json
python tools/chain_one_gpu.py > gpurun_out/r2_chain_1gpu_b.log 2>&1; cat gpurun_out/r2_chain_1gpu_b.log
