set -x
cd $GRAFT_REPO_ROOT
M="--set full --clock-control none --import-source on"
./build/gather_bench ncu > gpurun_out/r2_gather_bench_plain.log 2>&1 && ncu $M -k regex:gather -c 8 -o gpurun_out/r2_gather_bench ./build/gather_bench ncu > gpurun_out/r2_gather_bench_ncu.log 2>&1
python bench.py --steps 2 --warmup 1 --configs "" --no-cpu > gpurun_out/r2_b_plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2_launches.csv python bench.py --steps 2 --warmup 1 --configs "" --no-cpu > gpurun_out/r2_b_ncu.log 2>&1
for cfg in c1 c3b c3c; do
python tools/profile_one.py --config $cfg --kernel auto --runs 2 --flush > gpurun_out/r2_p_$cfg.log 2>&1 && ncu $M -k regex:"spmv_" -s 1 -c 1 -o gpurun_out/r2_${cfg}_auto python tools/profile_one.py --config $cfg --kernel auto --runs 2 --flush > gpurun_out/r2_p_${cfg}_ncu.log 2>&1
done
python tools/profile_one.py --config c5 --kernel auto --runs 1 > gpurun_out/r2_p_c5.log 2>&1 && ncu $M -k regex:"spmv_adaptive" -s 2 -c 2 -o gpurun_out/r2_c5_slab_pass python tools/profile_one.py --config c5 --kernel auto --runs 1 > gpurun_out/r2_p_c5_ncu.log 2>&1
python tools/batch_probe.py > gpurun_out/r2_batch_probe.log 2>&1 && ncu $M -k regex:"spmm_csr|gemm_lite" -c 3 -o gpurun_out/r2_batch python tools/batch_probe.py > gpurun_out/r2_batch_ncu.log 2>&1
ls -la gpurun_out/*.ncu-rep | tail -12
tail -2 gpurun_out/r2_gather_bench_plain.log
