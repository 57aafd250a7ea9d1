set -x
cd $GRAFT_REPO_ROOT
M="--set full --clock-control none --import-source on"
python bench.py --steps 2 --warmup 1 --configs "" --no-cpu > gpurun_out/r2_b_plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2_launches.csv python bench.py --steps 2 --warmup 1 --configs "" --no-cpu > gpurun_out/r2_b_ncu.log 2>&1
python tools/profile_one.py --config c2 --kernel auto --runs 2 > gpurun_out/r2_p_c2.log 2>&1 && ncu $M -k regex:"pb_expand|pb_reduce" -s 2 -c 2 -o gpurun_out/r2_c2_blocked_v11 python tools/profile_one.py --config c2 --kernel auto --runs 2 > gpurun_out/r2_p_c2_ncu.log 2>&1
python tools/profile_one.py --kernel gemv --rows 8192 --cols 4096 --runs 2 --flush > gpurun_out/r2_p_g3a.log 2>&1 && ncu $M -k regex:"gemv" -s 1 -c 1 -o gpurun_out/r2_g3a python tools/profile_one.py --kernel gemv --rows 8192 --cols 4096 --runs 2 --flush > gpurun_out/r2_p_g3a_ncu.log 2>&1
tail -3 gpurun_out/r2_p_c2.log gpurun_out/r2_p_g3a.log
ls -la gpurun_out/*.ncu-rep | tail -3
