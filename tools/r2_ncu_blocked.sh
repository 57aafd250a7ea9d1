set -x
cd $GRAFT_REPO_ROOT
M="--set full --clock-control none --import-source on"
python tools/profile_one.py --config c2 --kernel auto --runs 2 > gpurun_out/r2_p_c2.log 2>&1 && ncu $M -k regex:"pb_expand|pb_reduce" -s 2 -c 2 -o gpurun_out/r2_c2_blocked_${1:-v10} python tools/profile_one.py --config c2 --kernel auto --runs 2 > gpurun_out/r2_p_c2_ncu.log 2>&1
tail -3 gpurun_out/r2_p_c2.log
ls -la gpurun_out/*.ncu-rep | tail -3
