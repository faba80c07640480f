cd $GRAFT_REPO_ROOT
set -x
python __graft_entry__.py --smoke > gpurun_out/r2z_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/r2z_smoke.log
PROFILE_P=1e10 python tools/profile_r2.py > gpurun_out/r2z_plain.log 2>&1; echo "plain rc=$?"; cat gpurun_out/r2z_plain.log
PROFILE_P=1e10 timeout 900 ncu --set full --clock-control none --import-source on -k regex:'small_sweep_packed|path_kernel_tc|large_sweep_tc|large_sweep<|hist_var_fast' -c 14 -f -o gpurun_out/r2z_prof python tools/profile_r2.py > gpurun_out/r2z_ncu.log 2>&1; echo "ncu rc=$?"; tail -3 gpurun_out/r2z_ncu.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2z_bench_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/r2z_bench_under_ncu.json 2> gpurun_out/r2z_bench_under_ncu.err; echo "ncu launch list rc=$?"
ls -la gpurun_out/r2z_*
