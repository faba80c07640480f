cd $GRAFT_REPO_ROOT
bash tools/r2_profile.sh > gpurun_out/r2x_profile.log 2>&1; tail -3 gpurun_out/r2x_profile.log
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r2x_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2x_pytest.log; tail -4 gpurun_out/r2x_pytest.log
timeout 300 python tools/paths_wide_perf.py > gpurun_out/r2x_paths_wide.log 2>&1; cat gpurun_out/r2x_paths_wide.log
