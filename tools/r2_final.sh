cd $GRAFT_REPO_ROOT
bash tools/r2_profile.sh > gpurun_out/r2x_profile.log 2>&1; tail -4 gpurun_out/r2x_profile.log
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r2x_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2x_pytest.log; tail -4 gpurun_out/r2x_pytest.log
