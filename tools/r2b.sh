cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2i_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2i_pytest.log; tail -40 gpurun_out/r2i_pytest.log
