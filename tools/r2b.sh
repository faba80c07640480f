cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_round2_gpu.py tests/test_portfolios_gpu.py tests/test_large_tc_gpu.py tests/test_app_adapter_gpu.py -m gpu -q > gpurun_out/r2m_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2m_pytest.log; tail -40 gpurun_out/r2m_pytest.log
timeout 300 python tools/tc_bounds_perf.py > gpurun_out/r2m_tc_bounds.log 2>&1; echo "tc_bounds rc=$?"; cat gpurun_out/r2m_tc_bounds.log
