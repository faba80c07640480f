cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_round2_gpu.py tests/test_portfolios_gpu.py tests/test_large_tc_gpu.py tests/test_fuzz_gpu.py tests/test_canary_gpu.py -m gpu -q -x > gpurun_out/r2p_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2p_pytest.log; tail -30 gpurun_out/r2p_pytest.log
timeout 200 python tools/tcb_diag.py
