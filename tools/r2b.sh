cd $GRAFT_REPO_ROOT
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/r2w_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2w_pytest.log; tail -6 gpurun_out/r2w_pytest.log
timeout 900 python bench.py > gpurun_out/r2w_bench.json 2> gpurun_out/r2w_bench.err; echo "bench rc=$?"; tail -c 600 gpurun_out/r2w_bench.json; tail -5 gpurun_out/r2w_bench.err
bash tools/r2_profile.sh > gpurun_out/r2w_profile.log 2>&1; tail -12 gpurun_out/r2w_profile.log
