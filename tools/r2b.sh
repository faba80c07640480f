cd $GRAFT_REPO_ROOT
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r2v_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2v_pytest.log; tail -12 gpurun_out/r2v_pytest.log
timeout 900 python bench.py > gpurun_out/r2v_bench.json 2> gpurun_out/r2v_bench.err; echo "bench rc=$?"; tail -c 900 gpurun_out/r2v_bench.json; tail -5 gpurun_out/r2v_bench.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2v_bench_reference.json 2>&1; echo "ref rc=$?"; tail -c 600 gpurun_out/r2v_bench_reference.json
