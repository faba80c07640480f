cd $GRAFT_REPO_ROOT
N=${1:-2}
timeout 900 python -m pytest tests/test_multi_gpu.py tests/test_round2_gpu.py tests/test_app_adapter_gpu.py tests/test_sharded_gpu.py -q > gpurun_out/r2k_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2k_pytest.log; tail -15 gpurun_out/r2k_pytest.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/r2k_bench_n$N.json 2> gpurun_out/r2k_bench_n$N.err; echo "torchrun bench rc=$?"; tail -c 700 gpurun_out/r2k_bench_n$N.json; tail -3 gpurun_out/r2k_bench_n$N.err
timeout 600 python bench.py --gpus $N --single-process --steps 5 --warmup 3 > gpurun_out/r2k_bench_sp$N.json 2> gpurun_out/r2k_bench_sp$N.err; echo "single-process bench rc=$?"; tail -c 700 gpurun_out/r2k_bench_sp$N.json; tail -3 gpurun_out/r2k_bench_sp$N.err
