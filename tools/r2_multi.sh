cd $GRAFT_REPO_ROOT
N=${1:-2}
if [ "$N" = "2" ]; then timeout 600 python -m pytest tests/test_abi_gpu.py -x -q 2>&1 | tail -5; fi
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/r2t_bench_n$N.json 2> gpurun_out/r2t_bench_n$N.err; echo "torchrun bench rc=$?"; tail -c 700 gpurun_out/r2t_bench_n$N.json; tail -3 gpurun_out/r2t_bench_n$N.err
timeout 500 python bench.py --gpus $N --single-process --steps 10 --warmup 3 > gpurun_out/r2t_bench_sp$N.json 2> gpurun_out/r2t_bench_sp$N.err; echo "single-process bench rc=$?"; tail -c 700 gpurun_out/r2t_bench_sp$N.json; tail -3 gpurun_out/r2t_bench_sp$N.err
