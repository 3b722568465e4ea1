# 2 GPUs: the data-parallel parity tests and the bench line with the round-2 end-state kernels
F="grep -v -E Warning|vectorized_gather|warnings.warn|super().__init__"
( timeout 600 python -m pytest tests/test_gpu_dp.py -q -x -m gpu 2>&1 | $F | tail -4 ) > gpurun_out/r02_c39_dp_tests.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29539 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r02_c39_bench_2gpu.json 2> gpurun_out/r02_c39_bench_2gpu.err
cat gpurun_out/r02_c39_dp_tests.log; tail -c 300 gpurun_out/r02_c39_bench_2gpu.err
