F="grep -v -E Warning|vectorized_gather|warnings.warn|super().__init__"
( timeout 120 python tests/dev/persist_fwd.py 2>&1 | $F | tail -9 ) > gpurun_out/r02_c9_cs4_dev.log
( timeout 500 python -m pytest tests/test_gpu_dp.py tests/test_gpu_gru.py tests/test_gpu_mvae.py tests/test_gpu_latent.py tests/test_gpu_fullsize.py -q -x 2>&1 | $F | tail -8 ) > gpurun_out/r02_c9_tests.log
timeout 300 python bench.py --sections mvae,latent --steps 10 --no-cpu-baseline > gpurun_out/r02_c9_bench_1gpu.json 2>/dev/null
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 5 > gpurun_out/r02_c9_bench_2gpu.json 2> gpurun_out/r02_c9_bench_2gpu.err
for f in gpurun_out/r02_c9_*.log; do echo "== $f"; cut -c1-300 $f; done
tail -3 gpurun_out/r02_c9_bench_2gpu.err | cut -c1-300
ls -la gpurun_out/r02_c9*
