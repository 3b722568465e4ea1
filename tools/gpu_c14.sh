F="grep -v -E Warning|vectorized_gather|warnings.warn|super().__init__"
( timeout 600 python -m pytest tests -m gpu -q 2>&1 | $F | tail -8 ) > gpurun_out/r02_c14_gpu_tests.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r02_c14_bench.json 2> gpurun_out/r02_c14_bench.err
timeout 600 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r02_c14_bench_reference.json 2>/dev/null
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | $F | tail -4 > gpurun_out/r02_c14_smoke.log
for f in gpurun_out/r02_c14_*.log; do echo "== $f"; cut -c1-300 $f; done
tail -2 gpurun_out/r02_c14_bench.err | cut -c1-200; tail -c 600 gpurun_out/r02_c14_bench_reference.json
