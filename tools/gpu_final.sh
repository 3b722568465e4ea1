# end-of-round record: full GPU suite, smoke, full bench (with the CPU / cuDNN reference figures), the reference arm
F="grep -v -E Warning|vectorized_gather|warnings.warn|super().__init__"
( timeout 600 python -m pytest tests -m gpu -q 2>&1 | $F | tail -8 ) > gpurun_out/r02_final_gpu_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | $F | tail -4 > gpurun_out/r02_final_smoke.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r02_final_bench.json 2> gpurun_out/r02_final_bench.err
timeout 600 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r02_final_bench_reference.json 2>/dev/null
for f in gpurun_out/r02_final_*.log; do echo "== $f"; cut -c1-300 $f; done
tail -2 gpurun_out/r02_final_bench.err | cut -c1-200
