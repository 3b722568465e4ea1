F="grep -v -E Warning|vectorized_gather|warnings.warn|super().__init__"
( timeout 600 python -m pytest tests -m gpu -q 2>&1 | $F | tail -12 ) > gpurun_out/r02_c8_gpu_tests.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r02_c8_bench.json 2> gpurun_out/r02_c8_bench.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02_c8_bench_reference.json 2>/dev/null
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/r02_c8_launches_mvae.csv python bench.py --steps 2 --warmup 3 --sections mvae --no-cpu-baseline --profile-steps 2 > /dev/null 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gru_persist_fwd -c 6 -o gpurun_out/r02_c8_gru_fwd python bench.py --steps 2 --warmup 3 --sections mvae --no-cpu-baseline > /dev/null 2>&1
timeout 600 ncu --set full --clock-control none -k regex:gru_persist_bwd -c 6 -o gpurun_out/r02_c8_gru_bwd python bench.py --steps 2 --warmup 3 --sections mvae --no-cpu-baseline > /dev/null 2>&1
T=96 timeout 600 ncu --set full --clock-control none --import-source on -k regex:lstm_persist -c 3 -o gpurun_out/r02_c8_lstm python tests/dev/lstm_persist_time.py > /dev/null 2>&1
ls -la gpurun_out/ | tail -12
for f in gpurun_out/r02_c8_*.log; do echo "== $f"; cut -c1-300 $f; done
tail -3 gpurun_out/r02_c8_bench.err | cut -c1-300
