F="grep -v -E Warning|vectorized_gather|warnings.warn|super().__init__"
( timeout 400 python -m pytest tests/test_gpu_scripts.py -q -x 2>&1 | $F | tail -15 ) > gpurun_out/r02_c11_scripts.log
timeout 300 python bench.py --sections mvae --steps 10 --no-cpu-baseline > gpurun_out/r02_c11_bench_mvae.json 2> gpurun_out/r02_c11_bench.err
for f in gpurun_out/r02_c11_*.log; do echo "== $f"; cut -c1-400 $f; done
tail -3 gpurun_out/r02_c11_bench.err | cut -c1-300
python - <<'PY'
import json
try:
    d=json.load(open('gpurun_out/r02_c11_bench_mvae.json'))
    print(round(d['value']), d['ms_per_step']); print(d['roofline'])
except Exception as e: print('ERR',e)
PY
