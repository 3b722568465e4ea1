F="grep -v -E Warning|vectorized_gather|warnings.warn|super().__init__"
( timeout 300 python -m pytest tests/test_gpu_latent.py -q -x -k "graphed or inference" 2>&1 | $F | tail -15 ) > gpurun_out/r02_c10_graph_test.log
timeout 300 python bench.py --sections inpaint --steps 4 --no-cpu-baseline > gpurun_out/r02_c10_bench_inpaint.json 2> gpurun_out/r02_c10_bench.err
for f in gpurun_out/r02_c10_*.log; do echo "== $f"; cut -c1-400 $f; done
tail -5 gpurun_out/r02_c10_bench.err | cut -c1-300
python - <<'PY'
import json
try:
    d=json.load(open('gpurun_out/r02_c10_bench_inpaint.json'))
    i=d['inpaint']; print('inpaint graph',round(i['value']),'e2e',round(i['e2e']['value']),'eager',i['eager'])
except Exception as e: print('ERR',e)
PY
