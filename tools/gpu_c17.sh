F="grep -v -E Warning|vectorized_gather|warnings.warn|super().__init__"
( timeout 300 python -m pytest tests/test_gpu_gru.py -q -x -k lstm 2>&1 | $F | tail -4 ) > gpurun_out/r02_c17_lstm_tests.log
( timeout 300 python tests/dev/lstm_persist_time.py 2>&1 | $F | tail -8 ) > gpurun_out/r02_c17_lstm_time.log
( timeout 300 python -m pytest tests/test_gpu_fullsize.py -q -k "arnn" 2>&1 | $F | tail -4 ) > gpurun_out/r02_c17_arnn.log
timeout 300 python bench.py --sections mvae,arnn --steps 4 --no-cpu-baseline > gpurun_out/r02_c17_bench.json 2>/dev/null
for f in gpurun_out/r02_c17_*.log; do echo "== $f"; cut -c1-420 $f; done
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02_c17_bench.json'))
a=d['arnn_train']['modes']
for m in a: print(m, round(a[m]['ms_per_step'],2), {k:a[m]['kernels'][k] for k in ('lstm_layer_bwd_persist','lstm_layer_fwd_persist')})
PY
