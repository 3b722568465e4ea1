F="grep -v -E Warning|vectorized_gather|warnings.warn|super().__init__"
( timeout 600 python -m pytest tests -m gpu -q -x 2>&1 | $F | tail -6 ) > gpurun_out/r02_c19_gpu_tests.log
timeout 600 python bench.py --steps 10 --warmup 5 --no-cpu-baseline --sections mvae,latent,inpaint > gpurun_out/r02_c19_bench.json 2> gpurun_out/r02_c19_bench.err
for f in gpurun_out/r02_c19_*.log; do echo "== $f"; cut -c1-300 $f; done
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02_c19_bench.json'))
print('value',round(d['value']),'ms',round(d['ms_per_step'],3), {k:round(v['ms_per_step'],3) for k,v in d['modes'].items()})
print('inpaint',round(d['inpaint']['value']), 'latent', round(d['latent_train']['ms_per_step'],2))
for k in ('gru_prep_p','sum_slots','gemm_simt','gemm_umma_inproj_blocked','gru_layer_fwd_persist'): print(k, d['kernels'].get(k))
PY
