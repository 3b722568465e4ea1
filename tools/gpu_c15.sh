F="grep -v -E Warning|vectorized_gather|warnings.warn|super().__init__"
( timeout 600 python -m pytest tests -m gpu -q -x 2>&1 | $F | tail -6 ) > gpurun_out/r02_c15_gpu_tests.log
timeout 600 python bench.py --steps 10 --warmup 5 --no-cpu-baseline > gpurun_out/r02_c15_bench.json 2> gpurun_out/r02_c15_bench.err
for f in gpurun_out/r02_c15_*.log; do echo "== $f"; cut -c1-300 $f; done
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02_c15_bench.json'))
print('value',round(d['value']),'ms',round(d['ms_per_step'],3), {k:round(v['ms_per_step'],3) for k,v in d['modes'].items()})
print('inpaint',round(d['inpaint']['value']), 'latent', round(d['latent_train']['ms_per_step'],2), 'arnn', {k:round(v['ms_per_step'],2) for k,v in d['arnn_train']['modes'].items()})
for sec,k in (('kernels','gemm_umma_inproj_blocked'),('kernels','gemm_umma_nn_dgrad'),('kernels','gemm_umma_nt'),('kernels','gemm_umma_tn_wgrad')): print(k, d[sec][k])
a=d['arnn_train']['modes']['teacher_forced']['kernels']
for k in ('gemm_umma_inproj_blocked','gemm_umma_nn_dgrad','gemm_umma_tn_wgrad','gemm_umma_nt'): print('arnn',k,a[k])
i=d['inpaint']['kernels']
for k in ('gemm_umma_inproj_blocked','gemm_umma_nt'): print('inpaint',k,i[k])
PY
