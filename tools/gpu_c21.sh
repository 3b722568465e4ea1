for v in 1 0; do
IPN_SIDE_STREAM=$v timeout 300 python bench.py --steps 10 --warmup 5 --no-cpu-baseline --sections mvae > gpurun_out/r02_c21_bench_side$v.json 2>/dev/null
python - <<PY
import json
d=json.load(open('gpurun_out/r02_c21_bench_side$v.json'))
print('IPN_SIDE_STREAM=$v value',round(d['value']),'ms',round(d['ms_per_step'],3), {k:round(x['ms_per_step'],3) for k,x in d['modes'].items()}, 'wgrad', d['kernels']['gemm_umma_tn_wgrad'], 'bwd', d['kernels']['gru_layer_bwd_persist']['ms'])
PY
done
