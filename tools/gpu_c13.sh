timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 8 --steps 10 --warmup 5 > gpurun_out/r02_c13_bench_8gpu.json 2> gpurun_out/r02_c13_bench_8gpu.err
tail -3 gpurun_out/r02_c13_bench_8gpu.err | cut -c1-300
python - <<'PY'
import json
try:
    d=json.load(open('gpurun_out/r02_c13_bench_8gpu.json'))
    print('mvae',round(d['value']), d['ms_per_step'], 'e2e', round(d['e2e']['value'])); i=d['inpaint']; print('inpaint graph',round(i['value']),'e2e',round(i['e2e']['value']),'eager',round(i['eager']['value']),round(i['eager']['e2e_value']))
    l=d['latent_train']; print('latent',{k:(round(v['value']),round(v['ms_per_step'],2)) for k,v in l['modes'].items()})
    a=d['arnn_train']; print('arnn',{k:(round(v['value']),round(v['ms_per_step'],2)) for k,v in a['modes'].items()})
except Exception as e: print('ERR',e)
PY
