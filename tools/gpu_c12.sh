timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 6 --warmup 3 --sections mvae,inpaint > gpurun_out/r02_c12_bench_2gpu.json 2> gpurun_out/r02_c12_bench_2gpu.err
tail -3 gpurun_out/r02_c12_bench_2gpu.err | cut -c1-300
python - <<'PY'
import json
try:
    d=json.load(open('gpurun_out/r02_c12_bench_2gpu.json'))
    print(round(d['value']), d['ms_per_step']); i=d['inpaint']; print('inpaint graph',round(i['value']),'e2e',round(i['e2e']['value']),'eager',round(i['eager']['value']),round(i['eager']['e2e_value']))
except Exception as e: print('ERR',e)
PY
