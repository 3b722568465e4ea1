run() { tag=$1; shift; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --steps 20 --warmup 5 --no-cpu-baseline "$@" > gpurun_out/r02_c41_$tag.json 2> gpurun_out/r02_c41_$tag.err; }
run inpaint_latent --sections inpaint,latent
run latent20 --sections latent
python - <<'PY'
import json
for t in ('inpaint_latent','latent20'):
    try:
        d=json.loads(open(f'gpurun_out/r02_c41_{t}.json').read().strip().splitlines()[-1])
        print(t, round(d['ms_per_step'],3), {k:round(v['ms_per_step'],3) for k,v in d['latent_train']['modes'].items()})
    except Exception as e: print(t,'ERR',e)
PY
