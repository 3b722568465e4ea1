# 2 GPUs: which round-2 change costs the LatentRNN data-parallel step 4 ms? (latent section only, three settings)
run() { tag=$1; shift; env "$@" timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29540 bench.py --gpus 2 --steps 10 --warmup 3 --sections latent --no-cpu-baseline > gpurun_out/r02_c40_$tag.json 2> gpurun_out/r02_c40_$tag.err; }
run base A=1
run notick IPN_TICK_PERSIST=0
run nopaircs IPN_GPF_PAIRCS=0
python - <<'PY'
import json
for t in ('base','notick','nopaircs'):
    try:
        d=json.loads(open(f'gpurun_out/r02_c40_{t}.json').read().strip().splitlines()[-1])
        print(t, round(d['ms_per_step'],3), {k:round(v['ms_per_step'],3) for k,v in d['latent_train']['modes'].items()})
    except Exception as e: print(t,'ERR',e)
PY
