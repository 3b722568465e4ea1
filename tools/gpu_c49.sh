# 8 GPUs, MeasureVAE section only: which NCCL algorithm / channel count the gradient all-reduce gets, and two variants
run() { tag=$1; shift; env "$@" timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29549 bench.py --gpus 8 --steps 20 --warmup 5 --sections mvae --no-cpu-baseline > gpurun_out/r02_c49_$tag.json 2> gpurun_out/r02_c49_$tag.err; }
run base NCCL_DEBUG=INFO NCCL_DEBUG_SUBSYS=INIT,COLL,TUNING
grep -E "NVLS|Channel|channels|Algo|algo|nChannels|Using|Ring|Tree" gpurun_out/r02_c49_base.err | grep -v "^$" | sort | uniq -c | sort -rn | head -30 > gpurun_out/r02_c49_nccl_info.txt
run ch8 NCCL_MAX_NCHANNELS=8
run nvls NCCL_ALGO=NVLS
python - <<'PY'
import json
for t in ('base','ch8','nvls'):
    try:
        d=json.loads(open(f'gpurun_out/r02_c49_{t}.json').read().strip().splitlines()[-1])
        print(t, round(d['ms_per_step'],3), round(d['value']), {k:round(v['ms_per_step'],3) for k,v in d['modes'].items()})
    except Exception as e: print(t,'ERR',e)
PY
head -30 gpurun_out/r02_c49_nccl_info.txt | cut -c1-200
