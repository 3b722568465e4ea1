# 8 GPUs: the bench line of the round-2 end state (all sections), as the driver's scaling run launches it
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29542 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/r02_c42_bench_8gpu.json 2> gpurun_out/r02_c42_bench_8gpu.err
tail -c 300 gpurun_out/r02_c42_bench_8gpu.err
