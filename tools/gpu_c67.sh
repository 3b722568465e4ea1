# ncu full capture of the streamed-A form of the tick-decode kernel (replaces the capture of its first form)
CMD="python bench.py --steps 2 --warmup 1 --sections mvae --no-cpu-baseline"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:tick_decode_persist -c 1 -o gpurun_out/r02_c67_tick $CMD > gpurun_out/r02_c67_ncu.log 2>&1
ls -la gpurun_out/r02_c67_tick.ncu-rep; tail -2 gpurun_out/r02_c67_ncu.log
