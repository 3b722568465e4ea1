# ncu evidence for the round-2 end state: launch list of two MeasureVAE train steps (one teacher-forced, one argmax)
# and full captures of the persistent tick-decode kernel and of an encoder-layer forward launch
F="grep -v -E Warning|vectorized_gather|warnings.warn|super().__init__"
CMD="python bench.py --steps 2 --warmup 1 --sections mvae --no-cpu-baseline"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/r02_c37_launches.csv $CMD > gpurun_out/r02_c37_ncu_a.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:tick_decode_persist -c 1 -o gpurun_out/r02_c37_tick $CMD > gpurun_out/r02_c37_ncu_b.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gru_persist_fwd -s 4 -c 2 -o gpurun_out/r02_c37_grufwd $CMD > gpurun_out/r02_c37_ncu_c.log 2>&1
ls -la gpurun_out/r02_c37_*; tail -2 gpurun_out/r02_c37_ncu_b.log
