timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 2000 -c 700 --csv --log-file gpurun_out/r02_c16_launches_inpaint.csv python bench.py --steps 2 --warmup 3 --sections inpaint --no-cpu-baseline > /dev/null 2>&1
ls -la gpurun_out/r02_c16*
