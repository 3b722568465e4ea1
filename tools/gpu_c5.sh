F="grep -v -E Warning|vectorized_gather|warnings.warn|super().__init__"
( timeout 300 python -m pytest tests/test_gpu_fullsize.py tests/test_gpu_arnn.py -q -s -k arnn 2>&1 | $F | tail -15 ) > gpurun_out/r02_c5_arnn.log
( IPN_GPF_CS=1 timeout 120 python tests/dev/persist_fwd.py 2>&1 | $F | tail -15 ) > gpurun_out/r02_c5_cs_dev.log
( IPN_GPF_CS=1 timeout 300 python -m pytest tests/test_gpu_gru.py tests/test_gpu_mvae.py -q -x 2>&1 | $F | tail -8 ) > gpurun_out/r02_c5_cs_fwd.log
( IPN_GPF_CS=1 IPN_GPB_CS=1 timeout 300 python -m pytest tests/test_gpu_gru.py tests/test_gpu_mvae.py -q -x 2>&1 | $F | tail -8 ) > gpurun_out/r02_c5_cs_bwd.log
timeout 200 python bench.py --sections mvae,arnn --steps 6 --no-cpu-baseline > gpurun_out/r02_c5_bench_base.json 2>/dev/null
IPN_GPF_CS=1 timeout 200 python bench.py --sections mvae --steps 6 --no-cpu-baseline > gpurun_out/r02_c5_bench_csf.json 2>/dev/null
IPN_GPF_CS=1 IPN_GPB_CS=1 timeout 200 python bench.py --sections mvae --steps 6 --no-cpu-baseline > gpurun_out/r02_c5_bench_csfb.json 2>/dev/null
for f in gpurun_out/r02_c5_*.log; do echo "== $f"; cut -c1-300 $f; done
