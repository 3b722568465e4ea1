# robustness: compute-sanitizer memcheck over the persistent tick-decode kernel and the PAIR + column-split layer kernels
F="grep -v -E Warning|vectorized_gather|warnings.warn|super().__init__|NUMBER"
( timeout 500 compute-sanitizer --tool memcheck --launch-timeout 120 python -m pytest "tests/test_gpu_tick_persist.py" -q -x -m gpu -k "64-512-True or 64-256-False" 2>&1 | $F | tail -15 ) > gpurun_out/r02_c51_sanitizer_tick.log
( IPN_GPF_CS=2 IPN_GPB_PAIRCS=1 timeout 500 compute-sanitizer --tool memcheck --launch-timeout 120 python -m pytest tests/test_gpu_gru.py -q -x -m gpu -k "gru_layer_fwd_bwd" 2>&1 | $F | tail -15 ) > gpurun_out/r02_c51_sanitizer_gru.log
for f in gpurun_out/r02_c51_*.log; do echo "== $f"; cut -c1-250 $f; done
