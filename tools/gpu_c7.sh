F="grep -v -E Warning|vectorized_gather|warnings.warn|super().__init__"
( timeout 300 python -m pytest tests/test_gpu_gru.py -q -x -k lstm 2>&1 | $F | tail -6 ) > gpurun_out/r02_c7_lstm_tests.log
( timeout 300 python tests/dev/lstm_persist_time.py 2>&1 | $F | tail -12 ) > gpurun_out/r02_c7_lstm_time.log
( timeout 300 python -m pytest tests/test_gpu_fullsize.py tests/test_gpu_arnn.py -q -s -k "arnn" 2>&1 | $F | tail -8 ) > gpurun_out/r02_c7_arnn.log
for f in gpurun_out/r02_c7_*.log; do echo "== $f"; cut -c1-400 $f; done
