F="grep -v -E Warning|vectorized_gather|warnings.warn|super().__init__"
( timeout 300 python tests/dev/lstm_persist_time.py 2>&1 | $F | tail -12 ) > gpurun_out/r02_c6_lstm_time.log
( timeout 300 python -m pytest tests/test_gpu_fullsize.py -q -s -k "free_running_backward" 2>&1 | $F | tail -12 ) > gpurun_out/r02_c6_arnn_notf_bwd.log
for f in gpurun_out/r02_c6_*.log; do echo "== $f"; cut -c1-400 $f; done
