for d in 0; do echo "=== IPN_TICK_DBG=$d"; IPN_TICK_DBG=$d timeout 200 python tests/dev/tick_persist_counters.py 2>&1 | grep -v "vectorized_gather\|Warn\|warn"; done
