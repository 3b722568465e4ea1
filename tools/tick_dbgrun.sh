for d in 0 1 2 3; do echo "=== IPN_TICK_DBG=$d"; IPN_TICK_DBG=$d timeout 200 python tests/dev/tick_persist_counters.py 2>&1 | grep -v "vectorized_gather\|Warn\|warn" | sed -n 2,14p; done
