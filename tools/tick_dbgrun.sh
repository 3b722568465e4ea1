# IPN_TICK_DBG ablations of the persistent tick-decode kernel (1 no weight loads, 2 no MMAs, 32 no beat-projection loads,
# 64 no token-table gather, 128 no h_prev loads; 16 adds the wait counters): per-tick cycles and the L0-epilogue stamps
for d in 0 32 64 128 224; do echo "=== IPN_TICK_DBG=$d"; IPN_TICK_DBG=$d timeout 200 python tests/dev/tick_persist_counters.py 2>&1 | grep -v "vectorized_gather\|Warn\|warn" | grep -m8 "train=False\|mma_total  \|t=8 epi V token\|t=8 epi L0(t+1) c0"; done
