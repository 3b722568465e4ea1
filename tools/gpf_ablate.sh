# forward layer kernel, training encoder shape (B=4096, 2 directions: CTA pairs inside the column split): step period from
# the cycle stamps with A-loader ablations (IPN_GPF_DBG 32: no fence.proxy.async per k-block, 64: cta-scope barrier wait)
for d in 0 32 64 96; do echo "=== IPN_GPF_DBG=$d"; IPN_GPF_DBG=$d B=4096 NDIR=2 timeout 200 python tests/dev/persist_time.py 2>&1 | grep "aload kb0\|aload kb7\|mma c0 issued\|dbg=" | cut -c1-120; done
