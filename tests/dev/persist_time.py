"""dev: timing + wait-cycle counters of the persistent GRU forward kernel (env IPN_GPF_DBG selects ablations)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from inpaintnet_b200 import ops
from inpaintnet_b200.ops import Precision, F32

DEV = "cuda"
prec = Precision("bf16")
H, B, T, ndir = 512, int(os.environ.get("B", 4096)), 24, int(os.environ.get("NDIR", 2))
g = torch.Generator().manual_seed(0)
s = 1.0 / H ** 0.5
whh = [((torch.rand(3 * H, H, generator=g) * 2 - 1) * s).to(DEV).bfloat16().contiguous() for _ in range(ndir)]
bhh = [((torch.rand(3 * H, generator=g) * 2 - 1) * s).to(DEV) for _ in range(ndir)]
P = torch.randn(ndir, T * B, 3 * H, device=DEV).bfloat16()
hseq = torch.zeros(ndir, (T + 1) * B, H, dtype=torch.bfloat16, device=DEV)
gates = torch.zeros(ndir, T * B, ops.gates_cols(H), dtype=torch.bfloat16, device=DEV)
y = torch.zeros(T * B, ndir * H, dtype=torch.bfloat16, device=DEV)
NOSAVE = os.environ.get("NOSAVE", "0") != "0"      # inference: no saved gates
dirs = [ops.gru_dir(whh[d].data_ptr(), bhh[d].data_ptr(), hseq[d].data_ptr(), gates=0 if NOSAVE else gates[d].data_ptr(),
                    P=P[d].data_ptr(), ldP=3 * H, reverse=d, y_col0=d * H) for d in range(ndir)]
ncta = (B // 128) * ndir
timing = torch.zeros(65536, dtype=torch.int64, device=DEV)   # [0, 32768): per-CTA counters, [32768, ...): cycle stamps of CTA (0,0)
if os.environ.get("NOTIMING", "0") == "0":   # the wait counters cost ~150 cycles per clock64 read: off for clean launch times
    ops.lib().ipn_dbg_set_timing_buffer(timing.data_ptr())
ops.prof_enable(True)
for _ in range(4):
    ops.gru_layer_fwd(prec, T, B, H, dirs, y=y.data_ptr(), ld_y=ndir * H)
rep = ops.prof_report()
t = timing[:ncta * 16].view(ncta, 16).float().mean(0).cpu().tolist()
names = ["mma_total", "mma_wait_tmem_empty", "mma_wait_a_full", "mma_wait_w_full", "st_wait_ready", "st_store_time",
         "al_wait_a_free", "al_wait_h_stored", "epi_total", "epi_wait_tmem_full", "epi_wait_stg_free"]
print("dbg=%s B=%d ndir=%d" % (os.environ.get("IPN_GPF_DBG", "0"), B, ndir),
      {k: "%.3f ms (%d launches)" % (v["ms"] / v["launches"], v["launches"]) for k, v in rep.items()})
print("  per-step kcycles:", {n: round(v / T / 1000, 1) for n, v in zip(names, t)})

if os.environ.get("NOTIMING", "0") == "0":
    nch = 8
    ev_names = {0: {c * 3 + e: f"mma c{c} {n}" for c in range(nch) for e, n in enumerate(["tmem free", "last k-block landed", "issued"])},
                1: {c * 4 + e: f"epi c{c} {n}" for c in range(nch) for e, n in enumerate(["P loaded", "tmem_full", "math+staged", "-"])},
                2: {c * 3 + e: f"store c{c} {n}" for c in range(nch) for e, n in enumerate(["tile ready", "store complete", "signalled"])},
                3: {kb: f"aload kb{kb} ready to issue" for kb in range(8)}}
    trc = timing[32768:32768 + 4 * 2 * 32].view(4, 2, 32).cpu()
    rows = [(int(trc[r, tt, e]), f"s={8 + tt} {ev_names[r].get(e, str(e))}") for r in range(4) for tt in range(2) for e in range(32) if int(trc[r, tt, e])]
    rows.sort()
    for v, n in rows:
        print(f"      {v - rows[0][0]:8d}  {n}")
