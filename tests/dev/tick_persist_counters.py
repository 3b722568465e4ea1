"""dev: per-role wait-cycle counters of the persistent tick-decode kernel (B=4096, eval and train mode)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from inpaintnet_b200 import ops
from inpaintnet_b200.data import SyntheticFolkDataset
from inpaintnet_b200.measure_vae import MeasureVAE

DEV = "cuda"
V, B = 64, int(sys.argv[1]) if len(sys.argv) > 1 else 4096
m = MeasureVAE(SyntheticFolkDataset(num_notes=V)).to(DEV).set_precision("bf16")
m.decoder.teacher_forcing_prob = -1.0
tok = torch.randint(0, V, (B, 24)).to(DEV)
ncta = B // 128 * 4
timing = torch.zeros(4096 * 16, dtype=torch.int64, device=DEV)
names = ["mma_total", "mma_wait_tmem_empty", "mma_wait_a_full", "mma_wait_w_full", "mma_wait_lg_empty", "mma_total_ns (globaltimer)", "aload_wait_a_free",
         "aload_wait_exchange", "epi_total", "epi_wait_tmem_full", "epi_wait_stg_free", "epi_wait_lg_full", "epi_wait_hp", "mma_issue (elected lane)", "mma_commit (elected lane)", "mma k-block iterations, all"]
for train in (False, True):
    m.train(train)
    for it in range(3):
        if it == 2:
            timing.zero_()
            ops.lib().ipn_dbg_set_timing_buffer(timing.data_ptr())
        if train:
            m(tok, train=True)
        else:
            with torch.no_grad():
                m(tok, train=False)
        torch.cuda.synchronize()
    ops.lib().ipn_dbg_set_timing_buffer(0)
    t = timing.view(-1, 16)[:ncta].float()
    mean, mx = t.mean(0).cpu().tolist(), t.max(0).values.cpu().tolist()
    print(f"train={train}: per CTA, kcycles for 24 ticks (mean / max over {ncta} CTAs)")
    for i, n in enumerate(names):
        if n != "-" and mean[i] > 0:
            print(f"   {n:22s} {mean[i] / 1e3:9.1f} {mx[i] / 1e3:9.1f}   per tick {mean[i] / 24e3:7.2f}")
    # timeline of cluster 0 / rank 0, ticks 8 and 9 (cycle stamps; role 0 MMA, 1 epilogue thread 128, 2 store warp, 3 A loader)
    PH = ["A(t+1)", "Bh(t+1)", "Bx(t)", "V(t)"]
    ev_names = {0: {k * 4 + e: f"mma {PH[k]} {n}" for k in range(4) for e, n in enumerate(["tmem free", "first operands", "issued"])},
                1: {**{ci * 4 + e: f"epi L1 c{ci} {n}" for ci in range(2) for e, n in enumerate(["hp loaded", "tmem_full", "math done", "staged"])},
                    8: "epi V lg_full", 9: "epi V token",
                    **{16 + ci * 4 + e: f"epi L0(t+1) c{ci} {n}" for ci in range(2) for e, n in enumerate(["P+hp issued", "tmem_full", "math done", "staged"])}},
                2: {k * 8 + ci * 4 + e: f"store {['y0', 'h0', 'h1'][k]}(t) c{ci} {n}" for k in range(3) for ci in range(2) for e, n in enumerate(["tile ready", "store complete", "fenced", "signalled"])},
                3: {k * 4 + e: f"aload {PH[k]} {n}" for k in range(4) for e, n in enumerate(["a_free[0]", "kb0 issued", "all issued"])}}
    trc = timing[32768:32768 + 4 * 2 * 32].view(4, 2, 32).cpu()
    rows = []
    for role in range(4):
        for tt in range(2):
            for ev in range(32):
                v = int(trc[role, tt, ev])
                if v:
                    rows.append((v, f"t={8 + tt} {ev_names[role].get(ev, str(ev))}"))
    rows.sort()
    if rows:
        t0 = rows[0][0]
        for v, n in rows:
            print(f"      {v - t0:8d}  {n}")
