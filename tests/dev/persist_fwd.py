"""dev check: persistent GRU forward vs the per-step path (run on the GPU box)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from inpaintnet_b200 import ops
from inpaintnet_b200.ops import Precision, F32, BF16

DEV = "cuda"
prec = Precision("bf16")


def run(H, B, T, ndir, src, use_mask, persistent, seed=0, reps=0):
    g = torch.Generator().manual_seed(seed)
    s = 1.0 / H ** 0.5
    whh = [((torch.rand(3 * H, H, generator=g) * 2 - 1) * s).to(DEV).bfloat16().contiguous() for _ in range(ndir)]
    bhh = [((torch.rand(3 * H, generator=g) * 2 - 1) * s).to(DEV) for _ in range(ndir)]
    P = torch.randn(ndir, T * B, 3 * H, generator=g).to(DEV).bfloat16()
    Pb = torch.randn(ndir, B, 3 * H, generator=g).to(DEV).bfloat16()
    V = 11
    table = torch.randn(ndir, V, 3 * H, generator=g).to(DEV)
    tok = torch.randint(0, V, (T * B,), generator=g).int().to(DEV)
    pvec = torch.randn(ndir, 3 * H, generator=g).to(DEV)
    h0 = (torch.randn(ndir, B, H, generator=g) * 0.5).to(DEV).bfloat16()
    hseq = torch.zeros(ndir, (T + 1) * B, H, dtype=torch.bfloat16, device=DEV)
    hseq[0, :B] = h0[0]
    if ndir == 2:
        hseq[1, T * B:] = h0[1]
    gates = torch.zeros(ndir, T * B, ops.gates_cols(H), dtype=torch.bfloat16, device=DEV)
    y = torch.zeros(T * B, ndir * H, dtype=torch.bfloat16, device=DEV)
    mask = (torch.rand(T * B, ndir * H, generator=g) > 0.5).to(torch.uint8).to(DEV)
    fin = torch.zeros(B, ndir * H, dtype=torch.float32, device=DEV)
    dirs = []
    for d in range(ndir):
        kw = {}
        if src == "P":
            kw = dict(P=P[d].data_ptr(), ldP=3 * H)
        elif src == "table":
            kw = dict(table=table[d].data_ptr(), ld_table=3 * H, tok=tok.data_ptr(), table_rows=V)
        elif src == "pvec":
            kw = dict(pvec=pvec[d].data_ptr())
        elif src == "bcast+table":
            kw = dict(P=Pb[d].data_ptr(), ldP=3 * H, P_bcast=1, table=table[d].data_ptr(), ld_table=3 * H, tok=tok.data_ptr())
        dirs.append(ops.gru_dir(whh[d].data_ptr(), bhh[d].data_ptr(), hseq[d].data_ptr(), gates=gates[d].data_ptr(),
                                reverse=d, y_col0=d * H, final_col0=d * H, **kw))
    def call():
        return ops.gru_layer_fwd(prec, T, B, H, dirs, y=y.data_ptr(), ld_y=ndir * H, mask=mask.data_ptr() if use_mask else 0,
                                 ld_mask=ndir * H, mask_scale=2.0, final_out=fin.data_ptr(), final_dt=F32, ld_final=ndir * H,
                                 persistent=persistent)
    used = call()
    torch.cuda.synchronize()
    ms = None
    if reps:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            call()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
    return used, y.float().cpu(), hseq.float().cpu(), fin.cpu(), ms


ok = True
cases = [(64, 128, 3, 1, "P", False), (128, 256, 5, 2, "table", True), (512, 256, 4, 2, "P", True),
         (512, 128, 4, 1, "pvec", False), (256, 384, 6, 1, "bcast+table", True), (512, 4096, 24, 2, "P", False)]
if len(sys.argv) > 1:
    cases = cases[:int(sys.argv[1])]
for (H, B, T, ndir, src, um) in cases:
    big = B >= 4096
    u1, y1, h1, f1, ms1 = run(H, B, T, ndir, src, um, True, reps=5 if big else 0)
    u0, y0, h0, f0, ms0 = run(H, B, T, ndir, src, um, False, reps=5 if big else 0)
    assert u1 and not u0, (u1, u0)
    dy, dh, df = (y1 - y0).abs().max().item(), (h1 - h0).abs().max().item(), (f1 - f0).abs().max().item()
    good = dy < 3e-2 and dh < 3e-2 and df < 3e-2 and torch.isfinite(y1).all().item()
    ok = ok and good
    print(f"H={H} B={B} T={T} ndir={ndir} src={src} mask={um}: max|dy|={dy:.4f} max|dh|={dh:.4f} max|dfin|={df:.4f} "
          f"{'OK' if good else 'FAIL'}" + (f"  persist {ms1:.3f} ms  per-step {ms0:.3f} ms" if ms1 else ""), flush=True)
print("ALL OK" if ok else "FAILED")
