"""dev: timing of GRU layer fwd+bwd, persistent vs per-step (B=4096, H=512, T=24, bidirectional)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from inpaintnet_b200 import ops
from inpaintnet_b200.ops import Precision, F32

DEV = "cuda"
prec = Precision("bf16")
H, B, T, ndir = 512, int(os.environ.get("B", 4096)), 24, 2
g = torch.Generator().manual_seed(0)
s = 1.0 / H ** 0.5
whh = [((torch.rand(3 * H, H, generator=g) * 2 - 1) * s).to(DEV).bfloat16().contiguous() for _ in range(ndir)]
bhh = [((torch.rand(3 * H, generator=g) * 2 - 1) * s).to(DEV) for _ in range(ndir)]
P = torch.randn(ndir, T * B, 3 * H, device=DEV).bfloat16()
dY = (torch.randn(T * B, ndir * H, device=DEV) * 0.1).bfloat16()
dhn = torch.randn(ndir, B, H, device=DEV) * 0.1
res = {}
for persistent in (True, False):
    hseq = torch.zeros(ndir, (T + 1) * B, H, dtype=torch.bfloat16, device=DEV)
    gates = torch.zeros(ndir, T * B, ops.gates_cols(H), dtype=torch.bfloat16, device=DEV)
    y = torch.zeros(T * B, ndir * H, dtype=torch.bfloat16, device=DEV)
    dP = torch.zeros(ndir, T * B, 3 * H, dtype=torch.bfloat16, device=DEV)
    dGn = torch.zeros(ndir, T * B, H, dtype=torch.bfloat16, device=DEV)
    dh0 = torch.zeros(ndir, B, H, dtype=torch.float32, device=DEV)
    ws = torch.empty(2 * 2 * B * H, dtype=torch.float32, device=DEV)
    dirs = [ops.gru_dir(whh[d].data_ptr(), bhh[d].data_ptr(), hseq[d].data_ptr(), gates=gates[d].data_ptr(),
                        P=P[d].data_ptr(), ldP=3 * H, reverse=d, y_col0=d * H) for d in range(ndir)]
    bd = [ops.gru_bwd_dir(whh[d].data_ptr(), hseq[d].data_ptr(), gates[d].data_ptr(), dP[d].data_ptr(), dGn[d].data_ptr(),
                          dh_n=dhn[d].data_ptr(), ld_dhn=H, dh0=dh0[d].data_ptr(), dh0_dt=F32, ld_dh0=H, reverse=d,
                          y_col0=d * H) for d in range(ndir)]
    ops.prof_enable(False)
    for rep in range(3):
        if rep == 2:
            ops.prof_enable(True)
        pk = ops.gru_layer_fwd(prec, T, B, H, dirs, y=y.data_ptr(), ld_y=ndir * H, persistent=persistent)
        ops.gru_layer_bwd(prec, T, B, H, bd, ws.data_ptr(), dY=dY.data_ptr(), ld_dy=ndir * H, persistent=pk)
    rep = ops.prof_report()
    print("persistent" if persistent else "per-step", {k: "%.3f ms x%d" % (v["ms"], v["launches"]) for k, v in rep.items()}, flush=True)
    res[persistent] = (dP.float().cpu(), dGn.float().cpu(), dh0.cpu())
for name, a, b in zip(("dP", "dGn", "dh0"), res[True], res[False]):
    d = (a - b).abs().max().item()
    print(f"{name}: max|persist - per-step| = {d:.5f}  (max|ref| {b.abs().max().item():.4f})")
