"""dev: timing + per-role cycle counters of the persistent LSTM layer kernels (B=4096, H=256, T=384)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from inpaintnet_b200 import ops
from inpaintnet_b200.ops import Precision

DEV = "cuda"
prec = Precision("bf16")
H, B, T = int(os.environ.get("H", 256)), int(os.environ.get("B", 4096)), int(os.environ.get("T", 384))
NC = H // 64
g = torch.Generator().manual_seed(0)
s = 1.0 / H ** 0.5
whh = ((torch.rand(4 * H, H, generator=g) * 2 - 1) * s).to(DEV).bfloat16().contiguous()
bhh = ((torch.rand(4 * H, generator=g) * 2 - 1) * s).to(DEV)
P = (torch.randn(T * B, 4 * H, device=DEV) * 0.5).bfloat16()
hseq = torch.zeros((T + 1) * B, H, dtype=torch.bfloat16, device=DEV)
cseq = torch.zeros((T + 1) * B, H, dtype=torch.float32, device=DEV)
gates = torch.zeros(T * B, ops.lstm_gates_cols(H, True), dtype=torch.bfloat16, device=DEV)
dY = (torch.randn(T * B, H, device=DEV) * 0.1).bfloat16()
dP = torch.empty(T * B, 4 * H, dtype=torch.bfloat16, device=DEV)
ncta = (B // 128) * NC
timing = torch.zeros(ncta * 16, dtype=torch.int64, device=DEV)
ops.lib().ipn_dbg_set_timing_buffer(timing.data_ptr())


def fwd(save=True):
    ops.lstm_layer_fwd(prec, T, B, H, whh.data_ptr(), bhh.data_ptr(), P.data_ptr(), 4 * H, hseq.data_ptr(), cseq.data_ptr(),
                       gates=gates.data_ptr() if save else 0, P_blocked=1)


def bwd():
    ops.lstm_layer_bwd(prec, T, B, H, whh.data_ptr(), 0, 0, gates.data_ptr(), dY.data_ptr(), H, 0, dP.data_ptr(), 0, gates_persist=1)


for name, fn, names in (("fwd", fwd, {0: "mma_total", 1: "mma_wait_own_a", 2: "mma_wait_peer_a", 4: "epi_total", 5: "epi_wait_mma_all",
                                       6: "epi_wait_st_free", 7: "epi_work", 8: "epi_fence_signal"}),
                        ("fwd_nosave", lambda: fwd(False), None),
                        ("bwd", bwd, {4: "epi_total", 5: "wait_dy", 6: "wait_recv", 7: "e_phase", 8: "wait_mma_all", 9: "s_phase"})):
    timing.zero_()
    ops.prof_enable(True)
    for _ in range(3):
        fn()
    rep = ops.prof_report()
    ops.prof_enable(False)
    ms = {k: v["ms"] / v["launches"] for k, v in rep.items()}
    print(name, {k: "%.3f ms = %.2f us/step" % (v, v * 1e3 / T) for k, v in ms.items()})
    if names:
        t = timing.view(ncta, 16).float().mean(0).cpu().tolist()
        print("   per-step cycles:", {n: round(t[i] / T) for i, n in names.items()})
