"""dev: wait-cycle counters of the persistent GRU backward kernel's MMA issuer (encoder-layer shape)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from inpaintnet_b200 import ops
from inpaintnet_b200.ops import Precision, F32

DEV = "cuda"
prec = Precision("bf16")
H, B, T, ndir = 512, int(os.environ.get("B", 4096)), int(os.environ.get("T", 24)), int(os.environ.get("NDIR", 2))
g = torch.Generator().manual_seed(0)
s = 1.0 / H ** 0.5
whh = [((torch.rand(3 * H, H, generator=g) * 2 - 1) * s).to(DEV).bfloat16().contiguous() for _ in range(ndir)]
bhh = [((torch.rand(3 * H, generator=g) * 2 - 1) * s).to(DEV) for _ in range(ndir)]
P = torch.randn(ndir, T * B, 3 * H, device=DEV).bfloat16()
dY = (torch.randn(T * B, ndir * H, device=DEV) * 0.1).bfloat16()
hseq = torch.zeros(ndir, (T + 1) * B, H, dtype=torch.bfloat16, device=DEV)
gates = torch.zeros(ndir, T * B, ops.gates_cols(H), dtype=torch.bfloat16, device=DEV)
y = torch.zeros(T * B, ndir * H, dtype=torch.bfloat16, device=DEV)
dP = torch.zeros(ndir, T * B, 3 * H, dtype=torch.bfloat16, device=DEV)
dGn = torch.zeros(ndir, T * B, H, dtype=torch.bfloat16, device=DEV)
ws = torch.empty(2 * 2 * B * H, dtype=torch.float32, device=DEV)
dirs = [ops.gru_dir(whh[d].data_ptr(), bhh[d].data_ptr(), hseq[d].data_ptr(), gates=gates[d].data_ptr(),
                    P=P[d].data_ptr(), ldP=3 * H, reverse=d, y_col0=d * H) for d in range(ndir)]
bd = [ops.gru_bwd_dir(whh[d].data_ptr(), hseq[d].data_ptr(), gates[d].data_ptr(), dP[d].data_ptr(), dGn[d].data_ptr(),
                      reverse=d, y_col0=d * H) for d in range(ndir)]
pk = ops.gru_layer_fwd(prec, T, B, H, dirs, y=y.data_ptr(), ld_y=ndir * H)
ncta = (B // 128) * ndir
timing = torch.zeros(ncta * 16, dtype=torch.int64, device=DEV)
torch.cuda.synchronize()
ops.lib().ipn_dbg_set_timing_buffer(timing.data_ptr())
for _ in range(3):
    ops.gru_layer_bwd(prec, T, B, H, bd, ws.data_ptr(), dY=dY.data_ptr(), ld_dy=ndir * H, persistent=pk)
torch.cuda.synchronize()
t = timing.view(ncta, 16).float()
lead = t[t[:, 0] > 0]
names = ["mma_total", "wait_tmem_free(E phase)", "wait_a_full", "wait_w_full"]
print("B=%d T=%d ndir=%d  leaders=%d  per-step kcycles:" % (B, T, ndir, lead.shape[0]),
      {n: round(v / (T - 1) / 1000, 1) for n, v in zip(names, lead.mean(0)[:4].tolist())})
