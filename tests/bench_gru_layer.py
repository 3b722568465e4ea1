"""Micro-benchmark (not a test): encoder layer-0 GRU forward, 24 steps, both directions, B=4096, H=512."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from inpaintnet_b200 import ops
from inpaintnet_b200.ops import Precision, BF16, F32

prec = Precision("bf16")
T, B, H = 24, int(os.environ.get("B", 4096)), 512
dev = "cuda"
whh = [torch.randn(3 * H, H, device=dev).mul(0.04).bfloat16() for _ in range(2)]
bhh = [torch.zeros(3 * H, device=dev) for _ in range(2)]
P = torch.randn(2, T * B, 3 * H, device=dev).bfloat16()
hseq = torch.zeros(2, (T + 1) * B, H, device=dev, dtype=torch.bfloat16)
gates = torch.empty(2, T * B, ops.gates_cols(H), device=dev, dtype=torch.bfloat16)
y = torch.empty(T * B, 2 * H, device=dev, dtype=torch.bfloat16)
mask = (torch.rand(T * B, 2 * H, device=dev) > 0.5).to(torch.uint8)
dirs = [ops.gru_dir(whh[d].data_ptr(), bhh[d].data_ptr(), hseq[d].data_ptr(), gates=gates[d].data_ptr(), P=P[d].data_ptr(),
                    ldP=3 * H, reverse=d, y_col0=d * H) for d in range(2)]
def run():
    ops.gru_layer_fwd(prec, T, B, H, dirs, y=y.data_ptr(), ld_y=2 * H, mask=mask.data_ptr(), ld_mask=2 * H, mask_scale=2.0)
for _ in range(3): run()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5): run()
e1.record(); torch.cuda.synchronize()
us = e0.elapsed_time(e1) * 1e3 / (5 * T)
print(f"DBG={os.environ.get('IPN_DBG_EPI','0')} BR={os.environ.get('IPN_GRU_BR','128')} B={B}: {us:.1f} us/step  "
      f"{2*2*B*3*H*H/us/1e6:.1f} TFLOP/s")

if os.environ.get("TIMING"):
    ncta = 4 * ((B + int(os.environ.get('IPN_GRU_BR','128')) - 1) // int(os.environ.get('IPN_GRU_BR','128'))) * 2
    buf = torch.zeros(ncta * 8, dtype=torch.int64, device=dev)
    ops.lib().ipn_dbg_set_timing_buffer(buf.data_ptr())
    ops.gru_layer_fwd(prec, T, B, H, dirs, y=y.data_ptr(), ld_y=2 * H, mask=mask.data_ptr(), ld_mask=2 * H, mask_scale=2.0,
                      s_begin=5, s_end=6)
    torch.cuda.synchronize()
    ops.lib().ipn_dbg_set_timing_buffer(None)
    t = buf.view(ncta, 8).cpu().double()
    t0 = t[:, 0].min()
    names = ["start", "setup", "tma_issued", "mma_issued", "acc_ready", "epi_w2_done", "all_done"]
    print("CTAs", ncta, "kernel span us", (t[:, 6].max() - t0).item() / 1e3)
    for i, n in enumerate(names):
        rel = (t[:, i] - t[:, 0]) / 1e3
        print(f"  {n:12s} since CTA start: mean {rel.mean():8.2f} us  min {rel.min():8.2f}  max {rel.max():8.2f}")
    st = (t[:, 0] - t0) / 1e3
    print("  CTA start times: ", [round(x, 1) for x in st.sort().values[::max(1, ncta // 16)].tolist()])
