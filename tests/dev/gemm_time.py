"""dev: time the big hoisted GEMM shapes of the train step (bf16 tcgen05 core)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from inpaintnet_b200 import ops
from inpaintnet_b200.ops import Precision, F32, BF16

prec = Precision("bf16")
DEV = "cuda"
R = 98304
def t(fn, reps=5):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
X = torch.randn(R, 1024, device=DEV).bfloat16()
W = torch.randn(1536, 1024, device=DEV).bfloat16()
P = torch.empty(R, 1536, device=DEV, dtype=torch.bfloat16)
ms = t(lambda: ops.gemm(prec.core, prec.act, R, 1536, [(X.data_ptr(), 1024, 0, W.data_ptr(), 1024, 0, 1024)], P.data_ptr(), BF16, 1536))
print("NT  P = X[98304,1024] W[1536,1024]^T : %.3f ms  %.0f TFLOP/s" % (ms, 2 * R * 1536 * 1024 / ms / 1e9))
dP = torch.randn(R, 3072, device=DEV).bfloat16()
Wi = torch.randn(3072, 1024, device=DEV).bfloat16()
dX = torch.empty(R, 1024, device=DEV, dtype=torch.bfloat16)
ms = t(lambda: ops.gemm(prec.core, prec.act, R, 1024, [(dP.data_ptr(), 3072, 0, Wi.data_ptr(), 1024, 1, 3072)], dX.data_ptr(), BF16, 1024))
print("NN  dX = dP[98304,3072] W[3072,1024]  : %.3f ms  %.0f TFLOP/s" % (ms, 2 * R * 3072 * 1024 / ms / 1e9))
G = torch.zeros(1536, 1024, device=DEV)
ms = t(lambda: ops.gemm(prec.core, prec.act, 1536, 1024, [(dP.data_ptr(), 3072, 1, X.data_ptr(), 1024, 1, R)], G.data_ptr(), F32, 1024, accumulate=ops.ATOMIC_ADD))
print("TN  dW = dP[:, :1536]^T X             : %.3f ms  %.0f TFLOP/s" % (ms, 2 * R * 1536 * 1024 / ms / 1e9))
a = torch.randn(8192, 8192, device=DEV).bfloat16(); b = torch.randn(8192, 8192, device=DEV).bfloat16()
ms = t(lambda: torch.matmul(a, b.t()))
print("cuBLAS 8192^3 (reference point)       : %.3f ms  %.0f TFLOP/s" % (ms, 2 * 8192**3 / ms / 1e9))
c = torch.empty(8192, 8192, device=DEV, dtype=torch.bfloat16)
ms = t(lambda: ops.gemm(prec.core, prec.act, 8192, 8192, [(a.data_ptr(), 8192, 0, b.data_ptr(), 8192, 0, 8192)], c.data_ptr(), BF16, 8192))
print("ours   8192^3 NT                      : %.3f ms  %.0f TFLOP/s" % (ms, 2 * 8192**3 / ms / 1e9))
