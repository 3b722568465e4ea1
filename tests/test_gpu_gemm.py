"""GPU parity of the two GEMM cores (fp32 SIMT, bf16 tcgen05) through the C-ABI (ipn_gemm)
against a plain torch fp32 matmul of the same (bf16-rounded) operands."""
import pytest
import torch

pytestmark = pytest.mark.gpu

from inpaintnet_b200 import ops
from inpaintnet_b200.ops import (F32, BF16, CORE_SIMT, CORE_UMMA, ACT_NONE, ACT_SELU, ACT_RELU, STORE, ATOMIC_ADD,
                                 RMW_ADD, MUL_SELU_GRAD, MUL_KEEP_MASK)

DEV = "cuda"


def _mk(rows, cols, dt, seed, ld=None):
    g = torch.Generator().manual_seed(seed)
    ld = ld or (((cols + 7) // 8) * 8 if dt == BF16 else cols)
    t = torch.zeros(rows, ld)
    t[:, :cols] = torch.randn(rows, cols, generator=g)
    t = t.to(DEV).to(torch.bfloat16 if dt == BF16 else torch.float32)
    return t


def _run(core, dt, M, N, K, tA, tB, bias=False, act=ACT_NONE, out_dt=F32, accumulate=STORE, split_k=0, K2=0):
    # operands stored per layout
    A = _mk(K, M, dt, 1) if tA else _mk(M, K, dt, 1)
    B = _mk(K, N, dt, 2) if tB else _mk(N, K, dt, 2)
    segs = [(A.data_ptr(), A.shape[1], tA, B.data_ptr(), B.shape[1], tB, K)]
    def logical(X, trans, rows, k):
        return X.float()[:k, :rows].t() if trans else X.float()[:rows, :k]

    ref = logical(A, tA, M, K) @ logical(B, tB, N, K).t()
    keep = [A, B]
    if K2:
        A2 = _mk(K2, M, dt, 3) if tA else _mk(M, K2, dt, 3)
        B2 = _mk(K2, N, dt, 4) if tB else _mk(N, K2, dt, 4)
        segs.append((A2.data_ptr(), A2.shape[1], tA, B2.data_ptr(), B2.shape[1], tB, K2))
        ref = ref + logical(A2, tA, M, K2) @ logical(B2, tB, N, K2).t()
        keep += [A2, B2]
    bvec = None
    if bias:
        bvec = torch.randn(N, generator=torch.Generator().manual_seed(5)).to(DEV)
        ref = ref + bvec
    if act == ACT_SELU:
        ref = torch.nn.functional.selu(ref)
    elif act == ACT_RELU:
        ref = torch.relu(ref)
    tdt = torch.bfloat16 if out_dt == BF16 else torch.float32
    if accumulate == STORE:
        out = torch.full((M, N), float("nan"), device=DEV, dtype=tdt)
    else:
        out = torch.ones(M, N, device=DEV, dtype=tdt)
        ref = ref + 1.0
    ops.gemm(core, dt, M, N, segs, out.data_ptr(), out_dt, N, bias=bvec.data_ptr() if bias else 0, act=act,
             accumulate=accumulate, split_k=split_k)
    torch.cuda.synchronize()
    return out.float(), ref


SHAPES = [(128, 128, 64), (200, 136, 72), (37, 19, 10), (256, 64, 512), (130, 384, 200), (64, 1536, 16)]


@pytest.mark.parametrize("M,N,K", SHAPES)
@pytest.mark.parametrize("tA,tB", [(0, 0), (0, 1), (1, 1), (1, 0)])
def test_simt_fp32(M, N, K, tA, tB):
    out, ref = _run(CORE_SIMT, F32, M, N, K, tA, tB, bias=True, act=ACT_SELU)
    assert torch.allclose(out, ref, atol=1e-4, rtol=1e-4), (out - ref).abs().max()


def test_simt_modes():
    out, ref = _run(CORE_SIMT, F32, 100, 70, 333, 1, 1, accumulate=ATOMIC_ADD, split_k=4)
    assert torch.allclose(out, ref, atol=1e-3, rtol=1e-4)
    out, ref = _run(CORE_SIMT, F32, 100, 70, 50, 0, 0, accumulate=RMW_ADD, K2=30)
    assert torch.allclose(out, ref, atol=1e-4, rtol=1e-4)
    out, ref = _run(CORE_SIMT, BF16, 100, 72, 56, 0, 1, bias=True, out_dt=BF16)
    assert torch.allclose(out, ref, atol=0.1, rtol=2e-2)


UMMA_SHAPES = [(128, 128, 64), (128, 64, 128), (256, 256, 256), (200, 136, 72), (1000, 1536, 512), (96, 40, 520),
               (4, 32, 64), (333, 512, 1024)]


@pytest.mark.parametrize("M,N,K", UMMA_SHAPES)
def test_umma_nt(M, N, K):
    out, ref = _run(CORE_UMMA, BF16, M, N, K, 0, 0, bias=True)
    err = (out - ref).abs().max().item()
    assert err <= 2e-3 * max(1.0, ref.abs().max().item()), err


@pytest.mark.parametrize("M,N,K", UMMA_SHAPES)
def test_umma_nn_dgrad_layout(M, N, K):
    out, ref = _run(CORE_UMMA, BF16, M, N, K, 0, 1)
    err = (out - ref).abs().max().item()
    assert err <= 2e-3 * max(1.0, ref.abs().max().item()), err


@pytest.mark.parametrize("M,N,K", UMMA_SHAPES)
def test_umma_tn_wgrad_layout(M, N, K):
    out, ref = _run(CORE_UMMA, BF16, M, N, K, 1, 1, accumulate=ATOMIC_ADD)
    err = (out - ref).abs().max().item()
    assert err <= 2e-3 * max(1.0, ref.abs().max().item()), err


def test_umma_two_segments_and_bf16_out():
    out, ref = _run(CORE_UMMA, BF16, 300, 256, 192, 0, 1, K2=128, out_dt=BF16)
    err = (out - ref).abs().max().item()
    assert err <= 1e-2 * max(1.0, ref.abs().max().item()), err
    out, ref = _run(CORE_UMMA, BF16, 512, 512, 4096, 1, 1, accumulate=ATOMIC_ADD, split_k=8)
    err = (out - ref).abs().max().item()
    assert err <= 2e-3 * max(1.0, ref.abs().max().item()), err


def test_rowmap_split_and_mul():
    M, N, K = 24, 32, 16
    A, B = _mk(M, K, F32, 1), _mk(N, K, F32, 2)
    # row r -> (r % 4) * 6*N + (r // 4) * N   (a transposition of a 6 x 4 grid of rows)
    out = torch.zeros(M, N, device=DEV)
    ops.gemm(CORE_SIMT, F32, M, N, [(A.data_ptr(), K, 0, B.data_ptr(), K, 0, K)], out.data_ptr(), F32, N,
             rowmap=(1 << 30, 4, 0, N, 6 * N))
    ref = (A @ B.t()).view(6, 4, N).transpose(0, 1).reshape(M, N)
    torch.cuda.synchronize()
    assert torch.allclose(out, ref, atol=1e-4)
    # column split into two separate buffers
    out2 = torch.zeros(2, M, 16, device=DEV)
    ops.gemm(CORE_SIMT, F32, M, N, [(A.data_ptr(), K, 0, B.data_ptr(), K, 0, K)], out2.data_ptr(), F32, 16,
             split_cols=16, split_stride=M * 16)
    torch.cuda.synchronize()
    full = A @ B.t()
    assert torch.allclose(out2[0], full[:, :16], atol=1e-4) and torch.allclose(out2[1], full[:, 16:], atol=1e-4)
    # SELU-grad multiplier from a saved output, and keep-mask multiplier
    y = torch.nn.functional.selu(_mk(M, N, F32, 9))
    out3 = torch.zeros(M, N, device=DEV)
    ops.gemm(CORE_SIMT, F32, M, N, [(A.data_ptr(), K, 0, B.data_ptr(), K, 0, K)], out3.data_ptr(), F32, N,
             mul=(y.data_ptr(), F32, N, MUL_SELU_GRAD, 1.0))
    torch.cuda.synchronize()
    pre = torch.where(y > 0, y / 1.0507009873554805, torch.log1p(y / (1.0507009873554805 * 1.6732632423543772)))
    pre.requires_grad_()
    torch.nn.functional.selu(pre).backward(full)
    assert torch.allclose(out3, pre.grad, atol=1e-4, rtol=1e-4)
    m = (torch.rand(M, N, device=DEV) > 0.5).to(torch.uint8)
    out4 = torch.zeros(M, N, device=DEV)
    ops.gemm(CORE_SIMT, F32, M, N, [(A.data_ptr(), K, 0, B.data_ptr(), K, 0, K)], out4.data_ptr(), F32, N,
             mul=(m.data_ptr(), 2, N, MUL_KEEP_MASK, 2.0))
    torch.cuda.synchronize()
    assert torch.allclose(out4, full * m.float() * 2.0, atol=1e-4)
