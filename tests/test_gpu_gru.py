"""GPU parity of the GRU / LSTM layer drivers (fused step kernels) against the CPU oracle."""
import pytest
import torch

pytestmark = pytest.mark.gpu

from inpaintnet_b200 import ops
from inpaintnet_b200.ops import Precision, F32, BF16
from oracle import inpaintnet_oracle as O

DEV = "cuda"


def _weights(H, I, seed, G=3):
    g = torch.Generator().manual_seed(seed)
    s = 1.0 / H ** 0.5
    return dict(w_ih=(torch.rand(G * H, I, generator=g) * 2 - 1) * s, w_hh=(torch.rand(G * H, H, generator=g) * 2 - 1) * s,
                b_ih=(torch.rand(G * H, generator=g) * 2 - 1) * s, b_hh=(torch.rand(G * H, generator=g) * 2 - 1) * s)


def _to_dev(t, prec):
    return t.to(DEV).to(prec.tdt).contiguous()


@pytest.mark.parametrize("prec_name,H,B,T", [("fp32", 32, 5, 7), ("fp32", 40, 70, 3), ("bf16", 64, 130, 6),
                                              ("bf16", 512, 256, 4), ("bf16", 72, 9, 5), ("bf16", 128, 384, 5),
                                              ("bf16", 256, 128, 3)])
def test_gru_layer_fwd_bwd_bidirectional(prec_name, H, B, T):
    prec = Precision(prec_name)
    I = 24
    W = [_weights(H, I, 10 + d) for d in range(2)]
    g = torch.Generator().manual_seed(3)
    x = torch.randn(B, T, I, generator=g)
    h0 = torch.randn(2, B, H, generator=g) * 0.5
    keep = (torch.rand(B, T, 2 * H, generator=g) > 0.5).float()
    p_drop = 0.5
    # ---------------- oracle (CPU, autograd) ----------------
    Wr = [{k: v.clone().requires_grad_() for k, v in w.items()} for w in W]
    if prec_name == "bf16":  # the kernel sees bf16-rounded operands
        rnd = lambda t: t.to(torch.bfloat16).float()
    else:
        rnd = lambda t: t
    xr = x.clone().requires_grad_()
    h0r = h0.clone().requires_grad_()
    outs, hns = [], []
    for d in range(2):
        o, hn = O.gru_layer_dir(rnd(xr), rnd(h0r[d]), rnd(Wr[d]["w_ih"]), rnd(Wr[d]["w_hh"]), Wr[d]["b_ih"], Wr[d]["b_hh"],
                                reverse=(d == 1))
        outs.append(o)
        hns.append(hn)
    y_ref = torch.cat(outs, 2) * keep / (1 - p_drop)
    gy = torch.randn(B, T, 2 * H, generator=g)
    ghn = torch.randn(2, B, H, generator=g)
    (y_ref * gy).sum().add((torch.stack(hns) * ghn).sum()).backward()
    # ---------------- library ----------------
    es = prec.es
    P = torch.empty(2, T * B, 3 * H, dtype=prec.tdt, device=DEV)
    x_tm = _to_dev(x.transpose(0, 1).reshape(T * B, I), prec)            # time-major rows t*B+b
    wih = [_to_dev(W[d]["w_ih"], prec) for d in range(2)]
    whh = [_to_dev(W[d]["w_hh"], prec) for d in range(2)]
    bih = [W[d]["b_ih"].to(DEV) for d in range(2)]
    bhh = [W[d]["b_hh"].to(DEV) for d in range(2)]
    for d in range(2):
        ops.gemm(prec.core, prec.act, T * B, 3 * H, [(x_tm.data_ptr(), I, 0, wih[d].data_ptr(), I, 0, I)],
                 P[d].data_ptr(), prec.act, 3 * H, bias=bih[d].data_ptr())
    hseq = torch.zeros(2, (T + 1) * B, H, dtype=prec.tdt, device=DEV)
    hseq[0, :B] = _to_dev(h0[0], prec)
    hseq[1, T * B:] = _to_dev(h0[1], prec)
    gates = torch.empty(2, T * B, ops.gates_cols(H), dtype=prec.tdt, device=DEV)
    y = torch.empty(T * B, 2 * H, dtype=prec.tdt, device=DEV)
    mask = keep.transpose(0, 1).reshape(T * B, 2 * H).to(torch.uint8).to(DEV).contiguous()
    fin = torch.empty(B, 2 * H, dtype=torch.float32, device=DEV)
    dirs = [ops.gru_dir(whh[d].data_ptr(), bhh[d].data_ptr(), hseq[d].data_ptr(), gates=gates[d].data_ptr(),
                        P=P[d].data_ptr(), ldP=3 * H, reverse=d, y_col0=d * H, final_col0=d * H) for d in range(2)]
    pk = ops.gru_layer_fwd(prec, T, B, H, dirs, y=y.data_ptr(), ld_y=2 * H, mask=mask.data_ptr(), ld_mask=2 * H,
                           mask_scale=1 / (1 - p_drop), final_out=fin.data_ptr(), final_dt=F32, ld_final=2 * H)
    assert pk == (prec_name == "bf16" and H % 64 == 0 and H <= 512 and B % 128 == 0)
    torch.cuda.synchronize()
    tol = dict(atol=2e-5, rtol=1e-4) if prec_name == "fp32" else dict(atol=3e-2, rtol=3e-2)
    y_lib = y.float().cpu().view(T, B, 2 * H).transpose(0, 1)
    assert torch.allclose(y_lib, y_ref.detach(), **tol), (y_lib - y_ref).abs().max()
    hn_lib = fin.cpu()
    assert torch.allclose(hn_lib[:, :H], hns[0].detach(), **tol)
    assert torch.allclose(hn_lib[:, H:], hns[1].detach(), **tol)
    # ---------------- backward ----------------
    dY = _to_dev(gy.transpose(0, 1).reshape(T * B, 2 * H), prec)
    dhn = ghn.to(DEV).contiguous()
    dP = torch.empty(2, T * B, 3 * H, dtype=prec.tdt, device=DEV)
    dGn = torch.empty(2, T * B, H, dtype=prec.tdt, device=DEV)
    dh0 = torch.empty(2, B, H, dtype=torch.float32, device=DEV)
    ws = torch.empty(2 * 2 * B * H, dtype=torch.float32, device=DEV)
    bd = [ops.gru_bwd_dir(whh[d].data_ptr(), hseq[d].data_ptr(), gates[d].data_ptr(), dP[d].data_ptr(), dGn[d].data_ptr(),
                          dh_n=dhn[d].data_ptr(), ld_dhn=H, dh0=dh0[d].data_ptr(), dh0_dt=F32, ld_dh0=H, reverse=d,
                          y_col0=d * H) for d in range(2)]
    ops.gru_layer_bwd(prec, T, B, H, bd, ws.data_ptr(), dY=dY.data_ptr(), ld_dy=2 * H, mask=mask.data_ptr(),
                      ld_mask=2 * H, mask_scale=1 / (1 - p_drop), persistent=pk)
    # hoisted weight gradients
    for d in range(2):
        gW_hh = torch.zeros(3 * H, H, device=DEV)
        gW_ih = torch.zeros(3 * H, I, device=DEV)
        gb_ih = torch.zeros(3 * H, device=DEV)
        hprev = hseq[d].data_ptr() + (es * B * H if d == 1 else 0)
        ops.gemm(prec.core, prec.act, 2 * H, H, [(dP[d].data_ptr(), 3 * H, 1, hprev, H, 1, T * B)], gW_hh.data_ptr(), F32, H,
                 accumulate=ops.ATOMIC_ADD)
        ops.gemm(prec.core, prec.act, H, H, [(dGn[d].data_ptr(), H, 1, hprev, H, 1, T * B)],
                 gW_hh.data_ptr() + 4 * 2 * H * H, F32, H, accumulate=ops.ATOMIC_ADD)
        ops.gemm(prec.core, prec.act, 3 * H, I, [(dP[d].data_ptr(), 3 * H, 1, x_tm.data_ptr(), I, 1, T * B)],
                 gW_ih.data_ptr(), F32, I, accumulate=ops.ATOMIC_ADD)
        ops.colsum(dP[d].data_ptr(), prec.act, 3 * H, T * B, 3 * H, gb_ih.data_ptr())
        torch.cuda.synchronize()
        gt = dict(atol=1e-4, rtol=1e-3) if prec_name == "fp32" else dict(atol=6e-2, rtol=6e-2)
        scale = max(1.0, Wr[d]["w_hh"].grad.abs().max().item())
        assert (gW_hh.cpu() - Wr[d]["w_hh"].grad).abs().max() <= gt["atol"] * scale, "dW_hh"
        assert (gW_ih.cpu() - Wr[d]["w_ih"].grad).abs().max() <= gt["atol"] * max(1.0, Wr[d]["w_ih"].grad.abs().max().item()), "dW_ih"
        assert (gb_ih.cpu() - Wr[d]["b_ih"].grad).abs().max() <= gt["atol"] * max(1.0, Wr[d]["b_ih"].grad.abs().max().item()), "db_ih"
        assert (dh0[d].cpu() - h0r.grad[d]).abs().max() <= gt["atol"] * max(1.0, h0r.grad.abs().max().item()), "dh0"
    # input gradient dX = sum_d dP_d W_ih_d
    dX = torch.empty(T * B, I, dtype=torch.float32, device=DEV)
    ops.gemm(prec.core, prec.act, T * B, I, [(dP[0].data_ptr(), 3 * H, 0, wih[0].data_ptr(), I, 1, 3 * H),
                                            (dP[1].data_ptr(), 3 * H, 0, wih[1].data_ptr(), I, 1, 3 * H)],
             dX.data_ptr(), F32, I)
    torch.cuda.synchronize()
    dx_ref = xr.grad.transpose(0, 1).reshape(T * B, I)
    assert (dX.cpu() - dx_ref).abs().max() <= gt["atol"] * max(1.0, dx_ref.abs().max().item()), "dX"


@pytest.mark.parametrize("prec_name,H,B,T", [("fp32", 32, 6, 9), ("bf16", 64, 140, 8), ("bf16", 256, 64, 5)])
def test_lstm_layer_fwd_bwd(prec_name, H, B, T):
    prec = Precision(prec_name)
    I = 16
    W = _weights(H, I, 21, G=4)
    g = torch.Generator().manual_seed(4)
    x = torch.randn(B, T, I, generator=g)
    rnd = (lambda t: t.to(torch.bfloat16).float()) if prec_name == "bf16" else (lambda t: t)
    Wr = {k: v.clone().requires_grad_() for k, v in W.items()}
    xr = x.clone().requires_grad_()
    y_ref = O.lstm_layer(rnd(xr), rnd(Wr["w_ih"]), rnd(Wr["w_hh"]), Wr["b_ih"], Wr["b_hh"])
    gy = torch.randn(B, T, H, generator=g)
    (y_ref * gy).sum().backward()
    x_tm = _to_dev(x.transpose(0, 1).reshape(T * B, I), prec)
    wih, whh = _to_dev(W["w_ih"], prec), _to_dev(W["w_hh"], prec)
    bih, bhh = W["b_ih"].to(DEV), W["b_hh"].to(DEV)
    P = torch.empty(T * B, 4 * H, dtype=prec.tdt, device=DEV)
    ops.gemm(prec.core, prec.act, T * B, 4 * H, [(x_tm.data_ptr(), I, 0, wih.data_ptr(), I, 0, I)], P.data_ptr(), prec.act,
             4 * H, bias=bih.data_ptr())
    hseq = torch.zeros((T + 1) * B, H, dtype=prec.tdt, device=DEV)
    cseq = torch.zeros((T + 1) * B, H, dtype=torch.float32, device=DEV)
    gates = torch.empty(T * B, 4 * H, dtype=prec.tdt, device=DEV)
    y = torch.empty(T * B, H, dtype=prec.tdt, device=DEV)
    ops.lstm_layer_fwd(prec, T, B, H, whh.data_ptr(), bhh.data_ptr(), P.data_ptr(), 4 * H, hseq.data_ptr(), cseq.data_ptr(),
                       gates=gates.data_ptr(), y=y.data_ptr(), ld_y=H)
    torch.cuda.synchronize()
    tol = dict(atol=2e-5, rtol=1e-4) if prec_name == "fp32" else dict(atol=3e-2, rtol=3e-2)
    y_lib = y.float().cpu().view(T, B, H).transpose(0, 1)
    assert torch.allclose(y_lib, y_ref.detach(), **tol), (y_lib - y_ref).abs().max()
    dY = _to_dev(gy.transpose(0, 1).reshape(T * B, H), prec)
    dP = torch.empty(T * B, 4 * H, dtype=prec.tdt, device=DEV)
    ws = torch.empty(3 * B * H, dtype=torch.float32, device=DEV)
    ops.lstm_layer_bwd(prec, T, B, H, whh.data_ptr(), hseq.data_ptr(), cseq.data_ptr(), gates.data_ptr(), dY.data_ptr(), H, 0,
                       dP.data_ptr(), ws.data_ptr())
    gW_hh = torch.zeros(4 * H, H, device=DEV)
    gb = torch.zeros(4 * H, device=DEV)
    dX = torch.empty(T * B, I, dtype=torch.float32, device=DEV)
    ops.gemm(prec.core, prec.act, 4 * H, H, [(dP.data_ptr(), 4 * H, 1, hseq.data_ptr(), H, 1, T * B)], gW_hh.data_ptr(), F32, H,
             accumulate=ops.ATOMIC_ADD)
    ops.colsum(dP.data_ptr(), prec.act, 4 * H, T * B, 4 * H, gb.data_ptr())
    ops.gemm(prec.core, prec.act, T * B, I, [(dP.data_ptr(), 4 * H, 0, wih.data_ptr(), I, 1, 4 * H)], dX.data_ptr(), F32, I)
    torch.cuda.synchronize()
    a = 1e-4 if prec_name == "fp32" else 6e-2
    assert (gW_hh.cpu() - Wr["w_hh"].grad).abs().max() <= a * max(1.0, Wr["w_hh"].grad.abs().max().item())
    assert (gb.cpu() - Wr["b_ih"].grad).abs().max() <= a * max(1.0, Wr["b_ih"].grad.abs().max().item())
    dx_ref = xr.grad.transpose(0, 1).reshape(T * B, I)
    assert (dX.cpu() - dx_ref).abs().max() <= a * max(1.0, dx_ref.abs().max().item())


@pytest.mark.parametrize("H,B,T,rev", [(256, 128, 3, 0), (256, 384, 9, 1), (128, 256, 7, 0), (256, 4096, 24, 0)])
def test_lstm_persistent_cluster_layer_fwd_bwd(H, B, T, rev):
    """Persistent LSTM layer (thread-block cluster of H/64 CTAs per 128-row tile, W_hh resident in shared memory, h_t
    exchanged through distributed shared memory, cell state in registers): blocked two-segment input projection,
    forward outputs / final cell state and the backward pass (dP -> dW_hh, db, dX) against the oracle's autograd."""
    prec = Precision("bf16")
    assert ops.lstm_persist_eligible(prec, B, H)
    I1, I2 = 10, 64
    W = _weights(H, I1 + I2, 31, G=4)
    g = torch.Generator().manual_seed(9)
    x = torch.randn(B, T, I1 + I2, generator=g)
    rnd = lambda t: t.to(torch.bfloat16).float()
    Wr = {k: v.clone().requires_grad_() for k, v in W.items()}
    xr = x.clone().requires_grad_()
    xin = torch.flip(rnd(xr), [1]) if rev else rnd(xr)       # rev: the layer runs on the time-flipped sequence ...
    y_ref = O.lstm_layer(xin, rnd(Wr["w_ih"]), rnd(Wr["w_hh"]), Wr["b_ih"], Wr["b_hh"])
    if rev:
        y_ref = torch.flip(y_ref, [1])                        # ... and its output is flipped back (arnn_model.py:455-475)
    gy = torch.randn(B, T, H, generator=g)
    (y_ref * gy).sum().backward()
    x_seq = torch.flip(x, [1]) if rev else x
    x_tm = x_seq.transpose(0, 1).reshape(T * B, I1 + I2)
    X1 = torch.zeros(T * B, 16, dtype=torch.bfloat16, device=DEV)
    X1[:, :I1] = x_tm[:, :I1].to(DEV)
    X2 = x_tm[:, I1:].contiguous().to(device=DEV, dtype=torch.bfloat16)
    w1 = torch.zeros(4 * H, 16, dtype=torch.bfloat16, device=DEV)
    w1[:, :I1] = W["w_ih"][:, :I1].to(DEV)
    w2 = W["w_ih"][:, I1:].contiguous().to(device=DEV, dtype=torch.bfloat16)
    whh = _to_dev(W["w_hh"], prec)
    bih, bhh = W["b_ih"].to(DEV), W["b_hh"].to(DEV)
    P = torch.empty(T * B, 4 * H, dtype=torch.bfloat16, device=DEV)
    ops.lstm_inproj_blocked(X1.data_ptr(), 16, I1, w1.data_ptr(), 16, T * B, bih.data_ptr(), bhh.data_ptr(), H, P.data_ptr(),
                            X2=X2.data_ptr(), ldx2=I2, K2=I2, w_ih2=w2.data_ptr(), ldw2=I2)
    # blocked layout check of the projection itself: vec16(R, g, u) = ((R/128*4 + g)*(H/8) + u/8)*128 + R%128
    torch.cuda.synchronize()
    pre = (rnd(x_tm) @ rnd(W["w_ih"]).t() + W["b_ih"] + W["b_hh"]).view(T * B // 128, 128, 4, H // 8, 8)
    pre = pre * torch.tensor([0.5, 0.5, 1.0, 0.5]).view(1, 1, 4, 1, 1)
    blk = P.float().cpu().view(T * B // 128, 4, H // 8, 128, 8).permute(0, 3, 1, 2, 4)
    assert torch.allclose(blk, pre, atol=3e-2, rtol=2e-2), (blk - pre).abs().max()
    hseq = torch.empty((T + 1) * B, H, dtype=torch.bfloat16, device=DEV)
    hseq[:B].zero_()
    cseq = torch.empty((T + 1) * B, H, dtype=torch.float32, device=DEV)
    cseq[:B].zero_()
    gates = torch.empty(T * B, ops.lstm_gates_cols(H, True), dtype=torch.bfloat16, device=DEV)
    y = torch.empty(T * B, H, dtype=torch.bfloat16, device=DEV)
    ops.lstm_layer_fwd(prec, T, B, H, whh.data_ptr(), bhh.data_ptr(), P.data_ptr(), 4 * H, hseq.data_ptr(), cseq.data_ptr(),
                       gates=gates.data_ptr(), y=y.data_ptr(), ld_y=H, y_reverse_time=rev, P_blocked=1)
    torch.cuda.synchronize()
    tol = dict(atol=3e-2, rtol=3e-2)
    y_lib = y.float().cpu().view(T, B, H).transpose(0, 1)
    assert torch.allclose(y_lib, y_ref.detach(), **tol), (y_lib - y_ref).abs().max()
    h_lib = hseq[B:].float().cpu().view(T, B, H).transpose(0, 1)          # processing order
    assert torch.allclose(torch.flip(h_lib, [1]) if rev else h_lib, y_ref.detach(), **tol)
    # no-grad variant (no saved state) gives the same outputs
    y2 = torch.empty_like(y)
    hseq2, cseq2 = hseq.clone(), cseq.clone()
    ops.lstm_layer_fwd(prec, T, B, H, whh.data_ptr(), bhh.data_ptr(), P.data_ptr(), 4 * H, hseq2.data_ptr(), cseq2.data_ptr(),
                       y=y2.data_ptr(), ld_y=H, y_reverse_time=rev, P_blocked=1)
    torch.cuda.synchronize()
    assert torch.equal(y2, y) and torch.equal(cseq2[T * B:], cseq[T * B:])
    # ---- backward
    dY = _to_dev(gy.transpose(0, 1).reshape(T * B, H), prec)            # gradient wrt y in the order y was written
    dP = torch.empty(T * B, 4 * H, dtype=torch.bfloat16, device=DEV)
    ops.lstm_layer_bwd(prec, T, B, H, whh.data_ptr(), 0, 0, gates.data_ptr(), dY.data_ptr(), H, 0, dP.data_ptr(), 0,
                       y_reverse_time=rev, gates_persist=1)
    gW_hh = torch.zeros(4 * H, H, device=DEV)
    gb = torch.zeros(4 * H, device=DEV)
    gW2 = torch.zeros(4 * H, I2, device=DEV)
    ops.gemm(prec.core, prec.act, 4 * H, H, [(dP.data_ptr(), 4 * H, 1, hseq.data_ptr(), H, 1, T * B)], gW_hh.data_ptr(), F32, H,
             accumulate=ops.ATOMIC_ADD)
    ops.gemm(prec.core, prec.act, 4 * H, I2, [(dP.data_ptr(), 4 * H, 1, X2.data_ptr(), I2, 1, T * B)], gW2.data_ptr(), F32, I2,
             accumulate=ops.ATOMIC_ADD)
    ops.colsum(dP.data_ptr(), prec.act, 4 * H, T * B, 4 * H, gb.data_ptr())
    torch.cuda.synchronize()
    a = 6e-2
    for mine, ref, what in ((gW_hh, Wr["w_hh"].grad, "dW_hh"), (gb, Wr["b_ih"].grad, "db"), (gW2, Wr["w_ih"].grad[:, I1:], "dW_ih")):
        err = (mine.cpu() - ref).norm() / ref.norm().clamp_min(1e-9)
        assert err < a, (what, float(err))


@pytest.mark.parametrize("H,B,T,V", [(512, 512, 7, 64), (128, 256, 5, 11), (64, 128, 3, 128), (320, 384, 4, 90)])
def test_persistent_layer_token_table_gather_is_bit_exact(H, B, T, V):
    """Encoder layer 0: the input projection is a row of a token table.  With table_rows given the blocked P of the
    persistent kernel is gathered from a folded bf16 table (gru_gather_p_kernel); without it the generic relayout
    kernel builds it.  Same arithmetic and rounding: the layer outputs must be identical bit for bit."""
    from inpaintnet_b200 import ops
    from inpaintnet_b200.ops import Precision
    prec = Precision("bf16")
    ndir = 2
    g = torch.Generator().manual_seed(H + B)
    s = 1.0 / H ** 0.5
    whh = [((torch.rand(3 * H, H, generator=g) * 2 - 1) * s).to(DEV).bfloat16().contiguous() for _ in range(ndir)]
    bhh = [((torch.rand(3 * H, generator=g) * 2 - 1) * s).to(DEV) for _ in range(ndir)]
    table = torch.randn(ndir, V, 3 * H, generator=g).to(DEV)
    tok = torch.randint(0, V, (T * B,), generator=g).int().to(DEV)
    tok[:V] = torch.arange(V, dtype=torch.int32)          # every table row is used
    h0 = (torch.randn(ndir, B, H, generator=g) * 0.5).to(DEV).bfloat16()

    def run(rows):
        hseq = torch.zeros(ndir, (T + 1) * B, H, dtype=torch.bfloat16, device=DEV)
        hseq[0, :B] = h0[0]
        hseq[1, T * B:] = h0[1]
        y = torch.zeros(T * B, ndir * H, dtype=torch.bfloat16, device=DEV)
        dirs = [ops.gru_dir(whh[d].data_ptr(), bhh[d].data_ptr(), hseq[d].data_ptr(), reverse=d, y_col0=d * H,
                            table=table[d].data_ptr(), ld_table=3 * H, tok=tok.data_ptr(), table_rows=rows) for d in range(ndir)]
        assert ops.gru_layer_fwd(prec, T, B, H, dirs, y=y.data_ptr(), ld_y=ndir * H)   # True: the persistent kernel ran
        torch.cuda.synchronize()
        return y, hseq

    y_ref, h_ref = run(0)
    for _ in range(3):
        y, h = run(V)
        assert torch.equal(y, y_ref) and torch.equal(h, h_ref)
    assert y_ref.float().abs().max().item() > 0.1
