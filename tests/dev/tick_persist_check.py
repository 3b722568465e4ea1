"""dev: the persistent argmax tick-decode kernel (csrc/tick_persist.cu) against the per-tick launch path
(IPN_TICK_PERSIST=0) and the CPU oracle, small and full size, with and without dropout masks; timings."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from inpaintnet_b200 import engine
from inpaintnet_b200.data import SyntheticFolkDataset
from inpaintnet_b200.measure_vae import MeasureVAE
from oracle import inpaintnet_oracle as O
from tests.golden import recipe
from tests.test_gpu_fullsize import _lib_masks, subset_rows, rel_err, grad_errs

DEV = "cuda"
V, H, Z = 64, 512, 256
sd = recipe.make_state_dict(recipe.mvae_spec(V, 10, H, Z), 4321)


def run(B, train, persist, seed=29):
    os.environ["IPN_TICK_PERSIST"] = "1" if persist else "0"
    m = MeasureVAE(SyntheticFolkDataset(num_notes=V))
    m.load_state_dict(sd)
    m.to(DEV).set_precision("bf16")
    m.train(train)
    m.decoder.teacher_forcing_prob = -1.0
    g = torch.Generator().manual_seed(seed)
    tokens = torch.randint(0, V, (B, 24), generator=g)
    eps = torch.randn(B, Z, generator=g)
    enc = torch.rand(B, 24, 2 * H, generator=g) > 0.5
    beat = torch.rand(B, 4, H, generator=g) > 0.5
    tick = torch.rand(B, 24, H, generator=g) > 0.5
    masks = [x.to(torch.uint8) for x in _lib_masks(B, H, enc, beat, tick)]
    tok_d = tokens.to(DEV)
    m.zero_grad()
    rows = subset_rows(B) if B >= 1024 else torch.arange(0, B, 5)
    with engine.inject_noise(masks=masks if train else [], eps=[eps]):
        if train:
            w, s, zd, _, z, _ = m(tok_d, train=True)
        else:
            with torch.no_grad():
                w, s, zd, _, z, _ = m(tok_d, train=False)
    grads = None
    if train:
        rd = rows.to(DEV)
        loss = torch.nn.functional.cross_entropy(w[rd].reshape(-1, V), tok_d[rd].reshape(-1))
        loss.backward()
        grads = {k: p.grad.detach().float().cpu() for k, p in m.named_parameters() if p.grad is not None}
    torch.cuda.synchronize()
    return dict(w=w.detach().cpu(), s=s.cpu(), m=m, tokens=tokens, eps=eps, enc=enc, beat=beat, tick=tick, rows=rows, grads=grads)


def oracle_check(r, train):
    rows, tokens = r["rows"], r["tokens"]
    sdr = {k: v.clone().requires_grad_() for k, v in sd.items()}
    if train:
        drop = dict(enc=[r["enc"][rows].float()], beat=[r["beat"][rows].float()], tick=r["tick"][rows].float())
        mu_r, ls_r = O.encoder_forward(sdr, tokens[rows], 2, drop["enc"], 0.5)
    else:
        drop = dict(enc=None, beat=None, tick=None)
        mu_r, ls_r = O.encoder_forward(sdr, tokens[rows], 2, None, 0.0)
    z_r = mu_r + torch.exp(ls_r) * r["eps"][rows] if train else mu_r + torch.exp(ls_r) * r["eps"][rows]
    fed = r["s"][rows, 0]
    w_r, _ = O.decoder_forward(sdr, z_r, fed, True, 2, drop["beat"], drop["tick"], 0.5 if train else 0.0)
    err = rel_err(r["w"][rows], w_r.detach())
    top2 = w_r.detach().topk(2, dim=2).values
    strict = (top2[..., 0] - top2[..., 1]) > 5e-2
    ok = bool(((w_r.detach().argmax(2) == fed) | ~strict).all())
    out = dict(logit_rel_err=err, fed_is_argmax=ok)
    if train:
        loss_r = torch.nn.functional.cross_entropy(w_r.reshape(-1, V), tokens[rows].reshape(-1))
        loss_r.backward()
        errs = {}
        for k, v in sdr.items():
            if v.grad is None or k not in r["grads"] or not k.startswith("decoder"):
                continue
            errs[k] = ((r["grads"][k] - v.grad).norm() / v.grad.norm().clamp_min(1e-12)).item()
        out["max_grad_err"] = max(errs.values())
        out["worst"] = max(errs, key=errs.get)
    return out


def timed(B, train, persist, n=5):
    os.environ["IPN_TICK_PERSIST"] = "1" if persist else "0"
    m = MeasureVAE(SyntheticFolkDataset(num_notes=V))
    m.load_state_dict(sd)
    m.to(DEV).set_precision("bf16")
    m.train(train)
    m.decoder.teacher_forcing_prob = -1.0
    tok = torch.randint(0, V, (B, 24)).to(DEV)

    def f():
        if train:
            w = m(tok, train=True)[0]
        else:
            with torch.no_grad():
                w = m(tok, train=False)[0]
        return w
    for _ in range(3):
        f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        f()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


if __name__ == "__main__":
    sizes = [int(x) for x in (sys.argv[1] if len(sys.argv) > 1 else "256,4096").split(",")]
    for B in sizes:
        for train in (False, True):
            a = run(B, train, True)
            b = run(B, train, False)
            agree = (a["s"] == b["s"]).float().mean().item()
            same_rows = (a["s"] == b["s"]).all(2).all(1)
            dw = (a["w"][same_rows] - b["w"][same_rows]).abs().max().item() if same_rows.any() else float("nan")
            print(f"B={B} train={train}: tokens equal to the per-tick path {agree:.4f}, rows fully equal {same_rows.float().mean().item():.4f}, "
                  f"max |dw| on those {dw:.3e}", flush=True)
            print("   persistent vs oracle:", oracle_check(a, train), flush=True)
            print("   per-tick   vs oracle:", oracle_check(b, train), flush=True)
    for B in sizes:
        for train in (False, True):
            t1 = timed(B, train, True)
            t0 = timed(B, train, False)
            print(f"B={B} train={train}: forward {t0:.3f} ms per-tick launches -> {t1:.3f} ms persistent", flush=True)
