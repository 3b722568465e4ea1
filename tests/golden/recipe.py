"""Deterministic weight / input recipes shared by make_golden.py (run in the build container,
needs /root/reference) and by the parity tests (run anywhere, no reference needed).

Large models are NOT stored in the fixtures: both sides regenerate the same state_dict from
(shape spec, seed) with the CPU generator, which is deterministic for a given torch version.
"""
import math
import torch


def mvae_spec(V, E=10, H=512, Z=256, L=2):
    """Shapes of the reference MeasureVAE state_dict, L == 2."""
    s = {}
    for l in range(L):
        for sfx in ("", "_reverse"):
            I = E if l == 0 else 2 * H
            s[f"encoder.lstm.weight_ih_l{l}{sfx}"] = (3 * H, I)
            s[f"encoder.lstm.weight_hh_l{l}{sfx}"] = (3 * H, H)
            s[f"encoder.lstm.bias_ih_l{l}{sfx}"] = (3 * H,)
            s[f"encoder.lstm.bias_hh_l{l}{sfx}"] = (3 * H,)
    s["encoder.note_embedding_layer.weight"] = (V, E)
    for head in ("linear_mean", "linear_log_std"):
        s[f"encoder.{head}.0.weight"] = (2 * H, 2 * H * L)
        s[f"encoder.{head}.0.bias"] = (2 * H,)
        s[f"encoder.{head}.2.weight"] = (Z, 2 * H)
        s[f"encoder.{head}.2.bias"] = (Z,)
    s["decoder.b_0"] = (1,)
    s["decoder.x_0"] = (E,)
    s["decoder.note_embedding_layer.weight"] = (V, E)
    s["decoder.z_to_beat_rnn_input.0.weight"] = (H * L, Z)
    s["decoder.z_to_beat_rnn_input.0.bias"] = (H * L,)
    for l in range(L):
        s[f"decoder.rnn_beat.weight_ih_l{l}"] = (3 * H, 1 if l == 0 else H)
        s[f"decoder.rnn_beat.weight_hh_l{l}"] = (3 * H, H)
        s[f"decoder.rnn_beat.bias_ih_l{l}"] = (3 * H,)
        s[f"decoder.rnn_beat.bias_hh_l{l}"] = (3 * H,)
    s["decoder.beat_emb_to_tick_rnn_hidden.0.weight"] = (H * L, H)
    s["decoder.beat_emb_to_tick_rnn_hidden.0.bias"] = (H * L,)
    s["decoder.beat_emb_to_tick_rnn_input.0.weight"] = (H, H)
    s["decoder.beat_emb_to_tick_rnn_input.0.bias"] = (H,)
    for l in range(L):
        s[f"decoder.rnn_tick.weight_ih_l{l}"] = (3 * H, E + H if l == 0 else H)
        s[f"decoder.rnn_tick.weight_hh_l{l}"] = (3 * H, H)
        s[f"decoder.rnn_tick.bias_ih_l{l}"] = (3 * H,)
        s[f"decoder.rnn_tick.bias_hh_l{l}"] = (3 * H,)
    s["decoder.tick_emb_to_note_emb.0.weight"] = (V, H)
    s["decoder.tick_emb_to_note_emb.0.bias"] = (V,)
    return s


def latent_rnn_spec(Z=256, Hc=512, L=2, auto_reg=False, ablation=False):
    s = {}
    if not auto_reg:
        s["x_0"] = (1, 1, 1)
    for name, I, H in (("context_rnn_past", Z, Hc), ("context_rnn_future", Z, Hc),
                       ("generation_rnn", Z if auto_reg else 1, Hc if ablation else Hc * L)):
        for l in range(L):
            for sfx in ("", "_reverse"):
                s[f"{name}.weight_ih_l{l}{sfx}"] = (3 * H, I if l == 0 else 2 * H)
                s[f"{name}.weight_hh_l{l}{sfx}"] = (3 * H, H)
                s[f"{name}.bias_ih_l{l}{sfx}"] = (3 * H,)
                s[f"{name}.bias_hh_l{l}{sfx}"] = (3 * H,)
    s["generation_linear.weight"] = (Z, 2 * (Hc if ablation else Hc * L))
    s["generation_linear.bias"] = (Z,)
    return s


def make_state_dict(spec, seed, gain=1.0):
    """Values: 2-D+ tensors ~ N(0, gain * xavier std), 1-D tensors ~ U(-0.1, 0.1).
    Keys are visited in sorted order so the stream is independent of dict order."""
    g = torch.Generator().manual_seed(seed)
    sd = {}
    for k in sorted(spec):
        shp = spec[k]
        if len(shp) >= 2 and shp[0] * shp[-1] > 1:
            fan_out, fan_in = shp[0], shp[-1]
            std = gain * math.sqrt(2.0 / (fan_in + fan_out))
            sd[k] = torch.randn(shp, generator=g) * std
        else:
            sd[k] = (torch.rand(shp, generator=g) - 0.5) * 0.2
    return sd


def make_tokens(B, T, V, seed):
    g = torch.Generator().manual_seed(seed)
    return torch.randint(0, V, (B, T), generator=g)


def make_normal(shape, seed):
    g = torch.Generator().manual_seed(seed)
    return torch.randn(shape, generator=g)
