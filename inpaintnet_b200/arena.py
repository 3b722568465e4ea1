"""Flat parameter arena: every parameter of a model is a view into ONE fp32 buffer (trainable
parameters first), gradients likewise, so that the optimiser step is a single fused kernel and the
data-parallel gradient exchange is a single bucketed all-reduce over a contiguous range.  A bf16
"shadow" arena (the weight pack for the tcgen05 path) is refreshed lazily by one table-driven kernel
whenever a parameter changed.  state_dict()/load_state_dict() keep the reference layout because
the nn.Parameters themselves are untouched apart from where their storage lives."""
import ctypes as C

import numpy as np
import torch

from . import _lib as L
from . import ops


def _round_up(x, m):
    return (x + m - 1) // m * m


class ParamArena:
    # (param suffix that must be immediately followed in memory by another param suffix)
    ADJACENT = (("decoder.note_embedding_layer.weight", "decoder.x_0"),)

    def __init__(self, root: torch.nn.Module):
        named = []
        seen = set()
        for n, p in root.named_parameters():
            if id(p) in seen:
                continue
            seen.add(id(p))
            named.append((n, p))
        if not named:
            raise ValueError("module has no parameters")
        dev = named[0][1].device
        # order: trainable first, frozen after; honour adjacency pairs
        follow = {}
        for a, b in self.ADJACENT:
            for n, _ in named:
                if n.endswith(a):
                    pre = n[: len(n) - len(a)]
                    if any(m == pre + b for m, _ in named):
                        follow[n] = pre + b
        followers = set(follow.values())
        by_name = dict(named)
        ordered = []
        for want_grad in (True, False):
            for n, p in named:
                if p.requires_grad != want_grad or n in followers:
                    continue
                ordered.append((n, p))
                if n in follow:
                    ordered.append((follow[n], by_name[follow[n]]))
        self.root = root
        self.device = dev
        self.names = [n for n, _ in ordered]
        self.params = [p for _, p in ordered]
        self.offset = {}
        off = 0
        prev = None
        for n, p in ordered:
            if not (prev is not None and follow.get(prev) == n):
                off = _round_up(off, 4)  # 16-byte aligned unless glued to its predecessor
            self.offset[n] = off
            off += p.numel()
            prev = n
        self.total = _round_up(off, 4)
        trainable_end = 0
        for n, p in ordered:
            if p.requires_grad:
                trainable_end = max(trainable_end, self.offset[n] + p.numel())
        self.n_trainable = _round_up(trainable_end, 4)
        self.flat = torch.zeros(self.total, dtype=torch.float32, device=dev)
        self.grad = torch.zeros(self.total, dtype=torch.float32, device=dev)
        for n, p in ordered:
            o = self.offset[n]
            view = self.flat[o:o + p.numel()].view(p.shape)
            view.copy_(p.data.to(torch.float32))
            p.data = view
            p.grad = None
            p._ipn_arena = self
            p._ipn_name = n
        self.by_name = dict(ordered)
        self._rg_sig = self._requires_grad_signature()
        self.manual_version = 0
        self._shadow_entries = {}   # key -> (bf16 tensor, rows, cols, ld, full_cols, col0)
        self._shadow_items = None
        self._shadow_key = None
        self._derived = {}
        self._derived_key = None
        self.nan_flag = torch.zeros(1, dtype=torch.int32, device=dev)
        self.range_flag = torch.zeros(1, dtype=torch.int32, device=dev)
        self.rng_offset = 0

    # ------------------------------------------------------------------ validity / lookup
    def _requires_grad_signature(self):
        return tuple(p.requires_grad for p in self.params)

    def valid(self):
        """False once a parameter left the arena (model.to()/.float()) or the trainable set changed (freezing a
        sub-module after the arena exists would leave the trainable-first layout, n_trainable and the early-exchange
        ranges stale): arena_of() then rebuilds the arena, and FusedAdam carries its state over by name."""
        p0, p1 = self.params[0], self.params[-1]
        return (p0.data_ptr() == self.flat.data_ptr() + 4 * self.offset[self.names[0]]
                and p1.data_ptr() == self.flat.data_ptr() + 4 * self.offset[self.names[-1]]
                and p0.device == self.device and self._rg_sig == self._requires_grad_signature())

    def version_key(self):
        return (self.manual_version, sum(p._version for p in self.params))

    def invalidate(self):
        """Call after writing parameters through a path that does not bump tensor versions (`p.data.clamp_()`,
        an EMA swap through .data, ...): the bf16 shadow and the derived tables are rebuilt on the next forward."""
        self.manual_version += 1

    def fptr(self, name, elem_off=0):
        """device pointer into the fp32 master copy"""
        return self.flat.data_ptr() + 4 * (self.offset[name] + elem_off)

    def gptr(self, name, elem_off=0):
        """device pointer into the fp32 gradient arena; binds param.grad to its arena view (zeroing the
        region first when the gradient had been reset to None by a foreign optimiser)."""
        p = self.by_name[name]
        o = self.offset[name]
        if p.grad is None or p.grad.data_ptr() != self.grad.data_ptr() + 4 * o:
            view = self.grad[o:o + p.numel()].view(p.shape)
            view.zero_()
            p.grad = view
        return self.grad.data_ptr() + 4 * (o + elem_off)

    def wants_grad(self, name):
        return self.by_name[name].requires_grad

    def trainable_range(self, prefixes):
        """[lo, hi) element range of the gradient arena holding exactly the trainable parameters whose names start
        with one of `prefixes`, or None when there are none or they are not contiguous (early gradient exchange)."""
        prefixes = (prefixes,) if isinstance(prefixes, str) else tuple(prefixes)
        sel = [n for n, p in zip(self.names, self.params) if p.requires_grad and n.startswith(prefixes)]
        if not sel:
            return None
        lo = min(self.offset[n] for n in sel)
        hi = max(self.offset[n] + self.by_name[n].numel() for n in sel)
        for n, p in zip(self.names, self.params):
            if p.requires_grad and n not in sel and lo <= self.offset[n] < hi:
                return None
        return lo, min(_round_up(hi, 4), self.n_trainable)

    def zero_grad(self):
        self.grad[: self.n_trainable].zero_()
        for n, p in zip(self.names, self.params):
            if p.requires_grad:
                o = self.offset[n]
                if p.grad is None or p.grad.data_ptr() != self.grad.data_ptr() + 4 * o:
                    p.grad = self.grad[o:o + p.numel()].view(p.shape)

    # ------------------------------------------------------------------ bf16 shadow (weight pack)
    def _shadow_register(self, key):
        name, col0, cols = key
        p = self.by_name[name]
        rows = p.shape[0]
        full_cols = p.numel() // rows
        cols = full_cols - col0 if cols is None else cols
        ld = _round_up(cols, 8)
        buf = torch.zeros(rows * ld + 8, dtype=torch.bfloat16, device=self.device)  # own storage: stable pointer
        self._shadow_entries[key] = (buf, rows, cols, ld, full_cols, col0)
        self._shadow_items = None
        self._shadow_key = None

    def _shadow_table(self):
        n = len(self._shadow_entries)
        arr = (L.PackItem * n)()
        max_rows = max_ld = 1
        for i, (key, e) in enumerate(self._shadow_entries.items()):
            buf, rows, cols, ld, full_cols, col0 = e
            arr[i].src = self.fptr(key[0], col0)
            arr[i].ld_src = full_cols
            arr[i].dst = buf.data_ptr()
            arr[i].ld_dst = ld
            arr[i].rows = rows
            arr[i].cols = cols
            max_rows, max_ld = max(max_rows, rows), max(max_ld, ld)
        raw = np.frombuffer(memoryview(arr), dtype=np.uint8).copy()
        self._shadow_items = (torch.from_numpy(raw).to(self.device), n, max_rows, max_ld)

    def refresh(self):
        """Re-packs the bf16 shadow when a parameter changed since the last pack (one kernel launch).
        Call once at the start of every forward. Returns True if the pack ran."""
        if not self._shadow_entries:
            return False
        key = self.version_key()
        if key == self._shadow_key and self._shadow_items is not None:
            self._await(getattr(self, "_shadow_ready", None))
            return False
        if self._shadow_items is None:
            self._shadow_table()
        items, n, max_rows, max_ld = self._shadow_items
        ops.pack_bf16(items.data_ptr(), n, max_rows, max_ld)
        self._shadow_key = key
        self._shadow_ready = self._mark()
        return True

    # caches filled on one stream and read on another (micro-batches pipelined on two streams): the reader waits
    # for the event recorded behind the kernels that filled the cache
    def _mark(self):
        if self.flat.is_cuda:
            s = torch.cuda.current_stream()
            return s, s.record_event()
        return None

    @staticmethod
    def _await(mark):
        if mark is not None:
            cur = torch.cuda.current_stream()
            if cur != mark[0]:
                if torch.cuda.is_current_stream_capturing():
                    # a capturing stream cannot wait on an event recorded outside the capture; the caches were filled
                    # by eager work that the caller synchronised with before starting the capture (GraphedInpainter)
                    return
                cur.wait_event(mark[1])

    def w(self, prec, name, col0=0, cols=None):
        """(device pointer, leading dimension) of a 2-D weight block in the operand dtype of `prec`.
        bf16 blocks live in the shadow; a block requested for the first time is packed immediately."""
        if prec.act == L.F32:
            p = self.by_name[name]
            return self.fptr(name, col0), p.numel() // p.shape[0]
        key = (name, col0, cols)
        e = self._shadow_entries.get(key)
        if e is None:
            self._shadow_register(key)
            self.refresh()
            e = self._shadow_entries[key]
        return e[0].data_ptr(), e[3]

    # ------------------------------------------------------------------ derived tensors cache
    def derived(self, key, builder):
        """Small tensors derived from parameters (embedding x W_ih tables, the beat-GRU constant input
        projection); rebuilt when any parameter changed."""
        vk = self.version_key()
        if vk != self._derived_key:
            self._derived = {}
            self._derived_key = vk
        if key not in self._derived:
            self._derived[key] = (builder(), self._mark())
        else:
            self._await(self._derived[key][1])
        return self._derived[key][0]

    # ------------------------------------------------------------------ randomness
    def next_rng_offset(self, n_counters):
        o = self.rng_offset
        self.rng_offset += int(n_counters) + 1
        return o


def arena_of(module: torch.nn.Module) -> ParamArena:
    """Returns the arena that owns ALL parameters of `module`, creating one rooted at `module` if needed."""
    params = list(module.parameters())
    if not params:
        raise ValueError("module has no parameters")
    a = getattr(params[0], "_ipn_arena", None)
    if a is not None and a.valid() and getattr(params[-1], "_ipn_arena", None) is a:
        return a
    return ParamArena(module)


def prefix_of(module: torch.nn.Module, arena: ParamArena) -> str:
    """Name prefix of `module`'s parameters inside the arena (e.g. 'vae_model.encoder.')."""
    n, p = next(iter(module.named_parameters()))
    full = p._ipn_name
    assert full.endswith(n)
    return full[: len(full) - len(n)]
