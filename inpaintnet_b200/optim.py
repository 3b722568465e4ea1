"""Fused multi-tensor Adam over the flat parameter arena (one kernel for all parameters) with the
NaN guard of MeasureVAE/encoder.py:111-116 folded in, plus the data-parallel gradient exchange.
reference: utils/trainer.py:32-35 (torch.optim.Adam, lr=1e-4, defaults) and :165-177."""
import torch

from . import ops
from .arena import arena_of


class FusedAdam(torch.optim.Optimizer):
    def __init__(self, model, lr=1e-4, betas=(0.9, 0.999), eps=1e-8):
        self.model = model
        params = [p for p in model.parameters() if p.requires_grad]
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps))
        self._arena = None
        self._m = self._v = None
        self.step_count = 0

    def arena(self):
        a = arena_of(self.model)
        if a is not self._arena:
            old, old_m, old_v = self._arena, self._m, self._v
            self._arena = a
            self._m = torch.zeros(a.n_trainable, dtype=torch.float32, device=a.device)
            self._v = torch.zeros(a.n_trainable, dtype=torch.float32, device=a.device)
            if old is not None and self.step_count > 0:
                # the arena was rebuilt (model.to()/.float(), a sub-module frozen or unfrozen): carry the moments of
                # every parameter that is still trainable over by NAME; anything new starts from zero moments
                carried = 0
                was_trainable = dict(zip(old.names, old._rg_sig))   # the state the old layout was built for
                for n, p in zip(a.names, a.params):
                    if p.requires_grad and was_trainable.get(n, False) and old.by_name[n].numel() == p.numel():
                        o_new, o_old, k = a.offset[n], old.offset[n], p.numel()
                        self._m[o_new:o_new + k].copy_(old_m[o_old:o_old + k])
                        self._v[o_new:o_new + k].copy_(old_v[o_old:o_old + k])
                        carried += 1
                n_train = sum(1 for p in a.params if p.requires_grad)
                if carried != n_train:
                    import warnings
                    warnings.warn(f"FusedAdam: parameter arena rebuilt; Adam moments carried over for {carried} of {n_train} "
                                  f"trainable parameters, the others restart from zero moments (step count kept)")
        return a

    def zero_grad(self, set_to_none=False):
        self.arena().zero_grad()

    @torch.no_grad()
    def step(self, closure=None, grad_scale=1.0):
        a = self.arena()
        g = self.param_groups[0]
        self.step_count += 1
        # parameters whose grad was never produced keep a zero gradient (views are bound by zero_grad)
        ops.adam_step(a.flat.data_ptr(), a.grad.data_ptr(), self._m.data_ptr(), self._v.data_ptr(), a.n_trainable,
                      self.step_count, g["lr"], g["betas"][0], g["betas"][1], g["eps"], grad_scale, a.nan_flag.data_ptr())
        a.manual_version += 1

    def state_dict(self):
        """Adds what the reference never saved (no true resume): moments + step."""
        self.arena()
        return dict(step=self.step_count, m=self._m.clone(), v=self._v.clone(), param_groups=[
            {k: v for k, v in self.param_groups[0].items() if k != "params"}])

    def load_state_dict(self, sd):
        self.arena()
        self.step_count = int(sd["step"])
        self._m.copy_(sd["m"])
        self._v.copy_(sd["v"])
        for k, v in sd["param_groups"][0].items():
            self.param_groups[0][k] = v
