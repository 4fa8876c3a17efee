"""Adam over the network's flat parameter buffer: one kernel launch per step instead of torch.optim.Adam's foreach pass
over ~300 tensors (3 ms of host time per step at 256 windows -- more than the whole forward + backward on the GPU).

Drop-in for the optimizer line of the reference loop (denoise_train.py:24):

    optimizer = ecg_denoise_b200.optim.Adam(model, lr=1e-3)     # was: torch.optim.Adam(model.parameters(), lr=1e-3)

Same update rule and defaults as torch.optim.Adam (no weight decay, no amsgrad: the reference uses neither).  The
parameters a step updates are the ones that require grad when it runs; frozen parameters never move (their gradient
slice stays zero and so do their moments).  For whole training steps in one CUDA graph use engine.FusedTrainer."""
from __future__ import annotations

import torch

from . import _lib
from .engine import NetPlan


class Adam(torch.optim.Optimizer):
    def __init__(self, model: torch.nn.Module, lr: float = 1e-3, betas=(0.9, 0.999), eps: float = 1e-8):
        plan = getattr(model, "_plan", None)
        if not isinstance(plan, NetPlan):
            raise _lib.RalenetError("ecg_denoise_b200.optim.Adam takes the ralenet module itself (its parameters live "
                                    "in one flat buffer); use torch.optim.Adam for other modules")
        super().__init__([p for p in model.parameters()], dict(lr=lr, betas=betas, eps=eps))
        self.plan = plan
        self._m = self._v = self._step = None

    def _buffers(self):
        plan = self.plan
        dev = next(plan.net.parameters()).device
        plan.ensure(dev)
        if self._m is None or self._m.numel() != plan.flat.numel() or self._m.device != plan.flat.device:
            self._m, self._v = torch.zeros_like(plan.flat), torch.zeros_like(plan.flat)
            self._step = torch.zeros(1, device=plan.flat.device, dtype=torch.int32)
        return plan

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        plan = self._buffers()
        g = self.param_groups[0]
        from . import ops
        # gradients the backward pass accumulated sit in plan.flat_grad (the Parameters' .grad are views of it);
        # a parameter whose .grad is None (never reached by a backward) has a zero slice and zero moments: no update
        ops.adam_flat(plan.flat, plan.flat_grad, self._m, self._v, self._step, g["lr"], tuple(g["betas"]), g["eps"])
        return loss

    def zero_grad(self, set_to_none: bool = False):
        """one memset of the flat gradient buffer; the .grad views stay attached (set_to_none would make the next
        backward re-attach ~300 views)."""
        plan = self._buffers()
        plan.flat_grad.zero_()
        if set_to_none:
            for p in plan.params:
                p.grad = None

    def state_dict(self):
        sd = super().state_dict()
        sd["flat"] = {"m": self._m, "v": self._v, "step": self._step}
        return sd

    def load_state_dict(self, sd):
        flat = sd.get("flat")
        super().load_state_dict({k: v for k, v in sd.items() if k != "flat"})
        if flat is not None and flat["m"] is not None:
            self._buffers()
            self._m.copy_(flat["m"]); self._v.copy_(flat["v"]); self._step.copy_(flat["step"])
