"""clip_grad_norm_ + Adam as two kernels over the model's flat parameter / gradient buffers
(reference trainer_gmm.py:52,250-251: optim.Adam(lr), clip_grad_norm_(params, 1), step())."""
from __future__ import annotations

import torch

from ._lib import LIB, stream_ptr
from .ops import _p, _reduce_scratch


class FusedAdam(torch.optim.Optimizer):
    """Adam (betas (0.9,0.999), eps 1e-8, no weight decay) fused with global-norm clipping.

    `zero_grad()` is one memset of the flat gradient buffer; `step()` runs the norm reduction and
    the clip+update kernel.  In data-parallel runs `grad_sync` (a callable taking the flat gradient
    tensor) is invoked between backward and the norm: that is where the NCCL all-reduce goes.
    """

    def __init__(self, model, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, max_norm=1.0, grad_sync=None):
        self.model = model
        params = [p for _, p in model.live_parameters()]
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps))
        self.max_norm = max_norm
        self.grad_sync = grad_sync
        self.t = 0
        self._m = self._v = None
        self.last_norm = None          # device scalar: pre-clip global gradient norm of the last step

    def zero_grad(self, set_to_none: bool = False):
        self.model.zero_grad_flat()

    def _gather_grads(self, flat_grad):
        """Slow path: someone replaced p.grad (e.g. zero_grad(set_to_none=True)); copy back."""
        off = 0
        for _, p in self.model.live_parameters():
            n = p.numel()
            if p.grad is None:
                flat_grad[off:off + n].zero_()
            elif p.grad.data_ptr() != flat_grad.data_ptr() + off * 4:
                flat_grad[off:off + n].copy_(p.grad.reshape(-1))
            off += ((n + 3) // 4) * 4

    @torch.no_grad()
    def step(self, closure=None):
        flat, grad = self.model.flatten_parameters_()
        self._gather_grads(grad)
        if self.grad_sync is not None:
            self.grad_sync(grad)
        if self._m is None or self._m.numel() != flat.numel() or self._m.device != flat.device:
            self._m, self._v = torch.zeros_like(flat), torch.zeros_like(flat)
        self.t += 1
        g = self.param_groups[0]
        dev = flat.device
        st = stream_ptr(dev)
        norm = torch.empty(1, dtype=torch.float32, device=dev)
        scratch, nb = _reduce_scratch(dev)
        n = flat.numel()
        LIB.call("fn_grad_norm", _p(grad), n, _p(norm), _p(scratch), nb, st)
        LIB.call("fn_clip_adam", _p(flat), _p(grad), _p(self._m), _p(self._v), n, _p(norm),
                 float(self.max_norm if self.max_norm is not None else 3.0e38), float(g["lr"]),
                 float(g["betas"][0]), float(g["betas"][1]), float(g["eps"]), self.t, st)
        self.last_norm = norm
        return None

    def state_dict(self):
        return dict(t=self.t, m=self._m, v=self._v, lr=self.param_groups[0]["lr"])

    def load_state_dict(self, sd):
        self.t, self._m, self._v = sd["t"], sd["m"], sd["v"]
        self.param_groups[0]["lr"] = sd.get("lr", self.param_groups[0]["lr"])
