"""Fused Adam(amsgrad) + operand re-pack.

The reference trains with `optim.Adam(model.parameters(), lr, amsgrad=True)` (run.py:662) and, on this path, every
optimiser step is followed by a re-pack of the fp32 convolution weights into the 16-bit K-major operands of the next
forward. FusedAdam is a drop-in subclass of torch.optim.Adam -- same constructor, same per-parameter state
(`step`, `exp_avg`, `exp_avg_sq`, `max_exp_avg_sq`), so `state_dict()` / `load_state_dict()` and run.py's checkpoints
(run.py:436-445,559-569) are interchangeable with the stock optimiser -- whose `step()` sends every fp32 CUDA parameter
through vp3d_adam_step_multi (all tensors of a parameter group in ONE launch): one pass that updates p, m, v, vmax AND, for a convolution weight, writes the packed operand the
training forward has registered for it. Anything else (non-contiguous, other dtypes) falls through to torch's own
implementation. State steps live on the device (capturable), so the whole step can sit in a CUDA graph.
"""
import ctypes as C
import os

import torch
from torch.optim import adam as _adam

from . import native, ops


class FusedAdam(torch.optim.Adam):
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0, amsgrad=False, *, maximize=False,
                 min_numel=1):
        super().__init__(params, lr=lr, betas=betas, eps=eps, weight_decay=weight_decay, amsgrad=amsgrad,
                         maximize=maximize, capturable=True)
        self.min_numel = min_numel
        self._early_done = set()      # ids of the parameters already updated inside this step's backward

    def update_in_backward(self, on=True):
        """Apply the update of each parameter INSIDE the vp3d_b200 training backward, the moment its gradient is final
        and the backward no longer reads it (torch's "optimizer in backward" idea): the HBM-bound Adam pass then runs on
        its own stream beside the tensor-bound weight-gradient GEMMs instead of after the backward. `step()` still has to
        be called as run.py:487 does; it updates whatever the backward did not (other modules' parameters, gradients
        exchanged at the end of a data-parallel backward) and closes the step. The parameter values after step() are
        the same as without this option. Do not combine with gradient accumulation over several backward passes."""
        from . import training
        training.param_update_hook = self._update_now if on else None
        return self

    @torch.no_grad()
    def _update_now(self, pairs):
        """training.param_update_hook: [(parameter, gradient)] -> updated on the current stream."""
        by_id = {id(q): g for q, g in pairs}
        for group in self.param_groups:
            subset = [q for q in group['params'] if id(q) in by_id and id(q) not in self._early_done]
            if not subset:
                continue
            saved = [q.grad for q in subset]
            for q in subset:
                q.grad = by_id[id(q)]
            try:
                self._step_group(group, subset)
            finally:
                for q, g in zip(subset, saved):
                    q.grad = g            # autograd attaches the gradient itself when the backward returns
            self._early_done.update(id(q) for q in subset)

    def _is_big(self, p, g):
        """Tensors the native kernel updates: every contiguous fp32 CUDA parameter (the name is historical -- torch's
        capturable multi-tensor path costs ~0.3 ms per step on the 27 small BatchNorm / bias tensors alone)."""
        return (p.is_cuda and p.dtype == torch.float32 and p.numel() >= self.min_numel and p.is_contiguous() and
                g.dtype == torch.float32 and g.is_contiguous() and not g.is_sparse and p.data_ptr() % 16 == 0 and
                g.data_ptr() % 16 == 0)

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        for group in self.param_groups:
            self._step_group(group, [q for q in group['params'] if id(q) not in self._early_done])
        self._early_done.clear()
        return loss

    def _step_group(self, group, subset):
        """One Adam update of `subset` (parameters of `group` that have a gradient)."""
        if len(subset) != len(group['params']):
            group = dict(group)
            group['params'] = subset
        params, grads, exp_avgs, exp_avg_sqs, max_sqs, steps = [], [], [], [], [], []
        has_complex = self._init_group(group, params, grads, exp_avgs, exp_avg_sqs, max_sqs, steps)
        beta1, beta2 = group['betas']
        big = [i for i, (p, g) in enumerate(zip(params, grads)) if self._is_big(p, g)]
        big_set = set(big)
        small = [i for i in range(len(params)) if i not in big_set]
        pick = lambda lst, idx: [lst[i] for i in idx] if lst else []
        if small:
            _adam.adam(pick(params, small), pick(grads, small), pick(exp_avgs, small), pick(exp_avg_sqs, small),
                       pick(max_sqs, small), pick(steps, small), amsgrad=group['amsgrad'], has_complex=has_complex,
                       beta1=beta1, beta2=beta2, lr=group['lr'], weight_decay=group['weight_decay'], eps=group['eps'],
                       maximize=group['maximize'], foreach=group['foreach'], capturable=True,
                       differentiable=False, fused=group['fused'],
                       decoupled_weight_decay=group.get('decoupled_weight_decay', False))
        if not big:
            return
        torch._foreach_add_(pick(steps, big), 1)
        lr = group['lr']
        lr_dev = lr.data_ptr() if isinstance(lr, torch.Tensor) and lr.is_cuda else None
        args = (native.AdamArgs * len(big))()
        entries = []
        for slot, i in enumerate(big):
            p, g = params[i], grads[i]
            a = args[slot]
            a.p, a.g, a.m, a.v = p.data_ptr(), g.data_ptr(), exp_avgs[i].data_ptr(), exp_avg_sqs[i].data_ptr()
            a.vmax = max_sqs[i].data_ptr() if group['amsgrad'] else None
            a.n = p.numel()
            a.lr = float(lr) if lr_dev is None else 0.0
            a.beta1, a.beta2, a.eps, a.weight_decay = float(beta1), float(beta2), float(group['eps']), float(
                group['weight_decay'])
            a.step, a.lr_dev, a.maximize = steps[i].data_ptr(), lr_dev, int(bool(group['maximize']))
            reg = p.__dict__.get('_vp3d_packed')
            entry = None
            if reg and p.dim() == 3 and p.numel() % 4 == 0:
                # the training forward registered the operand(s) it packs from this weight: refresh the first in the
                # same pass, drop the others (they will be re-packed on demand)
                key = next(iter(reg))
                entry = reg[key]
                for k in list(reg):
                    if k != key:
                        del reg[k]
                dt, _rows_pad, k_pad = key
                a.packed, a.dtype = entry[0].data_ptr(), dt
                a.c_in, a.taps, a.k_pad = p.shape[1], p.shape[2], k_pad
            entries.append(entry)
        # ONE launch for all tensors of the group (on one device, one packed dtype); otherwise tensor by tensor
        devices = {params[i].device for i in big}
        dtypes = {args[k].dtype for k in range(len(big)) if args[k].packed}
        if len(devices) == 1 and len(dtypes) <= 1 and os.environ.get('VP3D_ADAM_MULTI', '1') != '0':
            with torch.cuda.device(params[big[0]].device):
                native.check(native.lib().vp3d_adam_step_multi(args, len(big), ops._stream()), 'adam_step_multi')
        else:
            for slot, i in enumerate(big):
                with torch.cuda.device(params[i].device):
                    native.check(native.lib().vp3d_adam_step(C.byref(args[slot]), ops._stream()), 'adam_step')
        for slot, i in enumerate(big):
            p = params[i]
            torch.autograd.graph.increment_version(p)     # p changed behind autograd's back
            if entries[slot] is not None:
                entries[slot][1] = p._version
