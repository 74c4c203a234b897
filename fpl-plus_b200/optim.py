"""``FusedAdam``: torch.optim.Adam(params, lr, weight_decay) -- the optimiser every FPL+ .cfg selects
(reference: PyMIC/pymic/net_run/get_optimizer.py:16-17, coupled L2 weight decay) -- as ONE kernel launch over all
parameters (csrc/adam.cu) instead of a multi-tensor library call.

Same observable behaviour as ``torch.optim.Adam``:

* parameters whose ``.grad`` is None are skipped and keep their own step counter (a single-domain step leaves the other
  domain's BatchNorm untouched, exactly as in the reference);
* ``state_dict()`` / ``load_state_dict()`` use torch's layout (per parameter ``step``, ``exp_avg``, ``exp_avg_sq``;
  ``param_groups`` with ``lr``, ``betas``, ``eps``, ``weight_decay``), so ``.pt`` checkpoints move between this class
  and ``torch.optim.Adam`` in both directions (agent_seg.py:793-805, :721-734);
* ``param_groups[i]['lr']`` may be a float or a 0-dim CUDA tensor; a tensor is read by the kernel on the device, so a
  step captured into a CUDA graph follows later ``lr.fill_()`` calls.

Moments live in two flat fp32 buffers per group (views of them are the ``exp_avg`` / ``exp_avg_sq`` state entries); the
per-call tensor table is cached per set of (parameter, gradient) pointers, which are stable when ``p.grad`` are views
of the network's master gradient buffer (net.py ``_deliver_grads``).
"""
import ctypes
import struct

import torch

from .ops import call, ptr, stream_ptr


class FusedAdam(torch.optim.Optimizer):
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0):
        if not torch.is_tensor(lr) and lr < 0.0:
            raise ValueError("Invalid learning rate: {}".format(lr))
        if not 0.0 <= betas[0] < 1.0 or not 0.0 <= betas[1] < 1.0:
            raise ValueError("Invalid beta parameters: {}".format(betas))
        if eps < 0.0 or weight_decay < 0.0:
            raise ValueError("Invalid eps / weight_decay: {} / {}".format(eps, weight_decay))
        # the extra keys keep param_groups interchangeable with torch.optim.Adam's state_dict
        defaults = dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay, amsgrad=False, maximize=False,
                        foreach=None, capturable=False, differentiable=False, fused=None, decoupled_weight_decay=False)
        super().__init__(params, defaults)
        self._flat = {}       # group index -> dict(m, v, steps, off, done)
        self._tables = {}     # (group index, ((param ptr, grad ptr), ...)) -> (segs tensor, chunks tensor, nseg, nchunks)

    # -- flat moment storage ---------------------------------------------------------------------
    def _group_storage(self, gi, group):
        st = self._flat.get(gi)
        params = group['params']
        dev = params[0].device
        if st is None or st["m"].device != dev:
            off, o = {}, 0
            for i, p in enumerate(params):
                off[i] = o
                o += (p.numel() + 3) // 4 * 4
            st = self._flat[gi] = {"m": torch.zeros(o, dtype=torch.float32, device=dev),
                                   "v": torch.zeros(o, dtype=torch.float32, device=dev),
                                   "steps": torch.zeros(len(params), dtype=torch.float32, device=dev),
                                   "done": torch.zeros(1, dtype=torch.int32, device=dev), "off": off}
            self._tables = {k: v for k, v in self._tables.items() if k[0] != gi}
            # adopt state that exists already (load_state_dict before the first step)
            for i, p in enumerate(params):
                s = self.state.get(p)
                if s:
                    self._adopt(st, i, p, s)
        return st

    def _views(self, st, i, p):
        o = st["off"][i]
        return st["m"][o:o + p.numel()].view_as(p), st["v"][o:o + p.numel()].view_as(p), st["steps"][i]

    def _adopt(self, st, i, p, s):
        """Copy a (loaded) state entry into the flat buffers and re-point it at the views."""
        m, v, step = self._views(st, i, p)
        if s.get("exp_avg") is not None and s["exp_avg"].data_ptr() != m.data_ptr():
            m.copy_(s["exp_avg"].to(m.device, torch.float32))
            v.copy_(s["exp_avg_sq"].to(v.device, torch.float32))
            step.fill_(float(s["step"]))
        s["exp_avg"], s["exp_avg_sq"], s["step"] = m, v, step

    def load_state_dict(self, state_dict):
        super().load_state_dict(state_dict)
        self._tables = {}
        for gi, group in enumerate(self.param_groups):
            if gi in self._flat:
                for i, p in enumerate(group['params']):
                    s = self.state.get(p)
                    if s:
                        self._adopt(self._flat[gi], i, p, s)

    # -- the step --------------------------------------------------------------------------------
    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        for gi, group in enumerate(self.param_groups):
            params = group['params']
            active = [i for i, p in enumerate(params) if p.grad is not None]
            if not active:
                continue
            st = self._group_storage(gi, group)
            key = (gi, tuple((params[i].data_ptr(), params[i].grad.data_ptr()) for i in active))
            tab = self._tables.get(key)
            if tab is None:
                tab = self._tables[key] = self._build_table(st, params, active)
                if len(self._tables) > 16:                      # foreign gradients change address every step
                    for k in list(self._tables)[:8]:
                        if k != key:
                            del self._tables[k]
            segs, chunks, nseg, nchunks = tab
            lr = group['lr']
            lr_dev = lr if torch.is_tensor(lr) and lr.is_cuda else None
            if lr_dev is not None and lr_dev.dtype != torch.float32:
                raise TypeError("a device learning rate must be float32")
            b1, b2 = group['betas']
            call("fpl_adam_multi_tensor", ptr(segs), nseg, ptr(chunks), nchunks, ptr(lr_dev),
                 0.0 if lr_dev is not None else float(lr), float(b1), float(b2), float(group['eps']),
                 float(group['weight_decay']), ptr(st["done"]), stream_ptr())
        return loss

    def _build_table(self, st, params, active):
        chunk = int(ctypes.c_int(_chunk_elems()).value)
        rows, chunks = bytearray(), []
        for row, i in enumerate(active):
            p = params[i]
            if p.dtype != torch.float32 or not p.is_cuda or not p.is_contiguous():
                raise TypeError("FusedAdam handles contiguous float32 CUDA parameters only")
            g = p.grad
            if g.dtype != torch.float32 or not g.is_contiguous() or g.is_sparse:
                raise TypeError("FusedAdam needs dense contiguous float32 gradients")
            s = self.state[p]
            if not s:
                m, v, step = self._views(st, i, p)
                s["step"], s["exp_avg"], s["exp_avg_sq"] = step, m, v
            elif s["exp_avg"].data_ptr() != self._views(st, i, p)[0].data_ptr():
                self._adopt(st, i, p, s)
            ptrs = (p.data_ptr(), g.data_ptr(), s["exp_avg"].data_ptr(), s["exp_avg_sq"].data_ptr(), s["step"].data_ptr())
            if any(a % 16 for a in ptrs[:4]):
                raise ValueError("FusedAdam: tensors must be 16-byte aligned")
            rows += struct.pack("<5Qii", *ptrs, p.numel(), 0)
            chunks += [(row, c0) for c0 in range(0, p.numel(), chunk)]
        dev = params[0].device
        segs = torch.frombuffer(rows, dtype=torch.uint8).to(dev)
        ch = torch.tensor(chunks, dtype=torch.int32).to(dev)
        return segs, ch, len(active), len(chunks)


_CHUNK = None


def _chunk_elems():
    global _CHUNK
    if _CHUNK is None:
        from . import lib as _lib
        _CHUNK = int(_lib.load().fpl_adam_chunk_elems())
    return _CHUNK
