"""Fused optimizer step for bf16 training (SURVEY 8f "next" #1): what the reference does after every backward
(`engine/trainer.py:857-884`: `multiply_grads` -> `clip_grad_norm` -> `optimizer.step`) through
`engine/optim/fp16_optimizer.py` + `engine/optim/adam.py`, as two multi-tensor launches of libofab
(`csrc/optim.cu`).  Same call surface as the reference's optimizer wrapper (`FairseqOptimizer`,
`fairseq_optimizer.py:15-182`): `multiply_grads(c)`, `clip_grad_norm(max_norm)`, `step()`, `zero_grad()`,
`get_lr()` / `set_lr()`, `state_dict()` / `load_state_dict()`.

State per parameter: fp32 master copy, exp_avg, exp_avg_sq (fp32) -- 12 bytes per element next to the bf16 parameter
and gradient.  There is no CPU fallback.
"""
import ctypes

import torch

from . import _lib


class FusedAdam:
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0):
        self.params = [p for p in params if p.requires_grad]
        if not self.params:
            raise ValueError("FusedAdam: no trainable parameters")
        for p in self.params:
            if not p.is_cuda or p.dtype != torch.bfloat16 or not p.is_contiguous():
                raise _lib.OfabError("FusedAdam needs contiguous CUDA bf16 parameters (no CPU fallback)")
        self.device = self.params[0].device
        self.lr, self.betas, self.eps, self.weight_decay = float(lr), (float(betas[0]), float(betas[1])), float(eps), float(weight_decay)
        self.num_updates = 0
        # one flat fp32 arena per state kind: masters / exp_avg / exp_avg_sq of parameter i are views at a 32-byte-aligned offset
        offs, total = [], 0
        for p in self.params:
            offs.append(total)
            total += (p.numel() + 7) // 8 * 8
        self._offsets, self._total = offs, total
        self._master = torch.zeros(total, dtype=torch.float32, device=self.device)
        self._m = torch.zeros(total, dtype=torch.float32, device=self.device)
        self._v = torch.zeros(total, dtype=torch.float32, device=self.device)
        for p, o in zip(self.params, offs):
            self._master[o:o + p.numel()].copy_(p.detach().reshape(-1))  # build_fp32_params (fp16_optimizer.py:44-77)
        chunk = _lib.lib().ofab_adam_chunk_elems()
        self._first_block, nb = [], 0
        for p in self.params:
            self._first_block.append(nb)
            nb += (p.numel() + chunk - 1) // chunk
        self._n_blocks = nb
        self._host_table = (_lib.AdamTensor * len(self.params))()
        # two pinned staging buffers: the host may run a whole step ahead of the device, so a buffer is rewritten only
        # after the async copy that last read it has completed (event)
        self._pinned = [torch.empty(ctypes.sizeof(self._host_table), dtype=torch.uint8).pin_memory() for _ in range(2)]
        self._pin_ev = [None, None]
        self._pin_i = 0
        self._dev_table = torch.empty(ctypes.sizeof(self._host_table), dtype=torch.uint8, device=self.device)
        self._partial = torch.empty(nb, dtype=torch.float32, device=self.device)
        self._norm = torch.zeros(1, dtype=torch.float32, device=self.device)
        self._factor = 1.0       # deferred multiply_grads factor (fp16_optimizer.py:170-172)
        self._max_norm = 0.0     # set by clip_grad_norm for the coming step
        self._table_ready = False

    # ---- views (tests, checkpoints)
    def master(self, i):
        return self._master[self._offsets[i]:self._offsets[i] + self.params[i].numel()].view_as(self.params[i])

    def exp_avg(self, i):
        return self._m[self._offsets[i]:self._offsets[i] + self.params[i].numel()].view_as(self.params[i])

    def exp_avg_sq(self, i):
        return self._v[self._offsets[i]:self._offsets[i] + self.params[i].numel()].view_as(self.params[i])

    def _stream(self):
        return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)

    def _upload_table(self):
        """Gradient tensors are new allocations after every backward: rebuild the pointer table (host, ~300 records) and
        send it with one async copy from pinned memory."""
        if self._table_ready:
            return
        t = self._host_table
        for i, p in enumerate(self.params):
            g = p.grad
            if g is not None and (g.dtype != torch.bfloat16 or not g.is_contiguous()):
                g = p.grad = g.to(torch.bfloat16).contiguous()
            o = self._offsets[i] * 4
            t[i].p, t[i].g = p.data_ptr(), (None if g is None else g.data_ptr())
            t[i].master, t[i].m, t[i].v = self._master.data_ptr() + o, self._m.data_ptr() + o, self._v.data_ptr() + o
            t[i].n, t[i].first_block = p.numel(), self._first_block[i]
        k = self._pin_i
        self._pin_i ^= 1
        if self._pin_ev[k] is not None:
            self._pin_ev[k].synchronize()
        ctypes.memmove(self._pinned[k].data_ptr(), ctypes.addressof(t), ctypes.sizeof(t))
        self._dev_table.copy_(self._pinned[k], non_blocking=True)
        self._pin_ev[k] = torch.cuda.Event()
        self._pin_ev[k].record()
        self._table_ready = True

    # ---- FairseqOptimizer surface
    def get_lr(self):
        return self.lr

    def set_lr(self, lr):
        self.lr = float(lr)

    def multiply_grads(self, c):
        """Deferred, like the reference's bf16 wrapper: folded into the step's single gradient multiply."""
        self._factor *= float(c)

    def clip_grad_norm(self, max_norm, aggregate_norm_fn=None):
        """Returns the gradient norm (after the deferred multiply factor) as a device tensor -- no host sync; the clip
        coefficient is applied inside step()."""
        if aggregate_norm_fn is not None:
            raise NotImplementedError("aggregate_norm_fn (sharded optimizers) is outside the data-parallel path")
        self._upload_table()
        _lib.call("ofab_grad_norm", ctypes.c_void_p(self._dev_table.data_ptr()), len(self.params), self._n_blocks,
                  ctypes.c_void_p(self._partial.data_ptr()), ctypes.c_void_p(self._norm.data_ptr()), self._stream())
        self._max_norm = float(max_norm)
        return self._norm[0] * self._factor

    def step(self):
        self._upload_table()
        self.num_updates += 1
        h = _lib.AdamHyper()
        h.lr, h.weight_decay, h.beta1_d, h.beta2_d = self.lr, self.weight_decay, self.betas[0], self.betas[1]
        h.beta1, h.beta2, h.eps = self.betas[0], self.betas[1], self.eps
        h.grad_scale, h.max_norm, h.step = self._factor, self._max_norm, self.num_updates
        h.norm = self._norm.data_ptr() if self._max_norm > 0 else None
        _lib.call("ofab_adam_step", ctypes.c_void_p(self._dev_table.data_ptr()), len(self.params), self._n_blocks, ctypes.byref(h), self._stream())
        self._factor, self._max_norm, self._table_ready = 1.0, 0.0, False

    def zero_grad(self):
        for p in self.params:
            p.grad = None
        self._table_ready = False

    def state_dict(self):
        """torch.optim-style layout, as the reference's fp32 optimizer emits through `fp32_optimizer.state_dict()`
        (engine/optim/fp16_optimizer.py:104-124 -> torch.optim.Optimizer.state_dict): `state[i] = {step, exp_avg, exp_avg_sq}`
        per parameter index (fp32, the parameter's shape) + one `param_groups` entry; the fp32 master copies ride along as
        `master_params` (the reference keeps them as the fp32 optimizer's own parameters)."""
        state = {i: {"step": self.num_updates, "exp_avg": self.exp_avg(i).clone(), "exp_avg_sq": self.exp_avg_sq(i).clone()}
                 for i in range(len(self.params))}
        group = {"lr": self.lr, "betas": self.betas, "eps": self.eps, "weight_decay": self.weight_decay, "amsgrad": False,
                 "params": list(range(len(self.params)))}
        return {"state": state, "param_groups": [group], "master_params": [self.master(i).clone() for i in range(len(self.params))]}

    def load_state_dict(self, sd):
        """Accepts the layout of state_dict() (and a reference / torch.optim Adam checkpoint of the same parameter order:
        `master_params` optional -- the masters are then rebuilt from the bf16 parameters).  Shapes are validated; the bf16
        parameters are re-synchronised from the masters."""
        groups = sd["param_groups"]
        if len(groups) != 1 or len(groups[0]["params"]) != len(self.params):
            raise ValueError(f"FusedAdam.load_state_dict: expected one param group of {len(self.params)} parameters")
        g = groups[0]
        self.lr, self.betas, self.eps, self.weight_decay = float(g["lr"]), tuple(float(b) for b in g["betas"]), float(g["eps"]), float(g["weight_decay"])
        state = sd["state"]
        steps = set()
        for i, p in enumerate(self.params):
            st = state.get(i, state.get(str(i)))
            if st is None:  # a parameter that never received a gradient has no state in torch's layout
                self.exp_avg(i).zero_()
                self.exp_avg_sq(i).zero_()
                continue
            for key, dst in (("exp_avg", self.exp_avg(i)), ("exp_avg_sq", self.exp_avg_sq(i))):
                if tuple(st[key].shape) != tuple(p.shape):
                    raise ValueError(f"FusedAdam.load_state_dict: state[{i}].{key} has shape {tuple(st[key].shape)}, parameter has {tuple(p.shape)}")
                dst.copy_(st[key].to(self.device, torch.float32))
            steps.add(int(st["step"]))
        if len(steps) > 1:
            raise ValueError(f"FusedAdam.load_state_dict: per-parameter step counts differ ({sorted(steps)}); one shared update count is kept")
        self.num_updates = steps.pop() if steps else 0
        masters = sd.get("master_params")
        for i, p in enumerate(self.params):
            if masters is not None:
                if tuple(masters[i].shape) != tuple(p.shape):
                    raise ValueError(f"FusedAdam.load_state_dict: master_params[{i}] has shape {tuple(masters[i].shape)}, parameter has {tuple(p.shape)}")
                self.master(i).copy_(masters[i].to(self.device, torch.float32))
                with torch.no_grad():
                    p.copy_(self.master(i))  # the bf16 parameter is always bf16(master)
            else:
                self.master(i).copy_(p.detach().float())
        self._table_ready = False
