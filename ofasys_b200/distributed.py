"""Data-parallel exchange step of the hot path: the path shards by batch, every rank
holds a full replica, and the only collective is the gradient average after backward
(reference: c10d DDP wrap, ofasys/distributed/distributed_model_dispatcher.py:49-75, then
`multiply_grads(world_size / sum(sample_size))`, ofasys/engine/trainer.py:857-860).

Gradients are packed into ~32 MB flat buckets and averaged with one all-reduce per bucket
(NCCL over NVLink 5 / NVSwitch on B200; gloo in the CPU tests).  A task that did not touch an
adaptor simply contributes zeros for it (the reference needs DDP's find_unused_parameters walk).
"""
from typing import Iterable, List, Optional

import torch
import torch.distributed as dist


def build_buckets(params: Iterable[torch.nn.Parameter], bucket_bytes: int = 32 << 20) -> List[List[torch.nn.Parameter]]:
    """Reverse registration order ~ the order gradients become ready in backward."""
    buckets, cur, size = [], [], 0
    for p in reversed([p for p in params if p.requires_grad]):
        cur.append(p)
        size += p.numel() * p.element_size()
        if size >= bucket_bytes:
            buckets.append(cur)
            cur, size = [], 0
    if cur:
        buckets.append(cur)
    return buckets


def allreduce_grads(buckets: List[List[torch.nn.Parameter]], group=None, scale: Optional[float] = None) -> None:
    """p.grad <- mean over ranks of p.grad (missing grads count as zeros); optional extra `scale`
    (the trainer's world_size / sum(sample_size))."""
    world = dist.get_world_size(group)
    backend = dist.get_backend(group)
    for bucket in buckets:
        grads = []
        for p in bucket:
            if p.grad is None:
                p.grad = torch.zeros_like(p)
            grads.append(p.grad)
        flat = torch.cat([g.reshape(-1) for g in grads])
        if backend == "nccl":
            dist.all_reduce(flat, op=dist.ReduceOp.AVG, group=group)
            if scale is not None:
                flat.mul_(scale)
        else:
            dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
            flat.mul_((scale or 1.0) / world)
        off = 0
        for g in grads:
            g.copy_(flat[off:off + g.numel()].view_as(g))
            off += g.numel()


class GradBuckets:
    """Flat-bucket gradient exchange for the CUDA path: per bucket ONE `ofab_multi_copy` launch packs the
    per-parameter gradients into a persistent flat buffer, one NCCL all-reduce averages it, one launch unpacks.
    The chunk tables are cached by gradient address (static under CUDA-graph replay), so a step costs
    3 launches per bucket instead of one copy kernel per parameter."""

    CHUNK = 256 << 10  # bytes per copy chunk (one thread block each)

    def __init__(self, params: Iterable[torch.nn.Parameter], bucket_bytes: int = 64 << 20):
        self.buckets = build_buckets(params, bucket_bytes)
        self._flat: List[Optional[torch.Tensor]] = [None] * len(self.buckets)
        self._tabs = [None] * len(self.buckets)  # (signature, pack table, unpack table)

    def _tables(self, i, grads):
        import numpy as np

        sig = tuple(g.data_ptr() for g in grads)
        if self._tabs[i] is not None and self._tabs[i][0] == sig:
            return self._tabs[i][1], self._tabs[i][2]
        g0 = grads[0]
        esz = g0.element_size()
        offs, off = [], 0
        for g in grads:
            assert g.dtype == g0.dtype and g.is_contiguous(), "bucket gradients must be contiguous and of one dtype"
            offs.append(off)
            off += (g.numel() * esz + 15) // 16 * 16
        if self._flat[i] is None or self._flat[i].numel() * esz != off:
            self._flat[i] = torch.zeros(off // esz, dtype=g0.dtype, device=g0.device)
        base = self._flat[i].data_ptr()
        rec = []
        for g, o in zip(grads, offs):
            nb = g.numel() * esz
            for c in range(0, nb, self.CHUNK):
                rec.append((g.data_ptr() + c, base + o + c, min(self.CHUNK, nb - c)))
        pack = np.asarray(rec, dtype=np.uint64)
        unpack = pack[:, [1, 0, 2]].copy()
        tp = torch.from_numpy(pack.view(np.int64)).to(g0.device)
        tu = torch.from_numpy(unpack.view(np.int64)).to(g0.device)
        self._tabs[i] = (sig, tp, tu)
        return tp, tu

    def allreduce(self, group=None, scale: Optional[float] = None) -> None:
        """p.grad <- mean over ranks (times `scale`); gradients must live on a CUDA device."""
        from . import _lib

        stream = torch.cuda.current_stream().cuda_stream
        for i, bucket in enumerate(self.buckets):
            grads = []
            for p in bucket:
                if p.grad is None:
                    p.grad = torch.zeros_like(p)
                grads.append(p.grad)
            tp, tu = self._tables(i, grads)
            flat = self._flat[i]
            _lib.call("ofab_multi_copy", tp.data_ptr(), tp.shape[0], stream)
            dist.all_reduce(flat, op=dist.ReduceOp.AVG, group=group)
            if scale is not None:
                flat.mul_(scale)
            _lib.call("ofab_multi_copy", tu.data_ptr(), tu.shape[0], stream)


class DataParallelModel(torch.nn.Module):
    """The data-parallel wrapper surface the reference's trainer expects (`Trainer.model`, engine/trainer.py:352-364,
    766-784; `FairseqOptimizer.all_reduce_grads`, engine/optim/fairseq_optimizer.py:99-103; the `legacy_ddp` wrapper of
    distributed_model_dispatcher.py:76-84): anything with `.forward`, `.parameters()`, `.no_sync()` and
    `.all_reduce_grads()`, forwarding every other attribute to the wrapped module like `ModuleProxyWrapper`.

      with model.no_sync():            # delayed-update loop / all but the last task batch of a step: accumulate locally
          loss.backward()
      model.all_reduce_grads()         # ONE exchange per step: p.grad <- mean over ranks (missing grads count as zeros)
      optimizer.multiply_grads(world_size / sum(sample_sizes))      # trainer.py:857-860

    The exchange is GradBuckets (flat buckets, NCCL over NVLink) on CUDA gradients and the plain bucketed all-reduce
    (gloo) on CPU ones.  Unused parameters -- a task that never touched an adaptor -- need no graph walk: they contribute
    zeros (the reference relies on DDP's find_unused_parameters)."""

    def __init__(self, module: torch.nn.Module, process_group=None, bucket_bytes: int = 64 << 20):
        super().__init__()
        self.module = module
        self.process_group = process_group
        self.bucket_bytes = bucket_bytes
        self.accumulate_grads = False
        self._buckets = None

    def __getattr__(self, name):
        try:
            return super().__getattr__(name)
        except AttributeError:
            return getattr(super().__getattr__("module"), name)

    def state_dict(self, *args, **kwargs):
        return self.module.state_dict(*args, **kwargs)

    def load_state_dict(self, *args, **kwargs):
        return self.module.load_state_dict(*args, **kwargs)

    def forward(self, *args, **kwargs):
        return self.module(*args, **kwargs)

    def no_sync(self):
        """Context manager: all_reduce_grads() inside it is a no-op (gradients keep accumulating locally)."""
        import contextlib

        @contextlib.contextmanager
        def ctx():
            old = self.accumulate_grads
            self.accumulate_grads = True
            try:
                yield
            finally:
                self.accumulate_grads = old

        return ctx()

    def all_reduce_grads(self, scale: Optional[float] = None):
        if self.accumulate_grads or not (dist.is_available() and dist.is_initialized()):
            return
        if dist.get_world_size(self.process_group) == 1 and scale is None:
            return
        params = [p for p in self.module.parameters() if p.requires_grad]
        if params and params[0].is_cuda:
            if self._buckets is None:
                self._buckets = GradBuckets(params, self.bucket_bytes)
            self._buckets.allreduce(self.process_group, scale)
        else:
            if self._buckets is None:
                self._buckets = build_buckets(params, self.bucket_bytes)
            allreduce_grads(self._buckets, self.process_group, scale)


# ----------------------------------------------------------------------------------------------------------------------
# Gradient arena: weight gradients are written by the backward GEMMs STRAIGHT into one flat buffer per model (no
# pack / unpack copies around the collective), cut into ~64 MB buckets in the order gradients become ready; a bucket is
# all-reduced (side stream) as soon as its last gradient has been produced, while the rest of the backward runs.
# Reference: c10d DDP's Reducer with gradient_as_bucket_view + per-bucket ready hooks
# (distributed_model_dispatcher.py:49-59, configs.py `gradient_as_bucket_view`, `bucket_cap_mb`).
# ----------------------------------------------------------------------------------------------------------------------
_ARENA_SLOTS = {}  # (data_ptr of the forward weight tensor, numel) -> (arena view, [params], arena)


def grad_slot(weight: torch.Tensor, shape):
    """Destination for the gradient of `weight` (a Parameter's tensor, or a packed view of several adjacent Parameters) inside
    the active gradient arena, or None (no arena / slot already used in this step / a gradient is already accumulated there --
    autograd then accumulates in place into the slot through p.grad)."""
    if not _ARENA_SLOTS:
        return None
    ent = _ARENA_SLOTS.get((weight.data_ptr(), weight.numel()))
    if ent is None:
        return None
    view, params, arena = ent
    if not arena.enabled or any(p.grad is not None for p in params) or any(id(p) in arena._handed for p in params):
        return None
    for p in params:
        arena._handed.add(id(p))
    return view.view(shape)


class GradArena:
    """Flat gradient storage + bucketed, overlapped exchange for one model replica.

        arena = GradArena(model.parameters())          # once (after the parameters are on the device / packed)
        arena.begin_step()                             # p.grad = None for all
        ... loss_k.backward() for all but the last task batch of the step (gradients accumulate locally) ...
        arena.arm()                                    # the next backward completes the step's gradients
        loss.backward()                                # GEMMs write dW into the arena; buckets all-reduce as they complete
        arena.finish()                                 # leftovers copied in, remaining buckets reduced, streams joined
        # now p.grad (views of the arena) hold the mean over ranks

    A bucket counts as complete when as many of its gradients have been produced in the armed pass as in the previous step's
    armed pass (learned: a pass need not touch every parameter -- an adaptor no batch uses, a text-only task); the first step
    reduces everything in finish().

    Parameters whose storage is adjacent (ops.pack_params: q|k|v, k|v) get adjacent slots in the same order, so the packed
    weight gradient of one GEMM is one contiguous slot.  Gradients that autograd had to sum from several consumers (the tied
    embedding) or that small kernels produce (LayerNorm, biases, tables) arrive as ordinary tensors and are copied into their
    slots by ONE multi-tensor copy per bucket.  A parameter without a gradient in a step (an adaptor no batch touched)
    contributes zeros."""

    def __init__(self, params: Iterable[torch.nn.Parameter], bucket_bytes: int = 64 << 20, group=None, overlap: bool = True):
        self.params = [p for p in params if p.requires_grad]
        assert self.params and all(p.is_cuda for p in self.params), "GradArena needs CUDA parameters"
        self.group = group
        self.overlap = overlap
        self.enabled = True
        dt = self.params[0].dtype
        assert all(p.dtype == dt for p in self.params), "one gradient dtype per arena"
        es = self.params[0].element_size()
        # adjacency groups (packed parameters): consecutive in memory inside one storage
        by_addr = sorted(self.params, key=lambda p: p.data_ptr())
        group_of, groups = {}, []
        for p in by_addr:
            if groups:
                last = groups[-1][-1]
                if (last.untyped_storage().data_ptr() == p.untyped_storage().data_ptr() and last.data_ptr() + last.numel() * es == p.data_ptr()
                        and p.dim() == last.dim() and p.shape[1:] == last.shape[1:]):
                    groups[-1].append(p)
                    group_of[id(p)] = len(groups) - 1
                    continue
            groups.append([p])
            group_of[id(p)] = len(groups) - 1
        # slot order = reverse registration order (~ the order gradients become ready), whole groups at once
        order, seen = [], set()
        for p in reversed(self.params):
            g = group_of[id(p)]
            if g not in seen:
                seen.add(g)
                order.append(groups[g])
        offs, off = {}, 0
        self.buckets = []  # [start_elem, end_elem, [params]]
        cur_start, cur = 0, []
        for grp in order:
            for p in grp:
                offs[id(p)] = off
                off += p.numel()
                cur.append(p)
            off = (off + 7) // 8 * 8  # 16-byte aligned group starts
            if (off - cur_start) * es >= bucket_bytes:
                self.buckets.append([cur_start, off, cur])
                cur_start, cur = off, []
        if cur:
            self.buckets.append([cur_start, off, cur])
        self.flat = torch.zeros(off, dtype=dt, device=self.params[0].device)
        self._slot = {id(p): self.flat[offs[id(p)]:offs[id(p)] + p.numel()].view(p.shape) for p in self.params}
        self._bucket_of = {}
        for bi, (_, _, ps) in enumerate(self.buckets):
            for p in ps:
                self._bucket_of[id(p)] = bi
        for grp in order:  # registry for the backward GEMMs: single parameters and whole packed groups
            for p in grp:
                _ARENA_SLOTS[(p.data_ptr(), p.numel())] = (self._slot[id(p)], [p], self)
            if len(grp) > 1:
                n = sum(p.numel() for p in grp)
                o = offs[id(grp[0])]
                _ARENA_SLOTS[(grp[0].data_ptr(), n)] = (self.flat[o:o + n], list(grp), self)
            # (packed sub-groups, e.g. k|v of a q|k|v-adjacent triple, are registered too)
            for a in range(len(grp)):
                for b in range(a + 2, len(grp) + 1):
                    if a == 0 and b == len(grp):
                        continue
                    n = sum(p.numel() for p in grp[a:b])
                    o = offs[id(grp[a])]
                    _ARENA_SLOTS[(grp[a].data_ptr(), n)] = (self.flat[o:o + n], list(grp[a:b]), self)
        self._handed = set()
        self._fired = [0] * len(self.buckets)
        self._expected = [None] * len(self.buckets)
        self._done = [False] * len(self.buckets)
        self._tabs = [None] * len(self.buckets)
        self.side = torch.cuda.Stream(device=self.flat.device)
        self._hooks = [p.register_post_accumulate_grad_hook(self._on_grad) for p in self.params]
        self._armed = False

    def close(self):
        for h in self._hooks:
            h.remove()
        for k in [k for k, v in _ARENA_SLOTS.items() if v[2] is self]:
            del _ARENA_SLOTS[k]

    # ---- one step
    def begin_step(self, arm: bool = False):
        for p in self.params:
            p.grad = None
        self._handed.clear()
        self._done = [False] * len(self.buckets)
        self._armed = False
        if arm:
            self.arm()

    def arm(self):
        """The next backward pass completes this step's gradients: buckets may be reduced as they fill."""
        self._fired = [0] * len(self.buckets)
        self._armed = True

    def _on_grad(self, p):
        if not self._armed:
            return
        bi = self._bucket_of[id(p)]
        self._fired[bi] += 1
        if self._fired[bi] == self._expected[bi] and self.overlap and dist.is_initialized() and dist.get_world_size(self.group) > 1:
            self._reduce_bucket(bi, async_side=True)

    def _settle_bucket(self, bi):
        """gradients of the bucket that did not land in their slots (small tensors; sums autograd formed from several
        consumers, e.g. the tied embedding): one multi-tensor copy (capturable: no host table); missing ones: zeros"""
        dst, src = [], []
        for p in self.buckets[bi][2]:
            slot = self._slot[id(p)]
            g = p.grad
            if g is None:
                slot.zero_()
            elif g.data_ptr() != slot.data_ptr():
                dst.append(slot)
                src.append(g if g.dtype == slot.dtype else g.to(slot.dtype))
            p.grad = slot
        if dst:
            torch._foreach_copy_(dst, src)

    def _reduce_bucket(self, bi, async_side):
        if self._done[bi]:
            return
        self._done[bi] = True
        self._settle_bucket(bi)
        if not (dist.is_initialized() and dist.get_world_size(self.group) > 1):
            return
        a, b, _ = self.buckets[bi]
        if async_side:
            cur = torch.cuda.current_stream()
            self.side.wait_stream(cur)
            with torch.cuda.stream(self.side):
                dist.all_reduce(self.flat[a:b], op=dist.ReduceOp.AVG, group=self.group)
        else:
            dist.all_reduce(self.flat[a:b], op=dist.ReduceOp.AVG, group=self.group)

    def finish(self, scale: Optional[float] = None):
        """After backward: settle + reduce every bucket not yet reduced, join the side stream; p.grad = arena views."""
        if self._armed:
            self._expected = list(self._fired)  # (static graphs: the same parameters fire in the same pass every step)
        self._armed = False
        for bi in range(len(self.buckets)):
            self._reduce_bucket(bi, async_side=self.overlap)
        torch.cuda.current_stream().wait_stream(self.side)
        if scale is not None:
            self.flat.mul_(scale)
