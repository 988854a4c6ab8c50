"""ZeRO-2 style data-parallel AdamW over the flat parameter buckets (script/zero2.json:16-22;
optimizer = adamw_torch, muffin/train/train_llava15.py:75; lr/wd/schedule flags of
script/train/llava15_train.sh:31-34).

Every bucket of ParamStore (embed | layer i | head | projector) is cut into `world` equal slices;
rank r owns slice r of every bucket: fp32 master weights and Adam moments exist only for owned
slices (12 B/param / world), gradients are reduce-scattered in place into the owned slice as soon
as a bucket's backward has finished (overlapping the rest of the backward on a side stream), the
fused AdamW kernel updates the slice and writes bf16 straight into the parameter buffer, and an
in-place all-gather redistributes the updated parameters.  With world == 1 the collectives vanish.

torch.distributed (NCCL) is the transport; nothing here computes on the host.
"""
import math

import torch
import torch.distributed as dist

from . import lib as _lib
from . import ops

_F32 = torch.float32


def cosine_lr(step, total_steps, base_lr, warmup_ratio=0.05):
    """HF get_cosine_schedule_with_warmup value used for optimizer step `step` (0-based)."""
    warm = math.ceil(total_steps * warmup_ratio)
    if step < warm:
        return base_lr * step / max(1, warm)
    prog = (step - warm) / max(1, total_steps - warm)
    return base_lr * max(0.0, 0.5 * (1.0 + math.cos(math.pi * prog)))


class OptBucket:
    """One trainable bucket as the optimizer sees it: flat bf16 parameter / gradient views of equal length
    (a multiple of 8*world), of which the first `decay_size` elements get weight decay."""

    def __init__(self, name, flat, grad, decay_size):
        self.name, self.flat, self.grad, self.decay_size = name, flat, grad, decay_size
        self.size = flat.numel()


def store_buckets(store, names=None):
    """OptBuckets of a ParamStore / LoraStore (optionally only the named ones)."""
    out = []
    for b in store.buckets:
        if names is None or b.name in names:
            out.append(OptBucket(b.name, store.flat[b.start:b.start + b.size], store.grad[b.start:b.start + b.size],
                                 b.decay_size))
    return out


class Zero2AdamW:
    def __init__(self, buckets, lr=5e-7, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.01, rank=0, world=1,
                 group=None, gather_order=None):
        if not isinstance(buckets, (list, tuple)):         # a ParamStore: every bucket is trainable
            buckets = store_buckets(buckets)
        self.buckets = list(buckets)
        self.index = {b.name: i for i, b in enumerate(self.buckets)}
        self.lr, self.betas, self.eps, self.wd = lr, betas, eps, weight_decay
        self.rank, self.world, self.group = rank, world, group
        self.step_count = 0
        dev = self.buckets[0].flat.device
        self.slices = []   # (bucket, s0, s1, offset into the owned fp32 state), s0/s1 relative to the bucket
        total = 0
        for b in self.buckets:
            assert b.size % (world * 8) == 0, (b.name, b.size)
            n = b.size // world
            self.slices.append((b, rank * n, (rank + 1) * n, total))
            total += n
        self.owned = total
        self.master = torch.empty(total, dtype=_F32, device=dev)
        self.exp_avg = torch.zeros(total, dtype=_F32, device=dev)
        self.exp_avg_sq = torch.zeros(total, dtype=_F32, device=dev)
        for b, s0, s1, o in self.slices:
            self.master[o:o + (s1 - s0)].copy_(b.flat[s0:s1])
        self.comm_stream = torch.cuda.Stream(device=dev) if world > 1 else None
        self.opt_stream = None
        self._gathered = {}
        self._reduced = {}
        self._done = set()
        # order in which the next forward needs the parameters (default: as given)
        self.gather_order = [self.index[n] for n in gather_order if n in self.index] if gather_order else \
            list(range(len(self.buckets)))

    # ---- gradient reduction (called from the backward as buckets complete) ----
    def reduce_bucket(self, bucket_index):
        if self.world == 1:
            return
        if isinstance(bucket_index, str):
            if bucket_index not in self.index:
                return
            bucket_index = self.index[bucket_index]
        b, s0, s1, _ = self.slices[bucket_index]
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream())
        with torch.cuda.stream(self.comm_stream):
            self.comm_stream.wait_event(ev)
            dist.reduce_scatter_tensor(b.grad[s0:s1], b.grad, op=dist.ReduceOp.SUM, group=self.group)
            done = torch.cuda.Event()
            done.record(self.comm_stream)
        self._reduced[bucket_index] = done

    def reduce_all(self):
        for i in range(len(self.slices)):
            self.reduce_bucket(i)

    # ---- optimizer step ----
    def _adamw_bucket(self, bi, lr):
        b, s0, s1, o = self.slices[bi]
        b1, b2 = self.betas
        dec_end = min(s1, b.decay_size)
        if dec_end > s0:
            n = dec_end - s0
            ops.adamw_step(self.master[o:o + n], self.exp_avg[o:o + n], self.exp_avg_sq[o:o + n],
                           b.grad[s0:dec_end], b.flat[s0:dec_end], lr, b1, b2, self.eps, self.wd, self.step_count)
        nd0 = max(s0, b.decay_size)
        if s1 > nd0:
            n = s1 - nd0
            oo = o + (nd0 - s0)
            ops.adamw_step(self.master[oo:oo + n], self.exp_avg[oo:oo + n], self.exp_avg_sq[oo:oo + n],
                           b.grad[nd0:s1], b.flat[nd0:s1], lr, b1, b2, self.eps, 0.0, self.step_count)

    def begin_step(self, lr=None):
        """Open an optimizer step whose buckets are updated one by one (`step_bucket`) while the backward
        is still running; `finish_step` updates whatever is left and closes the step."""
        self._lr = self.lr if lr is None else lr
        self.step_count += 1
        self._done = set()
        if self.opt_stream is None:
            self.opt_stream = torch.cuda.Stream(device=self.buckets[0].flat.device)

    def step_bucket(self, name):
        """AdamW (+ parameter all-gather) of one bucket on the optimizer side stream, ordered after everything
        enqueued so far on the current stream and after this bucket's gradient reduce-scatter. The caller
        guarantees the bucket's gradients are final and its parameters are no longer read this step."""
        bi = self.index.get(name, -1)
        if bi < 0 or bi in self._done:
            return
        self._done.add(bi)
        cur = torch.cuda.current_stream()
        ev = torch.cuda.Event()
        ev.record(cur)
        self.opt_stream.wait_event(ev)
        if self.world > 1:
            rs = self._reduced.pop(bi, None)
            if rs is not None:
                self.opt_stream.wait_event(rs)
        with torch.cuda.stream(self.opt_stream), _lib.use_stream(self.opt_stream):
            self._adamw_bucket(bi, self._lr)
            if self.world > 1:
                b, s0, s1, _ = self.slices[bi]
                dist.all_gather_into_tensor(b.flat, b.flat[s0:s1], group=self.group)
            done = torch.cuda.Event()
            done.record(self.opt_stream)
        self._gathered[bi] = done

    def finish_step(self):
        for bi in self.gather_order:
            self.step_bucket(self.buckets[bi].name)
        # single-GPU callers read the parameters right away on the main stream: order it after the side stream
        if self.world == 1:
            self.wait_all()

    def step(self, lr=None):
        """Whole optimizer step at once (no overlap with the backward)."""
        self.begin_step(lr)
        if self.world > 1:
            torch.cuda.current_stream().wait_stream(self.comm_stream)
        self.finish_step()

    def wait_bucket(self, bucket_index):
        """Make the current stream wait for bucket's parameter all-gather of the last step (no-op if none)."""
        if isinstance(bucket_index, str):
            if bucket_index not in self.index:
                raise KeyError("wait_bucket: %r is not a bucket of this optimizer (%s ...)"
                               % (bucket_index, ", ".join(list(self.index)[:4])))
            bucket_index = self.index[bucket_index]
        ev = self._gathered.pop(bucket_index, None)
        if ev is not None:
            torch.cuda.current_stream().wait_event(ev)

    def wait_all(self):
        for bi in list(self._gathered):
            self.wait_bucket(bi)

    def state_dict(self):
        return {"step": self.step_count, "master": self.master, "exp_avg": self.exp_avg,
                "exp_avg_sq": self.exp_avg_sq, "rank": self.rank, "world": self.world}

    def load_state_dict(self, sd):
        assert sd["world"] == self.world and sd["rank"] == self.rank
        self.step_count = sd["step"]
        self.master.copy_(sd["master"])
        self.exp_avg.copy_(sd["exp_avg"])
        self.exp_avg_sq.copy_(sd["exp_avg_sq"])
