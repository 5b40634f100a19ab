"""Batch-sharded data parallelism for the HOP step: one process per GPU, NCCL over NVLink/NVSwitch.

Replaces what the reference gets implicitly from HF Accelerate + DeepSpeed/DDP
(run_ted.py:110-112, 363-364):

* replicas start identical (rank-0 broadcast of parameters and buffers); BatchNorm statistics stay
  per rank, like the reference (no SyncBatchNorm);
* gradients are all-reduced in ~25 MB buckets on a side stream *while backward is still running*:
  each parameter's post-accumulate hook counts its bucket down; when the last member arrives ONE multi-tensor copy
  packs the bucket's gradients into its flat wire buffer (bf16 with ``grad_dtype=torch.bfloat16``: half the bytes on
  the wire; fp32 by default) and the all-reduce is launched;
* after the collective one multi-tensor copy per bucket unpacks into a flat fp32 buffer whose views become ``p.grad``
  (fp32 wire format: the wire buffer itself is handed over, no unpack);
* tensors the reference never gives a gradient (audio_encoder.*) are discovered on the
  first step and skipped afterwards;
* ``mapping_layer`` (183 MB of the 263 MB of gradients, produced last) is never all-reduced: its upstream
  gradient ``dSource`` (1500x768, 4.6 MB) is all-reduced instead and every rank forms the identical
  ``dW = dSource @ WE^T`` locally (valid because ``source`` is linear in ``W_map`` and WE is frozen and
  replicated) -- see ``Model.source_embeddings``.

``DataParallel.backward(loss)`` is the only call the training step needs (the reference's
``accelerator.backward``).  With world_size 1 (or no process group) it degrades to ``loss.backward()``.
"""
import torch
import torch.distributed as dist


class _Bucket:
    def __init__(self, params, device, wire_dtype=torch.float32):
        self.params = params
        self.numel = sum(p.numel() for p in params)
        self.flat = torch.zeros(self.numel, device=device, dtype=wire_dtype)          # what travels
        self.flat32 = self.flat if wire_dtype == torch.float32 else torch.zeros(self.numel, device=device, dtype=torch.float32)
        self.views, self.views32, off = [], [], 0
        for p in params:
            self.views.append(self.flat[off:off + p.numel()].view(p.shape))
            self.views32.append(self.flat32[off:off + p.numel()].view(p.shape))
            off += p.numel()
        self.pending = len(params)
        self.work = None
        self.event = None

    def pack(self):
        """Gradients -> wire buffer: one multi-tensor copy (members without a gradient this step count as zero)."""
        have = [(v, p.grad) for v, p in zip(self.views, self.params) if p.grad is not None]
        for v, p in zip(self.views, self.params):
            if p.grad is None:
                v.zero_()
        if have:
            torch._foreach_copy_([v for v, _ in have], [g for _, g in have])

    def widen(self):
        """Averaged wire buffer -> the flat fp32 buffer (no-op for an fp32 wire).  Issued on the communication stream right
        behind the bucket's all-reduce, so only the last bucket's copy is left for the end of backward."""
        if self.flat32 is not self.flat:
            self.flat32.copy_(self.flat)

    def unpack(self):
        """Hand the averaged fp32 views to the optimiser."""
        for v, p in zip(self.views32, self.params):
            if p.grad is not None:
                p.grad = v


class _SourceReducer:
    """What ``Model.set_source_grad_reducer`` receives.  Called (``r(t)``) it all-reduces ``t`` in stream order; on CUDA it
    also offers ``defer(t, finish)``: the all-reduce is issued on the communication stream at once, and
    ``_ModuleDP.end`` later waits for it and turns ``finish(t) -> (dW, db)`` into the mapping layer's gradients."""

    def __init__(self, owner, part):
        self.o, self.part = owner, part
        if owner.comm_stream is not None:
            self.defer = self._defer

    def __call__(self, t):
        return self.o._reduce_now(t)

    def _defer(self, t, finish):
        o = self.o
        ev = torch.cuda.Event()
        ev.record()
        with torch.cuda.stream(o.comm_stream):
            o.comm_stream.wait_event(ev)
            work = dist.all_reduce(t, op=o._avg, group=o.group, async_op=True)
        t.record_stream(o.comm_stream)
        o.stats['allreduce_bytes'] += t.numel() * t.element_size()
        self.part.deferred.append((t, work, finish))
        self.part.fired = True


class _ModuleDP:
    """Hooks + buckets of one module (generator and discriminator get one each, so a backward that only
    touches one of them only reduces that one)."""

    def __init__(self, module, owner):
        self.module, self.o = module, owner
        self.buckets = None
        self._seen, self._order, self._where = set(), [], {}
        self.fired = False
        self.pre_reduced = set()
        self.deferred = []
        if hasattr(module, 'set_source_grad_reducer'):          # the dSource trick (see module docstring)
            module.set_source_grad_reducer(_SourceReducer(owner, self))
            self.pre_reduced = set(id(p) for p in module.mapping_layer.parameters())
        seen = set()
        for p in module.parameters():
            if p.requires_grad and id(p) not in seen and id(p) not in self.pre_reduced:
                seen.add(id(p))
                p.register_post_accumulate_grad_hook(self._on_grad)

    def _build_buckets(self, used):
        device = used[0].device
        cur, size = [], 0
        self.buckets = []
        for p in used:
            cur.append(p); size += p.numel() * 4
            if size >= self.o.bucket_bytes:
                self.buckets.append(_Bucket(cur, device, self.o.grad_dtype)); cur, size = [], 0
        if cur:
            self.buckets.append(_Bucket(cur, device, self.o.grad_dtype))
        for bi, b in enumerate(self.buckets):
            for pi, p in enumerate(b.params):
                self._where[id(p)] = (bi, pi)

    def _on_grad(self, p):
        self.fired = True
        if self.buckets is None:                                # discovery step: record readiness order
            if id(p) not in self._seen:
                self._seen.add(id(p)); self._order.append(p)
            return
        where = self._where.get(id(p))
        if where is None:
            return
        b = self.buckets[where[0]]
        b.pending -= 1
        if b.pending == 0 and self.o.overlap:
            b.pack()
            self.o._launch(b)

    def begin(self):
        self.fired = False
        self.deferred = []
        if self.buckets is not None:
            for b in self.buckets:
                b.pending = len(b.params)
                b.work = None

    def _finish_deferred(self):
        for t, work, finish in self.deferred:                   # dSource: averaged by now (or soon: wait() orders the stream)
            work.wait()
            dw, db = finish(t)
            lin = self.module.mapping_layer
            for p, g in ((lin.weight, dw), (lin.bias, db)):
                if p.requires_grad:
                    p.grad = g if p.grad is None else p.grad + g
        self.deferred = []

    def end(self):
        if not self.fired:
            return
        if self.buckets is None and not self._order:            # nothing but the deferred dSource reduction this backward
            self._finish_deferred()
            return
        if self.buckets is None:
            # first active step: the backward was a plain one; bucket in readiness order and reduce now
            used = [p for p in self._order if p.grad is not None]
            self._build_buckets(used)
            for b in self.buckets:
                b.pack()
                self.o._launch(b)
        else:
            for b in self.buckets:                              # a bucket some member of which got no gradient this step
                if b.work is None:
                    b.pack()
                    self.o._launch(b)
        for b in self.buckets:
            b.work.wait()                                       # orders the current stream after the collective
        if self.o.comm_stream is not None:
            torch.cuda.current_stream().wait_stream(self.o.comm_stream)
        for b in self.buckets:                                  # already averaged: the all-reduce runs with ReduceOp.AVG
            b.unpack()                                          # hand the averaged views to the optimiser
        self._finish_deferred()


class DataParallel:
    def __init__(self, modules, bucket_mb=25, process_group=None, broadcast=True, grad_dtype=torch.float32, overlap=True):
        self.modules = list(modules) if isinstance(modules, (list, tuple)) else [modules]
        self.group = process_group
        self.world = dist.get_world_size(self.group) if dist.is_available() and dist.is_initialized() else 1
        self.bucket_bytes = int(bucket_mb * (1 << 20))
        self.grad_dtype = grad_dtype                            # wire format of the bucketed gradients (fp32 or bf16)
        self.overlap = overlap                                  # False: every bucket is reduced after backward (no SM sharing with it)
        self.stats = dict(allreduce_bytes=0, buckets=0)
        self.comm_stream = None
        self.parts = []
        self._avg = dist.ReduceOp.SUM
        if self.world > 1:
            dev = next(self.modules[0].parameters()).device
            if dev.type == 'cuda':
                self.comm_stream = torch.cuda.Stream(device=dev)
                self._avg = dist.ReduceOp.AVG                   # NCCL averages inside the collective: no separate scaling pass
            if broadcast:
                for m in self.modules:
                    for t in list(m.parameters()) + list(m.buffers()):
                        dist.broadcast(t.data, src=0, group=self.group)
            self.parts = [_ModuleDP(m, self) for m in self.modules]

    def _reduce_now(self, t):
        """Stream-ordered mean all-reduce used inside autograd for the small dSource tensor."""
        if self.world > 1:
            dist.all_reduce(t, op=self._avg, group=self.group)
            if self._avg is dist.ReduceOp.SUM:
                t.div_(self.world)
            self.stats['allreduce_bytes'] += t.numel() * t.element_size()
        return t

    def _launch(self, b):
        if self.comm_stream is not None:
            ev = torch.cuda.Event()
            ev.record()                                         # gradients of this bucket are complete here
            with torch.cuda.stream(self.comm_stream):
                self.comm_stream.wait_event(ev)
                b.work = dist.all_reduce(b.flat, op=self._avg, group=self.group, async_op=True)
                b.work.wait()                                   # orders the communication stream (not the host) behind the collective
                b.widen()
        else:
            b.work = dist.all_reduce(b.flat, op=self._avg, group=self.group, async_op=True)
            b.work.wait()
            if self._avg is dist.ReduceOp.SUM:                  # gloo (CPU tests) has no AVG: scale after the sum
                b.flat.div_(self.world)
            b.widen()
        self.stats['allreduce_bytes'] += b.numel * b.flat.element_size()
        self.stats['buckets'] += 1

    def backward(self, loss):
        """The reference's ``accelerator.backward(loss)``: backward + averaged gradients on every rank."""
        for part in self.parts:
            part.begin()
        loss.backward()
        for part in self.parts:
            part.end()

    def broadcast_buffers(self, src=0):
        """Checkpoint-time buffer sync (BatchNorm running stats are per rank during training)."""
        if self.world > 1:
            for m in self.modules:
                for t in m.buffers():
                    dist.broadcast(t.data, src=src, group=self.group)
