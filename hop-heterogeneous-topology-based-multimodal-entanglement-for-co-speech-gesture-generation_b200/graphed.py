"""Whole-step CUDA graph of the HOP training step.

The TED step at batch 128 is ~2400 kernel launches for ~20 ms of GPU work: issued one by one from Python it is bound by
the host (the step takes the same 26-27 ms at batch 32, 64 and 128).  ``GraphedTrainStep`` captures one call of
``train_llm_device`` -- generator forward, discriminator forward, random-speaker forward, backward with its side streams
and (with data parallelism) its bucketed NCCL all-reduces, both Adam steps -- into a ``torch.cuda.CUDAGraph`` and replays
it per batch: inputs are copied into static device buffers, the reported scalars are read back after the replay.

What makes the step capturable:
  * ``train_llm_device`` has no host synchronisation (one ``.tolist()`` *after* the replay);
  * every hand-written kernel takes its stream from the caller, its workspaces from PyTorch's allocator (the graph's
    private pool under capture), and forks / joins its side streams with events;
  * the attention dropout seed is a launch argument, so one ``hopk_dropout_epoch_advance`` is captured at the top of
    the step: the device-side epoch it bumps is added to the seed inside the kernels, giving every replay a fresh mask;
  * torch's own RNG consumers (``randn_like``, ``randperm``) are graph-safe Philox users;
  * optimisers must be built with ``capturable=True``.
"""
import torch

from ._lib import check, lib, stream_ptr
from .train_llm import finish_losses, train_llm_device


class GraphedTrainStep:
    def __init__(self, args, epoch, model, discriminator, model_optim, dis_optimizer, accelerator, example_batch, warmup=3):
        self.args, self.epoch = args, epoch
        self.model, self.discriminator = model, discriminator
        self.model_optim, self.dis_optimizer, self.accelerator = model_optim, dis_optimizer, accelerator
        self.static = [t.clone() for t in example_batch]
        launches0 = lib().hopk_launch_count()
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):                             # lazy initialisation, cuDNN plans, optimiser state, DP buckets
            for _ in range(warmup):
                self._run()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        launches1 = lib().hopk_launch_count()
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            check(lib().hopk_dropout_epoch_advance(0, stream_ptr()))
            self.names, self.out = self._run()
        self.launches_per_step = int(lib().hopk_launch_count() - launches1)     # hand-written kernels inside one replay
        self.warmup_launches = int(launches1 - launches0)

    def _run(self):
        return train_llm_device(self.args, self.epoch, *self.static, self.model, self.discriminator, self.model_optim,
                                self.dis_optimizer, self.accelerator)

    def launch(self, batch):
        """Enqueue one step (copy the batch into the static inputs, replay) without waiting for it."""
        for s, t in zip(self.static, batch):
            s.copy_(t, non_blocking=True)
        self.graph.replay()

    def result(self):
        """The reported scalars of the last launched step (the step's only host synchronisation)."""
        return finish_losses(self.names, self.out.tolist())

    # ---- pipelined use: the host stays one step ahead of the GPU (the replay of a 2000-node graph costs the CPU about a
    # millisecond; waiting for step i's scalars before launching step i+1 leaves the GPU idle for that long every step)
    def launch_async(self, batch):
        """``launch`` + an asynchronous device->host copy of the step's scalars into a pinned slot; returns a ticket."""
        if not hasattr(self, '_slots'):
            self._slots = [torch.empty(self.out.shape, dtype=self.out.dtype).pin_memory() for _ in range(2)]
            self._events = [torch.cuda.Event() for _ in range(2)]
            self._n = 0
        self.launch(batch)
        k = self._n % 2
        self._slots[k].copy_(self.out, non_blocking=True)
        self._events[k].record()
        self._n += 1
        return k

    def collect(self, ticket):
        """Host floats of the step behind ``ticket`` (waits for that step only; at most one newer step may be in flight)."""
        self._events[ticket].synchronize()
        return finish_losses(self.names, self._slots[ticket].tolist())

    def __call__(self, batch):
        self.launch(batch)
        return self.result()
