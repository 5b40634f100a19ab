"""CPU, world_size 2, gloo: the data-parallel engine (hop_b200/dp.py) averages gradients exactly like a
single process over the concatenated batch, skips never-used parameters, keeps buffers per rank and
implements the dSource trick (all-reduce the small upstream gradient, form dW_map locally)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp
import torch.nn as nn


class Toy(nn.Module):
    def __init__(self):
        super().__init__()
        torch.manual_seed(0)
        self.a = nn.Linear(6, 40)
        self.b = nn.Linear(40, 3)
        self.unused = nn.Linear(5, 5)                       # never gets a gradient (like audio_encoder.*)
        self.mapping_layer = nn.Linear(12, 7)               # the dSource trick target
        self.register_buffer('we', torch.randn(12, 4))
        self.c = nn.Linear(4, 1)
        self._red = None

    def set_source_grad_reducer(self, fn):
        self._red = fn

    def forward(self, x):
        from hop_b200.HOP import _SourceFn
        src = _SourceFn.apply(self.mapping_layer.weight, self.mapping_layer.bias, self.we, self._red, torch.float32)   # (7, 4)
        return self.b(torch.tanh(self.a(x))).sum(1) + self.c(src).sum() * x.mean(1)


def _free_port():
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        return s.getsockname()[1]


def _worker(rank, world, port, xs, ret, wire='fp32'):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    from hop_b200.dp import DataParallel
    m = Toy()
    if rank == 1:                                            # replicas must be re-synchronised from rank 0
        with torch.no_grad():
            for p in m.parameters():
                p.add_(1.0)
    eng = DataParallel([m], bucket_mb=0.0005, grad_dtype=torch.bfloat16 if wire == 'bf16' else torch.float32)
    out = {}
    for step in range(3):                                    # step 0 = discovery, 1-2 = bucketed/overlapped path
        m.zero_grad(set_to_none=True)
        loss = (m(xs[rank]) ** 2).mean()
        eng.backward(loss)
        out[step] = {k: (p.grad.clone() if p.grad is not None else None) for k, p in m.named_parameters()}
    out['stats'] = dict(eng.stats)
    ret[rank] = out
    dist.destroy_process_group()


@pytest.mark.parametrize('wire', ['fp32', 'bf16'])
def test_dp_two_ranks_match_single_process(wire):
    world = 2
    torch.manual_seed(1)
    xs = [torch.randn(5, 6) for _ in range(world)]
    with mp.Manager() as mgr:
        ret = mgr.dict()
        mp.spawn(_worker, args=(world, _free_port(), xs, ret, wire), nprocs=world, join=True)
        ret = dict(ret)
    # bf16 wire format: every rank's gradient is rounded to bf16 before the sum (2^-8 relative to the tensor's scale);
    # the dSource path stays fp32
    rtol, atol_rel = (1e-5, 0.0) if wire == 'fp32' else (1e-2, 1e-2)
    # single-process reference: mean over ranks of the per-rank mean losses
    m = Toy()
    loss = sum((m(x) ** 2).mean() for x in xs) / world
    loss.backward()
    ref = {k: p.grad for k, p in m.named_parameters()}
    for step in range(3):
        for rank in range(world):
            got = ret[rank][step]
            for k, g in ref.items():
                if g is None:
                    assert got[k] is None, k
                else:
                    assert got[k].dtype == torch.float32
                    assert torch.allclose(got[k], g, rtol=rtol, atol=1e-6 + atol_rel * float(g.abs().max())), (step, rank, k)
    # mapping_layer.weight (12*7 floats) was never put on the wire: only dSource (7*4) and the bucketed rest
    assert ret[0]['stats']['buckets'] >= 3
