#!/usr/bin/env python
"""bench.py -- HOP TED training step (BASELINE.json configs[1]) on N B200s, one JSON line on stdout.

    python bench.py --gpus N --steps K --warmup W            # our arm (kernels through the C ABI)
    python bench.py --impl reference --gpus N --steps K ...  # CPU arm: the oracle port of the reference step

A *step* is one call of the training step ``train_llm`` (reference train_eval/train_llm.py, epoch <= 10
semantics: generator forward + discriminator forward + random-speaker forward + backward + Adam) on a
synthetic batch of 128 TED-shaped samples per GPU; metric = samples/s.

  value         device-timed (CUDA events, max over ranks), inputs already resident in HBM
  e2e           the same step driven from pinned HOST buffers: H2D of the batch + D2H of the losses inside the timed region
  roofline      the dominant hand-written kernel group inside the timed steps (CUDA events around its C-ABI call)
  cpu_baseline  the oracle port of the same step (all of the reference's work, batch 128) on the host cores, bounded sample
  stock_cuda    the same step and the two hot blocks through stock PyTorch CUDA ops (oracle/hop_torch.py on the B200):
                the speed bar SURVEY 8(d) names; N = 1 only
  expressive    BASELINE configs[2] (43 joints / 42 bones, 126-dim pose) through the same path; N = 1 only
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time
import types

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402

SEED = 2021
PER_GPU_BATCH = 128
TRAFFIC_FILE = os.path.join(ROOT, 'profiles', 'r2_traffic.json')


# ------------------------------------------------------------------------------------------------ workload
def synthetic_batch(B, datasets, gen):
    """shapes/ranges of the reference's data loader, drawn from a seeded CPU generator."""
    pose = 27 if datasets == 'TED' else 126
    in_audio = 0.1 * torch.randn(B, 36267, generator=gen)
    melspec = -80.0 * torch.rand(B, 34, 128, generator=gen)
    text = torch.randint(0, 30522, (B, 34), generator=gen) * (torch.rand(B, 34, generator=gen) > 0.7)
    target = torch.clamp(0.3 * torch.randn(B, 34, pose, generator=gen), -1, 1)
    vid = torch.randint(0, 1370, (B,), generator=gen)
    return in_audio, melspec, text.long(), target, vid


def step_args(datasets):
    ted = datasets == 'TED'
    return types.SimpleNamespace(z_type='speaker', loss_regression_weight=600.0 if ted else 2100.0, loss_gan_weight=5.0,
                                 loss_kld_weight=0.6 if ted else 0.8, loss_reg_weight=0.4 if ted else 0.5)


def model_cfg(datasets):
    return types.SimpleNamespace(d_ff=128, llm_dim=768, use_gwnet=True, use_reprograme=True, d_model=128, n_heads=8,
                                 datasets=datasets)


class _Tok:
    eos_token = None
    pad_token = None

    def add_special_tokens(self, d):
        return None


class _Spk:
    n_words = 1370


def build_bert():
    from transformers import BertConfig, BertModel
    return BertModel(BertConfig(num_hidden_layers=6)).eval()


# ------------------------------------------------------------------------------------------------ clocks
class ClockSampler:
    FIELDS = 'clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,' \
             'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap'

    def __init__(self, index=0):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(['nvidia-smi', '-i', str(self.index), f'--query-gpu={self.FIELDS}',
                                          '--format=csv,noheader,nounits', '-lms', '100'], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(',')])

    def stop(self):
        if self.proc is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        self.proc.terminate()
        rows = [r for r in self.rows if len(r) >= 6 and r[0].isdigit()]
        if not rows:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['no samples']}
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        reasons = [n for i, n in enumerate(names) if any(r[2 + i].lower().startswith('active') for r in rows)]
        return {'sm_mhz': float(np.median([int(r[0]) for r in rows])), 'sm_max_mhz': float(rows[0][1]),
                'reasons': reasons, 'samples': len(rows)}


# ------------------------------------------------------------------------------------------------ CPU arm
CPU_SAMPLE_NOTE = ('training step(s) at batch {b} of the same synthetic workload: oracle/hop_torch.py::OracleTrainer(literal=True), a '
                   'functional PyTorch-CPU port of the reference step doing all of the reference\'s work -- generator forward with '
                   'the J-fold beat MLP (HOP.py:210) and the mapping GEMM per forward, discriminator forward (train_llm.py:43-44), '
                   'random-speaker forward with autograd recording, backward, Adam')


_ORACLE_INIT = {}


def oracle_trainer(datasets, device='cpu', capturable=False):
    """The port of the reference step on ``device`` (CPU baseline; on the GPU the stock-PyTorch-CUDA speed bar)."""
    import copy
    from oracle.hop_torch import OracleTrainer
    from hop_b200.HOP import Model                      # constructor only: gives the reference's 314-key state_dict
    from hop_b200.discriminator import ConvDiscriminator
    if datasets not in _ORACLE_INIT:                    # random-init weights of the reference architecture, built once on the CPU
        torch.manual_seed(SEED)
        bert = build_bert()
        m = Model(model_cfg(datasets), bert, _Tok(), _Spk()).float()
        _ORACLE_INIT[datasets] = ({k: v.detach().clone() for k, v in m.state_dict().items()}, bert,
                                  ConvDiscriminator(27 if datasets == 'TED' else 126))
    sd, bert, disc = _ORACLE_INIT[datasets]
    return OracleTrainer(sd, copy.deepcopy(bert), lr=4e-4 if datasets == 'TED' else 2e-4, datasets=datasets, device=device,
                         literal=True, discriminator=copy.deepcopy(disc), capturable=capturable)


def cpu_oracle_rate(datasets, batch, steps, warmup, threads=None):
    """samples/s of the oracle port of the step on the host cores."""
    threads = threads or os.cpu_count()
    torch.set_num_threads(threads)
    tr = oracle_trainer(datasets)
    gen = torch.Generator().manual_seed(SEED)
    batch_t = synthetic_batch(batch, datasets, gen)
    for _ in range(warmup):
        tr.step(*batch_t)
    t0 = time.perf_counter()
    for _ in range(steps):
        tr.step(*batch_t)
    dt = time.perf_counter() - t0
    return batch * steps / dt, dt / steps, torch.get_num_threads()


def run_reference(a):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    steps, warmup = max(1, a.steps), max(0, a.warmup)
    rate, sec, threads = cpu_oracle_rate(a.datasets, a.batch, steps, warmup)
    line = {'impl': 'reference', 'metric': 'HOP train samples/s', 'value': rate, 'unit': 'samples/s', 'n_gpus': a.gpus,
            'steps': steps, 'warmup': warmup, 'ms_per_step': sec * 1e3, 'higher_is_better': True,
            'scaling': 'weak', 'vs_baseline': None, 'dtype': 'fp32', 'data': 'synthetic',
            'config': workload_config(a, 1, a.datasets),
            'cpu_baseline': {'value': rate, 'unit': 'samples/s', 'cores': threads, 'kind': 'port',
                             'sample': f'{steps} ' + CPU_SAMPLE_NOTE.format(b=a.batch) + f' (after {warmup} warm-up)'},
            'e2e': {'value': rate, 'unit': 'samples/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}}
    emit(line)


def workload_config(a, world, datasets):
    return {'workload': f'HOP {datasets} full training step (train_llm, epoch<=10 semantics: 2 generator forwards + '
                        'discriminator forward + backward + Adam), random-init weights, frozen 6-layer BERT',
            'per_gpu_batch': a.batch, 'global_batch': a.batch * world, 'frames': 34, 'seed_frames': 16, 'audio_samples': 36267,
            'parallelism': f'dp{world}',
            'l2_policy': 'working set per step (weights + Adam state + activations, > 1.4 GB) exceeds the 126 MB L2; no flush'}


# ------------------------------------------------------------------------------------------------ our arm
class Workload:
    """Model + optimisers + data of one configuration (TED or TED_expressive) on this rank's GPU."""

    def __init__(self, a, datasets, dev, rank, engine_factory):
        from hop_b200.HOP import Model
        from hop_b200.discriminator import ConvDiscriminator
        from hop_b200.train_llm import train_llm
        self.a, self.datasets, self.dev = a, datasets, dev
        torch.manual_seed(SEED)
        bert = build_bert()
        self.model = Model(model_cfg(datasets), bert, _Tok(), _Spk()).float().to(dev).set_precision(a.precision)
        pose = 27 if datasets == 'TED' else 126
        self.disc = ConvDiscriminator(pose).to(dev)
        lr = 4e-4 if datasets == 'TED' else 2e-4              # OneCycleLR start value, never stepped
        # same Adam as the reference (run_ted.py: lr, betas (0.5, 0.999)); fused=True only changes how many kernels apply it
        self.gen_opt = torch.optim.Adam([p for p in self.model.parameters() if p.requires_grad], lr=lr, betas=(0.5, 0.999),
                                        fused=True, capturable=bool(a.graph))
        self.dis_opt = torch.optim.Adam(self.disc.parameters(), lr=lr, betas=(0.5, 0.999), fused=True, capturable=bool(a.graph))
        self.engine = engine_factory([self.model, self.disc])
        self.sargs = step_args(datasets)
        gen = torch.Generator().manual_seed(SEED + rank)
        torch.manual_seed(SEED + rank)
        self.host = [t.pin_memory() for t in synthetic_batch(a.batch, datasets, gen)]
        self.resident = [t.to(dev) for t in self.host]
        self.epoch = 11 if a.gan else 1
        self._train_llm = train_llm
        self.graphed, self.graph_error = None, None

    def step(self, batch):
        return self._train_llm(self.sargs, self.epoch, batch[0], batch[1], batch[2], batch[3], batch[4], self.model, self.disc,
                               self.gen_opt, self.dis_opt, self.engine)

    def capture(self):
        from hop_b200.graphed import GraphedTrainStep
        try:
            self.graphed = GraphedTrainStep(self.sargs, self.epoch, self.model, self.disc, self.gen_opt, self.dis_opt,
                                            self.engine, self.resident)
        except Exception as exc:                                # noqa: BLE001 -- report and measure the eager path instead
            self.graph_error = f'{type(exc).__name__}: {exc}'[:300]
            sys.stderr.write('CUDA graph capture failed, running eagerly: ' + self.graph_error + '\n')
            torch.cuda.synchronize()

    def release(self):
        if self.graphed is not None:
            self.graphed.graph.reset()
            self.graphed = None


def run_ours(a):
    import torch.distributed as dist
    from hop_b200 import _lib, profiler
    from hop_b200.dp import DataParallel

    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    assert torch.cuda.is_available(), 'bench.py (our arm) needs a CUDA device; there is no CPU fallback'
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    if world > 1:
        dist.init_process_group('nccl', device_id=dev)
    lib = _lib.lib()
    torch.backends.cuda.matmul.allow_tf32 = bool(a.tf32)
    torch.backends.cudnn.allow_tf32 = bool(a.tf32)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def allmax(vals):
        t = torch.tensor(vals, device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return [float(v) for v in t]

    def measure(W, steps, warmup, with_e2e, sampler=None):
        """Warm up, time the per-kernel groups on eager steps, capture the step as a CUDA graph, time `steps` replays."""
        losses = []
        for _ in range(warmup):
            losses.append(W.step(W.resident)['loss'])
        prof_steps = max(1, min(steps, 5))
        # Eagerly launched, the step is host-bound (about 2 us of GPU idle time inside every event pair).  Each profiled step
        # therefore starts behind a ~30 ms device-side spin: the host enqueues the step while the GPU waits, and the event
        # pairs then bracket back-to-back GPU execution only.
        spin = int(0.030 * 1.9e9)
        profiler.enable(True)
        for _ in range(prof_steps):
            torch.cuda._sleep(spin)
            losses.append(W.step(W.resident)['loss'])
        torch.cuda.synchronize()
        spans = profiler.summary()
        profiler.enable(True, fine=True)                            # second pass: every dense GEMM launch on its own (nested spans)
        for _ in range(prof_steps):
            torch.cuda._sleep(spin)
            losses.append(W.step(W.resident)['loss'])
        torch.cuda.synchronize()
        fine = profiler.summary()
        spans_work = profiler.work_summary()
        if 'gemm_tma' in fine:
            spans['gemm_tma'] = fine['gemm_tma']
        profiler.enable(False)
        run = W.step
        if a.graph:
            W.capture()
        ok = torch.tensor([0 if W.graphed is None else 1], device=dev)
        if world > 1:                                           # all ranks replay or none does
            dist.all_reduce(ok, op=dist.ReduceOp.MIN)
        if int(ok) == 0:
            W.release()
        else:
            run = W.graphed
        for _ in range(2):
            losses.append(run(W.resident)['loss'])
        # ---- timed region 1: device-resident inputs
        if sampler is not None:
            sampler.start()
        launches0 = lib.hopk_launch_count()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        if W.graphed is not None:                               # host one step ahead: scalars of step i read while step i+1 runs
            pend = None
            for _ in range(steps):
                tk = W.graphed.launch_async(W.resident)
                if pend is not None:
                    losses.append(W.graphed.collect(pend)['loss'])
                pend = tk
            e1.record()
            losses.append(W.graphed.collect(pend)['loss'])
        else:
            for _ in range(steps):
                losses.append(run(W.resident)['loss'])
            e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        launches = W.graphed.launches_per_step * steps if W.graphed is not None else lib.hopk_launch_count() - launches0
        res = {'ms': ms, 'launches': int(launches), 'spans': spans, 'spans_work': spans_work, 'prof_steps': prof_steps, 'losses': losses, 'out_len': 0,
               'e2e_ms': None}
        if not with_e2e:
            return res
        # ---- timed region 2: end to end from pinned host buffers (H2D of the batch, D2H of the loss scalars)
        # Like a DataLoader with pinned memory, the copy of batch i+1 is issued (on a copy stream) before the host blocks
        # on the scalars of step i, so it travels under the step; every batch is still copied and every result still read.
        copy_stream = torch.cuda.Stream()

        def stage():
            with torch.cuda.stream(copy_stream):
                tensors = [t.to(dev, non_blocking=True) for t in W.host]
                ev = torch.cuda.Event()
                ev.record(copy_stream)
            return tensors, ev

        def e2e_loop(n):
            nxt = stage()
            out = {}
            pend = None
            for i in range(n):
                batch, ev = nxt
                torch.cuda.current_stream().wait_event(ev)
                for t in batch:
                    t.record_stream(torch.cuda.current_stream())
                if W.graphed is not None:
                    tk = W.graphed.launch_async(batch)              # H2D-fed replay + async D2H of its scalars into a pinned slot
                    if i + 1 < n:
                        nxt = stage()
                    if pend is not None:
                        out = W.graphed.collect(pend)               # host floats of the step before: one D2H read per step
                    pend = tk
                else:
                    if i + 1 < n:
                        nxt = stage()
                    out = run(batch)
            if pend is not None:
                out = W.graphed.collect(pend)
            return out

        e2e_loop(max(3, warmup))                                    # untimed: copy stream, staging buffers, allocator pools
        torch.cuda.synchronize()
        barrier()
        t0 = time.perf_counter()
        out = e2e_loop(steps)
        torch.cuda.synchronize()
        barrier()
        res['e2e_ms'] = (time.perf_counter() - t0) * 1e3
        res['out_len'] = len(out)
        return res

    # dtype-1 runs put bf16 gradients on the wire (half the all-reduce bytes); the exact mode keeps fp32
    engine_factory = lambda mods: DataParallel(mods, grad_dtype=torch.bfloat16 if a.precision == 'bf16' else torch.float32,
                                               overlap=bool(a.dp_overlap), bucket_mb=a.dp_bucket_mb)
    W = Workload(a, a.datasets, dev, rank, engine_factory)
    if a.profile_step:
        # one eagerly launched step between cudaProfilerStart / Stop (ncu --profile-from-start off): the launch list of the
        # whole step, stock parts included (profiles/*launches*); prints no bench line
        for _ in range(max(3, a.warmup)):
            W.step(W.resident)
        torch.cuda.synchronize()
        torch.cuda.profiler.start()
        W.step(W.resident)
        torch.cuda.synchronize()
        torch.cuda.profiler.stop()
        return
    sampler = ClockSampler(local) if rank == 0 else None
    r = measure(W, a.steps, a.warmup, True, sampler)
    clocks = sampler.stop() if rank == 0 else None
    ms, e2e_ms = allmax([r['ms'], r['e2e_ms']])

    line = None
    if rank == 0:
        samples = a.batch * world * a.steps
        h2d = sum(t.numel() * t.element_size() for t in W.host)
        line = {'metric': 'HOP train samples/s', 'value': samples / (ms * 1e-3), 'unit': 'samples/s', 'n_gpus': world,
                'steps': a.steps, 'warmup': a.warmup, 'ms_per_step': ms / a.steps, 'higher_is_better': True,
                'scaling': 'weak', 'vs_baseline': None, 'dtype': 'fp32' if a.precision == 'fp32' else 'bf16+fp32', 'data': 'synthetic',
                'config': workload_config(a, world, a.datasets) | {
                    'tf32_library_gemms': bool(a.tf32), 'gan_phase': bool(a.gan), 'cuda_graph': W.graphed is not None,
                    'graph_error': W.graph_error,
                    'kernel_group_timing': f'{r["prof_steps"]} eagerly launched steps before the timed region, each queued behind a 30 ms device-side spin (host runs ahead: event pairs bracket GPU execution only)',
                    'precision': a.precision + (' (hand-written tcgen05 bf16 UMMA kernels, fp32 accumulate; remaining stock parts under bf16 autocast)'
                                                if a.precision == 'bf16' else ' (FFMA kernels, reference numerics)')},
                'clocks': clocks,
                'e2e': {'value': samples / (e2e_ms * 1e-3), 'unit': 'samples/s', 'h2d_bytes_per_step': h2d,
                        'd2h_bytes_per_step': 4 * r['out_len']},
                'gpu_launches': r['launches'],
                'loss_first_last': [r['losses'][0], r['losses'][-1]],
                'roofline': roofline(r['spans'], a, a.datasets, r['prof_steps'], r['spans_work']),
                'kernel_ms_per_step': {k: round(v[1] / r['prof_steps'], 4) for k, v in sorted(r['spans'].items())},   # gemm_tma is nested inside bert_* / linear_* / beat_* / mapping_*
                'dp': (W.engine.stats | {'wire_dtype': str(W.engine.grad_dtype).replace('torch.', '')}) if world > 1 else None}
    # ---- N = 1 extras: Expressive configuration, stock-PyTorch-CUDA speed bar, CPU baseline
    if world == 1:
        W.release()
        del W
        torch.cuda.empty_cache()
        if not a.no_expressive and a.datasets == 'TED':
            try:
                ksteps = max(3, min(a.steps, 10))
                WE = Workload(a, 'TED_expressive', dev, rank, engine_factory)
                re_ = measure(WE, ksteps, 3, True)
                line['expressive'] = {
                    'config': workload_config(a, 1, 'TED_expressive'), 'steps': ksteps, 'ms_per_step': re_['ms'] / ksteps,
                    'value': a.batch * ksteps / (re_['ms'] * 1e-3), 'unit': 'samples/s',
                    'e2e': {'value': a.batch * ksteps / (re_['e2e_ms'] * 1e-3), 'unit': 'samples/s'},
                    'cuda_graph': WE.graphed is not None, 'gpu_launches': re_['launches'],
                    'kernel_ms_per_step': {k: round(v[1] / re_['prof_steps'], 4) for k, v in sorted(re_['spans'].items())},
                    'roofline': roofline(re_['spans'], a, 'TED_expressive', re_['prof_steps'], re_['spans_work'])}
                WE.release()
                del WE
            except Exception as exc:                            # noqa: BLE001
                line['expressive'] = {'error': f'{type(exc).__name__}: {exc}'[:300]}
            torch.cuda.empty_cache()
        if not a.no_stock_cuda:
            try:
                line['stock_cuda'] = stock_cuda(a, dev)
            except Exception as exc:                            # noqa: BLE001
                line['stock_cuda'] = {'error': f'{type(exc).__name__}: {exc}'[:300]}
            torch.cuda.empty_cache()
        if not a.no_cpu_baseline:
            rate, sec, threads = cpu_oracle_rate(a.datasets, a.batch, a.cpu_steps, 1)
            line['cpu_baseline'] = {'value': rate, 'unit': 'samples/s', 'cores': threads, 'kind': 'port',
                                    'sample': f'{a.cpu_steps} ' + CPU_SAMPLE_NOTE.format(b=a.batch) + ' (after 1 warm-up), on the host cores'}
    if rank == 0:
        emit(line)
    if world > 1:
        # release the captured graph (it holds NCCL kernels) before the communicator; a watchdog guarantees the process
        # still ends if the NCCL teardown blocks (seen in round 1 with a live graph) -- the result line is already out
        torch.cuda.synchronize()
        dist.barrier()
        sys.stdout.flush(); sys.stderr.flush()
        t = threading.Timer(30.0, lambda: (sys.stderr.write('bench: NCCL teardown still blocked after 30 s, leaving\n'), os._exit(0)))
        t.daemon = True
        t.start()
        W.release()
        torch.cuda.synchronize()
        dist.destroy_process_group()
        t.cancel()


# ------------------------------------------------------------------------------------------------ stock PyTorch CUDA bar
def _time_cuda(fn, iters, warmup=3, graph=False):
    """ms per call of fn() (CUDA events).  graph=True: capture one call and time replays."""
    if graph:
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(warmup):
                fn()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            fn()
        call = g.replay
    else:
        call = fn
        for _ in range(warmup):
            fn()
    call()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        call()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def stock_cuda(a, dev):
    """The speed bar of SURVEY 8(d): the same training step, and the two hot blocks alone, through stock PyTorch CUDA ops
    (cuDNN / cuBLAS / ATen via oracle/hop_torch.py) on this B200 -- fp32 with TF32 off (the reference's numerics) and bf16
    autocast, launched eagerly and replayed as a CUDA graph.  ms per call; the oracle is the thing *measured against*."""
    from oracle import hop_torch
    out = {'note': 'stock PyTorch CUDA ops (cuDNN/cuBLAS/ATen through oracle/hop_torch.py) on the same GPU; ms per call; '
                   'fp32 = TF32 off (reference numerics), bf16 = torch.autocast(bfloat16)', 'batch': a.batch}
    tf32 = (torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32)
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    gen = torch.Generator().manual_seed(SEED)
    batch = [t.to(dev) for t in synthetic_batch(a.batch, a.datasets, gen)]
    iters = max(3, min(a.steps, 10))

    def variants(make_fn):
        res = {}
        for prec in ('fp32', 'bf16'):
            for mode in ('eager', 'graph'):
                key = f'{prec}_{mode}_ms'
                try:
                    fn = make_fn(prec, mode == 'graph')
                    res[key] = round(_time_cuda(fn, iters, graph=(mode == 'graph')), 4)
                except Exception as exc:                        # noqa: BLE001
                    res[key] = None
                    res[key + '_error'] = f'{type(exc).__name__}: {exc}'[:200]
                    torch.cuda.synchronize()
        return res

    # ---- whole training step
    def make_step(prec, graph):
        tr = oracle_trainer(a.datasets, device=dev, capturable=graph)

        def fn():
            with torch.autocast('cuda', dtype=torch.bfloat16, enabled=prec == 'bf16'):
                return tr.step_device(*batch)
        return fn
    out['step'] = variants(make_step)
    for k in list(out['step']):
        if k.endswith('_ms') and out['step'][k]:
            out['step'][k.replace('_ms', '_samples_per_s')] = round(a.batch / out['step'][k] * 1e3, 1)
    torch.cuda.empty_cache()

    # ---- the two hot blocks alone
    tr = oracle_trainer(a.datasets, device=dev)
    sd = tr.sd
    V = 9 if a.datasets == 'TED' else 42
    x = torch.randn(a.batch, 173, V, 16, device=dev)
    dout = torch.randn(a.batch, 173, V, 4, device=dev)
    gparams = [t for k, t in sd.items() if k.startswith('gwnet.') and t.requires_grad]

    def make_gw(bwd):
        def make(prec, graph):
            xx = x.clone().requires_grad_(True)

            def fn():
                with torch.autocast('cuda', dtype=torch.bfloat16, enabled=prec == 'bf16'):
                    y = hop_torch.gwnet_forward(sd, xx, training=True)
                if bwd:
                    torch.autograd.grad(y.float(), [xx] + gparams, dout, allow_unused=True)
            if not bwd:
                return lambda: torch.no_grad()(fn)()
            return fn
        return make
    out['gwnet_fwd'] = variants(make_gw(False))
    out['gwnet_fwd_bwd'] = variants(make_gw(True))

    mel = batch[1]
    src = torch.randn(1500, 768, device=dev, requires_grad=True)
    dy = torch.randn(a.batch, 34, 768, device=dev)
    rparams = [t for k, t in sd.items() if k.startswith('reprogramming_layer.')]

    def make_rp(bwd):
        def make(prec, graph):
            xx = mel.clone().requires_grad_(True)

            def fn():
                with torch.autocast('cuda', dtype=torch.bfloat16, enabled=prec == 'bf16'):
                    y = hop_torch.reprogramming_forward(sd, xx, src, src, 8, p_drop=0.1)
                if bwd:
                    torch.autograd.grad(y.float(), [xx, src] + rparams, dy)
            if not bwd:
                return lambda: torch.no_grad()(fn)()
            return fn
        return make
    out['reprogramming_fwd'] = variants(make_rp(False))
    out['reprogramming_fwd_bwd'] = variants(make_rp(True))
    torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32 = tf32
    return out


# ------------------------------------------------------------------------------------------------ roofline
def peaks():
    try:
        with open(os.path.join(ROOT, 'MEASURED_PEAKS.json')) as f:
            p = json.load(f)
        return p['hbm_gbs'], p['bf16_tflops_sustained'], 'measured'
    except (OSError, KeyError, ValueError):
        return 6650.0, 1400.0, 'fallback'


def measured_traffic():
    """dram__bytes_read.sum + dram__bytes_write.sum per call of each kernel group, from the committed ncu capture
    (profiles/r2_traffic.json, written by scripts/traffic_from_ncu.py together with the commit it was taken at; TED,
    B = 128).  Not measured in this run (ncu cannot run inside the bench); `traffic_source` says where it is from."""
    try:
        with open(TRAFFIC_FILE) as f:
            return json.load(f)
    except (OSError, ValueError):
        return {}


def roofline(spans, a, datasets, nsteps, spans_work=None):
    """Roofline position of every hand-written kernel group timed inside the step (CUDA events around the C-ABI calls);
    the top-level entry is the group that takes the most time per step -- the dense TMA + tcgen05 GEMM (`gemm_tma`: every
    hopk_gemm_bf16 launch of the encoder / projections / beat MLP / mapping layer, 2*M*N*K FLOPs each, timed per launch).

    Algorithmic work per call:
      gwnet forward  : fused-floor bytes of SURVEY 8(d) with s = 4 (fp32 activations), backward = 2x    -> HBM roofline
      attention      : forward 4*B*L*S*H*E FLOPs; backward 2.5x forward = 10*B*L*S*H*E (the recompute passes execute
                       14*B*L*S*H*E, reported as executed_flops)                                        -> tensor roofline
    bf16 precision: tcgen05 UMMA kernels; fp32 precision: FFMA kernels, still reported against the same peaks."""
    B, L, S, H, E = a.batch, 34, 1500, 8, 128
    hbm, tf, which = peaks()
    V, s = (9, 4) if datasets == 'TED' else (42, 4)
    floor_fwd = s * B * V * (173 * 16 + 64 * 16 + 64 * (88 + 76) + 2 * 8 * 64 * 4 + 173 * 4)
    blshe = float(B) * L * S * H * E
    work = {'xattn_bwd': ('tensor', 10.0 * blshe, 14.0 * blshe), 'xattn_fwd': ('tensor', 4.0 * blshe, 4.0 * blshe),
            'gwnet_fwd': ('hbm', float(floor_fwd), None), 'gwnet_bwd': ('hbm', 2.0 * floor_fwd, None)}
    traffic = measured_traffic() if (datasets == 'TED' and a.batch == 128) else {}
    groups = {}
    if spans_work and spans.get('gemm_tma') and a.precision == 'bf16':
        calls, total_ms = spans['gemm_tma']
        flops = spans_work['gemm_tma']
        ach = flops / (total_ms * 1e-3) / 1e12
        groups['gemm_tma'] = {'bound': 'tensor', 'achieved': ach, 'peak': tf, 'unit': 'TFLOP/s', 'frac': ach / tf,
                              'traffic': traffic.get('gemm_tma'), 'avg_ms': total_ms / calls, 'launches_timed': calls,
                              'ms_per_step': total_ms / nsteps, 'algorithmic_flops': flops / calls,
                              'arithmetic': 'bf16 tcgen05 UMMA (TMA operands, fp32 accumulate in TMEM); FLOPs = 2*M*N*K summed over the '
                                            'launches, time = sum of their CUDA-event durations'}
    for k, (bound, amount, executed) in work.items():
        if k not in spans:
            continue
        calls, total_ms = spans[k]
        sec = total_ms / calls * 1e-3
        if bound == 'tensor':
            ach, peak, unit = amount / sec / 1e12, tf, 'TFLOP/s'
            extra = {'algorithmic_flops': amount, 'executed_flops': executed,
                     'arithmetic': 'bf16 tcgen05 UMMA, fp32 accumulate in TMEM' if a.precision == 'bf16' else 'fp32 FFMA'}
        else:
            ach, peak, unit = amount / sec / 1e9, hbm, 'GB/s'
            extra = {'algorithmic_bytes': amount}
        groups[k] = {'bound': bound, 'achieved': ach, 'peak': peak, 'unit': unit, 'frac': ach / peak,
                     'traffic': traffic.get(k), 'avg_ms': total_ms / calls, 'launches_timed': calls,
                     'ms_per_step': total_ms / nsteps} | extra
    if not groups:
        return None
    top = max(groups, key=lambda k: groups[k]['ms_per_step'])
    out = {'kernel': top} | groups[top]
    out['peak_source'] = which + (' (hbm_gbs)' if groups[top]['bound'] == 'hbm' else ' (bf16_tflops_sustained)')
    out['traffic_source'] = traffic.get('source', None)
    out['groups'] = groups
    return out


_REAL_STDOUT = None


def _guard_stdout():
    """Library chatter (e.g. the 'NCCL version' banner) must not pollute the one-JSON-line contract: everything that
    writes to fd 1 during the run goes to stderr; emit() writes the JSON line to the real stdout at the end."""
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.fdopen(os.dup(1), 'w')
    os.dup2(2, 1)


def emit(line):
    _REAL_STDOUT.write(json.dumps(line) + '\n')
    _REAL_STDOUT.flush()


def main():
    _guard_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=10)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--datasets', default='TED', choices=['TED', 'TED_expressive'])
    ap.add_argument('--batch', type=int, default=PER_GPU_BATCH)
    ap.add_argument('--cpu-steps', type=int, default=2, help='timed steps of the bounded CPU sample inside our line')
    ap.add_argument('--gan', action='store_true', help='epoch > 10 variant (adds the discriminator step)')
    ap.add_argument('--tf32', type=int, default=0, help='allow TF32 in the stock cuBLAS/cuDNN parts (off = fp32 like the reference)')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-stock-cuda', action='store_true')
    ap.add_argument('--no-expressive', action='store_true')
    ap.add_argument('--dp-overlap', type=int, default=1, help='0: reduce the gradient buckets after backward instead of under it')
    ap.add_argument('--dp-bucket-mb', type=float, default=25)
    ap.add_argument('--profile-step', action='store_true', help='profile one eager step (for ncu --profile-from-start off) and exit')
    ap.add_argument('--graph', type=int, default=1, help='replay the whole training step as one CUDA graph (0 = launch eagerly)')
    ap.add_argument('--precision', default='bf16', choices=['bf16', 'fp32'],
                    help='bf16 (BASELINE configs[1]): stock cuBLAS/cuDNN parts under bf16 autocast; fp32: reference numerics')
    a = ap.parse_args()
    if a.impl == 'reference':
        run_reference(a)
    else:
        if a.warmup < 3:
            a.warmup = 3
        run_ours(a)


if __name__ == '__main__':
    main()
