#!/usr/bin/env python
"""bench.py -- HOP TED training step (BASELINE.json configs[1]) on N B200s, one JSON line on stdout.

    python bench.py --gpus N --steps K --warmup W            # our arm (kernels through the C ABI)
    python bench.py --impl reference --gpus N --steps K ...  # CPU arm: the oracle port of the reference step

A *step* is one call of the training step ``train_llm`` (reference train_eval/train_llm.py, epoch <= 10
semantics: generator forward + discriminator forward + random-speaker forward + backward + Adam) on a
synthetic batch of 128 TED-shaped samples per GPU; metric = samples/s.

  value    device-timed (CUDA events, max over ranks), inputs already resident in HBM
  e2e      the same step driven from pinned HOST buffers: H2D of the batch + D2H of the losses inside the timed region
  roofline the dominant hand-written kernel group inside the timed steps (CUDA events around its C-ABI call)
  cpu_baseline  the oracle port of the same step on the host cores (bounded sample), rank 0 at N=1 only
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
def cpu_oracle_rate(datasets, batch, steps, warmup, threads=None):
    """samples/s of the oracle port of the step on the host cores."""
    from oracle.hop_torch import OracleTrainer
    from hop_b200.HOP import Model                      # constructor only: gives the reference's 314-key state_dict
    threads = threads or os.cpu_count()
    torch.set_num_threads(threads)
    torch.manual_seed(SEED)
    bert = build_bert()
    m = Model(model_cfg(datasets), bert, _Tok(), _Spk()).float()
    tr = OracleTrainer(m.state_dict(), bert, lr=4e-4 if datasets == 'TED' else 2e-4, datasets=datasets)
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
    batch = a.cpu_batch
    rate, sec, threads = cpu_oracle_rate(a.datasets, batch, max(1, a.steps), min(a.warmup, 1))
    line = {'impl': 'reference', 'metric': 'HOP train samples/s', 'value': rate, 'unit': 'samples/s', 'n_gpus': a.gpus,
            'steps': max(1, a.steps), 'warmup': min(a.warmup, 1), 'ms_per_step': sec * 1e3, 'higher_is_better': True,
            'scaling': 'weak', 'vs_baseline': None, 'dtype': 'fp32', 'data': 'synthetic',
            'config': workload_config(a, 1) | {'cpu_sample_batch': batch},
            'cpu_baseline': {'value': rate, 'unit': 'samples/s', 'cores': threads, 'kind': 'port',
                             'sample': f'{max(1, a.steps)} training step(s) at batch {batch} of the same synthetic workload '
                                       '(oracle/hop_torch.py: functional PyTorch-CPU port of the reference step)'},
            'e2e': {'value': rate, 'unit': 'samples/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}}
    emit(line)


def workload_config(a, world):
    return {'workload': f'HOP {a.datasets} full training step (train_llm, epoch<=10 semantics: 2 generator forwards + '
                        'discriminator forward + backward + Adam), gwnet + reprogramming on hand-written kernels, '
                        'random-init weights, frozen 6-layer BERT',
            'per_gpu_batch': a.batch, 'global_batch': a.batch * world, 'frames': 34, 'seed_frames': 16, 'audio_samples': 36267,
            'parallelism': f'dp{world}',
            'l2_policy': 'working set per step (weights + Adam state + activations, > 1.4 GB) exceeds the 126 MB L2; no flush'}


# ------------------------------------------------------------------------------------------------ our arm
def run_ours(a):
    import torch.distributed as dist
    from hop_b200 import _lib, profiler
    from hop_b200.HOP import Model
    from hop_b200.discriminator import ConvDiscriminator
    from hop_b200.dp import DataParallel
    from hop_b200.train_llm import train_llm

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

    torch.manual_seed(SEED)
    bert = build_bert()
    model = Model(model_cfg(a.datasets), bert, _Tok(), _Spk()).float().to(dev).set_precision(a.precision)
    pose = 27 if a.datasets == 'TED' else 126
    disc = ConvDiscriminator(pose).to(dev)
    lr = 4e-4 if a.datasets == 'TED' else 2e-4                # OneCycleLR start value, never stepped
    # same Adam as the reference (run_ted.py: lr, betas (0.5, 0.999)); fused=True only changes how many kernels apply it
    gen_opt = torch.optim.Adam([p for p in model.parameters() if p.requires_grad], lr=lr, betas=(0.5, 0.999), fused=True,
                               capturable=bool(a.graph))
    dis_opt = torch.optim.Adam(disc.parameters(), lr=lr, betas=(0.5, 0.999), fused=True, capturable=bool(a.graph))
    engine = DataParallel([model, disc])
    sargs = step_args(a.datasets)
    gen = torch.Generator().manual_seed(SEED + rank)
    torch.manual_seed(SEED + rank)
    host = [t.pin_memory() for t in synthetic_batch(a.batch, a.datasets, gen)]
    resident = [t.to(dev) for t in host]
    epoch = 11 if a.gan else 1

    def step(batch):
        return train_llm(sargs, epoch, batch[0], batch[1], batch[2], batch[3], batch[4], model, disc, gen_opt, dis_opt, engine)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    losses = []
    for _ in range(a.warmup):
        losses.append(step(resident)['loss'])
    # ---- per-kernel-group timing: eager steps with CUDA events around every C-ABI call (feeds `roofline`)
    prof_steps = max(1, min(a.steps, 5))
    profiler.enable(True)
    for _ in range(prof_steps):
        losses.append(step(resident)['loss'])
    torch.cuda.synchronize()
    spans = profiler.summary()
    profiler.enable(False)
    # ---- the step as one CUDA graph (hop_b200/graphed.py); falls back to eager launches if capture is refused
    graphed, graph_error = None, None
    run = step
    if a.graph:
        try:
            from hop_b200.graphed import GraphedTrainStep
            graphed = GraphedTrainStep(sargs, epoch, model, disc, gen_opt, dis_opt, engine, resident)
            run = graphed
        except Exception as exc:                                # noqa: BLE001 -- report and measure the eager path instead
            graph_error = f'{type(exc).__name__}: {exc}'[:300]
            sys.stderr.write('CUDA graph capture failed, running eagerly: ' + graph_error + '\n')
            torch.cuda.synchronize()
    ok = torch.tensor([0 if graphed is None else 1], device=dev)
    if world > 1:                                               # all ranks replay or none does
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
    if int(ok) == 0:
        graphed, run = None, step
    for _ in range(2):
        losses.append(run(resident)['loss'])
    # ---- timed region 1: device-resident inputs
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    launches0 = lib.hopk_launch_count()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.steps):
        losses.append(run(resident)['loss'])
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    launches = graphed.launches_per_step * a.steps if graphed is not None else lib.hopk_launch_count() - launches0
    # ---- timed region 2: end to end from pinned host buffers (H2D of the batch, D2H of the loss scalars)
    # Like a DataLoader with pinned memory, the copy of batch i+1 is issued (on a copy stream) before the host blocks
    # on the scalars of step i, so it travels under the step; every batch is still copied and every result still read.
    copy_stream = torch.cuda.Stream()

    def stage():
        with torch.cuda.stream(copy_stream):
            tensors = [t.to(dev, non_blocking=True) for t in host]
            ev = torch.cuda.Event()
            ev.record(copy_stream)
        return tensors, ev

    barrier()
    t0 = time.perf_counter()
    nxt = stage()
    for i in range(a.steps):
        batch, ev = nxt
        torch.cuda.current_stream().wait_event(ev)
        for t in batch:
            t.record_stream(torch.cuda.current_stream())
        if graphed is not None:
            graphed.launch(batch)
            if i + 1 < a.steps:
                nxt = stage()
            out = graphed.result()                              # host floats: one D2H read per step
        else:
            if i + 1 < a.steps:
                nxt = stage()
            out = run(batch)
    barrier()
    e2e_s = time.perf_counter() - t0
    clocks = sampler.stop() if rank == 0 else None
    tms = torch.tensor([ms, e2e_s * 1e3], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(tms, op=dist.ReduceOp.MAX)
    ms, e2e_ms = float(tms[0]), float(tms[1])

    if rank == 0:
        samples = a.batch * world * a.steps
        h2d = sum(t.numel() * t.element_size() for t in host)
        line = {'metric': 'HOP train samples/s', 'value': samples / (ms * 1e-3), 'unit': 'samples/s', 'n_gpus': world,
                'steps': a.steps, 'warmup': a.warmup, 'ms_per_step': ms / a.steps, 'higher_is_better': True,
                'scaling': 'weak', 'vs_baseline': None, 'dtype': 'fp32' if a.precision == 'fp32' else 'bf16+fp32', 'data': 'synthetic',
                'config': workload_config(a, world) | {'tf32_library_gemms': bool(a.tf32), 'gan_phase': bool(a.gan),
                                                       'cuda_graph': graphed is not None, 'graph_error': graph_error,
                                                       'kernel_group_timing': f'{prof_steps} eagerly launched steps before the timed region',
                                                       'precision': a.precision + (' (gwnet + reprogramming on tcgen05 bf16 UMMA kernels, fp32 accumulate; '
                                                                                   'stock BERT/GRU/MLP parts under bf16 autocast)'
                                                                                   if a.precision == 'bf16' else ' (FFMA kernels, reference numerics)')},
                'clocks': clocks,
                'e2e': {'value': samples / (e2e_ms * 1e-3), 'unit': 'samples/s', 'h2d_bytes_per_step': h2d,
                        'd2h_bytes_per_step': 4 * len(out)},
                'gpu_launches': int(launches),
                'loss_first_last': [losses[0], losses[-1]],
                'roofline': roofline(spans, a, world, prof_steps),
                'kernel_ms_per_step': {k: round(v[1] / prof_steps, 4) for k, v in sorted(spans.items())},
                'dp': engine.stats if world > 1 else None}
        if world == 1 and not a.no_cpu_baseline:
            rate, sec, threads = cpu_oracle_rate(a.datasets, a.cpu_batch, 1, 1)
            line['cpu_baseline'] = {'value': rate, 'unit': 'samples/s', 'cores': threads, 'kind': 'port',
                                    'sample': f'1 training step at batch {a.cpu_batch} (after 1 warm-up) of the same '
                                              'synthetic workload, oracle/hop_torch.py on the host cores'}
        emit(line)
    if world > 1:
        torch.cuda.synchronize()
        dist.barrier()
        if graphed is not None:
            # the captured graph holds NCCL kernels: a regular communicator / interpreter teardown was seen to wait forever
            # after the result line had been printed, so leave without running destructors (everything is flushed)
            sys.stdout.flush(); sys.stderr.flush()
            os._exit(0)
        dist.destroy_process_group()


def peaks():
    try:
        with open(os.path.join(ROOT, 'MEASURED_PEAKS.json')) as f:
            p = json.load(f)
        return p['hbm_gbs'], p['bf16_tflops_sustained'], 'measured'
    except (OSError, KeyError, ValueError):
        return 6650.0, 1400.0, 'fallback'


def measured_traffic():
    """dram__bytes_read.sum + dram__bytes_write.sum per call of each kernel group, from the committed ncu capture
    (profiles/r1_traffic.json, written by scripts/traffic_from_ncu.py; TED, B = 128).  None when absent."""
    try:
        with open(os.path.join(ROOT, 'profiles', 'r1_traffic.json')) as f:
            return json.load(f)
    except (OSError, ValueError):
        return {}


def roofline(spans, a, world, nsteps):
    """Roofline position of every hand-written kernel group timed inside the step (CUDA events around the C-ABI calls);
    the top-level entry is the group that takes the most time per step.

    Algorithmic work per call:
      gwnet forward  : fused-floor bytes of SURVEY 8(d) (fp32 activations), backward = 2x        -> HBM roofline
      attention      : forward 4*B*L*S*H*E FLOPs; backward 14*B*L*S*H*E (both passes recompute)  -> tensor roofline
    bf16 precision: tcgen05 UMMA kernels; fp32 precision: FFMA kernels, still reported against the same peaks."""
    B, L, S, H, E = a.batch, 34, 1500, 8, 128
    hbm, tf, which = peaks()
    V, s = (9, 4) if a.datasets == 'TED' else (42, 4)
    floor_fwd = s * B * V * (173 * 16 + 64 * 16 + 64 * (88 + 76) + 2 * 8 * 64 * 4 + 173 * 4)
    work = {'xattn_bwd': ('tensor', 14.0 * B * L * S * H * E), 'xattn_fwd': ('tensor', 4.0 * B * L * S * H * E),
            'gwnet_fwd': ('hbm', float(floor_fwd)), 'gwnet_bwd': ('hbm', 2.0 * floor_fwd)}
    traffic = measured_traffic() if (a.datasets == 'TED' and a.batch == 128) else {}
    groups = {}
    for k, (bound, amount) in work.items():
        if k not in spans:
            continue
        calls, total_ms = spans[k]
        sec = total_ms / calls * 1e-3
        if bound == 'tensor':
            ach, peak, unit = amount / sec / 1e12, tf, 'TFLOP/s'
            extra = {'algorithmic_flops': amount,
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
    ap.add_argument('--cpu-batch', type=int, default=32, help='batch of the bounded CPU sample')
    ap.add_argument('--gan', action='store_true', help='epoch > 10 variant (adds the discriminator step)')
    ap.add_argument('--tf32', type=int, default=0, help='allow TF32 in the stock cuBLAS/cuDNN parts (off = fp32 like the reference)')
    ap.add_argument('--no-cpu-baseline', action='store_true')
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
