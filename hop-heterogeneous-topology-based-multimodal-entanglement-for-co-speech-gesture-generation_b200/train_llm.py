"""Mirror of the reference's training step ``train_eval/train_llm.py::train_llm`` (lines 9-98).

Same signature, same losses, same optimiser calls, same returned dict -- so the epoch loops of
run_ted.py:399-402 / run_expressive.py:446-449 can call it unchanged.  Differences that do not
change the arithmetic:
  * the batch-independent text prototypes (``mapping_layer(word_embeddings)``, 70 GFLOP) are computed
    once per step and shared by the 2-3 generator forwards (the reference recomputes them);
  * the random-speaker pass, whose outputs the reference only uses ``.detach()``-ed, runs under
    ``torch.no_grad()`` (its BatchNorm running-stat updates and dropout draws still happen);
  * the beat features and the Graph-WaveNet output -- a function of the audio, the seed poses and the weights only, with
    no dropout on that branch -- are computed by the first forward of the step and reused by the others (the reference
    recomputes identical values); the BatchNorm running statistics still receive one update per forward;
  * the random-speaker pass is issued on a second CUDA stream beside the discriminator forward and the regression terms
    (it depends on neither; its decoder recurrence leaves SMs idle); joined before the diversity term;
  * the 2-5 ``.item()`` host syncs are folded into one device->host read at the end.
``accelerator`` only needs ``.backward(loss)``: hop_b200.dp.DataParallel (single- and multi-GPU) or, on ONE GPU with the
bare module, HF Accelerate.  A DDP / DeepSpeed-wrapped model is refused (see ``_check_wrapper``).
"""
import contextlib

import torch
import torch.nn.functional as F


SHARE_STEP_FEATURES = True
OVERLAP_RANDOM_SPEAKER_PASS = True
FUSED_LOSSES = True
_SIDE = {}


def _side_stream(device):
    key = torch.device(device).index
    if key not in _SIDE:
        _SIDE[key] = torch.cuda.Stream(device=device)
    return _SIDE[key]


def add_noise(data):
    return data + torch.randn_like(data) * 0.1


def _forward(model, in_audio, log_melspec, text, pre_seq, vids, source, shared=None):
    if source is not None:
        return model.forecast(in_audio, log_melspec, text, pre_seq, vids, source=source, shared=shared)
    return model(in_audio, log_melspec, text, pre_seq, vids)


def train_llm(args, epoch, in_audio, log_melspec, text_token_padded, target_dir_vec, vid_indices,
              model, discriminator, model_optim, dis_optimizer, accelerator):
    names, dev = train_llm_device(args, epoch, in_audio, log_melspec, text_token_padded, target_dir_vec, vid_indices,
                                  model, discriminator, model_optim, dis_optimizer, accelerator)
    return finish_losses(names, dev.tolist())                   # one device->host read for all reported scalars


def finish_losses(names, host):
    ret = dict(zip(names, host))
    for k in ('KLD', 'DIV_REG'):                                # the reference drops falsy entries (`if kld:`)
        if k in ret and not ret[k]:
            del ret[k]
    return ret


def _check_wrapper(model, accelerator):
    """The step calls ``model.module.forecast`` directly (shared text prototypes / features), which bypasses the forward
    of a wrapping DistributedDataParallel / DeepSpeedEngine: their reducers would never be armed and the replicas
    would silently diverge.  Gradient averaging is hop_b200.dp.DataParallel's job here -- refuse anything else."""
    kind = type(model).__name__
    if kind in ('DistributedDataParallel', 'DeepSpeedEngine', 'FullyShardedDataParallel'):
        from .dp import DataParallel
        if not isinstance(accelerator, DataParallel):
            raise RuntimeError(f'hop_b200.train_llm: the model is wrapped in {kind}, whose gradient reduction this step '
                               'would bypass; pass the bare module and use hop_b200.dp.DataParallel as `accelerator`')


def train_llm_device(args, epoch, in_audio, log_melspec, text_token_padded, target_dir_vec, vid_indices,
                     model, discriminator, model_optim, dis_optimizer, accelerator):
    """The whole step without its single host read: returns (names, stacked device tensor of the reported scalars).
    Free of host synchronisation, so it can be captured in a CUDA graph (hop_b200.graphed.GraphedTrainStep)."""
    pre_seq = target_dir_vec[:, 0:16]
    dis_error = None
    _check_wrapper(model, accelerator)
    core = getattr(model, 'module', model)
    source = core.source_embeddings() if hasattr(core, 'source_embeddings') else None
    # beat features + Graph-WaveNet output are identical in every forward of a step (SHARE_STEP_FEATURES = False recomputes)
    shared = {} if (source is not None and SHARE_STEP_FEATURES) else None
    gan = epoch > 10 and args.loss_gan_weight > 0.0

    if gan:                                                     # discriminator step (train_llm.py:15-36)
        dis_optimizer.zero_grad()
        with torch.no_grad():
            outputs, *_ = _forward(core, in_audio, log_melspec, text_token_padded, pre_seq, vid_indices,
                                   None if source is None else source.detach(), shared)
        dis_real = discriminator(add_noise(target_dir_vec), text_token_padded)
        dis_fake = discriminator(add_noise(outputs.detach()), text_token_padded)
        dis_error = torch.sum(-torch.mean(torch.log(dis_real + 1e-8) + torch.log(1 - dis_fake + 1e-8)))
        accelerator.backward(dis_error)
        dis_optimizer.step()

    model_optim.zero_grad()                                     # generator step (train_llm.py:38-86)
    outputs, z_context, z_mu, z_logvar = _forward(core, in_audio, log_melspec, text_token_padded, pre_seq,
                                                   vid_indices, source, shared)
    want_rand = (args.z_type == 'speaker' or args.z_type == 'random') and args.loss_reg_weight > 0.0
    side = None
    if want_rand:
        if args.z_type == 'speaker':
            rand_idx = torch.randperm(vid_indices.shape[0], device=vid_indices.device)
            rand_vids = vid_indices[rand_idx]
        else:
            rand_vids = None
        # The random-speaker pass (no autograd graph) does not depend on the discriminator / regression terms below and its
        # decoder recurrence leaves SMs idle: on CUDA it runs on a second stream beside them (forked and joined with events,
        # so a CUDA-graph capture of the step records the same dependency structure).
        side = _side_stream(outputs.device) if (OVERLAP_RANDOM_SPEAKER_PASS and outputs.is_cuda) else None
        cur = torch.cuda.current_stream() if side is not None else None
        if side is not None:
            side.wait_stream(cur)
        with torch.cuda.stream(side) if side is not None else contextlib.nullcontext():
            with torch.no_grad():
                out_dir_vec_rand_vid, z_rand_vid, _, _ = _forward(core, in_audio, log_melspec, text_token_padded, pre_seq,
                                                                  rand_vids, None if source is None else source.detach(), shared)
    dis_output = discriminator(outputs, text_token_padded)      # computed every step, like train_llm.py:43-44
    gen_error = -torch.mean(torch.log(dis_output + 1e-8))
    kld = div_reg = None
    if want_rand and side is not None:
        cur.wait_stream(side)
        for t in (out_dir_vec_rand_vid, z_rand_vid):
            t.record_stream(cur)
    if FUSED_LOSSES and outputs.is_cuda:
        # one kernel forward + one backward for the regression / diversity / KLD terms (hop_b200/losses.py)
        from .losses import step_losses
        speaker = want_rand and args.z_type == 'speaker'
        loss, lv = step_losses(outputs, target_dir_vec,
                               out_dir_vec_rand_vid if want_rand else None, z_context.detach() if want_rand else None,
                               z_rand_vid if want_rand else None, z_mu if speaker else None, z_logvar if speaker else None,
                               args.loss_regression_weight, args.loss_reg_weight if want_rand else 0.0,
                               args.loss_kld_weight if speaker else 0.0)
        huber_loss = lv[1]
        if want_rand:
            div_reg = lv[2]
        if speaker:
            kld = lv[3]
    else:
        huber_loss = F.smooth_l1_loss(outputs / 0.1, target_dir_vec / 0.1) * 0.1
        if want_rand:
            beta = 0.05
            pose_l1 = F.smooth_l1_loss(outputs / beta, out_dir_vec_rand_vid.detach() / beta, reduction='none') * beta
            pose_l1 = pose_l1.sum(dim=1).sum(dim=1)
            pose_l1 = pose_l1.view(pose_l1.shape[0], -1).mean(1)
            z_l1 = F.l1_loss(z_context.detach(), z_rand_vid.detach(), reduction='none')
            z_l1 = z_l1.view(z_l1.shape[0], -1).mean(1)
            div_reg = torch.clamp(-(pose_l1 / (z_l1 + 1.0e-5)), min=-1000).mean()
            if args.z_type == 'speaker':
                kld = -0.5 * torch.mean(1 + z_logvar - z_mu.pow(2) - z_logvar.exp())
                loss = huber_loss * args.loss_regression_weight + div_reg * args.loss_reg_weight + kld * args.loss_kld_weight
            else:
                loss = huber_loss * args.loss_regression_weight + div_reg * args.loss_reg_weight
        else:
            loss = huber_loss * args.loss_regression_weight
    if gan:
        loss = loss + gen_error * args.loss_gan_weight

    accelerator.backward(loss)
    model_optim.step()

    names, vals = ['loss'], [args.loss_regression_weight * huber_loss.detach()]
    if kld is not None:
        names.append('KLD'); vals.append(args.loss_kld_weight * kld.detach())
    if div_reg is not None:
        names.append('DIV_REG'); vals.append(args.loss_reg_weight * div_reg.detach())
    if gan:
        names += ['gen', 'dis']; vals += [args.loss_gan_weight * gen_error.detach(), dis_error.detach()]
    return names, torch.stack([v.float().reshape(()) for v in vals])
