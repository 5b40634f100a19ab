"""CPU oracle (TEST INFRASTRUCTURE ONLY) for the whole HOP generator and its training step.

A functional plain-PyTorch restatement -- it runs on whatever device its tensors live on, the
tests and bench use it on the CPU -- of
  HOP.Model.forecast            reference model/HOP.py:181-252
  gwnet.forward                 reference model/gwnet.py:143-249
  ReprogrammingLayer.forward    reference model/HOP.py:271-299
  train_llm (generator step)    reference train_eval/train_llm.py:38-98
working directly on a ``state_dict`` (the reference's 314 keys) plus the frozen BERT module.
Gradients come from autograd, exactly like the reference obtains them.  It is written from
SURVEY.md (sections 3.2-3.4, Appendix A), deliberately *without* the reference's J-fold
repeated beat MLP (it applies SURVEY F9's equivalent gather), and is pinned against the
reference by tests/golden/make_golden.py::make_model (fixture hop_model_ted.npz).

Product code must not import this module.  Users: tests/, __graft_entry__.smoke,
bench.py's cpu_baseline and ``--impl reference`` legs (kind "port": the reference itself
cannot travel to the GPU box because /root/reference does not exist there).
"""
import torch
import torch.nn.functional as F

DILATIONS = (1, 2, 1, 2, 1, 2, 1, 2)


def gwnet_forward(sd, x, training=True, prefix='gwnet.', update_buffers=True, relu_masks=None, dilations=DILATIONS,
                  capture=None):
    """x (B, in, V, T) -> (B, out, V, T-12). ``sd`` maps names to tensors (leaf tensors for grads).

    ``relu_masks`` = (m_skip (B,S,V,Tl), m_end1 (B,E,V,Tl)), 0/1 tensors: the two head ReLUs (gwnet.py:240-243) become
    multiplications by these fixed masks.  A reduced-precision implementation flips a few gates whose pre-activation lies
    within rounding distance of zero; pinning the oracle to the implementation's own gate pattern separates that
    (discontinuous, measured apart) effect from the arithmetic error of every other operation."""
    g = lambda k: sd[prefix + k]
    rf = 1 + sum(dilations)
    if x.shape[3] < rf:
        x = F.pad(x, (rf - x.shape[3], 0, 0, 0))
    x = F.conv2d(x, g('start_conv.weight'), g('start_conv.bias'))
    A = torch.softmax(torch.relu(g('nodevec1') @ g('nodevec2')), dim=1)
    skip = None
    for i, d in enumerate(dilations):
        res = x
        f = torch.tanh(F.conv2d(res, g(f'filter_convs.{i}.weight'), g(f'filter_convs.{i}.bias'), dilation=(1, d)))
        s = torch.sigmoid(F.conv2d(res, g(f'gate_convs.{i}.weight'), g(f'gate_convs.{i}.bias'), dilation=(1, d)))
        y = f * s
        sk = F.conv2d(y, g(f'skip_convs.{i}.weight'), g(f'skip_convs.{i}.bias'))
        skip = sk if skip is None else sk + skip[..., -sk.shape[3]:]
        x1 = torch.einsum('ncvl,vw->ncwl', y, A)
        x2 = torch.einsum('ncvl,vw->ncwl', x1, A)
        h = F.conv2d(torch.cat([y, x1, x2], 1), g(f'gconv.{i}.mlp.mlp.weight'), g(f'gconv.{i}.mlp.mlp.bias'))
        u = h + res[..., -h.shape[3]:]
        rm, rv = g(f'bn.{i}.running_mean'), g(f'bn.{i}.running_var')
        if not update_buffers:
            rm, rv = rm.clone(), rv.clone()
        x = F.batch_norm(u, rm, rv, g(f'bn.{i}.weight'), g(f'bn.{i}.bias'), training, 0.1, 1e-5)
        if training and update_buffers:
            sd[prefix + f'bn.{i}.num_batches_tracked'] += 1
    if capture is not None:                                # the exact gate pattern, for flip counting
        capture['g0'] = (skip > 0).detach()
        capture['g1'] = (F.conv2d(F.relu(skip), g('end_conv_1.weight'), g('end_conv_1.bias')) > 0).detach()
    if relu_masks is None:
        x = F.relu(F.conv2d(F.relu(skip), g('end_conv_1.weight'), g('end_conv_1.bias')))
    else:
        x = F.conv2d(skip * relu_masks[0], g('end_conv_1.weight'), g('end_conv_1.bias')) * relu_masks[1]
    return F.conv2d(x, g('end_conv_2.weight'), g('end_conv_2.bias'))


def reprogramming_forward(sd, target, source, value, n_heads, p_drop=0.0, prefix='reprogramming_layer.'):
    g = lambda k: sd[prefix + k]
    B, L, _ = target.shape
    S = source.shape[0]
    q = F.linear(target, g('query_projection.weight'), g('query_projection.bias')).view(B, L, n_heads, -1)
    k = F.linear(source, g('key_projection.weight'), g('key_projection.bias')).view(S, n_heads, -1)
    v = F.linear(value, g('value_projection.weight'), g('value_projection.bias')).view(S, n_heads, -1)
    sc = torch.einsum('blhe,she->bhls', q, k) / q.shape[-1] ** 0.5
    a = F.dropout(torch.softmax(sc, dim=-1), p_drop, training=p_drop > 0)
    o = torch.einsum('bhls,she->blhe', a, v).reshape(B, L, -1)
    return F.linear(F.relu(o), g('out_projection.weight'), g('out_projection.bias'))


def _gru(sd, x, hidden=350, layers=4):
    """4-layer bidirectional GRU through torch's own kernel, fed from state_dict tensors."""
    flat = []
    for l in range(layers):
        for suf in ('', '_reverse'):
            flat += [sd[f'gru.weight_ih_l{l}{suf}'], sd[f'gru.weight_hh_l{l}{suf}'],
                     sd[f'gru.bias_ih_l{l}{suf}'], sd[f'gru.bias_hh_l{l}{suf}']]
    h0 = x.new_zeros(2 * layers, x.shape[0], hidden)
    # (input, hx, params, has_biases, num_layers, dropout, train, bidirectional, batch_first); dropout is 0, `train` only
    # tells cuDNN to keep its reserve space for backward
    out, _ = torch._VF.gru(x, h0, flat, True, layers, 0.0, torch.is_grad_enabled(), True, True)
    return out


def model_forward(sd, bert, in_audio, x_enc, text, pre_seq, vid, noise, n_heads=8, p_drop=0.0, training=True,
                  update_buffers=True, literal_beat=False):
    """HOP.Model.forward.  ``noise`` (B,16) replaces reparameterize's randn so both sides share it.
    ``literal_beat``: run the beat MLP on the J-fold repeated windows exactly like HOP.py:210-212 (same values as the
    de-duplicated gather, J times the work) -- what the reference itself executes, used by the timing baselines."""
    B = pre_seq.shape[0]
    J = pre_seq.shape[2] // 3
    e = F.linear(F.embedding(vid, sd['speaker_embedding.0.weight']), sd['speaker_embedding.1.weight'],
                 sd['speaker_embedding.1.bias'])
    z_mu = F.linear(e, sd['speaker_mu.weight'], sd['speaker_mu.bias'])
    z_logvar = F.linear(e, sd['speaker_logvar.weight'], sd['speaker_logvar.bias'])
    z = z_mu + noise * torch.exp(0.5 * z_logvar)
    we = sd['word_embeddings']
    text_emb = F.embedding(text.long(), we)
    source = F.linear(we.t(), sd['mapping_layer.weight'], sd['mapping_layer.bias']).t()
    enc = reprogramming_forward(sd, x_enc, source, source, n_heads, p_drop)
    h = F.linear(torch.cat([enc, text_emb], 2), sd['align_layer.weight'], sd['align_layer.bias'])
    dec = bert(inputs_embeds=h).last_hidden_state
    win = in_audio.unfold(1, 3400, 2191)
    if literal_beat:
        win = win.unsqueeze(1).repeat(1, J, 1, 1)                                  # (B, J, 16, 3400), HOP.py:210
    feat = F.linear(F.leaky_relu(F.linear(win, sd['beat.0.weight'], sd['beat.0.bias']), 0.2), sd['beat.2.weight'],
                    sd['beat.2.bias'])
    if literal_beat:
        feat_g = feat.reshape(B, 16, J, 170)                                       # the reference's .view
    else:
        idx = (torch.arange(16 * J, device=feat.device) % 16).view(16, J)
        feat_g = feat[:, idx]
    seq = torch.cat([pre_seq.reshape(B, 16, J, 3), feat_g], 3).permute(0, 3, 2, 1)
    feature = gwnet_forward(sd, seq, training, update_buffers=update_buffers)
    g_seq = feature[:, :3].reshape(B, 3 * J, -1).permute(0, 2, 1)
    beat = feature[:, 3:].reshape(B, 34, -1)
    seed = g_seq.new_zeros(B, 34, 3 * J + 1)
    seed[:, :g_seq.shape[1], :-1] = g_seq
    seed[:, :g_seq.shape[1], -1] = 1
    full = torch.cat([seed, beat, dec, z.unsqueeze(1).expand(B, 34, 16)], 2).contiguous()
    o = _gru(sd, full)
    o = o[..., :350] + o[..., 350:]
    o = F.linear(o, sd['out.0.weight'], sd['out.0.bias'])          # LeakyReLU(True) == slope 1 == identity
    o = F.linear(o, sd['out.3.weight'], sd['out.3.bias'])
    return o, z, z_mu, z_logvar


def generator_loss(out, out_rand, z, z_rand, z_mu, z_logvar, target, w_reg=600.0, w_div=0.4, w_kld=0.6):
    """Loss of the generator step for epoch <= 10 (train_llm.py:46-79): huber + diversity + KLD."""
    huber = F.smooth_l1_loss(out / 0.1, target / 0.1) * 0.1
    beta = 0.05
    pose_l1 = (F.smooth_l1_loss(out / beta, out_rand.detach() / beta, reduction='none') * beta).sum(1).sum(1)
    pose_l1 = pose_l1.view(pose_l1.shape[0], -1).mean(1)
    z_l1 = F.l1_loss(z.detach(), z_rand.detach(), reduction='none').view(z.shape[0], -1).mean(1)
    div = torch.clamp(-(pose_l1 / (z_l1 + 1.0e-5)), min=-1000).mean()
    kld = -0.5 * torch.mean(1 + z_logvar - z_mu.pow(2) - z_logvar.exp())
    return huber * w_reg + div * w_div + kld * w_kld, huber, div, kld


class OracleTrainer:
    """The reference's generator step (epoch <= 10 semantics, train_llm.py:38-98) on the functional oracle: generator
    forward, discriminator forward (train_llm.py:43-44, computed every step), random-speaker forward, losses, backward,
    Adam(0.5, 0.999) with constant lr.  Used as bench.py's CPU baseline / ``--impl reference`` arm (kind "port") and, on
    the GPU, as the stock-PyTorch-CUDA speed bar.

    ``literal=True`` (the timing baselines) does the work the reference does: the J-fold repeated beat MLP
    (HOP.py:210-212), the mapping GEMM and K/V projections recomputed in each forward, and the random-speaker forward
    with autograd recording (the reference only detaches its outputs afterwards).  ``literal=False`` keeps the lighter
    equivalents (same values) that the parity tests use.
    """

    def __init__(self, state_dict, bert, lr=4e-4, p_drop=0.1, datasets='TED', device='cpu', literal=False,
                 discriminator=None, capturable=False):
        self.bert = bert.to(device)
        self.sd = {}
        self.params = []
        frozen = ('llm_model.', 'word_embeddings')
        for k, v in state_dict.items():
            t = v.detach().clone().to(device)
            if t.is_floating_point() and not k.startswith(frozen) and 'running_' not in k:
                t.requires_grad_(True)
                self.params.append(t)
            self.sd[k] = t
        for p in self.bert.parameters():
            p.requires_grad_(False)
        self.opt = torch.optim.Adam(self.params, lr=lr, betas=(0.5, 0.999), capturable=capturable)
        self.p_drop = p_drop
        self.w = (600.0, 0.4, 0.6) if datasets == 'TED' else (2100.0, 0.5, 0.8)
        self.device, self.literal = device, literal
        self.disc = discriminator.to(device) if discriminator is not None else None

    def step_device(self, in_audio, x_enc, text, target, vid):
        """One step without host synchronisation; returns the loss tensor."""
        dev = self.device
        pre_seq = target[:, :16]
        self.opt.zero_grad(set_to_none=True)
        noise = torch.randn(target.shape[0], 16, device=dev)
        out, z, z_mu, z_lv = model_forward(self.sd, self.bert, in_audio, x_enc, text, pre_seq, vid, noise,
                                           p_drop=self.p_drop, literal_beat=self.literal)
        if self.disc is not None:                          # train_llm.py:43-44: computed, unused for epoch <= 10
            gen_error = -torch.mean(torch.log(self.disc(out, text) + 1e-8))   # noqa: F841
        rand_vid = vid[torch.randperm(vid.shape[0], device=dev)]
        with torch.set_grad_enabled(self.literal):
            out_r, z_r, _, _ = model_forward(self.sd, self.bert, in_audio, x_enc, text, pre_seq, rand_vid,
                                             torch.randn(target.shape[0], 16, device=dev), p_drop=self.p_drop,
                                             literal_beat=self.literal)
        loss, huber, div, kld = generator_loss(out, out_r, z, z_r, z_mu, z_lv, target, *self.w)
        loss.backward()
        self.opt.step()
        return loss.detach()

    def step(self, in_audio, x_enc, text, target, vid):
        return float(self.step_device(in_audio, x_enc, text, target, vid))
