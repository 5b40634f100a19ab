"""Generate the golden fixtures in tests/golden/ by EXECUTING THE REFERENCE modules.

Run in the build container only (needs /root/reference; the GPU box has no
reference tree -- tests read only the committed .npz files):

    python tests/golden/make_golden.py

For every case it (1) draws parameters/inputs from a numpy RandomState seed via
the oracle's ``init_params`` (so tests can regenerate them from the seed alone),
(2) loads them into the reference ``model.gwnet.gwnet`` /
``model.HOP.ReprogrammingLayer`` in float64, runs forward + autograd backward,
(3) asserts the numpy oracle agrees with the reference to 1e-9, and (4) stores
the reference's outputs.  Large gradient tensors are stored as a seeded random
sample of 512 entries plus their L2 norm to keep fixtures small.
"""
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
REF = os.environ.get('HOP_REFERENCE', '/root/reference')

from oracle import gwnet_np, reprog_np  # noqa: E402

SAMPLE = 512


def import_reference():
    """SURVEY Appendix C shim: stub the plotting / fasttext imports the reference pulls in."""
    sys.path.insert(0, REF)
    for name in ['matplotlib', 'matplotlib.pyplot', 'matplotlib.colors', 'seaborn', 'fasttext']:
        if name not in sys.modules:
            sys.modules[name] = types.ModuleType(name)
    sys.modules['matplotlib.colors'].LinearSegmentedColormap = object
    sys.modules['matplotlib'].pyplot = sys.modules['matplotlib.pyplot']
    sys.modules['matplotlib'].colors = sys.modules['matplotlib.colors']
    from model import gwnet as ref_gwnet
    from model import HOP as ref_hop
    return ref_gwnet, ref_hop


def sample_idx(name, n):
    rs = np.random.RandomState(abs(hash_name(name)) % (2 ** 31))
    return rs.randint(0, n, size=SAMPLE)


def hash_name(name):
    h = 2166136261
    for ch in name.encode():
        h = ((h ^ ch) * 16777619) & 0xFFFFFFFF
    return h


def pack_grad(out, key, g):
    g = np.asarray(g, np.float64).ravel()
    if g.size <= 1024:
        out['full:' + key] = g
    else:
        out['samp:' + key] = g[sample_idx(key, g.size)]
        out['norm:' + key] = np.array(np.linalg.norm(g))


GW_CASES = {
    # name: (seed, B, V, T, cfg)
    'gwnet_tiny': (11, 3, 5, 16, dict(in_dim=6, out_dim=7, residual=32, dilation=32, skip=16, end=24)),
    'gwnet_tiny_pad': (12, 2, 4, 10, dict(in_dim=3, out_dim=5, residual=32, dilation=32, skip=8, end=8)),
    'gwnet_ted': (13, 2, 9, 16, dict(in_dim=173, out_dim=173, residual=64, dilation=64, skip=256, end=512)),
    'gwnet_expr': (14, 1, 42, 16, dict(in_dim=173, out_dim=173, residual=64, dilation=64, skip=256, end=512)),
    'gwnet_ted_eval': (15, 2, 9, 16, dict(in_dim=173, out_dim=173, residual=64, dilation=64, skip=256, end=512)),
}


def gw_inputs(seed, B, V, T, cfg):
    rs = np.random.RandomState(seed)
    P = gwnet_np.init_params(rs, V, **cfg)
    x = rs.standard_normal((B, cfg['in_dim'], V, T))
    T_out = max(T, gwnet_np.receptive_field()) - gwnet_np.receptive_field() + 1
    dout = rs.standard_normal((B, cfg['out_dim'], V, T_out))
    return P, x, dout


def make_gwnet(ref_gwnet, name):
    seed, B, V, T, cfg = GW_CASES[name]
    training = not name.endswith('_eval')
    P, x, dout = gw_inputs(seed, B, V, T, cfg)
    m = ref_gwnet.gwnet(torch.device('cpu'), V, dropout=0, supports=None, gcn_bool=True, addaptadj=True,
                        aptinit=None, in_dim=cfg['in_dim'], out_dim=cfg['out_dim'],
                        residual_channels=cfg['residual'], dilation_channels=cfg['dilation'],
                        skip_channels=cfg['skip'], end_channels=cfg['end']).double()
    sd = {k: torch.from_numpy(np.asarray(v)) for k, v in P.items()}
    assert set(sd) == set(m.state_dict()), set(sd) ^ set(m.state_dict())
    m.load_state_dict(sd, strict=True)
    m.train(training)
    xt = torch.from_numpy(x).requires_grad_(True)
    out = m(xt)
    (out * torch.from_numpy(dout)).sum().backward()
    # oracle vs reference
    o_out, o_bufs, cache = gwnet_np.forward(P, x, training=training, keep=True)
    o_dx, o_G = gwnet_np.backward(P, cache, dout)
    err = lambda a, b: float(np.abs(a - b).max() / (np.abs(b).max() + 1e-6))
    assert err(o_out, out.detach().numpy()) < 1e-9, err(o_out, out.detach().numpy())
    assert err(o_dx, xt.grad.numpy()) < 1e-9, err(o_dx, xt.grad.numpy())
    fix = dict(out=out.detach().numpy())
    pack_grad(fix, 'dx', xt.grad.numpy())
    none_grads = []
    for k, p_ in m.named_parameters():
        if p_.grad is None:
            none_grads.append(k)
            assert k not in o_G, k
            continue
        e = err(o_G[k].reshape(p_.shape), p_.grad.numpy())
        assert e < 1e-8, (k, e)
        pack_grad(fix, k, p_.grad.numpy())
    fix['none_grads'] = np.array(none_grads)
    for k, v in m.state_dict().items():
        if 'running_' in k or 'num_batches' in k:
            fix['buf:' + k] = v.numpy()
            if training:
                assert err(o_bufs[k], v.numpy()) < 1e-12, k
    np.savez_compressed(os.path.join(HERE, name + '.npz'), **fix)
    print(name, 'ok; out', out.shape, 'none_grads', len(none_grads))


RP_CASES = {
    'reprog_tiny': (21, 3, 7, 50, dict(d_model=16, n_heads=2, d_keys=8, d_llm=24)),
    'reprog_hop': (22, 2, 34, 1500, dict(d_model=128, n_heads=8, d_keys=128, d_llm=768)),
}


def rp_inputs(seed, B, L, S, cfg):
    rs = np.random.RandomState(seed)
    P = reprog_np.init_params(rs, **cfg)
    x = rs.standard_normal((B, L, cfg['d_model']))
    src = rs.standard_normal((S, cfg['d_llm'])) * 0.5
    dY = rs.standard_normal((B, L, cfg['d_llm']))
    return P, x, src, dY


def make_reprog(ref_hop, name):
    seed, B, L, S, cfg = RP_CASES[name]
    P, x, src, dY = rp_inputs(seed, B, L, S, cfg)
    m = ref_hop.ReprogrammingLayer(cfg['d_model'], cfg['n_heads'], cfg['d_keys'], cfg['d_llm']).double()
    m.load_state_dict({k: torch.from_numpy(v) for k, v in P.items()}, strict=True)
    m.dropout.p = 0.0
    xt = torch.from_numpy(x).requires_grad_(True)
    st = torch.from_numpy(src).requires_grad_(True)
    y = m(xt, st, st)
    (y * torch.from_numpy(dY)).sum().backward()
    o_y, cache = reprog_np.forward(P, x, src, src, cfg['n_heads'], keep=True)
    o_dx, o_ds, o_dv, o_G = reprog_np.backward(P, cache, dY, cfg['n_heads'])
    err = lambda a, b: float(np.abs(a - b).max() / (np.abs(b).max() + 1e-6))
    assert err(o_y, y.detach().numpy()) < 1e-10
    assert err(o_dx, xt.grad.numpy()) < 1e-9
    assert err(o_ds + o_dv, st.grad.numpy()) < 1e-9
    fix = dict(out=y.detach().numpy())
    pack_grad(fix, 'dx', xt.grad.numpy())
    pack_grad(fix, 'dsource', st.grad.numpy())
    for k, p_ in m.named_parameters():
        assert err(o_G[k], p_.grad.numpy()) < 1e-9, k
        pack_grad(fix, k, p_.grad.numpy())
    np.savez_compressed(os.path.join(HERE, name + '.npz'), **fix)
    print(name, 'ok; out', y.shape)


# ---------------------------------------------------------------------------------------------- full model
class DummyTok:
    eos_token = None
    pad_token = None

    def add_special_tokens(self, d):
        return None


class DummySpk:
    n_words = 1370


MODEL_SEED = 2021
CHECK_PARAMS = ['gwnet.nodevec1', 'gwnet.start_conv.weight', 'gwnet.filter_convs.3.weight', 'gwnet.gconv.5.mlp.mlp.weight',
                'gwnet.end_conv_2.bias', 'reprogramming_layer.query_projection.weight',
                'reprogramming_layer.out_projection.bias', 'mapping_layer.bias', 'align_layer.weight', 'beat.0.weight',
                'beat.2.bias', 'gru.weight_ih_l0', 'gru.weight_hh_l3_reverse', 'out.3.weight', 'speaker_mu.weight',
                'speaker_embedding.0.weight', 'llm_model.encoder.layer.5.output.dense.weight', 'word_embeddings',
                'audio_encoder.feat_extractor.0.weight']
GRAD_PARAMS = ['gwnet.nodevec1', 'gwnet.nodevec2', 'gwnet.start_conv.weight', 'gwnet.filter_convs.0.weight',
               'gwnet.gate_convs.7.weight', 'gwnet.skip_convs.2.weight', 'gwnet.gconv.3.mlp.mlp.weight', 'gwnet.bn.4.weight',
               'gwnet.end_conv_1.weight', 'gwnet.end_conv_2.bias', 'reprogramming_layer.query_projection.weight',
               'reprogramming_layer.key_projection.weight', 'reprogramming_layer.value_projection.bias',
               'reprogramming_layer.out_projection.weight', 'mapping_layer.bias', 'mapping_layer.weight',
               'align_layer.weight', 'beat.0.weight', 'beat.2.bias', 'gru.weight_ih_l0', 'out.3.weight',
               'speaker_mu.weight']


def model_inputs(datasets='TED', B=2, seed=31):
    rs = np.random.RandomState(seed)
    pose = 27 if datasets == 'TED' else 126
    return dict(in_audio=(0.1 * rs.standard_normal((B, 36267))).astype(np.float32),
                x_enc=(-80 * rs.rand(B, 34, 128)).astype(np.float32),
                text=(rs.randint(0, 30522, (B, 34)) * (rs.rand(B, 34) > 0.7)).astype(np.int64),
                pre_seq=np.clip(0.3 * rs.standard_normal((B, 16, pose)), -1, 1).astype(np.float32),
                vid=rs.randint(0, 1370, (B,)).astype(np.int64),
                noise=rs.standard_normal((B, 16)).astype(np.float32),
                d_out=rs.standard_normal((B, 34, pose)).astype(np.float32),
                d_mu=rs.standard_normal((B, 16)).astype(np.float32),
                d_lv=rs.standard_normal((B, 16)).astype(np.float32))


def model_configs(datasets='TED'):
    return types.SimpleNamespace(d_ff=128, llm_dim=768, use_gwnet=True, use_reprograme=True, d_model=128, n_heads=8,
                                 datasets=datasets)


def build_bert():
    from transformers import BertConfig, BertModel
    return BertModel(BertConfig(num_hidden_layers=6)).eval()


def make_model(ref_hop, datasets='TED'):
    """Reference HOP.Model, fp32 (the reference's own precision), seed 2021, dropout p=0, shared reparameterize noise."""
    from model import embedding_net
    inp = model_inputs(datasets)
    torch.manual_seed(MODEL_SEED)
    bert = build_bert()
    m = ref_hop.Model(model_configs(datasets), bert, DummyTok(), DummySpk()).float()
    m.reprogramming_layer.dropout.p = 0.0
    noise = torch.from_numpy(inp['noise'])
    embedding_net.reparameterize = lambda mu, logvar: mu + noise * torch.exp(0.5 * logvar)
    sd = m.state_dict()
    fix = {'wsum:' + k: np.array([float(sd[k].double().sum()), float(sd[k].double().abs().sum())]) for k in CHECK_PARAMS}
    fix['n_keys'] = np.array(len(sd))
    fix['keys'] = np.array(sorted(sd.keys()))
    t = lambda k: torch.from_numpy(inp[k])
    out, z, z_mu, z_lv = m(t('in_audio'), t('x_enc'), t('text'), t('pre_seq'), t('vid'))
    loss = (out * t('d_out')).sum() + (z_mu * t('d_mu')).sum() + (z_lv * t('d_lv')).sum()
    loss.backward()
    fix.update(out=out.detach().numpy(), z=z.detach().numpy(), z_mu=z_mu.detach().numpy(), z_logvar=z_lv.detach().numpy())
    params = dict(m.named_parameters())
    for k in GRAD_PARAMS:
        pack_grad(fix, k, params[k].grad.numpy())
    fix['none_grads'] = np.array(sorted(k for k, p_ in params.items() if p_.requires_grad and p_.grad is None))
    name = 'hop_model_' + ('ted' if datasets == 'TED' else 'expr')
    np.savez_compressed(os.path.join(HERE, name + '.npz'), **fix)
    print(name, 'ok; keys', len(sd), 'out', out.shape, 'none_grads', len(fix['none_grads']))


# ---------------------------------------------------------------------------------------------- 200-step loss curve
CURVE_STEPS, CURVE_B = 200, 4


def curve_batch(step, datasets='TED', B=CURVE_B):
    """Synthetic batch of training step ``step`` (SURVEY 8(d) shapes), reproducible from numpy alone."""
    rs = np.random.RandomState(10_000 + step)
    pose = 27 if datasets == 'TED' else 126
    return dict(in_audio=(0.1 * rs.standard_normal((B, 36267))).astype(np.float32),
                melspec=(-80 * rs.rand(B, 34, 128)).astype(np.float32),
                text=(rs.randint(0, 30522, (B, 34)) * (rs.rand(B, 34) > 0.7)).astype(np.int64),
                target=np.clip(0.3 * rs.standard_normal((B, 34, pose)), -1, 1).astype(np.float32),
                vid=rs.randint(0, 1370, (B,)).astype(np.int64))


class NoiseSource:
    """Shared randomness of the step: reparameterize noise and the speaker permutation, in call order."""

    def __init__(self, B=CURVE_B):
        self.rs, self.B = np.random.RandomState(4242), B

    def noise(self):
        return torch.from_numpy(self.rs.standard_normal((self.B, 16)).astype(np.float32))

    def perm(self):
        return torch.from_numpy(self.rs.permutation(self.B).astype(np.int64))


def curve_args(datasets='TED'):
    ted = datasets == 'TED'
    return types.SimpleNamespace(z_type='speaker', loss_regression_weight=600.0 if ted else 2100.0, loss_gan_weight=5.0,
                                 loss_kld_weight=0.6 if ted else 0.8, loss_reg_weight=0.4 if ted else 0.5)


class PlainAccelerator:
    def backward(self, loss):
        loss.backward()


def make_loss_curve(ref_hop):
    """Run the REFERENCE train_llm (train_eval/train_llm.py) for 200 generator steps on the CPU, fp32."""
    import importlib.machinery
    torch.manual_seed(MODEL_SEED)
    bert = build_bert()
    for name in ['soundfile', 'librosa', 'lmdb']:            # data-loader imports pulled in by train_llm.py:2
        if name not in sys.modules:
            mod = types.ModuleType(name)
            mod.__spec__ = importlib.machinery.ModuleSpec(name, None)
            sys.modules[name] = mod
    from model import embedding_net
    from model.multimodal_context_net import ConvDiscriminator
    from train_eval import train_llm as ref_step
    m = ref_hop.Model(model_configs('TED'), bert, DummyTok(), DummySpk()).float()
    m.reprogramming_layer.dropout.p = 0.0
    disc = ConvDiscriminator(27)
    opt = torch.optim.Adam([p_ for p_ in m.parameters() if p_.requires_grad], lr=4e-4, betas=(0.5, 0.999))
    dopt = torch.optim.Adam(disc.parameters(), lr=4e-4, betas=(0.5, 0.999))
    src = NoiseSource()
    embedding_net.reparameterize = lambda mu, logvar: mu + src.noise() * torch.exp(0.5 * logvar)
    real_randperm = torch.randperm
    torch.randperm = lambda n, **kw: src.perm()
    rows = []
    try:
        for step in range(CURVE_STEPS):
            b = {k: torch.from_numpy(v) for k, v in curve_batch(step).items()}
            ret = ref_step.train_llm(curve_args(), 1, b['in_audio'], b['melspec'], b['text'], b['target'], b['vid'], m, disc,
                                     opt, dopt, PlainAccelerator())
            rows.append([ret.get('loss', 0.0), ret.get('KLD', 0.0), ret.get('DIV_REG', 0.0)])
            if step % 20 == 0:
                print('step', step, rows[-1], flush=True)
    finally:
        torch.randperm = real_randperm
    np.savez_compressed(os.path.join(HERE, 'loss_curve_ted.npz'), curve=np.array(rows, np.float64))
    print('loss_curve_ted ok')


if __name__ == '__main__':
    torch.manual_seed(0)
    ref_gwnet, ref_hop = import_reference()
    if len(sys.argv) > 1 and sys.argv[1] == 'curve':
        make_loss_curve(ref_hop)
        sys.exit(0)
    for n in GW_CASES:
        make_gwnet(ref_gwnet, n)
    for n in RP_CASES:
        make_reprog(ref_hop, n)
    make_model(ref_hop, 'TED')
    make_model(ref_hop, 'TED_expressive')
