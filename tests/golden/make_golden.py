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


if __name__ == '__main__':
    torch.manual_seed(0)
    ref_gwnet, ref_hop = import_reference()
    for n in GW_CASES:
        make_gwnet(ref_gwnet, n)
    for n in RP_CASES:
        make_reprog(ref_hop, n)
