"""CPU oracle (TEST INFRASTRUCTURE ONLY) for the Graph-WaveNet block of HOP.

This is a numpy float64 restatement of the algorithm in the reference's
``model/gwnet.py`` -- forward *and* a hand-derived backward -- written from the
maths, not from the reference's code.  Nothing in the
product package may import this module: only ``tests/``, ``__graft_entry__.smoke``
and ``bench.py``'s cpu_baseline / ``--impl reference`` legs use it, as a checker.

Parity status: the reference ships no golden vectors or tests for this path
(SURVEY.md section 4), so the oracle is pinned against outputs of the reference
module itself, executed in the build container by
``tests/golden/make_golden.py`` (fixtures under ``tests/golden/*.npz``) and
re-checked on every CPU test run by ``tests/test_oracle_golden.py``.

Reference lines restated (relative to /root/reference):
  adaptive adjacency   model/gwnet.py:161-164
  start conv           model/gwnet.py:144-149
  gated dilated conv   model/gwnet.py:186-200
  skip accumulate      model/gwnet.py:209-220
  nconv / gcn          model/gwnet.py:12-14, 33-46
  residual + BN        model/gwnet.py:233-237
  head                 model/gwnet.py:240-246

Layout everywhere: NCHW = (batch, channel, node, time), like the reference.
Parameters are a plain dict keyed by the reference's state_dict names.
"""
import numpy as np

BN_EPS = 1e-5
BN_MOMENTUM = 0.1


def layer_dilations(blocks=4, layers=2):
    """Dilation of each of the blocks*layers gated convs (gwnet.py:98-123)."""
    out = []
    for _ in range(blocks):
        d = 1
        for _ in range(layers):
            out.append(d)
            d *= 2
    return out


def receptive_field(blocks=4, layers=2, kernel_size=2):
    rf = 1
    for _ in range(blocks):
        scope = kernel_size - 1
        for _ in range(layers):
            rf += scope
            scope *= 2
    return rf


def _id(a):
    return a


def bf16_round(a):
    """Round-to-nearest-even to bfloat16 (returned as float64): the operand quantiser of the tensor-core path."""
    a32 = np.ascontiguousarray(a, dtype=np.float32)
    u = a32.view(np.uint32).astype(np.uint64)
    u = (u + 0x7FFF + ((u >> 16) & 1)) & 0xFFFF0000
    return u.astype(np.uint32).view(np.float32).astype(np.float64).reshape(a32.shape)


def _pw(x, w, b, q=_id):
    """1x1 conv: w is (O, C, 1, 1).  ``q`` quantises the two GEMM operands (identity = exact reference)."""
    return np.einsum('oc,bcvt->bovt', q(w[:, :, 0, 0]), q(x)) + b[None, :, None, None]


def adaptive_adjacency(e1, e2):
    """softmax(relu(E1 @ E2), dim=1)  -- gwnet.py:163."""
    z = e1 @ e2
    r = np.maximum(z, 0.0)
    r = r - r.max(axis=1, keepdims=True)
    p = np.exp(r)
    return p / p.sum(axis=1, keepdims=True), z


def _sigmoid(x):
    return 1.0 / (1.0 + np.exp(-x))


def forward(params, x_in, blocks=4, layers=2, training=True, keep=False, q=_id, relu_masks=None):
    """gwnet.forward (gwnet.py:143-249).

    Returns (out, new_buffers, cache).  ``new_buffers`` holds the updated
    ``bn.i.running_mean / running_var / num_batches_tracked`` (train mode).
    ``cache`` is what :func:`backward` needs (only when keep=True).

    ``q`` (default identity) is applied to both operands of every dense contraction -- start/gate/mlp/skip/end
    convs -- and nowhere else; with ``q=bf16_round`` this is the quantisation-aware reference of the bf16
    tensor-core mode (dtype 1: bf16 operands, fp32 accumulation, everything else fp32).

    ``relu_masks`` = (m_skip (B,S,V,Tl), m_end1 (B,E,V,Tl)) of 0/1 pins the gate pattern of the two head ReLUs
    (gwnet.py:240-243) to an implementation's own pattern (see oracle/hop_torch.py::gwnet_forward for why).
    """
    p = {k: np.asarray(v, dtype=np.float64) for k, v in params.items()}
    x = np.asarray(x_in, dtype=np.float64)
    dil = layer_dilations(blocks, layers)
    rf = receptive_field(blocks, layers)
    pad = 0
    if x.shape[3] < rf:                       # gwnet.py:144-148 left-pad in time
        pad = rf - x.shape[3]
        x = np.pad(x, ((0, 0), (0, 0), (0, 0), (pad, 0)))
    x0 = x
    x = _pw(x, p['start_conv.weight'], p['start_conv.bias'], q)
    A, Z = adaptive_adjacency(p['nodevec1'], p['nodevec2'])
    skip = None
    bufs = {}
    cache = dict(x0=x0, pad=pad, A=A, Z=Z, layers=[])
    for i, d in enumerate(dil):
        T = x.shape[3]
        To = T - d
        wf, wg = p[f'filter_convs.{i}.weight'], p[f'gate_convs.{i}.weight']
        xq, wfq, wgq = q(x), q(wf), q(wg)
        f = (np.einsum('oc,bcvt->bovt', wfq[:, :, 0, 0], xq[..., :To]) +
             np.einsum('oc,bcvt->bovt', wfq[:, :, 0, 1], xq[..., d:]) +
             p[f'filter_convs.{i}.bias'][None, :, None, None])
        g = (np.einsum('oc,bcvt->bovt', wgq[:, :, 0, 0], xq[..., :To]) +
             np.einsum('oc,bcvt->bovt', wgq[:, :, 0, 1], xq[..., d:]) +
             p[f'gate_convs.{i}.bias'][None, :, None, None])
        tf, sg = np.tanh(f), _sigmoid(g)
        y = tf * sg
        s = _pw(y, p[f'skip_convs.{i}.weight'], p[f'skip_convs.{i}.bias'], q)
        skip = s if skip is None else s + skip[..., -To:]
        x1 = np.einsum('ncvl,vw->ncwl', y, A)
        x2 = np.einsum('ncvl,vw->ncwl', x1, A)
        hcat = np.concatenate([y, x1, x2], axis=1)
        h = _pw(hcat, p[f'gconv.{i}.mlp.mlp.weight'], p[f'gconv.{i}.mlp.mlp.bias'], q)
        u = h + x[..., -To:]
        gam, bet = p[f'bn.{i}.weight'], p[f'bn.{i}.bias']
        if training:
            mu = u.mean(axis=(0, 2, 3))
            var = u.var(axis=(0, 2, 3))
            n = u.shape[0] * u.shape[2] * u.shape[3]
            bufs[f'bn.{i}.running_mean'] = ((1 - BN_MOMENTUM) * p[f'bn.{i}.running_mean'] + BN_MOMENTUM * mu)
            bufs[f'bn.{i}.running_var'] = ((1 - BN_MOMENTUM) * p[f'bn.{i}.running_var'] +
                                           BN_MOMENTUM * var * n / max(n - 1, 1))
            bufs[f'bn.{i}.num_batches_tracked'] = p[f'bn.{i}.num_batches_tracked'] + 1
        else:
            mu, var = p[f'bn.{i}.running_mean'], p[f'bn.{i}.running_var']
        rstd = 1.0 / np.sqrt(var + BN_EPS)
        xhat = (u - mu[None, :, None, None]) * rstd[None, :, None, None]
        xn = xhat * gam[None, :, None, None] + bet[None, :, None, None]
        if keep:
            cache['layers'].append(dict(x=x, tf=tf, sg=sg, y=y, x1=x1, x2=x2, xhat=xhat, rstd=rstd, d=d))
        x = xn
    g0 = (skip > 0) if relu_masks is None else (np.asarray(relu_masks[0]) > 0)
    r0 = skip * g0
    e1 = _pw(r0, p['end_conv_1.weight'], p['end_conv_1.bias'], q)
    g1 = (e1 > 0) if relu_masks is None else (np.asarray(relu_masks[1]) > 0)
    r1 = e1 * g1
    out = _pw(r1, p['end_conv_2.weight'], p['end_conv_2.bias'], q)
    if keep:
        cache.update(skip=skip, r0=r0, e1=e1, r1=r1, g0=g0, g1=g1, training=training)
    return out, bufs, cache


def _pw_bwd(x, w, dy, q=_id):
    """Backward of a 1x1 conv: returns (dx, dw(O,C,1,1), db).  The bias gradient rides the weight-gradient GEMM as
    an all-ones column in the kernels, so it sums the *quantised* dy."""
    dyq = q(dy)
    dx = np.einsum('oc,bovt->bcvt', q(w[:, :, 0, 0]), dyq)
    dw = np.einsum('bovt,bcvt->oc', dyq, q(x))[:, :, None, None]
    db = dyq.sum(axis=(0, 2, 3))
    return dx, dw, db


def backward(params, cache, dout, blocks=4, layers=2, q=_id):
    """Gradient of sum(out*dout) w.r.t. every parameter and the input.

    Returns (dx_in, grads) with grads keyed like the state_dict; tensors the
    reference never reaches (residual_convs.*, last layer's gconv/bn affine,
    SURVEY F7) are absent, mirroring ``p.grad is None``.
    """
    p = {k: np.asarray(v, dtype=np.float64) for k, v in params.items()}
    dil = layer_dilations(blocks, layers)
    L = len(dil)
    G = {}
    dout = np.asarray(dout, dtype=np.float64)
    dr1, G['end_conv_2.weight'], G['end_conv_2.bias'] = _pw_bwd(cache['r1'], p['end_conv_2.weight'], dout, q)
    de1 = dr1 * cache['g1']
    dr0, G['end_conv_1.weight'], G['end_conv_1.bias'] = _pw_bwd(cache['r0'], p['end_conv_1.weight'], de1, q)
    dskip = dr0 * cache['g0']                    # (B, S, V, T_last)
    Tl = dskip.shape[3]
    A = cache['A']
    M1 = np.zeros_like(A)
    M2 = np.zeros_like(A)
    dxn = None                                    # grad w.r.t. BN output of layer i (= input of layer i+1)
    for i in reversed(range(L)):
        c = cache['layers'][i]
        x, tf, sg, y, x1, x2, d = c['x'], c['tf'], c['sg'], c['y'], c['x1'], c['x2'], c['d']
        To = y.shape[3]
        C = y.shape[1]
        dy = np.zeros_like(y)
        dx = np.zeros_like(x)
        if dxn is not None:
            # BatchNorm backward (train-mode batch statistics or eval running stats)
            gam = p[f'bn.{i}.weight']
            xhat, rstd = c['xhat'], c['rstd']
            G[f'bn.{i}.weight'] = (dxn * xhat).sum(axis=(0, 2, 3))
            G[f'bn.{i}.bias'] = dxn.sum(axis=(0, 2, 3))
            if cache['training']:
                n = xhat.shape[0] * xhat.shape[2] * xhat.shape[3]
                m1 = G[f'bn.{i}.bias'] / n
                m2 = G[f'bn.{i}.weight'] / n
                du = (gam * rstd)[None, :, None, None] * (dxn - m1[None, :, None, None] - xhat * m2[None, :, None, None])
            else:
                du = (gam * rstd)[None, :, None, None] * dxn
            # residual x[..., -To:]
            dx[..., -To:] += du
            # gcn mlp on cat[y, x1, x2] and the two diffusion hops, in the form the kernels use
            # (mathematically identical to back-propagating hop by hop):
            #   P1 = A du, P2 = A P1 (node mixing with A, not A^T);  dy = [du|P1|P2] . Wm
            #   dWm = du^T [y|x1|x2];   dA = M1 + A^T M2 + M2 A^T  with  M_k = sum y (x) (du . Wm_k)
            wm = p[f'gconv.{i}.mlp.mlp.weight'][:, :, 0, 0]
            P1 = np.einsum('ncwl,vw->ncvl', du, A)
            P2 = np.einsum('ncwl,vw->ncvl', P1, A)
            duq, wmq = q(du), q(wm)
            hcat = np.concatenate([y, x1, x2], axis=1)
            G[f'gconv.{i}.mlp.mlp.weight'] = np.einsum('bovt,bcvt->oc', duq, q(hcat))[:, :, None, None]
            G[f'gconv.{i}.mlp.mlp.bias'] = duq.sum(axis=(0, 2, 3))
            dy += (np.einsum('oc,bovt->bcvt', wmq[:, :C], duq) +
                   np.einsum('oc,bovt->bcvt', wmq[:, C:2 * C], q(P1)) +
                   np.einsum('oc,bovt->bcvt', wmq[:, 2 * C:], q(P2)))
            M1 += np.einsum('ncvl,ncwl->vw', y, np.einsum('oc,bovt->bcvt', wmq[:, C:2 * C], duq))
            M2 += np.einsum('ncvl,ncwl->vw', y, np.einsum('oc,bovt->bcvt', wmq[:, 2 * C:], duq))
        # skip path: only the last Tl time steps of s_i reach the head
        ws = p[f'skip_convs.{i}.weight']
        ds = np.zeros((y.shape[0], ws.shape[0], y.shape[2], To))
        ds[..., -Tl:] = dskip
        dys, G[f'skip_convs.{i}.weight'], G[f'skip_convs.{i}.bias'] = _pw_bwd(y, ws, ds, q)
        dy += dys
        # gate
        df = dy * sg * (1 - tf * tf)
        dg = dy * tf * sg * (1 - sg)
        wf, wg = p[f'filter_convs.{i}.weight'], p[f'gate_convs.{i}.weight']
        xq = q(x)
        xa, xb = xq[..., :To], xq[..., d:]
        df, dg, wf, wg = q(df), q(dg), q(wf), q(wg)
        G[f'filter_convs.{i}.weight'] = np.stack([np.einsum('bovt,bcvt->oc', df, xa),
                                                   np.einsum('bovt,bcvt->oc', df, xb)], axis=-1)[:, :, None, :]
        G[f'gate_convs.{i}.weight'] = np.stack([np.einsum('bovt,bcvt->oc', dg, xa),
                                                 np.einsum('bovt,bcvt->oc', dg, xb)], axis=-1)[:, :, None, :]
        G[f'filter_convs.{i}.bias'] = df.sum(axis=(0, 2, 3))
        G[f'gate_convs.{i}.bias'] = dg.sum(axis=(0, 2, 3))
        dx[..., :To] += (np.einsum('oc,bovt->bcvt', wf[:, :, 0, 0], df) +
                         np.einsum('oc,bovt->bcvt', wg[:, :, 0, 0], dg))
        dx[..., d:] += (np.einsum('oc,bovt->bcvt', wf[:, :, 0, 1], df) +
                        np.einsum('oc,bovt->bcvt', wg[:, :, 0, 1], dg))
        dxn = dx
    # adaptive adjacency backward: A = softmax_row(relu(E1 E2))
    dA = M1 + A.T @ M2 + M2 @ A.T
    dR = A * (dA - (dA * A).sum(axis=1, keepdims=True))
    dZ = dR * (cache['Z'] > 0)
    G['nodevec1'] = dZ @ p['nodevec2'].T
    G['nodevec2'] = p['nodevec1'].T @ dZ
    dx0, G['start_conv.weight'], G['start_conv.bias'] = _pw_bwd(cache['x0'], p['start_conv.weight'], dxn, q)
    if cache['pad']:
        dx0 = dx0[..., cache['pad']:]
    return dx0, G


def init_params(rng, num_nodes, in_dim=173, out_dim=173, residual=64, dilation=64, skip=256, end=512,
                blocks=4, layers=2, scale=None):
    """Random parameters with the reference's state_dict names/shapes (gwnet.py:50-139).

    Drawn from ``rng`` (a numpy RandomState) so fixtures can be regenerated from a
    seed on any machine; magnitudes follow PyTorch's default fan-in scaling.
    """
    def u(shape, fan_in):
        b = 1.0 / np.sqrt(fan_in)
        return rng.uniform(-b, b, size=shape)
    P = {}
    P['nodevec1'] = rng.standard_normal((num_nodes, 10))
    P['nodevec2'] = rng.standard_normal((10, num_nodes))
    P['start_conv.weight'] = u((residual, in_dim, 1, 1), in_dim)
    P['start_conv.bias'] = u((residual,), in_dim)
    for i in range(blocks * layers):
        P[f'filter_convs.{i}.weight'] = u((dilation, residual, 1, 2), 2 * residual)
        P[f'filter_convs.{i}.bias'] = u((dilation,), 2 * residual)
        P[f'gate_convs.{i}.weight'] = u((dilation, residual, 1, 2), 2 * residual)
        P[f'gate_convs.{i}.bias'] = u((dilation,), 2 * residual)
        P[f'residual_convs.{i}.weight'] = u((residual, dilation, 1, 1), dilation)
        P[f'residual_convs.{i}.bias'] = u((residual,), dilation)
        P[f'skip_convs.{i}.weight'] = u((skip, dilation, 1, 1), dilation)
        P[f'skip_convs.{i}.bias'] = u((skip,), dilation)
        P[f'bn.{i}.weight'] = rng.uniform(0.5, 1.5, size=(residual,))
        P[f'bn.{i}.bias'] = rng.uniform(-0.2, 0.2, size=(residual,))
        P[f'bn.{i}.running_mean'] = rng.uniform(-0.1, 0.1, size=(residual,))
        P[f'bn.{i}.running_var'] = rng.uniform(0.8, 1.2, size=(residual,))
        P[f'bn.{i}.num_batches_tracked'] = np.array(3, dtype=np.int64)
        P[f'gconv.{i}.mlp.mlp.weight'] = u((residual, 3 * dilation, 1, 1), 3 * dilation)
        P[f'gconv.{i}.mlp.mlp.bias'] = u((residual,), 3 * dilation)
    P['end_conv_1.weight'] = u((end, skip, 1, 1), skip)
    P['end_conv_1.bias'] = u((end,), skip)
    P['end_conv_2.weight'] = u((out_dim, end, 1, 1), end)
    P['end_conv_2.bias'] = u((out_dim,), end)
    return P
