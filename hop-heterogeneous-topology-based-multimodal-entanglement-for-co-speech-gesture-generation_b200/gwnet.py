"""Drop-in mirror of the reference's ``model/gwnet.py`` on hand-written sm_100a kernels.

Same classes, constructor signatures, ``forward()`` contracts, attribute names and
``state_dict`` keys as the reference (``nconv`` gwnet.py:8, ``linear`` :16, ``gcn`` :24,
``gwnet`` :49), so ``model/HOP.py``'s ``gwnet.gwnet(device, num_nodes, ...)`` call
(HOP.py:143) and ``load_state_dict`` of reference checkpoints work unchanged.  Parameters
are created with the same initialisers in the same order, so a given ``torch.manual_seed``
yields the same initial weights as the reference.

All maths runs in libhopk.so (csrc/gwnet.cu, csrc/linear.cu) through the C ABI of
include/hopk.h; there is no eager fallback -- configurations the kernels do not cover raise
``NotImplementedError`` and CPU tensors raise ``RuntimeError``.
"""
import torch
import torch.nn as nn
import torch.nn.functional as F

from . import _lib, profiler
from ._lib import GwnetGrads, GwnetParams, GwnetShape, check, f32c, lib, ptr, stream_ptr


# ------------------------------------------------------------------------------------ nconv
class _NconvFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, A):
        x, A = f32c(x), f32c(A)
        N, C_, V, T = x.shape
        out = torch.empty_like(x)
        check(lib().hopk_nconv_fwd(ptr(x), ptr(A), ptr(out), N, C_, V, T, stream_ptr()))
        ctx.save_for_backward(x, A)
        return out

    @staticmethod
    def backward(ctx, dout):
        x, A = ctx.saved_tensors
        dout = f32c(dout)
        N, C_, V, T = x.shape
        dx = torch.empty_like(x) if ctx.needs_input_grad[0] else None
        dA = torch.empty_like(A) if ctx.needs_input_grad[1] else None
        check(lib().hopk_nconv_bwd(ptr(x), ptr(A), ptr(dout), ptr(dx), ptr(dA), N, C_, V, T, stream_ptr()))
        return dx, dA


class nconv(nn.Module):
    """einsum('ncvl,vw->ncwl') + contiguous  (reference gwnet.py:8-14)."""

    def __init__(self):
        super(nconv, self).__init__()

    def forward(self, x, A):
        return _NconvFn.apply(x, A)


# ------------------------------------------------------------------------------------ linear (1x1 conv)
class _Conv1x1Fn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, w, b):
        x = f32c(x)
        B, K, V, T = x.shape
        N = w.shape[0]
        y = torch.empty((B, N, V, T), device=x.device, dtype=torch.float32)
        check(lib().hopk_conv1x1_nchw_fwd(ptr(x), ptr(f32c(w)), ptr(b), ptr(y), B, K, N, V, T, stream_ptr()))
        ctx.save_for_backward(x, w)
        ctx.has_bias = b is not None
        return y

    @staticmethod
    def backward(ctx, dy):
        x, w = ctx.saved_tensors
        dy = f32c(dy)
        B, K, V, T = x.shape
        N = w.shape[0]
        dx = torch.empty_like(x) if ctx.needs_input_grad[0] else None
        dw = torch.empty_like(w, memory_format=torch.contiguous_format)
        db = torch.empty(N, device=x.device, dtype=torch.float32) if ctx.has_bias else None
        check(lib().hopk_conv1x1_nchw_bwd(ptr(x), ptr(f32c(w)), ptr(dy), ptr(dx), ptr(dw), ptr(db), B, K, N, V, T,
                                          stream_ptr()))
        return dx, dw, db


class linear(nn.Module):
    """1x1 Conv2d with bias (reference gwnet.py:16-22); the parameter lives in ``self.mlp``."""

    def __init__(self, c_in, c_out):
        super(linear, self).__init__()
        self.mlp = torch.nn.Conv2d(c_in, c_out, kernel_size=(1, 1), padding=(0, 0), stride=(1, 1), bias=True)

    def forward(self, x):
        return _Conv1x1Fn.apply(x, self.mlp.weight, self.mlp.bias)


# ------------------------------------------------------------------------------------ gcn
class gcn(nn.Module):
    """Order-k diffusion over the supports, concat, 1x1 mlp, dropout (reference gwnet.py:24-46).

    Stand-alone use composes the nconv / conv1x1 kernels; inside :class:`gwnet` the whole
    layer is driven by the fused per-layer path instead.
    """

    def __init__(self, c_in, c_out, dropout, support_len=3, order=2):
        super(gcn, self).__init__()
        self.nconv = nconv()
        c_in = (order * support_len + 1) * c_in
        self.mlp = linear(c_in, c_out)
        self.dropout = dropout
        self.order = order

    def forward(self, x, support):
        out = [x]
        for a in support:
            x1 = self.nconv(x, a)
            out.append(x1)
            for _ in range(2, self.order + 1):
                x2 = self.nconv(x1, a)
                out.append(x2)
                x1 = x2
        h = self.mlp(torch.cat(out, dim=1))
        return F.dropout(h, self.dropout, training=self.training)


# ------------------------------------------------------------------------------------ gwnet
class _GwnetFn(torch.autograd.Function):
    """One autograd node for the whole block: forward = hopk_gwnet_forward, backward = hopk_gwnet_backward."""

    @staticmethod
    def forward(ctx, mod, x, *params):
        shape = mod._shape(x)
        pstruct, keep = mod._param_struct(params)
        B, _, V, T = x.shape
        l = lib()
        t_out = l.hopk_gwnet_out_steps(shape)
        out = torch.empty((B, shape.out_dim, V, t_out), device=x.device, dtype=torch.float32)
        ws = torch.empty(l.hopk_gwnet_workspace_bytes(shape), device=x.device, dtype=torch.uint8)
        xs = _lib._I64x4(*x.stride())
        with profiler.span('gwnet_fwd'):
            check(l.hopk_gwnet_forward(shape, pstruct, ptr(x), xs, ptr(out), ptr(ws), stream_ptr()))
        # saved tensors (not plain attributes): autograd's version counters catch in-place edits of x / the parameters
        # between forward and backward, and a second backward without retain_graph raises autograd's own clear error
        ctx.save_for_backward(x, ws, *params)
        ctx.mod, ctx.shape = mod, shape
        if mod._keep_ws:                                   # tests / tools: inspect what backward will re-read
            mod._last_ws, mod._last_shape = ws, shape
        return out

    @staticmethod
    def backward(ctx, dout):
        mod, shape = ctx.mod, ctx.shape
        x, ws, *params = ctx.saved_tensors
        l = lib()
        dout = f32c(dout)
        pstruct, keep = mod._param_struct(params)
        L = shape.L
        # one flat buffer for every gradient the reference produces (the last layer's
        # gconv mlp and BatchNorm affine get none)
        live = [i for i, (name, layer) in enumerate(mod._param_names()) if not mod._is_dead(name, layer)]
        sizes = [params[i].numel() for i in live]
        flat = torch.empty(sum(sizes), device=x.device, dtype=torch.float32)
        grads = [None] * len(params)
        off = 0
        for i, n in zip(live, sizes):
            grads[i] = flat[off:off + n].view(params[i].shape)
            off += n
        g = GwnetGrads()
        g.flat, g.flat_bytes = ptr(flat), flat.numel() * 4
        for i, (name, layer) in enumerate(mod._param_names()):
            if layer is None:
                setattr(g, name, ptr(grads[i]))
            else:
                getattr(g, name)[layer] = ptr(grads[i]).value or 0
        dx = None
        if ctx.needs_input_grad[1]:
            dx = torch.empty((shape.B, shape.T, shape.V, shape.in_dim), device=x.device, dtype=torch.float32)
        scratch = torch.empty(l.hopk_gwnet_scratch_bytes(shape), device=x.device, dtype=torch.uint8)
        xs = _lib._I64x4(*x.stride())
        with profiler.span('gwnet_bwd'):
            check(l.hopk_gwnet_backward(shape, pstruct, ptr(x), xs, ptr(dout), ptr(ws), ptr(scratch), g, ptr(dx),
                                        stream_ptr()))
        if dx is not None:
            dx = dx.permute(0, 3, 2, 1)          # (B, T, V, C) rows layout -> (B, C, V, T) view
        return (None, dx, *grads)


class gwnet(nn.Module):
    """Graph WaveNet block (reference gwnet.py:49-249), same constructor and forward contract."""

    def __init__(self, device, num_nodes, dropout=0.3, supports=None, gcn_bool=True, addaptadj=True, aptinit=None,
                 in_dim=2, out_dim=12, residual_channels=32, dilation_channels=32, skip_channels=256, end_channels=512,
                 kernel_size=2, blocks=4, layers=2):
        super(gwnet, self).__init__()
        self.dropout = dropout
        self.blocks = blocks
        self.layers = layers
        self.gcn_bool = gcn_bool
        self.addaptadj = addaptadj
        self.kernel_size = kernel_size
        self.num_nodes = num_nodes

        self.filter_convs = nn.ModuleList()
        self.gate_convs = nn.ModuleList()
        self.residual_convs = nn.ModuleList()
        self.skip_convs = nn.ModuleList()
        self.bn = nn.ModuleList()
        self.gconv = nn.ModuleList()

        # creation order == reference order, so the RNG stream (and thus the init) matches
        self.start_conv = nn.Conv2d(in_channels=in_dim, out_channels=residual_channels, kernel_size=(1, 1))
        self.supports = supports
        receptive_field = 1
        self.supports_len = 0
        if supports is not None:
            self.supports_len += len(supports)
        if gcn_bool and addaptadj:
            if supports is None:
                self.supports = []
            if aptinit is None:
                self.nodevec1 = nn.Parameter(torch.randn(num_nodes, 10).to(device), requires_grad=True)
                self.nodevec2 = nn.Parameter(torch.randn(10, num_nodes).to(device), requires_grad=True)
            else:                                                     # SVD initialisation, gwnet.py:85-93
                m, p, n = torch.svd(aptinit)
                initemb1 = torch.mm(m[:, :10], torch.diag(p[:10] ** 0.5))
                initemb2 = torch.mm(torch.diag(p[:10] ** 0.5), n[:, :10].t())
                self.nodevec1 = nn.Parameter(initemb1.to(device), requires_grad=True)
                self.nodevec2 = nn.Parameter(initemb2.to(device), requires_grad=True)
            self.supports_len += 1

        self.dilations = []
        for _ in range(blocks):
            additional_scope = kernel_size - 1
            new_dilation = 1
            for _ in range(layers):
                self.filter_convs.append(nn.Conv2d(residual_channels, dilation_channels, kernel_size=(1, kernel_size),
                                                   dilation=new_dilation))
                self.gate_convs.append(nn.Conv2d(residual_channels, dilation_channels, kernel_size=(1, kernel_size),
                                                 dilation=new_dilation))
                self.residual_convs.append(nn.Conv2d(dilation_channels, residual_channels, kernel_size=(1, 1)))
                self.skip_convs.append(nn.Conv2d(dilation_channels, skip_channels, kernel_size=(1, 1)))
                self.bn.append(nn.BatchNorm2d(residual_channels))
                self.dilations.append(new_dilation)
                new_dilation *= 2
                receptive_field += additional_scope
                additional_scope *= 2
                if self.gcn_bool:
                    self.gconv.append(gcn(dilation_channels, residual_channels, dropout, support_len=self.supports_len))

        self.end_conv_1 = nn.Conv2d(skip_channels, end_channels, kernel_size=(1, 1), bias=True)
        self.end_conv_2 = nn.Conv2d(end_channels, out_dim, kernel_size=(1, 1), bias=True)
        self.receptive_field = receptive_field

        self._cfg = dict(in_dim=in_dim, out_dim=out_dim, C=residual_channels, D=dilation_channels, S=skip_channels,
                         E=end_channels)
        self._names = None
        self._keep_ws = False         # True: keep a reference to the last forward's workspace in self._last_ws
        self._last_ws = self._last_shape = None
        self.precision = 'fp32'       # 'fp32': FFMA, 1e-5 parity mode; 'bf16': tcgen05 bf16 operands, fp32 accumulate

    def set_precision(self, name):
        assert name in ('fp32', 'bf16')
        self.precision = name
        return self

    # ---- kernel plumbing ---------------------------------------------------------------------
    def _check_supported(self):
        c = self._cfg
        why = None
        if not (self.gcn_bool and self.addaptadj):
            why = 'gcn_bool=True and addaptadj=True are required'
        elif self.supports is None or len(self.supports) != 0:
            why = 'static supports are not implemented (HOP passes supports=None)'
        elif self.kernel_size != 2:
            why = 'kernel_size must be 2'
        elif c['C'] != c['D']:
            why = 'residual_channels must equal dilation_channels'
        elif self.dropout != 0 and self.training:
            why = 'gcn dropout > 0 in training mode is not implemented (HOP uses dropout=0)'
        elif self.blocks * self.layers > _lib.MAX_LAYERS:
            why = 'too many layers'
        elif any(bn.momentum is None or bn.momentum != self.bn[0].momentum or bn.eps != self.bn[0].eps or not bn.affine
                 or not bn.track_running_stats for bn in self.bn):
            why = 'the BatchNorm2d layers must share one momentum (not None) and eps, be affine and track running stats'
        if why:
            raise NotImplementedError('hop_b200.gwnet: ' + why + ' -- no eager fallback by design')

    def _param_names(self):
        """[(field name in HopkGwnetParams, layer index or None)] in the order of :meth:`_param_list`."""
        if self._names is None:
            names = [('nodevec1', None), ('nodevec2', None), ('start_w', None), ('start_b', None)]
            for i in range(self.blocks * self.layers):
                names += [('filter_w', i), ('filter_b', i), ('gate_w', i), ('gate_b', i), ('skip_w', i), ('skip_b', i),
                          ('mlp_w', i), ('mlp_b', i), ('bn_w', i), ('bn_b', i)]
            names += [('end1_w', None), ('end1_b', None), ('end2_w', None), ('end2_b', None)]
            self._names = names
        return self._names

    def _is_dead(self, name, layer):
        # the last layer's gcn + BatchNorm output is never consumed (reference gwnet.py:240 uses only
        # `skip`), so those four tensors get no gradient -- mirrored as grad None
        return layer == self.blocks * self.layers - 1 and name in ('mlp_w', 'mlp_b', 'bn_w', 'bn_b')

    def _param_list(self):
        ps = [self.nodevec1, self.nodevec2, self.start_conv.weight, self.start_conv.bias]
        for i in range(self.blocks * self.layers):
            ps += [self.filter_convs[i].weight, self.filter_convs[i].bias, self.gate_convs[i].weight,
                   self.gate_convs[i].bias, self.skip_convs[i].weight, self.skip_convs[i].bias,
                   self.gconv[i].mlp.mlp.weight, self.gconv[i].mlp.mlp.bias, self.bn[i].weight, self.bn[i].bias]
        ps += [self.end_conv_1.weight, self.end_conv_1.bias, self.end_conv_2.weight, self.end_conv_2.bias]
        return ps

    def _shape(self, x):
        c = self._cfg
        s = GwnetShape()
        s.B, s.V, s.T = x.shape[0], x.shape[2], x.shape[3]
        s.in_dim, s.out_dim, s.C, s.S, s.E = c['in_dim'], c['out_dim'], c['C'], c['S'], c['E']
        s.L = self.blocks * self.layers
        for i, d in enumerate(self.dilations):
            s.dil[i] = d
        s.rank = self.nodevec1.shape[1]
        s.training = 1 if self.training else 0
        s.dtype = 1 if self.precision == 'bf16' else 0
        s.bn_momentum, s.bn_eps = float(self.bn[0].momentum), float(self.bn[0].eps)
        return s

    def workspace_field(self, name, layer=0):
        """fp32 view (rows layout) of a saved activation of the last forward (needs ``self._keep_ws = True``)."""
        import ctypes as C
        off, nbytes, eb = C.c_size_t(), C.c_size_t(), C.c_int()
        check(lib().hopk_gwnet_ws_field(self._last_shape, name.encode(), layer, C.byref(off), C.byref(nbytes), C.byref(eb)))
        raw = self._last_ws[off.value:off.value + nbytes.value]
        return raw.view(torch.float32 if eb.value == 4 else torch.bfloat16)

    def _param_struct(self, params):
        p = GwnetParams()
        keep = []
        for (name, layer), t in zip(self._param_names(), params):
            t = f32c(t)
            keep.append(t)
            if layer is None:
                setattr(p, name, ptr(t))
            else:
                getattr(p, name)[layer] = t.data_ptr()
        for i, bn in enumerate(self.bn):
            p.bn_mean[i] = bn.running_mean.data_ptr()
            p.bn_var[i] = bn.running_var.data_ptr()
            p.bn_nbt[i] = bn.num_batches_tracked.data_ptr()
        return p, keep

    def forward(self, input):
        self._check_supported()
        if input.dim() != 4 or input.shape[1] != self._cfg['in_dim'] or input.shape[2] != self.num_nodes:
            raise ValueError(f'expected (B, {self._cfg["in_dim"]}, {self.num_nodes}, T), got {tuple(input.shape)}')
        if not input.is_cuda:
            raise RuntimeError('hop_b200.gwnet needs CUDA tensors (no CPU fallback)')
        if input.dtype != torch.float32:
            input = input.float()
        return _GwnetFn.apply(self, input, *self._param_list())
