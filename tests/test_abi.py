"""CPU: the C-ABI library builds/loads and exports every symbol include/hopk.h declares; the host-side
mirror keeps the reference's state_dict keys.  No kernel is launched here."""
import ctypes
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, 'include', 'hopk.h')).read()
    src = re.sub(r'/\*.*?\*/', '', src, flags=re.S)
    src = re.sub(r'#ifdef HOPK_DEBUG.*?#endif', '', src, flags=re.S)     # debug-only entry points are not in the release library
    return sorted(set(re.findall(r'\b(hopk_\w+)\s*\(', src)))


def test_library_exports_every_declared_symbol():
    from hop_b200 import _lib
    from hop_b200.build import build
    build()
    l = ctypes.CDLL(_lib.LIB_PATH)
    names = declared_symbols()
    assert len(names) >= 14
    for n in names:
        assert hasattr(l, n), f'{n} declared in include/hopk.h but not exported by libhopk.so'
        assert n in _lib.SIGNATURES, f'{n} has no ctypes signature in hop_b200/_lib.py'
    assert set(_lib.SIGNATURES) == set(names)
    assert _lib.lib().hopk_version() >= 200
    assert not hasattr(l, 'hopk_debug_set'), 'debug hooks must be compiled out of the release library'


def test_struct_sizes_match_header():
    from hop_b200 import _lib
    assert ctypes.sizeof(_lib.GwnetShape) == 4 * (9 + 16 + 3 + 2)
    assert ctypes.sizeof(_lib.GwnetParams) == 8 * (4 + 13 * 16 + 4)
    assert ctypes.sizeof(_lib.GwnetGrads) == 8 * (4 + 10 * 16 + 4 + 2)


def test_workspace_queries_run_on_host():
    from hop_b200 import _lib
    s = _lib.GwnetShape()
    s.B, s.V, s.T, s.in_dim, s.out_dim, s.C, s.S, s.E, s.L = 128, 9, 16, 173, 173, 64, 256, 512, 8
    for i, d in enumerate([1, 2] * 4):
        s.dil[i] = d
    s.rank = 10
    assert _lib.lib().hopk_gwnet_out_steps(s) == 4
    assert _lib.lib().hopk_gwnet_workspace_bytes(s) > 128 * 9 * 16 * 64 * 4
    s.T = 10                                     # shorter than the receptive field: left-padded to 13 -> 1 step
    assert _lib.lib().hopk_gwnet_out_steps(s) == 1


def test_cpu_tensors_are_refused():
    from hop_b200 import gwnet as G
    m = G.gwnet('cpu', 5, dropout=0, in_dim=4, out_dim=4)
    with pytest.raises(RuntimeError):
        m(torch.randn(2, 4, 5, 16))


GWNET_KEYS_PER_LAYER = ['filter_convs.{i}.weight', 'filter_convs.{i}.bias', 'gate_convs.{i}.weight', 'gate_convs.{i}.bias',
                        'residual_convs.{i}.weight', 'residual_convs.{i}.bias', 'skip_convs.{i}.weight',
                        'skip_convs.{i}.bias', 'bn.{i}.weight', 'bn.{i}.bias', 'bn.{i}.running_mean', 'bn.{i}.running_var',
                        'bn.{i}.num_batches_tracked', 'gconv.{i}.mlp.mlp.weight', 'gconv.{i}.mlp.mlp.bias']


def test_gwnet_state_dict_keys_and_shapes():
    """128 keys with the reference's names and shapes (incl. the dead residual_convs)."""
    from hop_b200 import gwnet as G
    m = G.gwnet('cpu', 9, dropout=0, in_dim=173, out_dim=173, residual_channels=64, dilation_channels=64,
                skip_channels=256, end_channels=512)
    sd = m.state_dict()
    want = {'nodevec1', 'nodevec2', 'start_conv.weight', 'start_conv.bias', 'end_conv_1.weight', 'end_conv_1.bias',
            'end_conv_2.weight', 'end_conv_2.bias'}
    for i in range(8):
        want |= {k.format(i=i) for k in GWNET_KEYS_PER_LAYER}
    assert set(sd) == want and len(sd) == 128
    assert sum(v.numel() for v in sd.values()) == 631017
    assert sd['filter_convs.3.weight'].shape == (64, 64, 1, 2)
    assert sd['gconv.0.mlp.mlp.weight'].shape == (64, 192, 1, 1)
    assert m.receptive_field == 13 and m.supports == [] and m.supports_len == 1


def test_reprogramming_state_dict():
    from hop_b200.HOP import ReprogrammingLayer
    m = ReprogrammingLayer(128, 8, 128, 768)
    sd = m.state_dict()
    assert set(sd) == {f'{n}_projection.{p}' for n in ('query', 'key', 'value', 'out') for p in ('weight', 'bias')}
    assert sum(v.numel() for v in sd.values()) == 2494208
