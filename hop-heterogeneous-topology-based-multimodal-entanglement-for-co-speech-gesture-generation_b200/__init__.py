"""hop_b200 -- B200-native (sm_100a) implementation of HOP's training hot path.

Mirrors the reference's module interface for that path only:
    hop_b200.gwnet   <->  model/gwnet.py   (nconv, linear, gcn, gwnet)
    hop_b200.HOP     <->  model/HOP.py     (ReprogrammingLayer, Model)
    hop_b200.train_llm <-> train_eval/train_llm.py (the training step that calls them)
Kernels: csrc/*.cu behind the C ABI in include/hopk.h (libhopk.so, loaded by _lib.py).
"""
__version__ = '0.1.0'
