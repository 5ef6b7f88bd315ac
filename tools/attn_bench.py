#!/usr/bin/env python
"""Micro-benchmark of the decode attention kernel (with / without folded RoPE) through the C-ABI.
usage: attn_bench.py [path/to/lib.so ...]   (default: the in-tree libit_b200.so)"""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
B, H, S, D, pos = 16, 32, 1024, 128, 511
nbuf = 6
kcs = [torch.randn(B, H, S, D, device="cuda", dtype=torch.bfloat16) * 0.5 for _ in range(nbuf)]
vcs = [torch.randn(B, H, S, D, device="cuda", dtype=torch.bfloat16) * 0.5 for _ in range(nbuf)]
q, k, v = [torch.randn(B, H, 1, D, device="cuda", dtype=torch.bfloat16) * 0.5 for _ in range(3)]
out = torch.empty_like(q)
p = torch.full((B, 1), pos, dtype=torch.int64, device="cuda")
ws = torch.empty(64 << 20, dtype=torch.uint8, device="cuda")
st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
P = lambda t: ctypes.c_void_p(t.data_ptr())
paths = sys.argv[1:]
if not paths:
    from infinitensor_b200 import _lib as L
    paths = [L.LIB_PATH]
for path in paths:
    lib = ctypes.CDLL(path)
    def plain(i):
        assert lib.it_b200_attention_kvcache(16, P(kcs[i % nbuf]), P(vcs[i % nbuf]), P(q), P(k), P(v), P(p), 7, P(out), B, H, S, D, P(ws), ctypes.c_int64(ws.numel()), st) == 0
    def rope(i):
        assert lib.it_b200_attention_kvcache_rope(16, P(kcs[i % nbuf]), P(vcs[i % nbuf]), P(q), P(k), P(v), P(p), 7, P(p), 7, P(out), B, H, S, D, P(ws), ctypes.c_int64(ws.numel()), st) == 0
    fns = [("plain", plain)] + ([("rope", rope)] if hasattr(lib, "it_b200_attention_kvcache_rope") else [])
    for rep in range(2):
        for name, fn in fns:
            for i in range(nbuf): fn(i)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for i in range(60): fn(i)
            e1.record(); e1.synchronize()
            us = e0.elapsed_time(e1) * 1e3 / 60
            gb = B * H * 2 * (pos + 1) * D * 2 / 1e9
            print(f"{os.path.basename(path)} {name}: {us:.2f} us  {gb / (us * 1e-6):.0f} GB/s", flush=True)
