#!/usr/bin/env python
"""Sweep (tile width, split-K) of the skinny decode GEMM per Llama shape through the C-ABI tuning hook.
   Grouped shapes (q/k/v, gate/up) run as the grouped launch the schedule emits."""
import ctypes
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from infinitensor_b200 import _lib as L

# (name, m, k, [n per group])
SHAPES = [("o", 16, 4096, [4096]), ("qkv", 16, 4096, [4096] * 3), ("gate_up", 16, 4096, [11008] * 2),
          ("down", 16, 11008, [4096]), ("logits", 16, 4096, [32000])]
P = lambda t: ctypes.c_void_p(t.data_ptr())


def bench(m, k, ns, ws, x, ys, reps=36):
    st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    nbuf = len(ws)
    ng = len(ns)
    VP = ctypes.c_void_p * ng
    IN = ctypes.c_int * ng

    def go(i):
        if ng == 1:
            L.check(L.lib.it_b200_matmul(16, P(x), P(ws[i % nbuf][0]), None, P(ys[0]), 1, m, ns[0], k, m * k, 0, 0, 0, 0,
                                         0, 0, 0x200, None, 0, st))
        else:
            L.check(L.lib.it_b200_matmul_grouped(16, P(x), ng, VP(*[w.data_ptr() for w in ws[i % nbuf]]),
                                                 VP(*[y.data_ptr() for y in ys]), IN(*ns), m, k, st))
    for i in range(nbuf):
        go(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(reps):
        go(i)
    e1.record()
    e1.synchronize()
    return e0.elapsed_time(e1) * 1e3 / reps


if __name__ == "__main__":
    for name, m, k, ns in SHAPES:
        nbuf = max(3, int(2.4e9 / (k * sum(ns) * 2)))
        ws = [[torch.randn(k, n, device="cuda", dtype=torch.bfloat16) * 0.02 for n in ns] for _ in range(nbuf)]
        x = torch.randn(m, k, device="cuda", dtype=torch.bfloat16)
        ys = [torch.empty(m, n, device="cuda", dtype=torch.bfloat16) for n in ns]
        gb = (k * sum(ns) + m * (k + sum(ns))) * 2 / 1e9
        L.lib.it_b200_tune_skinny(0, 0)
        auto = bench(m, k, ns, ws, x, ys)
        print(f"{name:8s} auto: {auto:7.2f} us {gb / auto * 1e6:6.0f} GB/s", flush=True)
        for nb in (1, 2):
            row = []
            for sk in range(1, 9):
                L.lib.it_b200_tune_skinny(nb, sk)
                try:
                    us = bench(m, k, ns, ws, x, ys)
                    row.append(f"{sk}:{us:6.2f}")
                except Exception as e:  # a configuration the launcher refuses
                    row.append(f"{sk}:  n/a ")
            print(f"{name:8s} nb={nb}  " + "  ".join(row), flush=True)
        del ws
        torch.cuda.empty_cache()
