#!/usr/bin/env python
"""One fp8-weight decode GEMM (gate/up shape: 16 x 4096 x 2 x 11008) and its bf16 twin, timed with CUDA events over rotating
weight buffers (>> L2); also the target of the ncu capture in profiles/."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from infinitensor_b200 import _lib as L

M, K, N, G, R = 16, 4096, 11008, 2, 6
x = torch.randn(M, K, device="cuda").bfloat16()
# codes of max-abs-scaled Gaussian weights (the realistic mix: ~1e-4 of them below 2^-6, the kernel's exact slow path)
def qweights():
    w = torch.randn(K, N, device="cuda")
    return (w / (w.abs().amax(dim=0) / 448.0)[None, :]).to(torch.float8_e4m3fn).view(torch.uint8).contiguous()
wq = [[qweights() for _ in range(G)] for _ in range(R)]
wb = [[torch.randn(K, N, device="cuda").bfloat16() * 0.02 for _ in range(G)] for _ in range(R)]
sc = [torch.rand(N, device="cuda") * 0.01 for _ in range(G)]
out = [torch.zeros(M, N, device="cuda").bfloat16() for _ in range(G)]
VP = ctypes.c_void_p
arr = lambda ts: (VP * len(ts))(*[t.data_ptr() for t in ts])
st = VP(torch.cuda.current_stream().cuda_stream)
Ns = L.i32arr([N] * G)

def fp8(i):
    L.check(L.lib.it_b200_matmul_fp8w(16, VP(x.data_ptr()), G, arr(wq[i % R]), arr(sc), arr(out), Ns, M, K, None, st))
def bf16(i):
    L.check(L.lib.it_b200_matmul_grouped(16, VP(x.data_ptr()), G, arr(wb[i % R]), arr(out), Ns, M, K, st))
cases = [("bf16", bf16, 2 * G * K * N, 0, 0)] + [(f"fp8 nb={nb} splitk={sk}", fp8, G * K * N, nb, sk) for nb in (1, 2) for sk in (0, 1, 2, 3, 4)]
for name, fn, bytes_, nb, sk in cases:
    L.lib.it_b200_tune_skinny(nb, sk)
    for i in range(3):
        fn(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(24):
        fn(i)
    e1.record(); e1.synchronize()
    us = e0.elapsed_time(e1) / 24 * 1e3
    print(f"{name}: {us:.1f} us per launch, {bytes_ / us / 1e3:.0f} GB/s of weight bytes")
