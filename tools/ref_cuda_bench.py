#!/usr/bin/env python
"""Same-box timing of the REFERENCE's CUDA kernels (oracle/_ref/libit_ref_cuda.so, PTX/SASS built for sm_100a from the
unmodified .cu files) against this repo's kernels, CUDA events, 50 launches each after 5 warm-ups.  Test infrastructure."""
import ctypes
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from infinitensor_b200 import _lib as L
from oracle import ref_cuda

R = ref_cuda.lib()
P = lambda t: ctypes.c_void_p(t.data_ptr())
st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def timed(fn, reps=50):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    e1.synchronize()
    return e0.elapsed_time(e1) * 1e3 / reps


rows = []
# RMSNorm fp16 [16, 4096]
x = torch.randn(16, 4096, device="cuda", dtype=torch.float16); w = torch.ones(4096, device="cuda", dtype=torch.float16); y = torch.empty_like(x)
rows.append(("RMSNorm f16 [16,4096]", timed(lambda: R.ref_cuda_rmsnorm(10, P(x), P(w), P(y), 16, 4096)),
             timed(lambda: L.lib.it_b200_rmsnorm(10, P(x), P(w), P(y), 16, 4096, st))))
# AttentionKVCache fp32 B=16 H=32 S=1024 p=511 (the reference kernel is fp32-only; rotating caches >> L2)
B, H, S, D, pos = 16, 32, 1024, 128, 511
caches = [(torch.randn(B, H, S, D, device="cuda") * 0.5, torch.randn(B, H, S, D, device="cuda") * 0.5) for _ in range(3)]
q, k, v = [torch.randn(B, H, 1, D, device="cuda") * 0.5 for _ in range(3)]
o = torch.empty_like(q)
p32 = torch.tensor([pos], dtype=torch.int32, device="cuda")
tmp_o = torch.empty(B * H * (S // 8 + 8) * D, device="cuda"); tmp_s = torch.empty(B * H * (S // 8 + 8), device="cuda")
wsb = L.lib.it_b200_attention_kvcache_workspace(B, H, S, D)
ws = torch.empty(int(wsb), dtype=torch.uint8, device="cuda")
i = [0]
def ref_att():
    kc, vc = caches[i[0] % 3]; i[0] += 1
    R.ref_cuda_attention_kvcache(P(kc), P(vc), P(q), P(k), P(v), P(p32), P(o), B, H, S, D, P(tmp_o), P(tmp_s))
def our_att():
    kc, vc = caches[i[0] % 3]; i[0] += 1
    L.lib.it_b200_attention_kvcache(1, P(kc), P(vc), P(q), P(k), P(v), P(p32), 6, P(o), B, H, S, D, P(ws), ctypes.c_int64(int(wsb)), st)
rows.append(("AttentionKVCache f32 B16 H32 S1024 p511 (268 MB)", timed(ref_att, 20), timed(our_att, 20)))
# Softmax fp16 [1,12,128,128] last axis
x = torch.randn(1, 12, 128, 128, device="cuda", dtype=torch.float16); y = torch.empty_like(x)
rows.append(("Softmax f16 [1,12,128,128] axis -1", timed(lambda: R.ref_cuda_softmax_f16(P(x), P(y), x.numel(), 128, 1)),
             timed(lambda: L.lib.it_b200_softmax(10, P(x), P(y), 12 * 128, 128, 1, st))))
# LayerNorm fp32 [1,128,768]
x = torch.randn(1, 128, 768, device="cuda"); s = torch.ones(768, device="cuda"); b = torch.zeros(768, device="cuda"); y = torch.empty_like(x)
rows.append(("LayerNorm f32 [1,128,768]", timed(lambda: R.ref_cuda_layernorm_f32(P(x), P(s), 1e-5, x.numel(), 768, 768, 1, P(y), P(b), 768)),
             timed(lambda: L.lib.it_b200_layernorm(1, P(x), P(s), P(b), P(y), 128, 768, 1, 768, 768, ctypes.c_float(1e-5), st))))
print("| op | reference CUDA kernel (us) | this repo (us) | speed-up |\n|---|---|---|---|")
for name, tr, to in rows:
    print(f"| {name} | {tr:.2f} | {to:.2f} | {tr / to:.2f}x |")
