#!/usr/bin/env python
"""Phase timeline of the persistent decode kernel on the BASELINE shape: per phase kind, how long the CTAs compute
(phase start -> own barrier arrival), how long the slowest takes, and the barrier gap to the next phase start."""
import ctypes
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch


def main():
    layers = int(sys.argv[1]) if len(sys.argv) > 1 else 8
    from infinitensor_b200 import backend as B, graphs as G, _lib as L
    cfg = G.LlamaConfig(layers=layers)
    rt = B.CudaRuntime(0)
    h = B.GraphHandler(rt)
    g = G.build_llama_decode(h, cfg)
    h.data_malloc()
    ts = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    gen = torch.Generator(device="cuda")
    for i, (name, (t, shape, kind, shard)) in enumerate(g.weights.items()):
        gen.manual_seed(i)
        tmp = torch.empty(t.shape(), dtype=torch.bfloat16, device="cuda").normal_(1.0 if kind == "norm" else 0.0, 0.02, generator=gen)
        L.check(L.lib.it_b200_copy(ctypes.c_void_p(tmp.data_ptr()), ctypes.c_void_p(t.device_ptr()), t.nbytes(), ts))
        torch.cuda.synchronize()
    g.input_ids.copyin_numpy(np.arange(cfg.batch, dtype=np.int64).reshape(-1, 1))
    g.position_ids.copyin_numpy(np.full((cfg.batch, 1), 511, np.int64))
    for _ in range(3):
        h.run()
    nph = layers * 5
    flat = torch.zeros(148 * nph * 2 + 148 * 16 + 64 * 8, dtype=torch.int64, device="cuda")
    buf = flat[:148 * nph * 2].view(148, nph, 2)
    L.lib.it_b200_decode_stack_trace(ctypes.c_void_p(buf.data_ptr()))
    h.run()
    L.lib.it_b200_decode_stack_trace(None)
    roles = flat[148 * nph * 2:148 * nph * 2 + 148 * 16].view(148, 8, 2).cpu().numpy().astype(np.float64)
    probe = flat[148 * nph * 2 + 148 * 16:].view(64, 8).cpu().numpy().astype(np.float64)
    if probe[:, 1].max() > 0:
        base = probe[0, 0]
        print("chunk probe, CTA 0, layer-1 q/k/v phase [us since the first]:  wait_empty  tma_issued  full  xform_done  mma_ready  mma_issued  x_issued")
        for i in list(range(0, 16)) + list(range(30, 42)):
            print(f"  chunk {i:2d}: " + "  ".join(f"{(probe[i, k] - base) / 1e3:9.2f}" for k in (0, 1, 2, 3, 4, 5, 6)))
    for ri, nm in ((0, "W producer"), (2, "MMA issuer"), (3, "epilogue/consumer thread 0"), (4, "transform/consumer warp 4")):
        w, tot = roles[:, ri, 0], roles[:, ri, 1]
        print(f"role {nm:28s} blocked on mbarriers {100 * np.median(w / np.maximum(tot, 1)):5.1f} % of {np.median(tot)/1.965e3:8.1f} us (median CTA)")
    t = buf.cpu().numpy().astype(np.float64)
    t0 = t[:, 0, 0].min()
    start, arrive = t[:, :, 0] - t0, t[:, :, 1] - t0
    names = ["qkv", "attn", "o", "gate_up", "down"]
    print(f"layers {layers}: kernel span {(np.nanmax(arrive[:, :-1]) )/1e3:.1f} us (to the last barrier arrival)")
    print("phase      start->arrive median / max [us]   arrive(max)->next start(max) [us]   bytes/phase [MB]   GB/s over (next_start - start)")
    d, f = cfg.d_model, cfg.ffn
    byts = {"qkv": 3 * d * d * 2, "attn": 2 * 16 * 32 * 512 * 128 * 2, "o": d * d * 2, "gate_up": 2 * d * f * 2, "down": d * f * 2}
    for k, nm in enumerate(names):
        cm, cx, gap, tot = [], [], [], []
        for li in range(1, layers):  # skip the first layer (cold)
            p = li * 5 + k
            if p + 1 >= nph:
                continue
            cm.append(np.median(arrive[:, p] - start[:, p])); cx.append((arrive[:, p] - start[:, p]).max())
            gap.append(start[:, p + 1].max() - arrive[:, p].max())
            tot.append(start[:, p + 1].max() - start[:, p].max())
        if cm:
            print(f"{nm:8s}   {np.mean(cm)/1e3:8.2f} / {np.mean(cx)/1e3:8.2f}             {np.mean(gap)/1e3:8.2f}                     {byts[nm]/1e6:8.1f}          {byts[nm]/np.mean(tot):8.1f}")
    per_layer = (start[:, 5 * (layers - 1)].max() - start[:, 5].max()) / max(1, layers - 2)
    print(f"per layer: {per_layer/1e3:.1f} us  -> 32 layers {per_layer*32/1e6:.3f} ms")


if __name__ == "__main__":
    main()
