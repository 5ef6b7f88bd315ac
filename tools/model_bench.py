#!/usr/bin/env python
"""Whole-graph timings of the two non-headline BASELINE configs (parity-test cases, not bench lines):
   C2 GPT-2-small (B=1, S=128, fp16) and C4 ResNet-50 (B=64, fp16), CUDA-graph replay, device-timed.
   python tools/model_bench.py [gpt2] [resnet] [--batch N]"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from infinitensor_b200 import backend as B, graphs as G

_PEAKS_PATH = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")
# driver-written measured peaks; fallback = the figures of /opt/skills/guides/B200_PROFILING.md's pool (copy 6.5 TB/s, bf16 1.7 PF)
PEAKS = json.load(open(_PEAKS_PATH)) if os.path.exists(_PEAKS_PATH) else {"hbm_gbs": 6500.0, "bf16_tflops": 1700.0}


def timed(h, rt, reps, warm=5):
    if os.environ.get("MB_PROFILE"):  # under ncu: two eager passes, no graph replay
        h.run()
        h.run()
        return 1.0, 0
    for _ in range(warm):
        h.run_with_cudagraph()
    h.sync()
    l0 = rt.kernel_launches()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    st = torch.cuda.ExternalStream(rt.stream())
    with torch.cuda.stream(st):
        e0.record()
        for _ in range(reps):
            h.launch_cudagraph_async()
        e1.record()
    e1.synchronize()
    return e0.elapsed_time(e1) / reps, (rt.kernel_launches() - l0) // max(reps, 1)


def gpt2():
    cfg = G.GPT2Config()
    rt = B.CudaRuntime(0)
    h = B.GraphHandler(rt)
    g = G.build_gpt2(h, cfg)
    h.data_malloc()
    G.fill_gpt2_weights_host(g)
    ids = np.random.default_rng(1).integers(0, cfg.vocab, size=(cfg.batch, cfg.seq)).astype(np.int64)
    pos = np.arange(cfg.seq, dtype=np.int64).reshape(1, -1).repeat(cfg.batch, 0)
    g.input_ids.copyin_numpy(ids)
    g.position_ids.copyin_numpy(pos)
    ms, _ = timed(h, rt, 50)
    steps = h.schedule()
    wbytes = (12 * 7087872 + 50257 * 768 + 1024 * 768) * 2
    print(json.dumps({"config": "C2 gpt2-small B=1 S=128 fp16 (no LM head, as the reference exports it)", "ms_per_forward": round(ms, 4),
                      "tokens_per_s": round(cfg.batch * cfg.seq / ms * 1e3, 1), "scheduled_steps": len(steps),
                      "weight_bytes": wbytes, "hbm_floor_ms": round(wbytes / PEAKS["hbm_gbs"] / 1e6, 4)}))


def resnet(batch):
    cfg = G.ResNetConfig(batch=batch)
    rt = B.CudaRuntime(0)
    h = B.GraphHandler(rt)
    g = G.build_resnet50(h, cfg)
    h.data_malloc()
    G.fill_resnet_weights_host(g)
    x = np.random.default_rng(3).standard_normal((cfg.batch, 3, cfg.image, cfg.image)).astype(np.float32)
    g.input.copyin_numpy(G.to_storage(x, cfg.dtype))
    ms, _ = timed(h, rt, 10, warm=3)
    flop = 2 * 4.09e9 * batch
    tf = flop / ms / 1e9
    out = G.from_storage(g.out.copyout_numpy(), cfg.dtype)
    print(json.dumps({"config": f"C4 resnet50 B={batch} fp16", "ms_per_batch": round(ms, 3),
                      "images_per_s": round(batch / ms * 1e3, 1), "tflops": round(tf, 1),
                      "frac_of_measured_dense_peak": round(tf / PEAKS["bf16_tflops"], 4),
                      "scheduled_steps": len(h.schedule()), "finite": bool(np.isfinite(out).all())}))


if __name__ == "__main__":
    args = [a for a in sys.argv[1:] if not a.startswith("--")] or ["gpt2", "resnet"]
    batch = 64
    if "--batch" in sys.argv:
        batch = int(sys.argv[sys.argv.index("--batch") + 1])
    if "gpt2" in args:
        gpt2()
    if "resnet" in args:
        resnet(batch)
