#!/usr/bin/env python
"""Per-layer micro-benchmark of the ResNet-50 convolutions (batch 64, fp16) through the C-ABI, CUDA events.
   columns: the NCHW im2col + GEMM path (it_b200_conv2d_fused), the NHWC implicit-GEMM path (it_b200_conv2d_nhwc), and cuDNN through
   torch (channels_last, what the reference's convCudnn dispatches to, conv only -- no BN / ReLU) when `cudnn` is passed.
   python tools/conv_bench.py [cudnn] [--batch N]"""
import ctypes
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from infinitensor_b200 import _lib as L

# (C, H, F, R, stride, pad, count in ResNet-50, residual)
LAYERS = [
    (64, 56, 64, 1, 1, 0, 1, 0), (64, 56, 64, 3, 1, 1, 3, 0), (64, 56, 256, 1, 1, 0, 4, 1), (256, 56, 64, 1, 1, 0, 2, 0),
    (256, 56, 128, 1, 1, 0, 1, 0), (128, 56, 128, 3, 2, 1, 1, 0), (128, 28, 512, 1, 1, 0, 4, 1), (256, 56, 512, 1, 2, 0, 1, 0),
    (512, 28, 128, 1, 1, 0, 3, 0), (128, 28, 128, 3, 1, 1, 3, 0),
    (512, 28, 256, 1, 1, 0, 1, 0), (256, 28, 256, 3, 2, 1, 1, 0), (256, 14, 1024, 1, 1, 0, 6, 1), (512, 28, 1024, 1, 2, 0, 1, 0),
    (1024, 14, 256, 1, 1, 0, 5, 0), (256, 14, 256, 3, 1, 1, 5, 0),
    (1024, 14, 512, 1, 1, 0, 1, 0), (512, 14, 512, 3, 2, 1, 1, 0), (512, 7, 2048, 1, 1, 0, 3, 1), (1024, 14, 2048, 1, 2, 0, 1, 0),
    (2048, 7, 512, 1, 1, 0, 2, 0), (512, 7, 512, 3, 1, 1, 2, 0),
]


def vp(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else None


def timeit(fn, reps=int(os.environ.get("CONV_BENCH_REPS", "20")), warm=int(os.environ.get("CONV_BENCH_WARM", "3"))):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    e1.synchronize()
    return e0.elapsed_time(e1) * 1e3 / reps


def main():
    B = 64
    if "--batch" in sys.argv:
        B = int(sys.argv[sys.argv.index("--batch") + 1])
    with_cudnn = "cudnn" in sys.argv
    global LAYERS
    if os.environ.get("CONV_BENCH_LAYERS"):  # indices into LAYERS, e.g. "1,2,9,21"
        LAYERS = [LAYERS[int(i)] for i in os.environ["CONV_BENCH_LAYERS"].split(",")]
    st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    dt = 10  # ITB_F16
    tot = {"nchw": 0.0, "nhwc": 0.0, "cudnn": 0.0, "floor": 0.0}
    print(f"{'layer (C,H,F,R,s) x count':>30s} {'GFLOP':>7s} {'MB':>7s} {'nchw us':>9s} {'nhwc us':>9s} {'TF/s':>7s} {'GB/s':>7s}"
          + (f" {'cudnn us':>9s}" if with_cudnn else ""))
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    for (C, H, F, R, s, pad, cnt, has_res) in LAYERS:
        OH = (H + 2 * pad - R) // s + 1
        x = torch.randn(B, C, H, H, device="cuda", dtype=torch.float16)
        xh = x.permute(0, 2, 3, 1).contiguous()
        w = torch.randn(F, C, R, R, device="cuda", dtype=torch.float16) * 0.05
        y = torch.empty(B, F, OH, OH, device="cuda", dtype=torch.float16)
        res = torch.randn(B, F, OH, OH, device="cuda", dtype=torch.float16) if has_res else None
        bn = [torch.rand(F, device="cuda", dtype=torch.float32) + 0.5 for _ in range(4)]
        wsb = max(int(L.lib.it_b200_conv2d_workspace(dt, B, C, H, H, F, R, R, pad, pad, s, s, 1, 1, 1)),
                  int(L.lib.it_b200_conv2d_nhwc_workspace(dt, C, F, R, R)), 16)
        ws = torch.empty(wsb, dtype=torch.uint8, device="cuda")

        def nchw():
            flush.zero_()
            rc = L.lib.it_b200_conv2d_fused(dt, vp(x), vp(w), vp(y), B, C, H, H, F, R, R, pad, pad, s, s, 1, 1, 1, vp(bn[0]), vp(bn[1]),
                                            vp(bn[2]), vp(bn[3]), 1e-5, vp(res), 1, vp(ws), wsb, st)
            assert rc == 0, L.last_error()

        def nhwc():
            flush.zero_()
            L.check(L.lib.it_b200_conv2d_nhwc(dt, vp(xh), vp(w), vp(y), 1, B, C, H, H, F, R, R, pad, pad, s, s, 1, 1, vp(bn[0]),
                                              vp(bn[1]), vp(bn[2]), vp(bn[3]), 1e-5, vp(res), 1, vp(ws), wsb, st))

        def fl():
            flush.zero_()

        t_fl = timeit(fl)
        t0 = timeit(nchw) - t_fl
        t1 = timeit(nhwc) - t_fl
        flop = 2.0 * B * OH * OH * F * C * R * R
        byt = 2.0 * (B * C * H * H + B * F * OH * OH * (2 if has_res else 1) + F * C * R * R)
        floor = max(flop / 1711.7e12, byt / 6567e9) * 1e6
        line = (f"{str((C, H, F, R, s)) + ' x' + str(cnt):>30s} {flop / 1e9:7.2f} {byt / 1e6:7.1f} {t0:9.1f} {t1:9.1f} "
                f"{flop / t1 / 1e6:7.1f} {byt / t1 / 1e3:7.0f}")
        tot["nchw"] += t0 * cnt
        tot["nhwc"] += t1 * cnt
        tot["floor"] += floor * cnt
        if with_cudnn:
            xc = x.to(memory_format=torch.channels_last)
            wc = w.to(memory_format=torch.channels_last)

            def cd():
                flush.zero_()
                torch.nn.functional.conv2d(xc, wc, None, s, pad)

            t2 = timeit(cd) - t_fl
            tot["cudnn"] += t2 * cnt
            line += f" {t2:9.1f}"
        print(line, flush=True)
    print(f"sum over the 52 bottleneck convs (x count): nchw {tot['nchw']:.0f} us, nhwc {tot['nhwc']:.0f} us, roofline floor {tot['floor']:.0f} us"
          + (f", cudnn (conv only) {tot['cudnn']:.0f} us" if with_cudnn else ""))


if __name__ == "__main__":
    main()
