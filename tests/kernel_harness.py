"""Test-side plumbing: torch owns device memory, ctypes calls the C-ABI launchers.

Nothing here computes: every function moves host arrays to the device, calls one
it_b200_* launcher from include/it_b200.h on raw pointers, and brings the result back.
"""
import ctypes

import numpy as np
import torch

from infinitensor_b200 import _lib as L

F32, F16, BF16 = 1, 10, 16
TORCH_DT = {F32: torch.float32, F16: torch.float16, BF16: torch.bfloat16}
UNARY = {"relu": 0, "sigmoid": 1, "tanh": 2, "gelu": 3, "silu": 4, "erf": 5, "neg": 6, "abs": 7, "sqrt": 8,
         "hardsigmoid": 9, "hardswish": 10, "exp": 11}
BINARY = {"add": 0, "sub": 1, "mul": 2, "div": 3, "pow": 4, "min": 5, "max": 6, "less": 7, "equal": 8, "greater": 9}


def dev(a, dt=F32):
    """float32 numpy -> cuda tensor stored as dt."""
    return torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32)).cuda().to(TORCH_DT[dt]).contiguous()


def raw(a):
    """any numpy array -> cuda tensor of the same dtype (bit-exact movers)."""
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def host(t):
    return t.float().cpu().numpy() if t.dtype in (torch.float16, torch.bfloat16) else t.cpu().numpy()


def ptr(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else None


def stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def sync():
    torch.cuda.synchronize()


def bstrides(shape, out_shape):
    r = len(out_shape)
    shape = (1,) * (r - len(shape)) + tuple(shape)
    st, acc = [], 1
    for d in reversed(shape):
        st.append(acc)
        acc *= d
    st = list(reversed(st))
    return [0 if shape[i] == 1 and out_shape[i] != 1 else st[i] for i in range(r)]


def unary(name, x, dt=F32):
    xd = dev(x, dt)
    y = torch.empty_like(xd)
    L.check(L.lib.it_b200_unary(UNARY[name], dt, ptr(xd), ptr(y), xd.numel(), stream()))
    sync()
    return host(y)


def binary(name, a, b, dt=F32):
    ad, bd = dev(a, dt), dev(b, dt)
    out_shape = tuple(np.broadcast_shapes(a.shape, b.shape))
    cmp_ = BINARY[name] >= 7
    c = torch.empty(out_shape, dtype=torch.uint8 if cmp_ else TORCH_DT[dt], device="cuda")
    dims = list(out_shape) or [1]
    L.check(L.lib.it_b200_binary(BINARY[name], dt, ptr(ad), ptr(bd), ptr(c), len(dims), L.i64arr(dims),
                                 L.i64arr(bstrides(a.shape, tuple(dims))), L.i64arr(bstrides(b.shape, tuple(dims))),
                                 stream()))
    sync()
    return host(c)


def softmax(x, axis, dt=F32):
    xd = dev(x, dt)
    y = torch.empty_like(xd)
    axis %= x.ndim
    outer = int(np.prod(x.shape[:axis])) if axis else 1
    inner = int(np.prod(x.shape[axis + 1:])) if axis + 1 < x.ndim else 1
    L.check(L.lib.it_b200_softmax(dt, ptr(xd), ptr(y), outer, x.shape[axis], inner, stream()))
    sync()
    return host(y)


def layer_norm(x, scale, bias, eps, axis, dt=F32):
    xd, sd = dev(x, dt), dev(scale, dt)
    bd = dev(bias, dt) if bias is not None else None
    y = torch.empty_like(xd)
    axis %= x.ndim
    outer = int(np.prod(x.shape[:axis])) if axis else 1
    inner = int(np.prod(x.shape[axis + 1:])) if axis + 1 < x.ndim else 1
    L.check(L.lib.it_b200_layernorm(dt, ptr(xd), ptr(sd), ptr(bd), ptr(y), outer, x.shape[axis], inner,
                                    sd.numel(), bd.numel() if bd is not None else 0, eps, stream()))
    sync()
    return host(y)


def rms_norm(x, w, dt=F32, const_w=False):
    xd, wd = dev(x, dt), dev(w, dt)
    y = torch.empty_like(xd)
    fn = L.lib.it_b200_rmsnorm_constw if const_w else L.lib.it_b200_rmsnorm
    L.check(fn(dt, ptr(xd), ptr(wd), ptr(y), xd.numel() // x.shape[-1], x.shape[-1], stream()))
    sync()
    return host(y)


def rope(pos, x, dt=F32, pos_np_dtype=np.int64):
    xd = dev(x, dt)
    pd = raw(np.asarray(pos).astype(pos_np_dtype))
    y = torch.empty_like(xd)
    code = {np.int64: 7, np.int32: 6, np.uint32: 12}[pos_np_dtype]
    B, S, dm = x.shape
    L.check(L.lib.it_b200_rope(dt, ptr(pd), code, ptr(xd), ptr(y), B, S, dm, 128, stream()))
    sync()
    return host(y)


def transpose(x, perm):
    xd = raw(x)
    y = torch.empty([x.shape[p] for p in perm], dtype=xd.dtype, device="cuda")
    L.check(L.lib.it_b200_transpose(x.dtype.itemsize, ptr(xd), ptr(y), x.ndim, L.i64arr(x.shape), L.i32arr(perm),
                                    stream()))
    sync()
    return y.cpu().numpy()


def concat(xs, dim):
    ds = [raw(x) for x in xs]
    shape = list(xs[0].shape)
    shape[dim] = sum(x.shape[dim] for x in xs)
    out = torch.empty(shape, dtype=ds[0].dtype, device="cuda")
    outer = int(np.prod(shape[:dim])) if dim else 1
    inner = int(np.prod(shape[dim + 1:])) if dim + 1 < len(shape) else 1
    parts = (ctypes.c_void_p * len(ds))(*[d.data_ptr() for d in ds])
    L.check(L.lib.it_b200_concat(xs[0].dtype.itemsize, len(ds), parts, L.i64arr([x.shape[dim] for x in xs]), ptr(out),
                                 outer, inner, stream()))
    sync()
    return out.cpu().numpy()


def split(x, dim, sizes):
    xd = raw(x)
    outs = []
    for s in sizes:
        shp = list(x.shape)
        shp[dim] = s
        outs.append(torch.empty(shp, dtype=xd.dtype, device="cuda"))
    outer = int(np.prod(x.shape[:dim])) if dim else 1
    inner = int(np.prod(x.shape[dim + 1:])) if dim + 1 < x.ndim else 1
    parts = (ctypes.c_void_p * len(outs))(*[d.data_ptr() for d in outs])
    L.check(L.lib.it_b200_split(x.dtype.itemsize, len(outs), parts, L.i64arr(sizes), ptr(xd), outer, inner, stream()))
    sync()
    return [o.cpu().numpy() for o in outs]


def gather(x, idx, axis):
    xd, idd = raw(x), raw(idx)
    axis %= x.ndim
    shp = list(x.shape[:axis]) + list(idx.shape) + list(x.shape[axis + 1:])
    out = torch.empty(shp, dtype=xd.dtype, device="cuda")
    outer = int(np.prod(x.shape[:axis])) if axis else 1
    inner = int(np.prod(x.shape[axis + 1:])) if axis + 1 < x.ndim else 1
    code = 7 if idx.dtype == np.int64 else 6
    L.check(L.lib.it_b200_gather(x.dtype.itemsize, code, ptr(xd), ptr(idd), ptr(out), outer, x.shape[axis], inner,
                                 idx.size, stream()))
    sync()
    return out.cpu().numpy()


def where(c, x, y):
    cd, xd, yd = raw(c.astype(np.uint8)), raw(x), raw(y)
    shp = tuple(np.broadcast_shapes(c.shape, x.shape, y.shape))
    out = torch.empty(shp, dtype=xd.dtype, device="cuda")
    L.check(L.lib.it_b200_where(x.dtype.itemsize, ptr(cd), ptr(xd), ptr(yd), ptr(out), len(shp), L.i64arr(shp),
                                L.i64arr(bstrides(c.shape, shp)), L.i64arr(bstrides(x.shape, shp)),
                                L.i64arr(bstrides(y.shape, shp)), stream()))
    sync()
    return out.cpu().numpy()


def expand(x, dims):
    xd = raw(x)
    shp = tuple(np.broadcast_shapes(x.shape, tuple(dims)))
    out = torch.empty(shp, dtype=xd.dtype, device="cuda")
    L.check(L.lib.it_b200_expand(x.dtype.itemsize, ptr(xd), ptr(out), len(shp), L.i64arr(shp),
                                 L.i64arr(bstrides(x.shape, shp)), stream()))
    sync()
    return out.cpu().numpy()


def pad_slice(x, out_shape, start, step):
    xd = raw(x)
    out = torch.empty(tuple(out_shape), dtype=xd.dtype, device="cuda")
    L.check(L.lib.it_b200_pad_slice(x.dtype.itemsize, ptr(xd), ptr(out), x.ndim, L.i64arr(x.shape),
                                    L.i64arr(out_shape), L.i64arr(start), L.i64arr(step), stream()))
    sync()
    return out.cpu().numpy()


def reduce(kind, x, axes, keep, dt=F32):
    xd = dev(x, dt)
    axes = list(range(x.ndim)) if axes is None else [a % x.ndim for a in axes]
    mask = [1 if i in axes else 0 for i in range(x.ndim)]
    oshape = [1 if m else d for d, m in zip(x.shape, mask)] if keep else [d for d, m in zip(x.shape, mask) if not m]
    out = torch.empty(oshape or [1], dtype=xd.dtype, device="cuda")
    L.check(L.lib.it_b200_reduce(dt, 1 if kind == "mean" else 0, ptr(xd), ptr(out), x.ndim, L.i64arr(x.shape),
                                 L.i32arr(mask), stream()))
    sync()
    return host(out)


def pool2d(kind, x, kh, kw, dh, dw, ph, pw, sh, sw, dt=F32):
    xd = dev(x, dt)
    N, C, H, W = x.shape
    OH = (H + 2 * ph - dh * (kh - 1) - 1) // sh + 1
    OW = (W + 2 * pw - dw * (kw - 1) - 1) // sw + 1
    out = torch.empty((N, C, OH, OW), dtype=xd.dtype, device="cuda")
    L.check(L.lib.it_b200_pool2d(dt, 1 if kind == "max" else 0, ptr(xd), ptr(out), N, C, H, W, kh, kw, dh, dw, ph, pw,
                                 sh, sw, OH, OW, stream()))
    sync()
    return host(out)


def pool2d_nhwc(kind, x, kh, kw, dh, dw, ph, pw, sh, sw, dt=F16):
    """x and the result are NCHW numpy arrays; permuted to / from NHWC on the host."""
    N, C, H, W = x.shape
    xd = dev(np.ascontiguousarray(x.transpose(0, 2, 3, 1)), dt)
    OH = (H + 2 * ph - dh * (kh - 1) - 1) // sh + 1
    OW = (W + 2 * pw - dw * (kw - 1) - 1) // sw + 1
    out = torch.empty((N, OH, OW, C), dtype=xd.dtype, device="cuda")
    L.check(L.lib.it_b200_pool2d_nhwc(dt, 1 if kind == "max" else 0, ptr(xd), ptr(out), N, C, H, W, kh, kw, dh, dw, ph, pw,
                                      sh, sw, OH, OW, stream()))
    sync()
    return np.ascontiguousarray(host(out).transpose(0, 3, 1, 2))


def batch_norm(x, mean, var, scale, bias, eps, dt=F32):
    xd = dev(x, dt)
    ms = [dev(np.asarray(v, np.float32)) for v in (mean, var, scale, bias)]
    out = torch.empty_like(xd)
    N, C = x.shape[:2]
    HW = int(np.prod(x.shape[2:])) if x.ndim > 2 else 1
    L.check(L.lib.it_b200_batchnorm(dt, ptr(xd), ptr(ms[0]), ptr(ms[1]), ptr(ms[2]), ptr(ms[3]), ptr(out), N, C, HW,
                                    eps, stream()))
    sync()
    return host(out)


def batch_norm_relu(x, mean, var, scale, bias, eps, dt=F32):
    xd = dev(x, dt)
    ms = [dev(np.asarray(v, np.float32)) for v in (mean, var, scale, bias)]
    out = torch.empty_like(xd)
    N, C = x.shape[:2]
    HW = int(np.prod(x.shape[2:])) if x.ndim > 2 else 1
    L.check(L.lib.it_b200_batchnorm_relu(dt, ptr(xd), ptr(ms[0]), ptr(ms[1]), ptr(ms[2]), ptr(ms[3]), ptr(out), N, C, HW,
                                         eps, stream()))
    sync()
    return host(out)


def conv2d_fused(x, w, ph, pw, sh, sw, dh, dw, bn, eps, residual, relu, dt=F16, y_nhwc=False):
    """Returns the fused result, or None when the C-ABI reports the shape as not taken (rc 2).  y_nhwc: the kernel writes y (and
    reads the residual) as [N, OH, OW, F]; arguments and result here stay NCHW (permuted on the host)."""
    xd, wd = dev(x, dt), dev(w, dt)
    N, C, H, W = x.shape
    F, Cg, R, S = w.shape
    groups = C // Cg
    OH = (H + 2 * ph - dh * (R - 1) - 1) // sh + 1
    OW = (W + 2 * pw - dw * (S - 1) - 1) // sw + 1
    y = torch.empty((N, OH, OW, F) if y_nhwc else (N, F, OH, OW), dtype=xd.dtype, device="cuda")
    ms = [dev(np.asarray(v, np.float32)) for v in bn] if bn is not None else [None] * 4
    rd = None
    if residual is not None:
        rd = dev(np.ascontiguousarray(residual.transpose(0, 2, 3, 1)) if y_nhwc else residual, dt)
    wsb = L.lib.it_b200_conv2d_workspace(dt, N, C, H, W, F, R, S, ph, pw, sh, sw, dh, dw, groups)
    ws = torch.empty(max(int(wsb), 16), dtype=torch.uint8, device="cuda")
    fn = L.lib.it_b200_conv2d_fused_nhwc_out if y_nhwc else L.lib.it_b200_conv2d_fused
    rc = fn(dt, ptr(xd), ptr(wd), ptr(y), N, C, H, W, F, R, S, ph, pw, sh, sw, dh, dw, groups,
                                    ptr(ms[0]), ptr(ms[1]), ptr(ms[2]), ptr(ms[3]), eps, ptr(rd), int(relu), ptr(ws),
                                    int(wsb), stream())
    if rc == 2:
        return None
    L.check(rc)
    sync()
    return np.ascontiguousarray(host(y).transpose(0, 3, 1, 2)) if y_nhwc else host(y)


def conv2d_stem(x, w, ph, pw, sh, sw, bn=None, eps=1e-5, relu=False, dt=F16):
    """NCHW x (<= 4 channels) -> NHWC y through the stem kernel; the result is permuted back to NCHW on the host."""
    N, C, H, W = x.shape
    F, _, R, S = w.shape
    OH, OW = (H + 2 * ph - R) // sh + 1, (W + 2 * pw - S) // sw + 1
    xd, wd = dev(x, dt), dev(w, dt)
    y = torch.empty((N, OH, OW, F), dtype=xd.dtype, device="cuda")
    ms = [dev(np.asarray(v, np.float32)) for v in bn] if bn is not None else [None] * 4
    assert L.lib.it_b200_conv2d_stem_supported(dt, C, F, R, S, ph, pw, sh, sw, 1, 1, 1) == 1
    L.check(L.lib.it_b200_conv2d_stem(dt, ptr(xd), ptr(wd), ptr(y), N, C, H, W, F, R, S, ph, pw, sh, sw, ptr(ms[0]), ptr(ms[1]),
                                      ptr(ms[2]), ptr(ms[3]), eps, int(relu), stream()))
    sync()
    return np.ascontiguousarray(host(y).transpose(0, 3, 1, 2))


def conv2d_nhwc(x, w, ph, pw, sh, sw, dh, dw, bn=None, eps=1e-5, residual=None, relu=False, dt=F16, y_nhwc=True):
    """x, residual and the result are NCHW numpy arrays; the permutation to / from NHWC happens here (host side)."""
    N, C, H, W = x.shape
    F, _, R, S = w.shape
    OH = (H + 2 * ph - dh * (R - 1) - 1) // sh + 1
    OW = (W + 2 * pw - dw * (S - 1) - 1) // sw + 1
    xd, wd = dev(np.ascontiguousarray(x.transpose(0, 2, 3, 1)), dt), dev(w, dt)
    y = torch.empty((N, OH, OW, F) if y_nhwc else (N, F, OH, OW), dtype=xd.dtype, device="cuda")
    ms = [dev(np.asarray(v, np.float32)) for v in bn] if bn is not None else [None] * 4
    rd = None
    if residual is not None:
        rd = dev(np.ascontiguousarray(residual.transpose(0, 2, 3, 1)) if y_nhwc else residual, dt)
    assert L.lib.it_b200_conv2d_nhwc_supported(dt, C, F, R, S, ph, pw, sh, sw, dh, dw, 1) == 1
    wsb = L.lib.it_b200_conv2d_nhwc_workspace(dt, C, F, R, S)
    ws = torch.empty(max(int(wsb), 16), dtype=torch.uint8, device="cuda")
    L.check(L.lib.it_b200_conv2d_nhwc(dt, ptr(xd), ptr(wd), ptr(y), int(y_nhwc), N, C, H, W, F, R, S, ph, pw, sh, sw, dh, dw,
                                      ptr(ms[0]), ptr(ms[1]), ptr(ms[2]), ptr(ms[3]), eps, ptr(rd), int(relu), ptr(ws), int(wsb),
                                      stream()))
    sync()
    out = host(y)
    return np.ascontiguousarray(out.transpose(0, 3, 1, 2)) if y_nhwc else out


def matmul(a, b, bias=None, transA=False, transB=False, dt=F32, act=0):
    ad, bd = dev(a, dt), dev(b, dt)
    m = a.shape[-1] if transA else a.shape[-2]
    k = a.shape[-2] if transA else a.shape[-1]
    n = b.shape[-2] if transB else b.shape[-1]
    batch = tuple(np.broadcast_shapes(a.shape[:-2], b.shape[:-2]))
    nb = int(np.prod(batch)) if batch else 1
    na = int(np.prod(a.shape[:-2])) if a.ndim > 2 else 1
    nbb = int(np.prod(b.shape[:-2])) if b.ndim > 2 else 1
    sa = 0 if (na == 1 and nb > 1) else m * k
    sb = 0 if (nbb == 1 and nb > 1) else n * k
    c = torch.empty(batch + (m, n), dtype=ad.dtype, device="cuda")
    biasd, bs = None, [0, 0, 0]
    if bias is not None:
        biasd = dev(bias, dt)
        st = bstrides(bias.shape, batch + (m, n))
        # fold leading batch dims into one stride (only full or broadcast supported here)
        full = bstrides(batch + (m, n), batch + (m, n))
        bs = [0 if not batch or all(s == 0 for s in st[:-2]) else m * n, st[-2], st[-1]]
    L.check(L.lib.it_b200_matmul(dt, ptr(ad), ptr(bd), ptr(biasd), ptr(c), nb, m, n, k, sa, sb, int(transA),
                                 int(transB), bs[0], bs[1], bs[2], act, None, 0, stream()))
    sync()
    return host(c)


def matmul_fused(a, b, bias, residual, act, dt=F16, transB=False):
    """MatMul (+ bias) -> [act: 0 none, 1 relu, 4 gelu] -> [+ residual] through it_b200_matmul_fused; None when the kernel answers 2."""
    ad, bd = dev(a, dt), dev(b, dt)
    m, k = a.shape
    n = b.shape[0] if transB else b.shape[1]
    c = torch.empty((m, n), dtype=ad.dtype, device="cuda")
    biasd = dev(bias, dt) if bias is not None else None
    resd = dev(residual, dt) if residual is not None else None
    rc = L.lib.it_b200_matmul_fused(dt, ptr(ad), ptr(bd), ptr(biasd), ptr(resd), ptr(c), 1, m, n, k, m * k, 0, 0, int(transB), 0, 0, 1,
                                    act | 0x200, stream())
    if rc == 2:
        return None
    L.check(rc)
    sync()
    return host(c)


def matmul_grouped(x, ws, dt=BF16):
    xd = dev(x, dt)
    wd = [dev(w, dt) for w in ws]
    m, k = x.shape
    outs = [torch.empty((m, w.shape[1]), dtype=xd.dtype, device="cuda") for w in ws]
    W = (ctypes.c_void_p * len(ws))(*[t.data_ptr() for t in wd])
    C = (ctypes.c_void_p * len(ws))(*[t.data_ptr() for t in outs])
    L.check(L.lib.it_b200_matmul_grouped(dt, ptr(xd), len(ws), W, C, L.i32arr([w.shape[1] for w in ws]), m, k, stream()))
    sync()
    return [host(o) for o in outs]


def silu_mul(g, u, dt=BF16):
    gd, ud = dev(g, dt), dev(u, dt)
    o = torch.empty_like(gd)
    L.check(L.lib.it_b200_silu_mul(dt, ptr(gd), ptr(ud), ptr(o), gd.numel(), stream()))
    sync()
    return host(o)


def conv2d(x, w, ph, pw, sh, sw, dh, dw, dt=F32):
    xd, wd = dev(x, dt), dev(w, dt)
    N, C, H, W = x.shape
    F, Cg, R, S = w.shape
    groups = C // Cg
    OH = (H + 2 * ph - dh * (R - 1) - 1) // sh + 1
    OW = (W + 2 * pw - dw * (S - 1) - 1) // sw + 1
    y = torch.empty((N, F, OH, OW), dtype=xd.dtype, device="cuda")
    wsb = L.lib.it_b200_conv2d_workspace(dt, N, C, H, W, F, R, S, ph, pw, sh, sw, dh, dw, groups)
    ws = torch.empty(max(int(wsb), 16), dtype=torch.uint8, device="cuda")
    L.check(L.lib.it_b200_conv2d(dt, ptr(xd), ptr(wd), ptr(y), N, C, H, W, F, R, S, ph, pw, sh, sw, dh, dw, groups,
                                 ptr(ws), int(wsb), stream()))
    sync()
    return host(y)


def attention_kvcache(kc, vc, q, k, v, pos, dt=F32, pos_np_dtype=np.int64, flags=0):
    """Returns (out, kcache_after, vcache_after) as float32 numpy.  `pos`: one int (the reference's rule: element 0 for every
    row) or a per-row sequence (flags gets ITB_POS_PER_ROW = 0x100)."""
    kcd, vcd = dev(kc, dt), dev(vc, dt)
    qd, kd, vd = dev(q, dt), dev(k, dt), dev(v, dt)
    B, H, Smax, D = kc.shape
    if np.ndim(pos) > 0:
        flags |= 0x100
    pd = raw(np.atleast_1d(np.asarray(pos)).astype(pos_np_dtype))
    code = {np.int64: 7, np.int32: 6, np.uint32: 12}[pos_np_dtype] | flags
    out = torch.empty_like(qd)
    wsb = L.lib.it_b200_attention_kvcache_workspace(B, H, Smax, D)
    ws = torch.empty(max(int(wsb), 16), dtype=torch.uint8, device="cuda")
    L.check(L.lib.it_b200_attention_kvcache(dt, ptr(kcd), ptr(vcd), ptr(qd), ptr(kd), ptr(vd), ptr(pd), code,
                                            ptr(out), B, H, Smax, D, ptr(ws), int(wsb), stream()))
    sync()
    return host(out), host(kcd), host(vcd)
