"""CPU oracle for the InfiniTensor operator-kernel hot path.

TEST INFRASTRUCTURE ONLY -- see the header of it_oracle.c.  Only tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg may
import this package; the product (infinitensor_b200) never does.

Floating-point ops are restated in C (it_oracle.c, one function per reference
kernel, each citing the reference file:line).  Pure data-movement ops
(Transpose, Concat, Split, Gather, Reshape, Cast, Where, Expand, Slice, Pad,
Reduce) are restated with numpy, whose semantics are the ONNX semantics the
reference's shape rules implement; each cites the reference kernel it pins.

Every fp function takes/returns float32 arrays whose values are exactly
representable in the storage dtype `dt` (ONNX enum: 1 f32, 10 f16, 16 bf16).
"""
from __future__ import annotations

import ctypes
import os
import subprocess
from pathlib import Path

import numpy as np

_HERE = Path(__file__).resolve().parent
_LIB_PATH = _HERE / "_build" / "libit_oracle.so"

F32, F16, BF16 = 1, 10, 16
INT8, UINT8, INT32, INT64, BOOL = 3, 2, 6, 7, 9

UNARY = {"relu": 0, "sigmoid": 1, "tanh": 2, "gelu": 3, "silu": 4, "erf": 5, "neg": 6,
         "abs": 7, "sqrt": 8, "hardsigmoid": 9, "hardswish": 10, "exp": 11}
BINARY = {"add": 0, "sub": 1, "mul": 2, "div": 3, "pow": 4, "min": 5, "max": 6,
          "less": 7, "equal": 8, "greater": 9}


def build(force: bool = False) -> Path:
    """Compile it_oracle.c -> oracle/_build/libit_oracle.so (gcc, OpenMP)."""
    src = _HERE / "it_oracle.c"
    if _LIB_PATH.exists() and not force and _LIB_PATH.stat().st_mtime >= src.stat().st_mtime:
        return _LIB_PATH
    _LIB_PATH.parent.mkdir(exist_ok=True)
    cc = "/usr/bin/gcc" if os.path.exists("/usr/bin/gcc") else "gcc"
    cmd = [cc, "-O3", "-march=x86-64-v3", "-ffp-contract=off", "-fopenmp", "-shared", "-fPIC",
           "-o", str(_LIB_PATH), str(src), "-lm"]
    subprocess.run(cmd, check=True)
    return _LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = ctypes.CDLL(str(_LIB_PATH))
        _lib.orc_round.restype = ctypes.c_float
        _lib.orc_round.argtypes = [ctypes.c_float, ctypes.c_int]
    return _lib


def _fp(a):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_float))


def _f32c(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _i64arr(v):
    return (ctypes.c_int64 * len(v))(*[int(x) for x in v])


def round_to(a, dt):
    """Round a float32 array to the nearest value representable in dt (RNE)."""
    a = _f32c(a).copy()
    if dt != F32:
        lib().orc_round_array(_fp(a), ctypes.c_int64(a.size), ctypes.c_int(dt))
    return a


# ---------------------------------------------------------------- fp ops (C)
def matmul(a, b, bias=None, transA=False, transB=False, dt=F32):
    """src/operators/matmul.cc:26-49 + src/kernels/cuda/matmul.cc:66-174."""
    a, b = _f32c(a), _f32c(b)
    m = a.shape[-1] if transA else a.shape[-2]
    k = a.shape[-2] if transA else a.shape[-1]
    n = b.shape[-2] if transB else b.shape[-1]
    kb = b.shape[-1] if transB else b.shape[-2]
    assert k == kb, "matmul: K mismatch"
    batch = np.broadcast_shapes(a.shape[:-2], b.shape[:-2])
    nb = int(np.prod(batch)) if batch else 1
    na = int(np.prod(a.shape[:-2])) if a.ndim > 2 else 1
    nbb = int(np.prod(b.shape[:-2])) if b.ndim > 2 else 1
    # reference batch rule (matmul.cc:124-137): full batch or stride-0 broadcast
    assert na in (1, nb) and nbb in (1, nb), "matmul: unsupported batch broadcast"
    sa = 0 if na == 1 and nb > 1 else m * k
    sb = 0 if nbb == 1 and nb > 1 else n * k
    out_shape = tuple(batch) + (m, n)
    c = np.empty(out_shape, dtype=np.float32)
    bias_full = None
    if bias is not None:
        bias_full = np.ascontiguousarray(np.broadcast_to(_f32c(bias), out_shape))
    lib().orc_matmul(_fp(a), _fp(b), _fp(bias_full) if bias_full is not None else None, _fp(c),
                     ctypes.c_int64(nb), ctypes.c_int64(sa), ctypes.c_int64(sb),
                     m, n, k, int(transA), int(transB), dt)
    return c


def conv_out_hw(H, W, R, S, ph, pw, sh, sw, dh, dw):
    """src/operators/conv.cc:85-114."""
    return ((H + 2 * ph - dh * (R - 1) - 1) // sh + 1, (W + 2 * pw - dw * (S - 1) - 1) // sw + 1)


def conv2d(x, w, ph, pw, sh, sw, dh, dw, dt=F32):
    x, w = _f32c(x), _f32c(w)
    N, C, H, W = x.shape
    F, Cg, R, S = w.shape
    groups = C // Cg
    OH, OW = conv_out_hw(H, W, R, S, ph, pw, sh, sw, dh, dw)
    y = np.empty((N, F, OH, OW), dtype=np.float32)
    lib().orc_conv2d(_fp(x), _fp(w), _fp(y), N, C, H, W, F, R, S, ph, pw, sh, sw, dh, dw, groups, dt)
    return y


def attention_kvcache(kcache, vcache, q, k, v, pos, dt=F32):
    """In-place on kcache / vcache (must be float32 C-contiguous). Returns out [B,H,1,D]."""
    assert kcache.dtype == np.float32 and kcache.flags.c_contiguous
    assert vcache.dtype == np.float32 and vcache.flags.c_contiguous
    q, k, v = _f32c(q), _f32c(k), _f32c(v)
    B, H, Smax, D = kcache.shape
    out = np.empty(q.shape, dtype=np.float32)
    lib().orc_attention_kvcache(_fp(kcache), _fp(vcache), _fp(q), _fp(k), _fp(v),
                                ctypes.c_int64(int(pos)), _fp(out), B, H, Smax, D, dt)
    return out


def _axis_view(shape, axis):
    axis = axis % len(shape)
    outer = int(np.prod(shape[:axis])) if axis > 0 else 1
    inner = int(np.prod(shape[axis + 1:])) if axis + 1 < len(shape) else 1
    return outer, int(shape[axis]), inner


def softmax(x, axis, dt=F32):
    x = _f32c(x)
    outer, dim, inner = _axis_view(x.shape, axis)
    y = np.empty_like(x)
    lib().orc_softmax(_fp(x), _fp(y), ctypes.c_int64(outer), dim, ctypes.c_int64(inner), dt)
    return y


def layer_norm(x, scale, bias=None, eps=1e-5, axis=-1, dt=F32):
    x, scale = _f32c(x), _f32c(scale)
    outer, dim, inner = _axis_view(x.shape, axis)
    y = np.empty_like(x)
    bsz = 0
    bp = None
    if bias is not None:
        bias = _f32c(bias)
        bsz, bp = bias.size, _fp(bias)
    lib().orc_layernorm(_fp(x), _fp(scale), bp, _fp(y), ctypes.c_int64(outer), dim,
                        ctypes.c_int64(inner), int(scale.size), int(bsz), ctypes.c_float(eps), dt)
    return y


def rms_norm(x, w, dt=F32):
    x, w = _f32c(x), _f32c(w)
    hidden = x.shape[-1]
    y = np.empty_like(x)
    lib().orc_rmsnorm(_fp(x), _fp(w), _fp(y), ctypes.c_int64(x.size // hidden), hidden, dt)
    return y


def rope(pos, x, dim_head=128, dt=F32):
    x = _f32c(x)
    B, S, dm = x.shape
    p = np.ascontiguousarray(pos, dtype=np.int64).reshape(B, S)
    y = np.empty_like(x)
    lib().orc_rope(p.ctypes.data_as(ctypes.POINTER(ctypes.c_int64)), _fp(x), _fp(y), B, S, dm, dim_head, dt)
    return y


def unary(name, x, dt=F32):
    x = _f32c(x)
    y = np.empty_like(x)
    lib().orc_unary(UNARY[name], _fp(x), _fp(y), ctypes.c_int64(x.size), dt)
    return y


# ---- FP8 E4M3 ("FN": 1 sign, 4 exponent bits with bias 7, 3 mantissa bits, no infinities, 0x7f / 0xff = NaN, max 448) -- the
# OCP 8-bit format of ONNX FLOAT8E4M3FN (type 17).  Restated from the format definition; the reference has no fp8 code.
def _e4m3_table():
    v = np.zeros(256, np.float32)
    for c in range(256):
        s, e, m = c >> 7, (c >> 3) & 15, c & 7
        if e == 15 and m == 7:
            val = np.nan
        elif e == 0:
            val = (m / 8.0) * 2.0 ** -6
        else:
            val = (1.0 + m / 8.0) * 2.0 ** (e - 7)
        v[c] = -val if s else val
    return v


E4M3 = _e4m3_table()


def e4m3_decode(codes):
    return E4M3[np.asarray(codes, np.uint8)]


def e4m3_quantize(x):
    """float32 -> E4M3 codes, round to nearest, ties to the even code, saturating at +-448 (no NaN / inf inputs)."""
    x = np.asarray(x, np.float32)
    mag = np.minimum(np.abs(x).astype(np.float64), 448.0)
    pos = E4M3[:127].astype(np.float64)            # codes 0..126: 0 .. 448, increasing
    hi = np.clip(np.searchsorted(pos, mag, side="left"), 0, 126)
    lo = np.clip(hi - 1, 0, 126)
    dlo, dhi = mag - pos[lo], pos[hi] - mag
    pick_hi = (dhi < dlo) | ((dhi == dlo) & (hi % 2 == 0))
    code = np.where(pick_hi, hi, lo).astype(np.uint8)
    return np.where(np.signbit(x), code | 0x80, code).astype(np.uint8)


def quantize_weight_fp8(w):
    """per-output-column symmetric quantisation of W [K, N]: scale[n] = max|W[:, n]| / 448; returns (codes uint8 [K,N], scale f32 [N])"""
    w = np.asarray(w, np.float32)
    scale = np.maximum(np.abs(w).max(axis=0), 1e-12).astype(np.float32) / np.float32(448.0)
    return e4m3_quantize(w / scale[None, :]), scale


def dequantize_fp8(codes, scale, dt=F32):
    """DequantizeLinear: T(e4m3(code) * scale[col]) (one rounding to the storage type)"""
    return round_to(e4m3_decode(codes) * np.asarray(scale, np.float32)[None, :], dt)


def matmul_fp8w(x, codes, scale, dt=F32):
    """X . (Wq * scale): fp32 accumulation over the EXACT code values, the column scale applied to the sum (what the GEMM's
    epilogue does), one rounding to the storage type."""
    x = _f32c(x)
    acc = x.astype(np.float64) @ e4m3_decode(codes).astype(np.float64)
    return round_to((acc * np.asarray(scale, np.float64)[None, :]).astype(np.float32), dt)


def unary_alpha(name, x, alpha, dt=F32):
    """LeakyRelu: x > 0 ? x : alpha * x (reference unary.cu:157-165); Elu: x >= 0 ? x : alpha * (expf(x) - 1) (unary.cu:97-106);
    fp32 arithmetic on the (already rounded) inputs, one rounding to the storage type."""
    x = _f32c(x)
    a = np.float32(alpha)
    if name == "leakyrelu":
        y = np.where(x > 0, x, a * x)
    else:
        y = np.where(x >= 0, x, a * (np.exp(x, dtype=np.float32) - np.float32(1)))
    return round_to(y.astype(np.float32), dt)


def _bstrides(shape, out_shape):
    r = len(out_shape)
    shape = (1,) * (r - len(shape)) + tuple(shape)
    st, acc = [], 1
    for d in reversed(shape):
        st.append(acc)
        acc *= d
    st = list(reversed(st))
    return [0 if shape[i] == 1 and out_shape[i] != 1 else st[i] for i in range(r)]


def binary(name, a, b, dt=F32):
    a, b = _f32c(a), _f32c(b)
    out_shape = np.broadcast_shapes(a.shape, b.shape)
    c = np.empty(out_shape, dtype=np.float32)
    dims = list(out_shape) if out_shape else [1]
    sa = _bstrides(a.shape, tuple(dims)) if out_shape else [0]
    sb = _bstrides(b.shape, tuple(dims)) if out_shape else [0]
    lib().orc_binary(BINARY[name], _fp(a), _fp(b), _fp(c), len(dims), _i64arr(dims), _i64arr(sa),
                     _i64arr(sb), dt)
    return c


def pool_out(H, k, d, p, s, ceil_mode):
    """src/operators/pooling.cc: floor/ceil((H + 2p - d(k-1) - 1)/s) + 1."""
    num = H + 2 * p - d * (k - 1) - 1
    return (-(-num // s) if ceil_mode else num // s) + 1


def pool2d(kind, x, kh, kw, dh, dw, ph, pw, sh, sw, ceil_mode=0, dt=F32):
    x = _f32c(x)
    N, C, H, W = x.shape
    OH, OW = pool_out(H, kh, dh, ph, sh, ceil_mode), pool_out(W, kw, dw, pw, sw, ceil_mode)
    y = np.empty((N, C, OH, OW), dtype=np.float32)
    lib().orc_pool2d(1 if kind == "max" else 0, _fp(x), _fp(y), N, C, H, W, kh, kw, dh, dw, ph, pw,
                     sh, sw, OH, OW, dt)
    return y


def batch_norm(x, mean, var, scale, bias, eps=1e-5, dt=F32):
    x = _f32c(x)
    N, C = x.shape[:2]
    HW = int(np.prod(x.shape[2:])) if x.ndim > 2 else 1
    y = np.empty_like(x)
    lib().orc_batchnorm(_fp(x), _fp(_f32c(mean)), _fp(_f32c(var)), _fp(_f32c(scale)), _fp(_f32c(bias)),
                        _fp(y), N, C, ctypes.c_int64(HW), ctypes.c_float(eps), dt)
    return y


# ------------------------------------------------- data-movement ops (numpy, bit-exact)
def transpose(x, perm):
    """src/kernels/cuda/transpose.cu:10-24 (N-d permute)."""
    return np.ascontiguousarray(np.transpose(x, perm))


def concat(xs, dim):
    """src/kernels/cuda/split_concat.cu:28-50."""
    return np.concatenate(xs, axis=dim)


def split(x, axis, num_or_ratio):
    """src/operators/split.cc:6-58: int num -> num-1 pieces of dim//num and a last piece that
    also takes the remainder; list -> sizes proportional to ratio."""
    if isinstance(num_or_ratio, int):
        num = num_or_ratio
        piece = x.shape[axis] // num
        last = x.shape[axis] - piece * num
        num_or_ratio = [piece] * (num - 1) + [piece + last] if last > 0 else [piece] * num
    tot = sum(num_or_ratio)
    unit = x.shape[axis] // tot
    idx = np.cumsum([r * unit for r in num_or_ratio])[:-1]
    return [np.ascontiguousarray(p) for p in np.split(x, idx, axis=axis)]


def gather(data, indices, axis):
    """src/kernels/cuda/gather.cu:31-39, include/cuda/gather.h:7-55 (negative idx not wrapped there; ONNX wraps)."""
    return np.take(data, np.asarray(indices), axis=axis)


def reshape(x, shape):
    """src/kernels/cuda/reshape.cc:4-21 (a copy)."""
    return np.ascontiguousarray(x).reshape(shape).copy()


def where(cond, x, y):
    """src/kernels/cuda/where.cu:20-41: out = cond ? x : y, 3-way broadcast."""
    return np.where(np.asarray(cond).astype(bool), x, y)


def expand(x, dims):
    """src/kernels/cuda/expand.cu:10-49."""
    return np.ascontiguousarray(np.broadcast_to(x, np.broadcast_shapes(x.shape, tuple(dims))))


def slice_(x, starts, ends, axes=None, steps=None):
    """src/operators/slice.cc:61-70 (ONNX Slice incl. steps; the CUDA kernel ignores steps, quirk q14)."""
    axes = list(range(len(starts))) if axes is None else axes
    steps = [1] * len(starts) if steps is None else steps
    sl = [slice(None)] * x.ndim
    for s, e, a, st in zip(starts, ends, axes, steps):
        sl[a] = slice(s, e, st)
    return np.ascontiguousarray(x[tuple(sl)])


def pad(x, pads, axes=None):
    """src/kernels/cuda/pad_slice.cu:6-47: constant zero padding, pads = [begins..., ends...]."""
    axes = list(range(x.ndim)) if axes is None else axes
    n = len(axes)
    pw = [(0, 0)] * x.ndim
    for i, a in enumerate(axes):
        pw[a] = (pads[i], pads[i + n])
    return np.pad(x, pw)


def reduce(kind, x, axes=None, keepdims=True, dt=F32):
    """src/kernels/cuda/reduce.cc:7-125 (cudnnReduceTensor ADD / AVG)."""
    ax = None if axes is None else tuple(axes)
    x64 = np.asarray(x, dtype=np.float64)
    r = x64.sum(axis=ax, keepdims=keepdims) if kind == "sum" else x64.mean(axis=ax, keepdims=keepdims)
    return round_to(r.astype(np.float32), dt)


def cast_f32_to(x, to):
    """src/kernels/cuda/unary.cu:145-154 (cub::CastOp == static_cast): f32->i32/i8 truncate toward zero."""
    x = _f32c(x)
    if to == F16:
        return x.astype(np.float16)
    if to == INT32:
        return np.trunc(x).astype(np.int32)
    if to == INT8:
        return np.trunc(x).astype(np.int8)
    if to == F32:
        return x
    raise NotImplementedError(to)


def all_reduce(kind, xs, dt=F32):
    """src/kernels/cuda/all_reduce.cc:8-63: every rank ends with op over all ranks' tensors."""
    acc = np.asarray(xs[0], dtype=np.float32).copy()
    for x in xs[1:]:
        x = np.asarray(x, dtype=np.float32)
        acc = {"sum": acc + x, "avg": acc + x, "prod": acc * x, "min": np.minimum(acc, x),
               "max": np.maximum(acc, x)}[kind]
    if kind == "avg":
        acc = acc / len(xs)
    return round_to(acc, dt)
