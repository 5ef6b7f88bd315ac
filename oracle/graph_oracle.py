"""OracleHandler -- a CPU executor with the GraphHandler call surface, built on the oracle ops.

TEST INFRASTRUCTURE ONLY (see oracle/it_oracle.c header).  It lets tests / smoke() / the
`bench.py --impl reference` leg run the SAME graph description (infinitensor_b200/graphs.py) on the
host: every handler call records an op with its inferred output shape, run() executes them in
insertion order with the oracle functions.  Floating tensors are held as float32 values that are
exactly representable in the tensor's storage dtype.
"""
from __future__ import annotations

import numpy as np

import oracle as O

F32, F16, BF16 = 1, 10, 16
_FLOAT = (F32, F16, BF16)
_NP = {17: np.uint8, 1: np.float32, 2: np.uint8, 3: np.int8, 6: np.int32, 7: np.int64, 9: np.bool_, 10: np.float16, 12: np.uint32,
       16: np.uint16}


def _bf16_bits_to_f32(b):
    return (b.astype(np.uint32) << 16).view(np.float32)


def _f32_to_bf16_bits(a):
    u = np.ascontiguousarray(a, dtype=np.float32).view(np.uint32)
    return ((u + (((u >> 16) & 1) + 0x7FFF)) >> 16).astype(np.uint16)


class OTensor:
    def __init__(self, dims, dtype):
        self.dims, self.dt = [int(d) for d in dims], int(dtype)
        self.value = None
        self.is_weight = self.is_input = self.is_output = False

    def shape(self):
        return list(self.dims)

    def dtype(self):
        return self.dt

    def set_weight(self):
        self.is_weight = True

    def set_input(self):
        self.is_input = True

    def set_output(self):
        self.is_output = True

    def copyin_numpy(self, arr):
        arr = np.asarray(arr)
        assert list(arr.shape) == self.dims, (arr.shape, self.dims)
        if self.dt == BF16:
            self.value = _bf16_bits_to_f32(arr.astype(np.uint16)).copy()
        elif self.dt == F16:
            self.value = arr.astype(np.float32)
        else:
            self.value = np.array(arr, dtype=_NP[self.dt])

    def set_f32(self, arr):
        """values given as float32, rounded to the storage dtype"""
        self.value = O.round_to(np.asarray(arr, np.float32).reshape(self.dims), self.dt) if self.dt in _FLOAT else arr

    def f32(self):
        return self.value

    def copyout_numpy(self):
        if self.dt == BF16:
            return _f32_to_bf16_bits(self.value)
        if self.dt == F16:
            return self.value.astype(np.float16)
        return self.value


class OracleHandler:
    def __init__(self, runtime=None):
        self.ops = []
        self.tensors = []

    def tensor(self, dims, dtype):
        t = OTensor(dims, dtype)
        self.tensors.append(t)
        return t

    def _out(self, given, dims, dtype):
        if given is not None:
            assert given.dims == [int(d) for d in dims], (given.dims, dims)
            return given
        return self.tensor(dims, dtype)

    def _rec(self, fn, ins, outs):
        self.ops.append((fn, ins, outs))
        return outs[0] if len(outs) == 1 else outs

    # ---- operators
    def matmul(self, a, b, y, transA, transB, bias, act, matmul_compute_type="default", w_scale=None):
        m = a.dims[-1] if transA else a.dims[-2]
        n = b.dims[-2] if transB else b.dims[-1]
        batch = list(np.broadcast_shapes(tuple(a.dims[:-2]), tuple(b.dims[:-2])))
        out = self._out(y, batch + [m, n], a.dt)
        if w_scale is not None:  # FP8 E4M3 weight codes + per-column scale: X . (codes * scale), scale applied to the fp32 sum
            return self._rec(lambda: O.matmul_fp8w(np.asarray(a.value, np.float32).reshape(-1, a.dims[-1]), b.value, w_scale.value,
                                                   a.dt).reshape(out.dims), [a, b, w_scale], [out])
        return self._rec(lambda: O.matmul(a.value, b.value, None if bias is None else bias.value, transA, transB, a.dt),
                         [a, b], [out])

    def conv(self, x, w, y, ph, pw, sh, sw, dh, dw):
        N, C, H, W = x.dims
        Fo, _, R, S = w.dims
        OH, OW = O.conv_out_hw(H, W, R, S, ph, pw, sh, sw, dh, dw)
        out = self._out(y, [N, Fo, OH, OW], x.dt)
        return self._rec(lambda: O.conv2d(x.value, w.value, ph, pw, sh, sw, dh, dw, x.dt), [x, w], [out])

    def batchNormalization(self, x, y, mean, var, scale, bias, momentum, eps, training):
        out = self._out(y, x.dims, x.dt)
        return self._rec(lambda: O.batch_norm(x.value, mean.value, var.value, scale.value, bias.value, eps, x.dt), [x], [out])

    def layerNormalization(self, x, scale, y, bias, eps, axis, stash_type):
        out = self._out(y, x.dims, x.dt)
        return self._rec(lambda: O.layer_norm(x.value, scale.value, None if bias is None else bias.value, eps, axis, x.dt),
                         [x], [out])

    def RMSNorm(self, x, w, y):
        out = self._out(y, x.dims, x.dt)
        return self._rec(lambda: O.rms_norm(x.value, w.value, x.dt), [x, w], [out])

    def _pool(self, kind, x, y, kh, kw, dh, dw, ph, pw, sh, sw, ceil):
        N, C, H, W = x.dims
        out = self._out(y, [N, C, O.pool_out(H, kh, dh, ph, sh, ceil), O.pool_out(W, kw, dw, pw, sw, ceil)], x.dt)
        return self._rec(lambda: O.pool2d(kind, x.value, kh, kw, dh, dw, ph, pw, sh, sw, ceil, x.dt), [x], [out])

    def maxPool(self, x, y, *a):
        return self._pool("max", x, y, *a)

    def avgPool(self, x, y, *a):
        return self._pool("avg", x, y, *a)

    def _binary(self, name, a, b, c):
        dims = list(np.broadcast_shapes(tuple(a.dims), tuple(b.dims)))
        cmp_ = name in ("less", "equal", "greater")
        out = self._out(c, dims, 9 if cmp_ else a.dt)
        if cmp_:
            return self._rec(lambda: O.binary(name, a.value, b.value, a.dt).astype(np.bool_), [a, b], [out])
        return self._rec(lambda: O.binary(name, a.value, b.value, a.dt), [a, b], [out])

    def add(self, a, b, c): return self._binary("add", a, b, c)
    def sub(self, a, b, c): return self._binary("sub", a, b, c)
    def mul(self, a, b, c): return self._binary("mul", a, b, c)
    def div(self, a, b, c): return self._binary("div", a, b, c)
    def pow(self, a, b, c): return self._binary("pow", a, b, c)
    def min(self, a, b, c): return self._binary("min", a, b, c)
    def max(self, a, b, c): return self._binary("max", a, b, c)
    def less(self, a, b, c): return self._binary("less", a, b, c)

    def _unary(self, name, x, y):
        out = self._out(y, x.dims, x.dt)
        return self._rec(lambda: O.unary(name, x.value, x.dt), [x], [out])

    def leakyRelu(self, x, y, alpha):
        out = self._out(y, x.dims, x.dt)
        return self._rec(lambda: O.unary_alpha("leakyrelu", x.value, alpha, x.dt), [x], [out])

    def elu(self, x, y, alpha):
        out = self._out(y, x.dims, x.dt)
        return self._rec(lambda: O.unary_alpha("elu", x.value, alpha, x.dt), [x], [out])

    def relu(self, x, y): return self._unary("relu", x, y)
    def silu(self, x, y): return self._unary("silu", x, y)
    def gelu(self, x, y): return self._unary("gelu", x, y)
    def sigmoid(self, x, y): return self._unary("sigmoid", x, y)
    def tanh(self, x, y): return self._unary("tanh", x, y)
    def erf(self, x, y): return self._unary("erf", x, y)
    def abs(self, x, y): return self._unary("abs", x, y)
    def sqrt(self, x, y): return self._unary("sqrt", x, y)
    def neg(self, x, y): return self._unary("neg", x, y)
    def hardSigmoid(self, x, y): return self._unary("hardsigmoid", x, y)
    def hardSwish(self, x, y): return self._unary("hardswish", x, y)

    def identity(self, x, y):
        out = self._out(y, x.dims, x.dt)
        return self._rec(lambda: x.value.copy(), [x], [out])

    def softmax(self, x, y, axis):
        out = self._out(y, x.dims, x.dt)
        return self._rec(lambda: O.softmax(x.value, axis, x.dt), [x], [out])

    def flatten(self, x, y, axis):
        a = int(np.prod(x.dims[:axis])) if axis > 0 else 1
        out = self._out(y, [a, int(np.prod(x.dims)) // a], x.dt)
        return self._rec(lambda: O.reshape(x.value, out.dims), [x], [out])

    def transpose(self, x, y, perm):
        out = self._out(y, [x.dims[p] for p in perm], x.dt)
        return self._rec(lambda: O.transpose(x.value, perm), [x], [out])

    def depthToSpace(self, x, y, blocksize, mode):
        """reference src/operators/transpose.cc:68-107 (shape rule) + src/kernels/cuda/transpose.cc:48-90 (the rank-6 permute)"""
        if isinstance(mode, bytes):
            mode = mode.decode()
        n, c, hh, ww = x.dims
        b = int(blocksize)
        out = self._out(y, [n, c // (b * b), hh * b, ww * b], x.dt)

        def run():
            v = np.asarray(x.value)
            if mode == "CRD":
                return np.ascontiguousarray(v.reshape(n, c // (b * b), b, b, hh, ww).transpose(0, 1, 4, 2, 5, 3)).reshape(out.dims)
            return np.ascontiguousarray(v.reshape(n, b, b, c // (b * b), hh, ww).transpose(0, 3, 4, 1, 5, 2)).reshape(out.dims)
        return self._rec(run, [x], [out])

    def reshape(self, x, y, shape):
        shape = list(np.empty(x.dims, dtype=np.bool_).reshape(shape).shape)
        out = self._out(y, shape, x.dt)
        return self._rec(lambda: O.reshape(x.value, shape), [x], [out])

    def squeeze(self, x, y, axes):
        dims = [d for i, d in enumerate(x.dims) if not ((i in [a % len(x.dims) for a in axes]) if axes else d == 1)]
        out = self._out(y, dims, x.dt)
        return self._rec(lambda: O.reshape(x.value, dims), [x], [out])

    def unsqueeze(self, x, y, axes):
        dims = list(np.expand_dims(np.empty(x.dims, np.bool_), tuple(axes)).shape)
        out = self._out(y, dims, x.dt)
        return self._rec(lambda: O.reshape(x.value, dims), [x], [out])

    def concat(self, inputs, y, dim):
        dims = list(inputs[0].dims)
        dims[dim] = sum(t.dims[dim] for t in inputs)
        out = self._out(y, dims, inputs[0].dt)
        return self._rec(lambda: O.concat([t.value for t in inputs], dim), list(inputs), [out])

    def split(self, x, outputs, axis, numOrRatio):
        probe = O.split(np.empty(x.dims, np.bool_), axis, numOrRatio)
        outs = [self._out(None if outputs is None else outputs[i], list(p.shape), x.dt) for i, p in enumerate(probe)]
        self.ops.append((lambda: O.split(x.value, axis, numOrRatio), [x], outs))
        return outs

    def gather(self, data, indices, y, axis):
        dims = data.dims[:axis] + indices.dims + data.dims[axis + 1:]
        out = self._out(y, dims, data.dt)
        return self._rec(lambda: O.gather(data.value, indices.value, axis), [data, indices], [out])

    def attentionKVCache(self, kc, vc, q, k, v, pos, y, per_row_positions=False):
        out = self._out(y, q.dims, q.dt)

        def run():
            # in-place append into the cache INPUTS (attention_kvcache.cu:49-53,89-93)
            if per_row_positions:  # the kernel of attention_kvcache.cu:8-145 applied to each batch row with its own position
                pv = np.asarray(pos.value).ravel()
                qv, kv, vv = (np.ascontiguousarray(t.value, np.float32) for t in (q, k, v))
                return np.concatenate([O.attention_kvcache(kc.value[b:b + 1], vc.value[b:b + 1], qv[b:b + 1], kv[b:b + 1],
                                                           vv[b:b + 1], int(pv[b]), q.dt) for b in range(q.dims[0])])
            return O.attention_kvcache(kc.value, vc.value, q.value, k.value, v.value, int(np.asarray(pos.value).ravel()[0]), q.dt)
        return self._rec(run, [kc, vc, q, k, v, pos], [out])

    def RoPE(self, pos, x, y):
        out = self._out(y, x.dims, x.dt)
        return self._rec(lambda: O.rope(pos.value, x.value, 128, x.dt), [pos, x], [out])

    def reduceMean(self, x, y, axes, keepdims):
        dims = list(np.empty(x.dims, np.bool_).sum(axis=None if axes is None else tuple(axes), keepdims=keepdims).shape) or [1]
        out = self._out(y, dims, x.dt)
        return self._rec(lambda: O.reduce("mean", x.value, axes, keepdims, x.dt).reshape(dims), [x], [out])

    def reduceSum(self, x, y, axes, keepdims):
        dims = list(np.empty(x.dims, np.bool_).sum(axis=None if axes is None else tuple(axes), keepdims=keepdims).shape) or [1]
        out = self._out(y, dims, x.dt)
        return self._rec(lambda: O.reduce("sum", x.value, axes, keepdims, x.dt).reshape(dims), [x], [out])

    def slice(self, x, y, starts, ends, axes, steps):
        dims = list(O.slice_(np.empty(x.dims, np.bool_), starts, ends, axes, steps).shape)
        out = self._out(y, dims, x.dt)
        return self._rec(lambda: O.slice_(x.value, starts, ends, axes, steps), [x], [out])

    def pad(self, x, y, pads, axes):
        dims = list(O.pad(np.empty(x.dims, np.bool_), pads, axes).shape)
        out = self._out(y, dims, x.dt)
        return self._rec(lambda: O.pad(x.value, pads, axes), [x], [out])

    def cast(self, x, y, to):
        out = self._out(y, x.dims, to)

        def run():
            if to in _FLOAT:
                return O.round_to(np.asarray(x.value, np.float32), to)
            return O.cast_f32_to(x.value, to)
        return self._rec(run, [x], [out])

    def expand(self, x, y, dims):
        odims = list(np.broadcast_shapes(tuple(x.dims), tuple(dims)))
        out = self._out(y, odims, x.dt)
        return self._rec(lambda: O.expand(x.value, dims), [x], [out])

    def where(self, xx, yy, cond, y):
        dims = list(np.broadcast_shapes(tuple(xx.dims), tuple(yy.dims), tuple(cond.dims)))
        out = self._out(y, dims, xx.dt)
        return self._rec(lambda: O.where(cond.value, xx.value, yy.value), [xx, yy, cond], [out])

    def allReduceSum(self, x, y):
        """all_reduce.cc:8-63 on the host: a gloo all-reduce across the torch.distributed world (CPU tests
        of the tensor-parallel path run one OracleHandler per rank)."""
        out = self._out(y, x.dims, x.dt)

        def run():
            import torch
            import torch.distributed as dist
            if not (dist.is_available() and dist.is_initialized()):
                raise RuntimeError("allReduceSum on the oracle needs an initialised torch.distributed (gloo) world")
            t = torch.from_numpy(np.ascontiguousarray(x.value, dtype=np.float32).copy())
            dist.all_reduce(t, op=dist.ReduceOp.SUM)
            return O.round_to(t.numpy(), x.dt)
        return self._rec(run, [x], [out])

    # ---- runtime
    def data_malloc(self, *a, **k):
        pass

    def topo_sort(self):
        return True

    def shape_infer(self):
        pass

    def run(self):
        for fn, ins, outs in self.ops:
            res = fn()
            if len(outs) == 1:
                outs[0].value = res
            else:
                for o, r in zip(outs, res):
                    o.value = r

    run_with_cudagraph = run

    def operators(self):
        return [None] * len(self.ops)
