"""Golden vectors transcribed from the reference's own kernel tests.

Each case cites the reference test (path:line under /root/reference) that holds the
expected values.  Only DATA is transcribed (shapes, generator kind, expected
numbers); inputs are rebuilt here from the reference's generator semantics
(include/utils/data_generator.h:30-102: Incremental = 0,1,2,..; One = 1; Val<k> = k).
Used by tests/test_oracle_golden.py (pins the oracle, CPU) and
tests/test_gpu_golden.py (pins the CUDA kernels through the C-ABI, GPU).
"""
import numpy as np


def inc(shape, dtype=np.float32):
    return np.arange(int(np.prod(shape)), dtype=dtype).reshape(shape)


def ones(shape, dtype=np.float32):
    return np.ones(shape, dtype=dtype)


def val(shape, v, dtype=np.float32):
    return np.full(shape, v, dtype=dtype)


F = np.float32

# test/kernels/cuda/test_cuda_matmul.cc:47-66 (first two also test/kernels/intelcpu/test_mkl_matmul.cc:32-41)
MATMUL = [
    dict(a=inc((1, 3, 5)), b=ones((1, 5, 2)), tA=False, tB=False, out=[10, 10, 35, 35, 60, 60]),
    dict(a=inc((2, 3, 4)), b=inc((2, 3, 2)), tA=True, tB=False,
         out=[40, 52, 46, 61, 52, 70, 58, 79, 400, 448, 424, 475, 448, 502, 472, 529]),
    dict(a=inc((2, 3, 5)), b=inc((5, 2)), tA=False, tB=False,
         out=[60, 70, 160, 195, 260, 320, 360, 445, 460, 570, 560, 695]),
    dict(a=inc((2, 5, 3)), b=inc((5, 2)), tA=True, tB=False,
         out=[180, 210, 200, 235, 220, 260, 480, 585, 500, 610, 520, 635]),
    dict(a=inc((3, 5)), b=inc((5, 2)), tA=False, tB=False, out=[60, 70, 160, 195, 260, 320]),
]

# test/core/test_graph.cc:75-102: uint32 MatMul known answer on the native CPU runtime
MATMUL_CORE = dict(a=np.array([[1, 2, 3], [4, 5, 6]], F), b=np.array([[1, 2, 3, 4], [5, 6, 7, 8], [9, 10, 11, 12]], F),
                   out=[38, 44, 50, 56, 83, 98, 113, 128])

# test/kernels/cuda/test_cuda_conv.cc:48-54 : 1x3x4x4 (*) 2x3x3x3, p(1,1) s(2,1) d(1,2)
CONV = [
    dict(x=ones((1, 3, 4, 4)), w=ones((2, 3, 3, 3)), args=(1, 1, 2, 1, 1, 2), out=[12, 12, 18, 18, 12, 12, 18, 18]),
    dict(x=inc((1, 3, 4, 4)), w=inc((2, 3, 3, 3)), args=(1, 1, 2, 1, 1, 2),
         out=[4794, 4386, 8199, 7506, 11274, 10542, 20835, 19656]),
]

# test/kernels/cuda/test_cuda_softmax.cc:67-132
_sm_in = inc((2, 3, 2, 2))
SOFTMAX = [
    dict(x=_sm_in, axis=0, dt=1, out=[6.14417422e-06] * 12 + [9.99993801e-01] * 12),
    dict(x=_sm_in, axis=1, dt=1,
         out=([3.29320435e-04] * 4 + [1.79802869e-02] * 4 + [9.81690347e-01] * 4) * 2),
    dict(x=_sm_in, axis=2, dt=1, out=[0.11920292, 0.11920292, 0.88079703, 0.88079703] * 6),
    dict(x=_sm_in, axis=3, dt=1, out=[0.26894143, 0.73105860] * 12),
    dict(x=val((2, 3, 2, 2), 2), axis=0, dt=10, out=[0.5] * 24),
    dict(x=val((2, 3, 2, 2), 2), axis=1, dt=10, out=[0.333252] * 24),  # exact fp16 rounding of 1/3 (quirk q11)
]

# test/kernels/cuda/test_cuda_layernorm.cc:150-224 ; input [2,3,2,3] incremental, axis=3, eps=1e-5
_ln_in = inc((2, 3, 2, 3))
LAYERNORM = [
    dict(x=_ln_in, scale=[0.3, 0.2, 0.5], bias=[0, 0, 0], axis=3, dt=1, out=[-0.3674207, 0.0, 0.6123678] * 12),
    dict(x=_ln_in, scale=[0.3, 0.2, 0.5], bias=[0.3, 0.2, 0.5], axis=3, dt=1,
         out=[-0.0674207, 0.2, 1.1123679] * 12),
    dict(x=_ln_in, scale=[0.3], bias=[0.3, 0.2, 0.5], axis=3, dt=1, out=[-0.0674207, 0.2, 0.8674207] * 12),
    dict(x=_ln_in, scale=[0.3, 0.2, 0.5], bias=None, axis=3, dt=1, out=[-0.3674207, 0.0, 0.6123678] * 12),
    dict(x=val((2, 3, 2, 3), 2), scale=[2, 2, 2], bias=[2, 2, 2], axis=3, dt=10, out=[2.0] * 36),
]

# test/kernels/cuda/test_cuda_attention.cc:10-43 : 1x1x1x128 all ones, position 0 -> all ones
ATTENTION = dict(B=1, H=1, S=1, D=128, pos=0, out=[1.0] * 128)

# test/kernels/cuda/test_cuda_rope.cc:10-35 : ones, pos 1 -> cos(10000^(-2c/128)), c = 0..31.
# The reference test feeds a 32-wide tensor and reads its rotate-half partner out of
# bounds (quirk q2); the well-defined equivalent is a 128-wide head whose upper half is 0.
ROPE_COS = [0.540302, 0.647906, 0.731761, 0.796458, 0.846009, 0.883756, 0.912396, 0.934062, 0.950415,
            0.962739, 0.972014, 0.978989, 0.98423, 0.988167, 0.991122, 0.99334, 0.995004, 0.996253,
            0.99719, 0.997892, 0.998419, 0.998815, 0.999111, 0.999333, 0.9995, 0.999625, 0.999719,
            0.999789, 0.999842, 0.999881, 0.999911, 0.999933]

# test/kernels/cuda/test_cuda_element_wise.cc:47-69
ELEMENTWISE = [
    dict(op="add", a=inc((1, 2, 2, 3)), b=inc((1, 2, 2, 3)), out=[0, 2, 4, 6, 8, 10, 12, 14, 16, 18, 20, 22]),
    dict(op="sub", a=inc((1, 2, 2, 3)), b=inc((1, 2, 2, 3)), out=[0] * 12),
    dict(op="mul", a=inc((1, 2, 2, 3)), b=inc((1, 2, 2, 3)), out=[0, 1, 4, 9, 16, 25, 36, 49, 64, 81, 100, 121]),
    dict(op="div", a=ones((1, 2, 2, 3)), b=ones((1, 2, 2, 3)), out=[1] * 12),
    dict(op="min", a=inc((1, 2, 2, 3)), b=inc((1, 2, 2, 3)), out=list(range(12))),
    dict(op="max", a=inc((1, 2, 2, 3)), b=inc((1, 2, 2, 3)), out=list(range(12))),
    dict(op="pow", a=inc((1, 2, 2, 1)), b=inc((1, 2, 2, 1)), out=[1, 1, 4, 27]),
]

# test/kernels/cuda/test_cuda_pooling.cc:45-58 ; KDPS = kh,kw,dh,dw,ph,pw,sh,sw
POOL = [
    dict(kind="max", x=inc((1, 2, 5, 5)), kdps=(3, 3, 1, 1, 1, 1, 2, 2),
         out=[6, 8, 9, 16, 18, 19, 21, 23, 24, 31, 33, 34, 41, 43, 44, 46, 48, 49]),
    dict(kind="avg", x=inc((1, 2, 5, 5)), kdps=(3, 3, 1, 1, 1, 1, 2, 2),
         out=[1.333333, 3.0, 2.666667, 7.0, 12.0, 9.0, 8.0, 13.0, 9.333333, 12.44444, 19.666667, 13.777778,
              23.666667, 37.0, 25.666667, 19.111111, 29.666667, 20.444444]),
]

# test/kernels/cuda/test_cuda_batch_norm.cc:10-53
BATCHNORM = dict(x=inc((1, 3, 2, 2)), mean=[1, 6, 9], var=[4, 1, 9], scale=[1, 1, 1], bias=[0, 0, 0], eps=0.0,
                 out=[-0.5, 0, 0.5, 1, -2, -1, 0, 1, -0.333333, 0, 0.333333, 0.666667])

# test/kernels/cuda/test_cuda_reduce.cc:42-83
_r1 = [5, 1, 20, 2, 30, 1, 40, 2, 55, 1, 60, 2]
REDUCE = [
    dict(kind="mean", x=np.array(_r1, F).reshape(3, 2, 2), axes=None, keep=True, out=[18.25]),
    dict(kind="mean", x=np.array(_r1, F).reshape(1, 3, 2, 2, 1), axes=None, keep=False, out=[18.25]),
    dict(kind="mean", x=inc((2, 3, 2, 2)), axes=[1, 2], keep=False, out=[5, 6, 17, 18]),
    dict(kind="mean", x=inc((2, 3, 2, 2, 1)), axes=[1, 2], keep=True, out=[5, 6, 17, 18]),
    dict(kind="sum", x=ones((3, 2, 2)), axes=None, keep=True, out=[12]),
    dict(kind="sum", x=ones((1, 3, 2, 2, 1)), axes=None, keep=False, out=[12]),
    dict(kind="sum", x=inc((2, 3, 2, 2)), axes=[1, 2], keep=False, out=[30, 36, 102, 108]),
    dict(kind="sum", x=inc((2, 3, 2, 2, 1)), axes=[1, 2], keep=True, out=[30, 36, 102, 108]),
]

# test/kernels/cuda/test_cuda_transpose.cc:37-39 ; perm {0,2,1,3}
TRANSPOSE = dict(x=inc((1, 2, 3, 4)), perm=(0, 2, 1, 3),
                 out=[0, 1, 2, 3, 12, 13, 14, 15, 4, 5, 6, 7, 16, 17, 18, 19, 8, 9, 10, 11, 20, 21, 22, 23])

# test/kernels/cuda/test_cuda_concat.cc:62-160
CONCAT = [
    dict(xs=[inc((2, 2, 3, 1)), ones((2, 2, 1, 1)), ones((2, 2, 2, 1))], dim=2,
         out=[0, 1, 2, 1, 1, 1, 3, 4, 5, 1, 1, 1, 6, 7, 8, 1, 1, 1, 9, 10, 11, 1, 1, 1]),
    dict(xs=[inc((1, 3)), ones((1, 3)), inc((1, 3))], dim=0, out=[0, 1, 2, 1, 1, 1, 0, 1, 2]),
    dict(xs=[inc((2, 2, 3, 1, 2)), ones((2, 2, 1, 1, 2)), ones((2, 2, 2, 1, 2))], dim=2,
         out=[0, 1, 2, 3, 4, 5, 1, 1, 1, 1, 1, 1, 6, 7, 8, 9, 10, 11, 1, 1, 1, 1, 1, 1,
              12, 13, 14, 15, 16, 17, 1, 1, 1, 1, 1, 1, 18, 19, 20, 21, 22, 23, 1, 1, 1, 1, 1, 1]),
]

# test/kernels/cuda/test_cuda_split.cc:14-47 : [2,10,2,1] split(axis 1, num 3) -> 3,3,4
SPLIT = dict(x=inc((2, 10, 2, 1)), axis=1, num=3,
             outs=[[0, 1, 2, 3, 4, 5, 20, 21, 22, 23, 24, 25], [6, 7, 8, 9, 10, 11, 26, 27, 28, 29, 30, 31],
                   [12, 13, 14, 15, 16, 17, 18, 19, 32, 33, 34, 35, 36, 37, 38, 39]])

# test/kernels/cuda/test_cuda_gather.cc:177-275
GATHER = [
    dict(x=np.array([1, 2, 3, 4, 5, 6], F).reshape(3, 2), idx=np.array([0, 1, 1, 2], np.int32).reshape(2, 2), axis=0,
         out=[1, 2, 3, 4, 3, 4, 5, 6]),
    dict(x=inc((3, 3)), idx=np.array([0, 2], np.int32).reshape(1, 2), axis=1, out=[0, 2, 3, 5, 6, 8]),
    dict(x=inc((2, 4, 2)), idx=np.array([0, 3, 1], np.int32).reshape(3, 1), axis=1,
         out=[0, 1, 6, 7, 2, 3, 8, 9, 14, 15, 10, 11]),
    dict(x=inc((2, 4, 2)), idx=np.array([0, 3, 1], np.int64).reshape(3, 1), axis=1,
         out=[0, 1, 6, 7, 2, 3, 8, 9, 14, 15, 10, 11]),
]

# test/kernels/cuda/test_cuda_where.cc:86-149
WHERE = [
    dict(x=inc((2, 2, 3, 1)), y=np.zeros((2, 2, 3, 1), F),
         c=np.array([0, 1, 1, 0, 0, 0, 1, 1, 0, 1, 1, 1], np.uint8).reshape(2, 2, 3, 1),
         out=[0., 1., 2., 0., 0., 0., 6., 7., 0., 9., 10., 11.]),
    dict(x=inc((2, 1, 1, 3)), y=ones((1, 2, 1, 1)), c=np.array([0, 1, 1, 0, 0, 0], np.uint8).reshape(2, 1, 3, 1),
         out=[1., 1., 1., 0., 1., 2., 0., 1., 2., 1., 1., 1., 0., 1., 2., 0., 1., 2.,
              1., 1., 1., 1., 1., 1., 1., 1., 1., 1., 1., 1., 1., 1., 1., 1., 1., 1.]),
    dict(x=inc((3,)), y=inc((2, 3, 1)), c=np.array([0, 1, 1, 0, 0, 0], np.uint8).reshape(2, 1, 3, 1),
         out=[0., 0., 0., 0., 1., 2., 0., 1., 2., 3., 3., 3., 0., 1., 2., 0., 1., 2., 0., 0., 0., 1., 1., 1.,
              2., 2., 2., 3., 3., 3., 4., 4., 4., 5., 5., 5.]),
]

# test/kernels/cuda/test_cuda_expand.cc:11-38
EXPAND = dict(x=inc((2, 1, 2, 1)), dims=(2, 2, 2, 3),
              out=[0, 0, 0, 1, 1, 1, 0, 0, 0, 1, 1, 1, 2, 2, 2, 3, 3, 3, 2, 2, 2, 3, 3, 3])

# test/kernels/cuda/test_cuda_pad.cc:9-38 : pads {1,0,1,1} on axes {0,3}
PAD = dict(x=inc((1, 2, 3, 2)), pads=[1, 0, 1, 1], axes=[0, 3],
           out=[0] * 18 + [0, 1, 0, 2, 3, 0, 4, 5, 0, 6, 7, 0, 8, 9, 0, 10, 11, 0] + [0] * 18)

# test/kernels/cuda/test_cuda_slice.cc:9-38 : starts {1,1} ends {2,5} axes {0,3}
SLICE = dict(x=inc((3, 2, 1, 5)), starts=[1, 1], ends=[2, 5], axes=[0, 3], out=[11, 12, 13, 14, 16, 17, 18, 19])

# test/kernels/cuda/test_cuda_all_reduce.cc:38-106 : world 2, rank tensors {2,3} and {5,6}
ALLREDUCE = [
    dict(kind="sum", xs=[[2., 3.], [5., 6.]], out=[7., 9.]),
    dict(kind="prod", xs=[[2., 3.], [5., 6.]], out=[10., 18.]),
    dict(kind="min", xs=[[2., 3.], [5., 6.]], out=[2., 3.]),
    dict(kind="max", xs=[[2., 3.], [5., 6.]], out=[5., 6.]),
    dict(kind="avg", xs=[[2., 3.], [5., 6.]], out=[3.5, 4.5]),
]
