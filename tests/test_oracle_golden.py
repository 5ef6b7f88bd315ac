"""Pin the CPU oracle against the reference's own golden vectors (CPU, no GPU).

Comparator mirrors TensorObj::equalData (include/core/tensor.h:198-221): relative 1e-6
for exactly-computed integers, and the printed precision of the golden (6-7 digits) elsewhere.
"""
import numpy as np
import pytest

import oracle
from tests.golden import reference_vectors as G


def close(got, exp, rel=2e-6, abs_=2e-7):
    got = np.asarray(got, dtype=np.float64).ravel()
    exp = np.asarray(exp, dtype=np.float64).ravel()
    assert got.shape == exp.shape, (got.shape, exp.shape)
    np.testing.assert_allclose(got, exp, rtol=rel, atol=abs_)


@pytest.mark.parametrize("c", G.MATMUL)
def test_matmul_golden(c):
    close(oracle.matmul(c["a"], c["b"], transA=c["tA"], transB=c["tB"]), c["out"])


def test_matmul_core_known_answer():
    close(oracle.matmul(G.MATMUL_CORE["a"], G.MATMUL_CORE["b"]), G.MATMUL_CORE["out"])


def test_matmul_bias_and_transb():
    rng = np.random.default_rng(0)
    a = rng.standard_normal((2, 5, 7)).astype(np.float32)
    b = rng.standard_normal((6, 7)).astype(np.float32)
    bias = rng.standard_normal((6,)).astype(np.float32)
    ref = a.astype(np.float64) @ b.astype(np.float64).T + bias
    close(oracle.matmul(a, b, bias=bias, transB=True), ref, rel=1e-5, abs_=1e-5)


@pytest.mark.parametrize("c", G.CONV)
def test_conv_golden(c):
    close(oracle.conv2d(c["x"], c["w"], *c["args"]), c["out"])


@pytest.mark.parametrize("c", G.SOFTMAX)
def test_softmax_golden(c):
    close(oracle.softmax(c["x"], c["axis"], dt=c["dt"]), c["out"], rel=1e-6 if c["dt"] == 1 else 1e-6, abs_=1e-9)


@pytest.mark.parametrize("c", G.LAYERNORM)
def test_layernorm_golden(c):
    bias = None if c["bias"] is None else np.array(c["bias"], np.float32)
    got = oracle.layer_norm(c["x"], np.array(c["scale"], np.float32), bias, eps=1e-5, axis=c["axis"], dt=c["dt"])
    close(got, c["out"], rel=1e-6, abs_=2e-7)


def test_attention_golden():
    c = G.ATTENTION
    kc = np.zeros((c["B"], c["H"], c["S"], c["D"]), np.float32)
    vc = np.zeros_like(kc)
    one = np.ones((c["B"], c["H"], 1, c["D"]), np.float32)
    out = oracle.attention_kvcache(kc, vc, one, one, one, c["pos"])
    close(out, c["out"])
    # in-place append contract (attention_kvcache.cu:49-53, 89-93)
    assert np.array_equal(kc[:, :, 0], one[:, :, 0]) and np.array_equal(vc[:, :, 0], one[:, :, 0])


def test_attention_matches_stable_softmax():
    rng = np.random.default_rng(3)
    B, H, S, D, pos = 2, 3, 64, 128, 37
    kc = (rng.standard_normal((B, H, S, D)) * 0.5).astype(np.float32)
    vc = (rng.standard_normal((B, H, S, D)) * 0.5).astype(np.float32)
    q, k, v = [(rng.standard_normal((B, H, 1, D)) * 0.5).astype(np.float32) for _ in range(3)]
    kc2, vc2 = kc.copy(), vc.copy()
    out = oracle.attention_kvcache(kc2, vc2, q, k, v, pos)
    kk = kc.astype(np.float64); vv = vc.astype(np.float64)
    kk[:, :, pos] = k[:, :, 0]; vv[:, :, pos] = v[:, :, 0]
    s = np.einsum("bhd,bhsd->bhs", q[:, :, 0].astype(np.float64), kk[:, :, :pos + 1]) / np.sqrt(128.0)
    p = np.exp(s - s.max(-1, keepdims=True)); p /= p.sum(-1, keepdims=True)
    ref = np.einsum("bhs,bhsd->bhd", p, vv[:, :, :pos + 1])
    close(out[:, :, 0], ref, rel=2e-5, abs_=2e-6)
    assert np.array_equal(kc2[:, :, pos], k[:, :, 0])
    assert np.array_equal(kc2[:, :, pos + 1:], kc[:, :, pos + 1:])


def test_rope_golden():
    x = np.zeros((1, 1, 128), np.float32)
    x[..., :64] = 1.0
    out = oracle.rope(np.array([[1]]), x)
    close(out[0, 0, :32], G.ROPE_COS, rel=2e-6, abs_=1e-6)


def test_rope_all_rows():
    rng = np.random.default_rng(1)
    x = rng.standard_normal((3, 2, 256)).astype(np.float32)
    pos = np.array([[0, 5], [7, 9], [100, 1023]])
    out = oracle.rope(pos, x)
    xx = x.reshape(3, 2, 2, 128).astype(np.float64)
    inv = 10000.0 ** (-np.arange(64) * 2 / 128.0)
    ang = pos[..., None, None] * inv
    ref = np.concatenate([xx[..., :64] * np.cos(ang) - xx[..., 64:] * np.sin(ang),
                          xx[..., 64:] * np.cos(ang) + xx[..., :64] * np.sin(ang)], -1).reshape(3, 2, 256)
    close(out, ref, rel=1e-4, abs_=2e-4)  # float powf/cos on large angles


def test_rmsnorm_restatement():
    """No reference test exists (parity unpinned); pin the formula of rms_norm.cu:36-54."""
    rng = np.random.default_rng(2)
    x = rng.standard_normal((5, 64)).astype(np.float32)
    w = rng.standard_normal((64,)).astype(np.float32)
    ref = x.astype(np.float64) / np.sqrt((x.astype(np.float64) ** 2).mean(-1, keepdims=True) + 1e-5) * w
    close(oracle.rms_norm(x, w), ref, rel=1e-5, abs_=1e-6)
    # bf16: rounding to T before the weight multiply (rms_norm.cu:52)
    xb, wb = oracle.round_to(x, oracle.BF16), oracle.round_to(w, oracle.BF16)
    r = 1.0 / np.sqrt((xb.astype(np.float32) ** 2).sum(-1, keepdims=True, dtype=np.float32) / 64 + np.float32(1e-5))
    ref_b = oracle.round_to(oracle.round_to(xb * r.astype(np.float32), oracle.BF16) * wb, oracle.BF16)
    got = oracle.rms_norm(xb, wb, dt=oracle.BF16)
    assert np.mean(got == ref_b) > 0.99  # summation-order ulps only


@pytest.mark.parametrize("c", G.ELEMENTWISE)
def test_elementwise_golden(c):
    close(oracle.binary(c["op"], c["a"], c["b"]), c["out"])


def test_binary_broadcast():
    rng = np.random.default_rng(4)
    a = rng.standard_normal((2, 1, 3, 4)).astype(np.float32)
    b = rng.standard_normal((5, 1, 4)).astype(np.float32)
    assert np.array_equal(oracle.binary("add", a, b), a + b)
    assert np.array_equal(oracle.binary("mul", a, np.float32(3.0).reshape(())), a * np.float32(3.0))
    assert np.array_equal(oracle.binary("less", a, b), (a < b).astype(np.float32))


@pytest.mark.parametrize("c", G.POOL)
def test_pool_golden(c):
    close(oracle.pool2d(c["kind"], c["x"], *c["kdps"]), c["out"], rel=2e-6)


def test_batchnorm_golden():
    c = G.BATCHNORM
    close(oracle.batch_norm(c["x"], c["mean"], c["var"], c["scale"], c["bias"], c["eps"]), c["out"], rel=2e-6)


@pytest.mark.parametrize("c", G.REDUCE)
def test_reduce_golden(c):
    close(oracle.reduce(c["kind"], c["x"], c["axes"], c["keep"]), c["out"])


def test_unary_formulas():
    x = np.arange(12, dtype=np.float32).reshape(1, 2, 2, 3) - 4
    from math import erf
    close(oracle.unary("gelu", x), [0.5 * v * (1 + erf(v / np.sqrt(2))) for v in x.ravel()], rel=1e-6, abs_=1e-7)
    close(oracle.unary("silu", x), x / (1 + np.exp(-x.astype(np.float64))), rel=1e-6, abs_=1e-7)
    close(oracle.unary("relu", x), np.maximum(x, 0))
    close(oracle.unary("hardswish", x), x * np.clip(x / 6 + 0.5, 0, 1), rel=1e-6, abs_=1e-7)
    close(oracle.unary("hardsigmoid", x), np.clip(0.2 * x + 0.5, 0, 1), rel=1e-6, abs_=1e-7)
    close(oracle.unary("sqrt", np.abs(x)), np.sqrt(np.abs(x)))


# ---- data-movement ops: bit-exact
def test_transpose_golden():
    c = G.TRANSPOSE
    assert oracle.transpose(c["x"], c["perm"]).ravel().tolist() == c["out"]


@pytest.mark.parametrize("c", G.CONCAT)
def test_concat_golden(c):
    assert oracle.concat(c["xs"], c["dim"]).ravel().tolist() == c["out"]


def test_split_golden():
    c = G.SPLIT
    outs = oracle.split(c["x"], c["axis"], c["num"])
    assert [o.ravel().tolist() for o in outs] == c["outs"]


@pytest.mark.parametrize("c", G.GATHER)
def test_gather_golden(c):
    assert oracle.gather(c["x"], c["idx"], c["axis"]).ravel().tolist() == c["out"]


@pytest.mark.parametrize("c", G.WHERE)
def test_where_golden(c):
    assert oracle.where(c["c"], c["x"], c["y"]).ravel().tolist() == c["out"]


def test_expand_pad_slice_golden():
    assert oracle.expand(G.EXPAND["x"], G.EXPAND["dims"]).ravel().tolist() == G.EXPAND["out"]
    assert oracle.pad(G.PAD["x"], G.PAD["pads"], G.PAD["axes"]).ravel().tolist() == G.PAD["out"]
    c = G.SLICE
    assert oracle.slice_(c["x"], c["starts"], c["ends"], c["axes"]).ravel().tolist() == c["out"]


@pytest.mark.parametrize("c", G.ALLREDUCE)
def test_allreduce_golden(c):
    close(oracle.all_reduce(c["kind"], [np.array(x, np.float32) for x in c["xs"]]), c["out"])


def test_rounding_helpers():
    import torch
    x = np.random.default_rng(5).standard_normal(4096).astype(np.float32) * 100
    assert np.array_equal(oracle.round_to(x, oracle.BF16), torch.from_numpy(x).bfloat16().float().numpy())
    assert np.array_equal(oracle.round_to(x, oracle.F16), torch.from_numpy(x).half().float().numpy())


def test_fp8_e4m3_codec_known_answers():
    """FP8 E4M3 (OCP FN) codec of the oracle pinned on the format's published corner values: 0x7e = 448 (max), 0x08 = 2^-6 (min
    normal), 0x01 = 2^-9 (min subnormal), 0x38 = 1.0, 0xc0 = -2.0; quantisation rounds to nearest-even and saturates."""
    import oracle
    d = oracle.e4m3_decode
    assert d([0x7E])[0] == 448.0 and d([0x08])[0] == 2.0 ** -6 and d([0x01])[0] == 2.0 ** -9 and d([0x38])[0] == 1.0 and d([0xC0])[0] == -2.0
    assert np.isnan(d([0x7F])[0]) and d([0x00])[0] == 0.0
    q = oracle.e4m3_quantize
    assert q([1.0])[0] == 0x38 and q([-2.0])[0] == 0xC0 and q([1000.0])[0] == 0x7E and q([-1e9])[0] == 0xFE
    assert q([1.0625])[0] == 0x38 and q([1.1875])[0] == 0x3A  # ties: 1.0625 between 1.0 (even) and 1.125; 1.1875 between 1.125 and 1.25 (even)
    allc = np.array([c for c in range(256) if c not in (0x7F, 0xFF)], np.uint8)
    assert np.array_equal(q(d(allc)) & 0x7F, allc & 0x7F)  # every code is a fixed point (the sign of zero aside)
    w = np.random.default_rng(0).standard_normal((64, 32)).astype(np.float32)
    codes, scale = oracle.quantize_weight_fp8(w)
    assert np.abs(oracle.dequantize_fp8(codes, scale) - w).max() <= np.abs(w).max() * 2.0 ** -4 * 1.01
