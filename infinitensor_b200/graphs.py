"""Synthetic graph builders for the BASELINE.json configs, written against the GraphHandler API only
(the same calls OnnxStub makes, reference pyinfinitensor/src/pyinfinitensor/onnx.py:136-1117), so they
need no `onnx` package.  Every builder takes a duck-typed handler: `backend.GraphHandler` (the product)
or `oracle.graph_oracle.OracleHandler` (the CPU checker) -- one graph description, two executors.

  build_llama_decode   -- config C3/C5: Llama-7B-shape KV-cache decode step (op mix of
                          examples/python/llama_kvcache_inference.py:45-80 with fused RMSNorm / RoPE /
                          AttentionKVCache nodes), optionally tensor-parallel-sharded exactly where
                          examples/distributed/parallel_opt.py:9-247 cuts the graph.
  build_gpt2           -- config C2: GPT-2-small forward, B=1 S=128 (op mix of run_pytorch.py:89-108 export)
  build_resnet50       -- config C4: ResNet-50 forward (conv + BN + relu + pooling + fc)
  build_matmul         -- config C1: single MatMul
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np

F32, F16, BF16, I64, I32 = 1, 10, 16, 7, 6
FP8 = 17  # FP8 E4M3 weight codes
LINEAR = 0


# ---------------------------------------------------------------- storage helpers (host side)
def f32_to_bf16_bits(a: np.ndarray) -> np.ndarray:
    """float32 -> bf16 bit pattern (uint16), round-to-nearest-even."""
    u = np.ascontiguousarray(a, dtype=np.float32).view(np.uint32)
    r = ((u >> 16) & 1) + 0x7FFF
    return ((u + r) >> 16).astype(np.uint16)


def bf16_bits_to_f32(b: np.ndarray) -> np.ndarray:
    return (b.astype(np.uint32) << 16).view(np.float32)


def to_storage(a: np.ndarray, dtype: int) -> np.ndarray:
    """float32 values -> array in the tensor's storage dtype (what copyin_numpy expects)."""
    if dtype == F32:
        return np.ascontiguousarray(a, dtype=np.float32)
    if dtype == F16:
        return np.ascontiguousarray(a, dtype=np.float32).astype(np.float16)
    if dtype == BF16:
        return f32_to_bf16_bits(a)
    return np.ascontiguousarray(a)


def from_storage(a: np.ndarray, dtype: int) -> np.ndarray:
    if dtype == BF16:
        return bf16_bits_to_f32(a)
    if dtype == F16:
        return a.astype(np.float32)
    return a


# ---------------------------------------------------------------- Llama decode (C3 / C5)
@dataclass
class LlamaConfig:
    layers: int = 32
    d_model: int = 4096
    heads: int = 32
    head_dim: int = 128
    ffn: int = 11008
    vocab: int = 32000
    s_max: int = 1024
    batch: int = 16
    dtype: int = BF16
    fp8_weights: bool = False  # the seven projection matrices of every layer + the logits head as FP8 E4M3 codes + f32 column scales

    @staticmethod
    def tiny(dtype=BF16, layers=2, batch=4):
        return LlamaConfig(layers=layers, d_model=256, heads=2, head_dim=128, ffn=512, vocab=1000, s_max=64,
                           batch=batch, dtype=dtype)

    def weight_elems(self, world: int = 1) -> int:
        d, f = self.d_model, self.ffn
        per_layer = (4 * d * d + 3 * d * f) // world + 2 * d
        return self.layers * per_layer + d + self.vocab * d + self.vocab * d  # + embedding table

    def algorithmic_bytes(self, pos: int, world: int = 1) -> int:
        """weights read once + KV read/append + ~12 activation tensors per layer."""
        e = 2 if self.dtype in (F16, BF16) else 4
        d, f = self.d_model, self.ffn
        proj = self.layers * ((4 * d * d + 3 * d * f) // world) + d * self.vocab
        norms = self.layers * 2 * d + d
        kv = self.layers * 2 * self.batch * (self.heads // world) * self.head_dim * (pos + 2)
        act = self.layers * 12 * self.batch * d
        if self.fp8_weights:  # one byte per projection weight + 4 bytes per output column
            cols = self.layers * ((3 * d + 2 * f) // world + 2 * d) + self.vocab
            return proj + 4 * cols + (norms + kv + act) * e
        return (proj + norms + kv + act) * e


@dataclass
class LlamaGraph:
    cfg: LlamaConfig
    input_ids: object = None
    position_ids: object = None
    logits: object = None
    k_caches: list = field(default_factory=list)
    v_caches: list = field(default_factory=list)
    weights: dict = field(default_factory=dict)  # name -> (tensor, shape, kind, shard)


def build_llama_decode(h, cfg: LlamaConfig, world: int = 1, rank: int = 0) -> LlamaGraph:
    """One decode step.  world > 1 applies the reference's tensor-parallel cut (parallel_opt.py):
    q/k/v + gate/up column-split, o/down row-split followed by AllReduceSum, KV cache split by head,
    embedding / norms / logits replicated."""
    assert cfg.heads % world == 0 and cfg.ffn % world == 0
    B, d, H, dh, f, V, dt = cfg.batch, cfg.d_model, cfg.heads // world, cfg.head_dim, cfg.ffn // world, cfg.vocab, cfg.dtype
    dl = H * dh  # local attention width
    g = LlamaGraph(cfg)

    def weight(name, shape, kind, shard=None):
        t = h.tensor(list(shape), dt)
        t.set_weight()
        g.weights[name] = (t, tuple(shape), kind, shard)
        return t

    def linear(x_, name, shape, shard=None):
        """x . W: 16-bit weight, or (cfg.fp8_weights) FP8 E4M3 codes + per-column f32 scale dequantised inside the GEMM"""
        if not cfg.fp8_weights:
            return h.matmul(x_, weight(name, shape, "proj", shard), None, False, False, None, LINEAR)
        assert world == 1, "fp8 weights: single GPU"
        wq = h.tensor(list(shape), FP8)
        wq.set_weight()
        sc = h.tensor([shape[1]], F32)
        sc.set_weight()
        g.weights[name] = (wq, tuple(shape), "proj_fp8", shard)
        g.weights[name + ".scale"] = (sc, (shape[1],), "scale", None)
        return h.matmul(x_, wq, None, False, False, None, LINEAR, w_scale=sc)

    g.input_ids = h.tensor([B, 1], I64)
    g.position_ids = h.tensor([B, 1], I64)
    g.input_ids.set_input()
    g.position_ids.set_input()
    emb = weight("embed", (V, d), "embed")
    x = h.gather(emb, g.input_ids, None, 0)  # [B,1,d]
    for li in range(cfg.layers):
        p = f"l{li}."
        kc = h.tensor([B, H, cfg.s_max, dh], dt)
        vc = h.tensor([B, H, cfg.s_max, dh], dt)
        kc.set_input()
        vc.set_input()
        g.k_caches.append(kc)
        g.v_caches.append(vc)
        hn = h.RMSNorm(x, weight(p + "ln1", (d,), "norm"), None)
        q = linear(hn, p + "wq", (d, dl), ("col", d))
        k = linear(hn, p + "wk", (d, dl), ("col", d))
        v = linear(hn, p + "wv", (d, dl), ("col", d))
        q = h.RoPE(g.position_ids, q, None)
        k = h.RoPE(g.position_ids, k, None)

        def heads(t):
            t = h.reshape(t, None, [B, 1, H, dh])
            return h.transpose(t, None, [0, 2, 1, 3])  # [B,H,1,dh]

        attn = h.attentionKVCache(kc, vc, heads(q), heads(k), heads(v), g.position_ids, None)
        a = h.reshape(h.transpose(attn, None, [0, 2, 1, 3]), None, [B, 1, dl])
        o = linear(a, p + "wo", (dl, d), ("row", d))
        if world > 1:
            o = h.allReduceSum(o, None)
        x = h.add(x, o, None)
        hn = h.RMSNorm(x, weight(p + "ln2", (d,), "norm"), None)
        gate = linear(hn, p + "wg", (d, f), ("col", cfg.ffn))
        up = linear(hn, p + "wu", (d, f), ("col", cfg.ffn))
        m = h.mul(h.silu(gate, None), up, None)
        dn = linear(m, p + "wd", (f, d), ("row", cfg.ffn))
        if world > 1:
            dn = h.allReduceSum(dn, None)
        x = h.add(x, dn, None)
    xf = h.RMSNorm(x, weight("ln_f", (d,), "norm"), None)
    g.logits = linear(xf, "lm_head", (d, V))
    g.logits.set_output()
    return g


def llama_weight_values(name: str, shape, kind: str, seed: int = 0) -> np.ndarray:
    """SURVEY.md 8(d) synthetic inputs: weights ~ N(0, 0.02^2), norm weights 1 + N(0, 0.02^2); a
    per-tensor stream so every rank / executor draws identical full tensors."""
    rng = np.random.default_rng([seed, abs(hash_name(name))])
    w = rng.standard_normal(shape, dtype=np.float32) * np.float32(0.02)
    if kind == "norm":
        w = w + np.float32(1.0)
    return w


def hash_name(name: str) -> int:
    hv = 2166136261
    for ch in name.encode():
        hv = ((hv ^ ch) * 16777619) & 0xFFFFFFFF
    return hv


def shard_weight(full: np.ndarray, shard, world: int, rank: int) -> np.ndarray:
    """parallel_opt.py:21-59 slicing: column split -> W[:, r*N/w:(r+1)*N/w]; row split -> W[r*K/w:(r+1)*K/w, :]."""
    if shard is None or world == 1:
        return full
    mode = shard[0]
    if mode == "col":
        n = full.shape[1] // world
        return np.ascontiguousarray(full[:, rank * n:(rank + 1) * n])
    n = full.shape[0] // world
    return np.ascontiguousarray(full[rank * n:(rank + 1) * n, :])


def quantize_fp8_host(w: np.ndarray):
    """per-output-column symmetric FP8 E4M3 quantisation (same rule as oracle.quantize_weight_fp8, restated so that the product
    never imports the oracle): codes uint8 [K, N], scale f32 [N] = max|column| / 448; round to nearest, ties to even."""
    w = np.asarray(w, np.float32)
    scale = np.maximum(np.abs(w).max(axis=0), 1e-12).astype(np.float32) / np.float32(448.0)
    x = (w / scale[None, :]).astype(np.float32)
    table = np.zeros(127, np.float64)
    for c in range(127):
        e, m = (c >> 3) & 15, c & 7
        table[c] = (m / 8.0) * 2.0 ** -6 if e == 0 else (1.0 + m / 8.0) * 2.0 ** (e - 7)
    mag = np.minimum(np.abs(x).astype(np.float64), 448.0)
    hi = np.clip(np.searchsorted(table, mag, side="left"), 0, 126)
    lo = np.clip(hi - 1, 0, 126)
    dlo, dhi = mag - table[lo], table[hi] - mag
    code = np.where((dhi < dlo) | ((dhi == dlo) & (hi % 2 == 0)), hi, lo).astype(np.uint8)
    return np.where(np.signbit(x), code | 0x80, code).astype(np.uint8), scale


def fill_llama_weights_host(g: LlamaGraph, world: int = 1, rank: int = 0, seed: int = 0):
    """Small configs only (host RNG): full tensors are drawn, then sharded like parallel_opt.py does."""
    for name, (t, shape, kind, shard) in g.weights.items():
        if kind == "scale":
            continue  # written together with its codes
        if kind == "proj_fp8":
            codes, scale = quantize_fp8_host(llama_weight_values(name, list(shape), "proj", seed))
            t.copyin_numpy(codes)
            g.weights[name + ".scale"][0].copyin_numpy(scale)
            continue
        full_shape = list(shape)
        if shard is not None and world > 1:
            full_shape[1 if shard[0] == "col" else 0] *= world
        w = shard_weight(llama_weight_values(name, full_shape, kind, seed), shard, world, rank)
        t.copyin_numpy(to_storage(w, g.cfg.dtype))


def llama_cache_values(cfg: LlamaConfig, layer: int, which: str, world: int = 1, rank: int = 0, seed: int = 2):
    """KV cache pre-filled with N(0,1)*0.5; full-head tensor then the rank's head slice."""
    rng = np.random.default_rng([seed, layer, 0 if which == "k" else 1])
    full = rng.standard_normal((cfg.batch, cfg.heads, cfg.s_max, cfg.head_dim), dtype=np.float32) * np.float32(0.5)
    hl = cfg.heads // world
    return np.ascontiguousarray(full[:, rank * hl:(rank + 1) * hl])


# ---------------------------------------------------------------- single MatMul (C1)
def build_matmul(h, m=512, n=512, k=512, dtype=F32):
    a = h.tensor([m, k], dtype)
    b = h.tensor([k, n], dtype)
    a.set_input()
    b.set_input()
    c = h.matmul(a, b, None, False, False, None, LINEAR)
    c.set_output()
    return a, b, c


# ---------------------------------------------------------------- GPT-2 small (C2)
@dataclass
class GPT2Config:
    layers: int = 12
    d_model: int = 768
    heads: int = 12
    ffn: int = 3072
    vocab: int = 50257
    n_pos: int = 1024
    seq: int = 128
    batch: int = 1
    dtype: int = F16
    lm_head: bool = False  # the reference export is GPT2Model without LM head (run_pytorch.py:47,78-81)

    @staticmethod
    def tiny(dtype=F16):
        return GPT2Config(layers=2, d_model=64, heads=2, ffn=128, vocab=500, n_pos=64, seq=16, batch=1, dtype=dtype)


@dataclass
class GPT2Graph:
    cfg: GPT2Config
    input_ids: object = None
    position_ids: object = None
    out: object = None
    weights: dict = field(default_factory=dict)
    consts: dict = field(default_factory=dict)  # name -> (tensor, float32 ndarray)


def build_gpt2(h, cfg: GPT2Config) -> GPT2Graph:
    B, S, d, H, f, dt = cfg.batch, cfg.seq, cfg.d_model, cfg.heads, cfg.ffn, cfg.dtype
    dh = d // H
    g = GPT2Graph(cfg)

    def weight(name, shape, kind):
        t = h.tensor(list(shape), dt)
        t.set_weight()
        g.weights[name] = (t, tuple(shape), kind, None)
        return t

    def const(name, arr):
        t = h.tensor(list(arr.shape), dt)
        t.set_weight()
        g.consts[name] = (t, np.asarray(arr, np.float32))
        return t

    g.input_ids = h.tensor([B, S], I64)
    g.position_ids = h.tensor([B, S], I64)
    g.input_ids.set_input()
    g.position_ids.set_input()
    x = h.add(h.gather(weight("wte", (cfg.vocab, d), "embed"), g.input_ids, None, 0),
              h.gather(weight("wpe", (cfg.n_pos, d), "embed"), g.position_ids, None, 0), None)  # [B,S,d]
    # causal mask as an additive 0/-inf constant: the frontend's rewrite of Where(mask, x, -inf) (onnx.py:1055-1081)
    big_neg = -65504.0 if dt == F16 else -3.0e38
    mask = const("mask", np.triu(np.full((S, S), big_neg, np.float32), 1).reshape(1, 1, S, S))
    scale = const("scale", np.array([np.sqrt(dh)], np.float32))
    for li in range(cfg.layers):
        p = f"h{li}."
        hn = h.layerNormalization(x, weight(p + "ln1.w", (d,), "norm"), None, weight(p + "ln1.b", (d,), "bias"), 1e-5, -1, 1)
        qkv = h.matmul(hn, weight(p + "attn.w", (d, 3 * d), "proj"), None, False, False,
                       weight(p + "attn.b", (3 * d,), "bias"), LINEAR)
        q, k, v = h.split(qkv, None, 2, 3)

        def heads(t):
            return h.transpose(h.reshape(t, None, [B, S, H, dh]), None, [0, 2, 1, 3])  # [B,H,S,dh]

        q, k, v = heads(q), heads(k), heads(v)
        kt = h.transpose(k, None, [0, 1, 3, 2])
        s = h.div(h.matmul(q, kt, None, False, False, None, LINEAR), scale, None)
        pr = h.softmax(h.add(s, mask, None), None, -1)
        a = h.matmul(pr, v, None, False, False, None, LINEAR)  # [B,H,S,dh]
        a = h.reshape(h.transpose(a, None, [0, 2, 1, 3]), None, [B, S, d])
        o = h.matmul(a, weight(p + "proj.w", (d, d), "proj"), None, False, False, weight(p + "proj.b", (d,), "bias"), LINEAR)
        x = h.add(x, o, None)
        hn = h.layerNormalization(x, weight(p + "ln2.w", (d,), "norm"), None, weight(p + "ln2.b", (d,), "bias"), 1e-5, -1, 1)
        m = h.matmul(hn, weight(p + "fc.w", (d, f), "proj"), None, False, False, weight(p + "fc.b", (f,), "bias"), LINEAR)
        m = h.gelu(m, None)
        m = h.matmul(m, weight(p + "fc2.w", (f, d), "proj"), None, False, False, weight(p + "fc2.b", (d,), "bias"), LINEAR)
        x = h.add(x, m, None)
    x = h.layerNormalization(x, weight("ln_f.w", (d,), "norm"), None, weight("ln_f.b", (d,), "bias"), 1e-5, -1, 1)
    if cfg.lm_head:
        x = h.matmul(x, g.weights["wte"][0], None, False, True, None, LINEAR)  # tied embedding
    g.out = x
    g.out.set_output()
    return g


def fill_gpt2_weights_host(g: GPT2Graph, seed: int = 0):
    for name, (t, shape, kind, _) in g.weights.items():
        w = llama_weight_values(name, shape, "norm" if kind == "norm" else "proj", seed)
        if kind == "bias":
            w = w * np.float32(0.5)
        t.copyin_numpy(to_storage(w, g.cfg.dtype))
    for name, (t, arr) in g.consts.items():
        t.copyin_numpy(to_storage(arr, g.cfg.dtype))


# ---------------------------------------------------------------- ResNet-50 (C4)
@dataclass
class ResNetConfig:
    batch: int = 64
    image: int = 224
    dtype: int = F16
    blocks: tuple = (3, 4, 6, 3)
    width: int = 64
    classes: int = 1000

    @staticmethod
    def tiny(dtype=F16):
        return ResNetConfig(batch=2, image=32, dtype=dtype, blocks=(1, 1, 1, 1), width=8, classes=10)


@dataclass
class ResNetGraph:
    cfg: ResNetConfig
    input: object = None
    out: object = None
    weights: dict = field(default_factory=dict)  # name -> (tensor, shape, kind, dtype)


def build_resnet50(h, cfg: ResNetConfig) -> ResNetGraph:
    """conv -> BatchNormalization -> Relu bottlenecks exactly as torchvision's export lowers through the
    frontend (Conv bias-free, BN as its own node, GlobalAveragePool -> AveragePool k=(H,W), onnx.py:489-503,
    Gemm -> MatMul(transB) + bias)."""
    dt = cfg.dtype
    g = ResNetGraph(cfg)

    def weight(name, shape, kind, dtype=None):
        t = h.tensor(list(shape), dtype or dt)
        t.set_weight()
        g.weights[name] = (t, tuple(shape), kind, dtype or dt)
        return t

    def conv_bn(x, name, cin, cout, k, stride, pad, relu=True):
        w = weight(name + ".w", (cout, cin, k, k), "conv")
        y = h.conv(x, w, None, pad, pad, stride, stride, 1, 1)
        y = h.batchNormalization(y, None, weight(name + ".bn.mean", (cout,), "bn_mean", F32),
                                 weight(name + ".bn.var", (cout,), "bn_var", F32),
                                 weight(name + ".bn.g", (cout,), "bn_g", F32), weight(name + ".bn.b", (cout,), "bn_b", F32),
                                 0.9, 1e-5, False)
        return h.relu(y, None) if relu else y

    g.input = h.tensor([cfg.batch, 3, cfg.image, cfg.image], dt)
    g.input.set_input()
    w0 = cfg.width
    x = conv_bn(g.input, "stem", 3, w0, 7, 2, 3)
    x = h.maxPool(x, None, 3, 3, 1, 1, 1, 1, 2, 2, 0)
    cin = w0
    for si, nb in enumerate(cfg.blocks):
        mid = w0 * (2 ** si)
        cout = mid * 4
        for bi in range(nb):
            stride = 2 if (bi == 0 and si > 0) else 1
            p = f"s{si}b{bi}"
            idn = x
            y = conv_bn(x, p + ".c1", cin, mid, 1, 1, 0)
            y = conv_bn(y, p + ".c2", mid, mid, 3, stride, 1)
            y = conv_bn(y, p + ".c3", mid, cout, 1, 1, 0, relu=False)
            if bi == 0:
                idn = conv_bn(x, p + ".down", cin, cout, 1, stride, 0, relu=False)
            x = h.relu(h.add(y, idn, None), None)
            cin = cout
    hw = x.shape()[2]
    x = h.avgPool(x, None, hw, hw, 1, 1, 0, 0, 1, 1, 0)
    x = h.flatten(x, None, 1)
    x = h.matmul(x, weight("fc.w", (cfg.classes, cin), "fc"), None, False, True, weight("fc.b", (cfg.classes,), "bias"),
                 LINEAR)
    g.out = x
    g.out.set_output()
    return g


def resnet_weight_values(name, shape, kind, seed=0):
    rng = np.random.default_rng([seed, abs(hash_name(name))])
    if kind == "conv":  # Kaiming-normal
        fan_in = int(np.prod(shape[1:]))
        return rng.standard_normal(shape, dtype=np.float32) * np.float32(np.sqrt(2.0 / fan_in))
    if kind == "fc":
        return rng.standard_normal(shape, dtype=np.float32) * np.float32(0.02)
    if kind == "bn_var":
        return np.float32(1.0) + np.abs(rng.standard_normal(shape, dtype=np.float32)) * np.float32(0.1)
    if kind == "bn_g":
        return np.float32(1.0) + rng.standard_normal(shape, dtype=np.float32) * np.float32(0.1)
    return rng.standard_normal(shape, dtype=np.float32) * np.float32(0.1)  # bn_mean, bn_b, bias


def fill_resnet_weights_host(g: ResNetGraph, seed: int = 0):
    for name, (t, shape, kind, dtype) in g.weights.items():
        t.copyin_numpy(to_storage(resnet_weight_values(name, shape, kind, seed), dtype))
