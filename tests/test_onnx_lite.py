"""ONNX ingestion without the onnx package: wire format, node lowering, export -> import round trips.
CPU tier: graphs run on the oracle's handler; the GPU tier test lives in test_gpu_graph.py."""
import os
import struct
import sys

import numpy as np
import pytest


def _free_port():
    """an unused TCP port for the torchrun rendezvous of the gloo world-2 tests"""
    import socket
    with socket.socket() as sk:
        sk.bind(("127.0.0.1", 0))
        return sk.getsockname()[1]

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))  # (also run as a torchrun worker script)
from infinitensor_b200 import onnx_lite as X

F32, F16, BF16, I64 = 1, 10, 16, 7


def test_wire_roundtrip_all_attribute_and_tensor_forms():
    rng = np.random.default_rng(0)
    w = rng.standard_normal((3, 4)).astype(np.float32)
    h16 = rng.standard_normal((2, 5)).astype(np.float16)
    b16 = rng.integers(0, 65535, size=(7,)).astype(np.uint16)
    node = X.Node("Foo", ["a", "", "w"], ["y", "z"], "n0",
                  {"i": -3, "f": 0.25, "s": b"half_pixel", "ints": [1, -2, 3 << 40], "floats": [0.5, -1.5],
                   "t": X.TensorProto("tv", [2], I64, np.array([9, -9], np.int64))})
    g = X.Graph([node], [X.TensorProto("w", [3, 4], F32, w), X.TensorProto("h", [2, 5], F16, h16),
                         X.TensorProto("b", [7], BF16, b16)],
                [X.ValueInfo("a", F32, [0, 4])], [X.ValueInfo("y", F16, [2, 5])])
    m = X.load_model(X.save_model(X.Model(g, opset=17)))
    assert m.opset == 17 and m.ir_version == 8
    n = m.graph.nodes[0]
    assert (n.op_type, n.inputs, n.outputs, n.name) == ("Foo", ["a", "", "w"], ["y", "z"], "n0")
    assert n.attrs["i"] == -3 and n.attrs["f"] == 0.25 and n.attrs["s"] == b"half_pixel"
    assert n.attrs["ints"] == [1, -2, 3 << 40] and n.attrs["floats"] == [0.5, -1.5]
    assert n.attrs["t"].array.tolist() == [9, -9]
    got = {t.name: t for t in m.graph.initializers}
    assert np.array_equal(got["w"].array, w) and np.array_equal(got["h"].array, h16) and np.array_equal(got["b"].array, b16)
    assert got["h"].array.dtype == np.float16 and got["b"].array.dtype == np.uint16
    assert m.graph.inputs[0].dims == [0, 4] and m.graph.outputs[0].elem_type == F16


def test_typed_repeated_fields_as_other_writers_emit_them():
    """float_data (packed), int64_data (packed), fp16 in int32_data, unpacked repeated ints: not just raw_data."""
    def ld(no, b): return X._key(no, 2) + X._enc_varint(len(b)) + b
    dims = X._key(1, 0) + X._enc_varint(3)
    t_f = dims + X._key(2, 0) + X._enc_varint(F32) + ld(4, struct.pack("<3f", 1.5, -2.0, 3.25)) + ld(8, b"f")
    t_i = ld(1, X._enc_varint(3)) + X._key(2, 0) + X._enc_varint(I64) + ld(7, b"".join(X._enc_varint(v) for v in (7, -1, 1 << 33)))
    h = np.array([1.0, -0.5, 65504.0], np.float16)
    t_h = dims + X._key(2, 0) + X._enc_varint(F16) + b"".join(X._key(5, 0) + X._enc_varint(int(v)) for v in h.view(np.uint16))
    assert X.TensorProto.parse(t_f).array.tolist() == [1.5, -2.0, 3.25]
    assert X.TensorProto.parse(t_i).array.tolist() == [7, -1, 1 << 33]
    assert np.array_equal(X.TensorProto.parse(t_h).array, h)


def _oracle():
    from oracle.graph_oracle import OracleHandler
    return OracleHandler()


def test_cnn_lowering_matches_direct_oracle_calls():
    """Conv(+bias, asymmetric pads) -> BatchNorm -> Relu -> MaxPool -> GlobalAveragePool -> Flatten -> Gemm(transB) -> Softmax,
    lowered with the reference frontend's conventions (quirk ledger q12)."""
    import oracle as O
    rng = np.random.default_rng(1)
    x = rng.standard_normal((2, 3, 9, 9)).astype(np.float32)
    w = (rng.standard_normal((8, 3, 3, 3)) * 0.2).astype(np.float32)
    cb = rng.standard_normal(8).astype(np.float32)
    bn = [rng.uniform(0.5, 1.5, 8).astype(np.float32), rng.standard_normal(8).astype(np.float32) * 0.1,
          rng.standard_normal(8).astype(np.float32) * 0.1, rng.uniform(0.5, 1.5, 8).astype(np.float32)]  # scale bias mean var
    fc, fb = (rng.standard_normal((5, 8)) * 0.3).astype(np.float32), rng.standard_normal(5).astype(np.float32)
    T = lambda n, a: X.TensorProto(n, list(a.shape), F32, a)
    nodes = [X.Node("Conv", ["x", "w", "cb"], ["c"], "", {"pads": [1, 0, 2, 1], "strides": [1, 1], "dilations": [1, 1]}),
             X.Node("BatchNormalization", ["c", "s", "b", "m", "v"], ["n"], "", {"epsilon": 1e-5}),
             X.Node("Relu", ["n"], ["r"]), X.Node("Dropout", ["r"], ["d"]),
             X.Node("MaxPool", ["d"], ["p"], "", {"kernel_shape": [2, 2], "strides": [2, 2]}),
             X.Node("GlobalAveragePool", ["p"], ["g"]), X.Node("Flatten", ["g"], ["f"], "", {"axis": 1}),
             X.Node("Gemm", ["f", "fc", "fb"], ["l"], "", {"transB": 1}), X.Node("Softmax", ["l"], ["y"], "", {"axis": -1})]
    g = X.Graph(nodes, [T("w", w), T("cb", cb), T("s", bn[0]), T("b", bn[1]), T("m", bn[2]), T("v", bn[3]), T("fc", fc),
                        T("fb", fb)], [X.ValueInfo("x", F32, [2, 3, 9, 9])], [X.ValueInfo("y", F32, [2, 5])])
    stub = X.OnnxStub(X.save_model(X.Model(g)), handler=_oracle())
    stub.inputs["x"].copyin_numpy(x)
    stub.run()
    xp = np.pad(x, ((0, 0), (0, 0), (1, 2), (0, 1)))
    c = O.binary("add", O.conv2d(xp, w, 0, 0, 1, 1, 1, 1), cb.reshape(1, 8, 1, 1))
    r = np.maximum(O.batch_norm(c, bn[2], bn[3], bn[0], bn[1], 1e-5), 0)
    p = O.pool2d("max", r, 2, 2, 1, 1, 0, 0, 2, 2)
    gap = O.pool2d("avg", p, p.shape[2], p.shape[3], 1, 1, 0, 0, 1, 1).reshape(2, 8)
    ref = O.softmax(O.matmul(gap, fc, fb, False, True), -1)
    np.testing.assert_array_equal(stub.outputs["y"].copyout_numpy(), ref)
    assert len(stub.handler.ops) == 12  # + Pad, Reshape(bias), Add; Dropout -> Identity


@pytest.mark.parametrize("dtype", [F32, BF16])
def test_llama_decode_export_import_roundtrip(dtype):
    """build_llama_decode through OnnxExporter -> ONNX bytes -> OnnxStub on a fresh handler: same schedule, same logits."""
    from infinitensor_b200 import graphs as G
    cfg = G.LlamaConfig(layers=2, d_model=256, heads=2, head_dim=128, ffn=384, vocab=96, s_max=16, batch=3, dtype=dtype)
    exp = X.OnnxExporter(_oracle())
    ge = G.build_llama_decode(exp, cfg)
    exp.data_malloc()
    G.fill_llama_weights_host(ge)          # weight uploads are captured as initializers
    blob = exp.save()
    m = X.load_model(blob)
    assert {n.op_type for n in m.graph.nodes} >= {"MatMul", "RMSNorm", "RoPE", "AttentionKVCache", "Gather", "Reshape", "Silu", "Mul", "Add"}
    stub = X.OnnxStub(blob, handler=_oracle())
    oh = _oracle()                          # the same graph built directly
    gd = G.build_llama_decode(oh, cfg)
    oh.data_malloc()
    G.fill_llama_weights_host(gd)
    rng = np.random.default_rng(5)
    names_in = [t.name for t in exp._inputs]
    assert len(stub.inputs) == len(names_in)
    # feed token ids, positions and caches: the builder's input order is the exporter's
    gi = [t for t in oh.tensors if t.is_input]
    assert len(gi) == len(names_in)
    for name, t in zip(names_in, gi):
        shp, dt = t.shape(), t.dtype()
        if dt == I64:
            v = rng.integers(0, 7, size=shp).astype(np.int64)
        else:
            v = G.to_storage((rng.standard_normal(shp) * 0.5).astype(np.float32), dt)
        t.copyin_numpy(v)
        stub.inputs[name].copyin_numpy(v)
    oh.run()
    stub.run()
    out_direct = [t for t in oh.tensors if t.is_output]
    assert len(out_direct) == len(stub.outputs)
    for t, (name, s) in zip(out_direct, stub.outputs.items()):
        np.testing.assert_array_equal(t.copyout_numpy(), s.copyout_numpy())


def test_exported_graph_plans_and_schedules_like_the_direct_build():
    """Through the C++ host (planning-only runtime): the imported ONNX graph gets the same fused schedule."""
    from infinitensor_b200 import backend as B, graphs as G
    rt = B.HostPlanRuntime()
    cfg = G.LlamaConfig(layers=2, d_model=512, heads=4, head_dim=128, ffn=1024, vocab=128, s_max=32, batch=16)
    h = B.GraphHandler(rt)
    G.build_llama_decode(h, cfg)
    exp = X.OnnxExporter(B.GraphHandler(rt))
    G.build_llama_decode(exp, cfg)
    stub_h = B.GraphHandler(rt)
    X.OnnxStub(exp.save(), handler=stub_h, upload=False)  # plan only: the planning runtime has no device to upload to
    assert stub_h.schedule() == h.schedule()
    res = B.GraphHandler(rt)
    expr = X.OnnxExporter(res)
    G.build_resnet50(expr, G.ResNetConfig(batch=2, image=64))
    h2 = B.GraphHandler(rt)
    X.OnnxStub(expr.save(), handler=h2, upload=False)
    assert h2.schedule() == res.schedule() and sum(x.startswith("ConvBnAct") for x in h2.schedule()) == 53


def test_unsupported_and_dynamic_inputs_are_refused_loudly():
    g = X.Graph([X.Node("Resize", ["x"], ["y"])], [], [X.ValueInfo("x", F32, [1, 3, 4, 4])], [X.ValueInfo("y", F32, [1, 3, 8, 8])])
    with pytest.raises(NotImplementedError, match="Resize"):
        X.OnnxStub(X.Model(g), handler=_oracle())
    g = X.Graph([X.Node("Reshape", ["x", "s"], ["y"])], [], [X.ValueInfo("x", F32, [2, 3]), X.ValueInfo("s", I64, [1])],
                [X.ValueInfo("y", F32, [6])])
    with pytest.raises(NotImplementedError, match="constant"):
        X.OnnxStub(X.Model(g), handler=_oracle())
    g = X.Graph([X.Node("Relu", ["nope"], ["y"])], [], [X.ValueInfo("x", F32, [2])], [X.ValueInfo("y", F32, [2])])
    with pytest.raises(ValueError, match="missing input"):
        X.OnnxStub(X.Model(g), handler=_oracle())


def _roundtrip(build, fill, cfg):
    exp = X.OnnxExporter(_oracle())
    ge = build(exp, cfg)
    exp.data_malloc()
    fill(ge)
    blob = exp.save()
    stub = X.OnnxStub(blob, handler=_oracle())
    oh = _oracle()
    gd = build(oh, cfg)
    oh.data_malloc()
    fill(gd)
    rng = np.random.default_rng(11)
    from infinitensor_b200 import graphs as G
    for name, t in zip([t.name for t in exp._inputs], [t for t in oh.tensors if t.is_input]):
        shp, dt = t.shape(), t.dtype()
        v = rng.integers(0, 7, size=shp).astype(np.int64) if dt == I64 else G.to_storage(rng.standard_normal(shp).astype(np.float32), dt)
        t.copyin_numpy(v)
        stub.inputs[name].copyin_numpy(v)
    oh.run()
    stub.run()
    outs = [t for t in oh.tensors if t.is_output]
    assert len(outs) == len(stub.outputs) >= 1
    for t, s_ in zip(outs, stub.outputs.values()):
        np.testing.assert_array_equal(t.copyout_numpy(), s_.copyout_numpy())
    return X.load_model(blob)


def test_gpt2_and_resnet_export_import_roundtrip():
    from infinitensor_b200 import graphs as G
    m = _roundtrip(G.build_gpt2, G.fill_gpt2_weights_host, G.GPT2Config.tiny(F32))
    assert {"LayerNormalization", "Softmax", "Transpose", "Gelu", "Gather"} <= {n.op_type for n in m.graph.nodes}
    m = _roundtrip(G.build_resnet50, G.fill_resnet_weights_host, G.ResNetConfig.tiny(F32))
    assert {"Conv", "BatchNormalization", "Relu", "MaxPool", "AveragePool", "Gemm"} <= {n.op_type for n in m.graph.nodes}


def test_constant_folding_replaces_onnxsim_for_exporter_shape_arithmetic():
    """Shape -> Gather -> Unsqueeze -> Concat -> Reshape and a Range / Less / Cast / Mul mask builder (what torch.onnx.export
    leaves behind and the reference removes with onnxsim) are evaluated at load time; only real compute reaches the handler."""
    import oracle as O
    C = lambda name, arr: X.Node("Constant", [], [name], "", {"value": X.TensorProto("", list(np.shape(arr)), I64 if np.asarray(arr).dtype == np.int64 else F32, np.asarray(arr))})
    nodes = [
        X.Node("Shape", ["x"], ["s"]), C("i0", np.array(0, np.int64)), X.Node("Gather", ["s", "i0"], ["b"], "", {"axis": 0}),
        X.Node("Unsqueeze", ["b"], ["b1"], "", {"axes": [0]}), C("m1", np.array([-1], np.int64)),
        X.Node("Concat", ["b1", "m1"], ["shp"], "", {"axis": 0}), X.Node("Reshape", ["x", "shp"], ["y"]),
        X.Node("Shape", ["y"], ["sy"], "", {"start": 1}), C("zero", np.array(0, np.int64)), C("one", np.array(1, np.int64)),
        X.Node("Squeeze", ["sy"], ["n"]), X.Node("Range", ["zero", "n", "one"], ["r"]), C("half", np.array(12, np.int64)),
        X.Node("Less", ["r", "half"], ["lt"]), X.Node("Cast", ["lt"], ["ltf"], "", {"to": F32}), C("neg", np.array(-2.0, np.float32)),
        X.Node("Mul", ["ltf", "neg"], ["bias"]), X.Node("Add", ["y", "bias"], ["z"]), X.Node("Softmax", ["z"], ["out"], "", {"axis": -1}),
        # a constant-only branch nobody consumes must not create device tensors
        X.Node("ConstantOfShape", ["shp0"], ["dead"], "", {"value": X.TensorProto("", [1], F32, np.array([3.0], np.float32))}),
    ]
    g = X.Graph(nodes, [X.TensorProto("shp0", [2], I64, np.array([4, 5], np.int64))], [X.ValueInfo("x", F32, [2, 3, 8])],
                [X.ValueInfo("out", F32, [2, 24])])
    stub = X.OnnxStub(X.save_model(X.Model(g)), handler=_oracle())
    assert stub.folded == ["Shape", "Gather", "Unsqueeze", "Concat", "Shape", "Squeeze", "Range", "Less", "Cast", "Mul", "ConstantOfShape"]
    assert len(stub.handler.ops) == 3                      # Reshape, Add, Softmax
    assert sorted(stub._data) == ["bias"]                  # the only constant that became a device weight
    x = np.random.default_rng(2).standard_normal((2, 3, 8)).astype(np.float32)
    stub.inputs["x"].copyin_numpy(x)
    stub.run()
    bias = np.where(np.arange(24) < 12, np.float32(-2.0), np.float32(0.0)).astype(np.float32)
    np.testing.assert_array_equal(stub.outputs["out"].copyout_numpy(), O.softmax(O.binary("add", x.reshape(2, 24), bias), -1))


def test_fold_helpers_follow_onnx_semantics():
    f = X.OnnxStub._FOLD
    a = np.arange(24, dtype=np.int64).reshape(2, 3, 4)
    assert np.array_equal(X.OnnxStub._slice_np([a, np.array([1]), np.array([2 ** 40]), np.array([2]), np.array([2])], {}), a[:, :, 1::2])
    assert np.array_equal(X.OnnxStub._slice_np([a, np.array([-1]), np.array([-5]), np.array([1]), np.array([-1])], {}), a[:, ::-1, :][:, :3, :])
    assert f["Div"]([np.array([7, -7], np.int64), np.array([2, 2], np.int64)], {}).tolist() == [3, -3]  # truncation, not floor
    assert f["Expand"]([np.array([[1], [2]], np.int64), np.array([1, 3], np.int64)], {}).tolist() == [[1, 1, 1], [2, 2, 2]]
    assert f["Reshape"]([a, np.array([0, -1], np.int64)], {}).shape == (2, 12)
    assert f["Squeeze"]([np.zeros((1, 3, 1))], {}).shape == (3,) and f["Squeeze"]([np.zeros((1, 3, 1)), np.array([0])], {}).shape == (3, 1)
    assert f["ConstantOfShape"]([np.array([2, 2], np.int64)], {}).dtype == np.float32
    assert f["Where"]([np.array([True, False]), np.array([1, 2]), np.array([3, 4])], {}).tolist() == [1, 4]


def test_hf_style_gpt2_block_with_exporter_shape_arithmetic():
    """One GPT-2 block the way torch.onnx.export writes it (Conv1D as MatMul + Add, Split, Shape-driven Reshapes, a causal mask
    sliced out of a constant tril buffer with Shape / Sub / Slice, Where, the tanh GELU spelled with Pow / Tanh): loads without
    onnx / onnxsim, folds every shape computation, and matches a direct numpy transcription."""
    rng = np.random.default_rng(7)
    S, D, H, P = 6, 16, 2, 8
    dh = D // H
    W = lambda *s: (rng.standard_normal(s) * 0.2).astype(np.float32)
    init = {"ln_g": 1 + W(D) * 0.1, "ln_b": W(D) * 0.1, "w_qkv": W(D, 3 * D), "b_qkv": W(3 * D), "w_o": W(D, D), "b_o": W(D),
            "w_fc": W(D, 4 * D), "b_fc": W(4 * D), "w_pr": W(4 * D, D), "b_pr": W(D),
            "tril": np.tril(np.ones((1, 1, P, P), np.bool_)), "neg": np.array(-1e4, np.float32),
            "scale": np.array(np.sqrt(dh), np.float32), "three": np.array(3.0, np.float32), "c0": np.array(0.044715, np.float32),
            "c1": np.array(np.sqrt(2 / np.pi), np.float32), "one": np.array(1.0, np.float32), "halfc": np.array(0.5, np.float32)}
    I = lambda v: X.TensorProto("", list(np.shape(v)), I64, np.asarray(v, np.int64))
    C = lambda name, v: X.Node("Constant", [], [name], "", {"value": I(v)})
    N = X.Node
    nodes = [
        N("LayerNormalization", ["x", "ln_g", "ln_b"], ["h"], "", {"axis": -1, "epsilon": 1e-5}),
        N("MatMul", ["h", "w_qkv"], ["qkv0"]), N("Add", ["qkv0", "b_qkv"], ["qkv"]),
        N("Split", ["qkv", "split3"], ["q", "k", "v"], "", {"axis": 2}), C("split3", [D, D, D]),
        # new_shape = concat(shape(q)[:-1], [H, dh])
        N("Shape", ["q"], ["qs"]), C("i0", 0), C("i1", 1), N("Gather", ["qs", "i0"], ["bdim"]), N("Gather", ["qs", "i1"], ["sdim"]),
        N("Unsqueeze", ["bdim", "ax0"], ["b1"]), N("Unsqueeze", ["sdim", "ax0"], ["s1"]), C("ax0", [0]), C("hd", [H, dh]),
        N("Concat", ["b1", "s1", "hd"], ["hshape"], "", {"axis": 0}),
        N("Reshape", ["q", "hshape"], ["q4"]), N("Reshape", ["k", "hshape"], ["k4"]), N("Reshape", ["v", "hshape"], ["v4"]),
        N("Transpose", ["q4"], ["qt"], "", {"perm": [0, 2, 1, 3]}), N("Transpose", ["k4"], ["kt"], "", {"perm": [0, 2, 3, 1]}),
        N("Transpose", ["v4"], ["vt"], "", {"perm": [0, 2, 1, 3]}),
        N("MatMul", ["qt", "kt"], ["sc0"]), N("Div", ["sc0", "scale"], ["sc"]),
        # causal mask = tril[:, :, k_len - q_len : k_len, :k_len]
        N("Shape", ["sc"], ["scs"]), C("im2", -2), C("im1", -1), N("Gather", ["scs", "im2"], ["qlen"]), N("Gather", ["scs", "im1"], ["klen"]),
        N("Sub", ["klen", "qlen"], ["off"]), N("Unsqueeze", ["off", "ax0"], ["off1"]), N("Unsqueeze", ["klen", "ax0"], ["kl1"]),
        C("ax2", [2]), C("ax3", [3]), C("z1", [0]), C("st1", [1]),
        N("Slice", ["tril", "off1", "kl1", "ax2", "st1"], ["m0"]), N("Slice", ["m0", "z1", "kl1", "ax3", "st1"], ["mask"]),
        N("Where", ["mask", "sc", "neg"], ["scm"]), N("Softmax", ["scm"], ["pr"], "", {"axis": -1}),
        N("MatMul", ["pr", "vt"], ["ctx"]), N("Transpose", ["ctx"], ["ctxt"], "", {"perm": [0, 2, 1, 3]}),
        C("dmodel", [D]), N("Concat", ["b1", "s1", "dmodel"], ["mshape"], "", {"axis": 0}), N("Reshape", ["ctxt", "mshape"], ["merged"]),
        N("MatMul", ["merged", "w_o"], ["o0"]), N("Add", ["o0", "b_o"], ["o"]), N("Add", ["x", "o"], ["r1"]),
        N("MatMul", ["r1", "w_fc"], ["f0"]), N("Add", ["f0", "b_fc"], ["f"]),
        N("Pow", ["f", "three"], ["f3"]), N("Mul", ["f3", "c0"], ["f3c"]), N("Add", ["f", "f3c"], ["inner"]), N("Mul", ["inner", "c1"], ["arg"]),
        N("Tanh", ["arg"], ["th"]), N("Add", ["th", "one"], ["th1"]), N("Mul", ["f", "halfc"], ["fh"]), N("Mul", ["fh", "th1"], ["gelu"]),
        N("MatMul", ["gelu", "w_pr"], ["p0"]), N("Add", ["p0", "b_pr"], ["pp"]), N("Add", ["r1", "pp"], ["y"]),
    ]
    dt = lambda v: 9 if v.dtype == np.bool_ else F32
    g = X.Graph(nodes, [X.TensorProto(k, list(v.shape), dt(v), v) for k, v in init.items()],
                [X.ValueInfo("x", F32, [1, S, D])], [X.ValueInfo("y", F32, [1, S, D])])
    blob = X.save_model(X.Model(g))
    # the C++ host accepts the lowered graph too (shape inference + memory plan on the planning-only runtime)
    from infinitensor_b200 import backend as B
    hp = B.GraphHandler(B.HostPlanRuntime())
    sp = X.OnnxStub(blob, handler=hp, upload=False)
    hp.data_malloc()
    assert sp.outputs["y"].shape() == [1, S, D] and not any(s_.startswith("Single:Shape") for s_ in hp.schedule())
    stub = X.OnnxStub(blob, handler=_oracle())
    lowered = {"Shape", "Gather", "Unsqueeze", "Concat", "Sub", "Slice"}
    assert lowered <= set(stub.folded) and stub.folded.count("Slice") == 2
    x = rng.standard_normal((1, S, D)).astype(np.float32)
    stub.inputs["x"].copyin_numpy(x)
    stub.run()
    # float64 transcription
    f8 = {k: v.astype(np.float64) for k, v in init.items() if v.dtype == np.float32}
    xx = x.astype(np.float64)
    mu, var = xx.mean(-1, keepdims=True), xx.var(-1, keepdims=True)
    hh = (xx - mu) / np.sqrt(var + 1e-5) * f8["ln_g"] + f8["ln_b"]
    qkv = hh @ f8["w_qkv"] + f8["b_qkv"]
    q, k, v = [t.reshape(1, S, H, dh).transpose(0, 2, 1, 3) for t in np.split(qkv, 3, axis=2)]
    sc = q @ k.transpose(0, 1, 3, 2) / np.sqrt(dh)
    sc = np.where(init["tril"][:, :, :S, :S], sc, -1e4)
    pr = np.exp(sc - sc.max(-1, keepdims=True)); pr /= pr.sum(-1, keepdims=True)
    o = (pr @ v).transpose(0, 2, 1, 3).reshape(1, S, D) @ f8["w_o"] + f8["b_o"]
    r1 = xx + o
    f = r1 @ f8["w_fc"] + f8["b_fc"]
    gelu = 0.5 * f * (1 + np.tanh(np.sqrt(2 / np.pi) * (f + 0.044715 * f ** 3)))
    ref = r1 + gelu @ f8["w_pr"] + f8["b_pr"]
    np.testing.assert_allclose(stub.outputs["y"].copyout_numpy(), ref, rtol=2e-4, atol=2e-5)


def test_simplify_removes_constant_subgraphs():
    from oracle.graph_oracle import OracleHandler
    nodes = [X.Node("Shape", ["x"], ["s"]), X.Node("Constant", [], ["i"], "", {"value": X.TensorProto("", [], I64, np.array(1, np.int64))}),
             X.Node("Gather", ["s", "i"], ["n"]), X.Node("Unsqueeze", ["n"], ["n1"], "", {"axes": [0]}),
             X.Node("Constant", [], ["m1"], "", {"value": X.TensorProto("", [1], I64, np.array([-1], np.int64))}),
             X.Node("Concat", ["n1", "m1"], ["shp"], "", {"axis": 0}), X.Node("Reshape", ["x", "shp"], ["y"]), X.Node("Relu", ["y"], ["z"])]
    g = X.Graph(nodes, [], [X.ValueInfo("x", F32, [2, 3, 4])], [X.ValueInfo("z", F32, [3, 8])])
    m = X.simplify(X.Model(g), OracleHandler)
    assert [n.op_type for n in m.graph.nodes] == ["Reshape", "Relu"]
    assert {t.name: t.array.tolist() for t in m.graph.initializers} == {"shp": [3, -1]}
    assert X.tensor_shapes(m, OracleHandler)["z"] == [3, 8]
    m2 = X.load_model(X.save_model(m))  # and it still round-trips through the wire format
    assert [n.op_type for n in m2.graph.nodes] == ["Reshape", "Relu"]
    stub = X.OnnxStub(X.Model(g), handler=OracleHandler())
    m3 = X.load_model(stub.to_onnx())       # OnnxStub.to_onnx(): the same folded model
    assert [n.op_type for n in m3.graph.nodes] == ["Reshape", "Relu"] and [t.name for t in m3.graph.initializers] == ["shp"]


def _tp_gpt2_inputs(cfg):
    ids = np.random.default_rng(1).integers(0, cfg.vocab, size=(cfg.batch, cfg.seq)).astype(np.int64)
    return ids, np.arange(cfg.seq, dtype=np.int64).reshape(1, -1).repeat(cfg.batch, 0)


def _tp_onnx_worker():
    """one rank of a world-size-2 gloo run: the UNSHARDED Llama ONNX file, rewritten by parallel_model for this rank"""
    import os
    import torch.distributed as dist
    from infinitensor_b200 import graphs as G
    from oracle.graph_oracle import OracleHandler
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    if os.environ.get("TP_MODEL") == "gpt2":
        cfg = G.GPT2Config.tiny(1)
        exp = X.OnnxExporter(OracleHandler())
        ge = G.build_gpt2(exp, cfg)
        exp.data_malloc()
        G.fill_gpt2_weights_host(ge)
        sharded = X.parallel_model(exp.save(), world, rank, OracleHandler)
        assert [n.op_type for n in sharded.graph.nodes].count("ReduceSum") == 2 * cfg.layers
        stub = X.OnnxStub(X.save_model(sharded), handler=OracleHandler())
        ids, pos = _tp_gpt2_inputs(cfg)
        stub.inputs[ge.input_ids.name].copyin_numpy(ids)
        stub.inputs[ge.position_ids.name].copyin_numpy(pos)
        stub.run()
        np.save(os.environ["TP_OUT"] + f".{rank}.npy", stub.outputs[ge.out.name].f32())
        dist.barrier()
        return
    cfg = G.LlamaConfig.tiny(dtype=1, layers=2, batch=3)
    exp = X.OnnxExporter(OracleHandler())
    ge = G.build_llama_decode(exp, cfg)
    exp.data_malloc()
    G.fill_llama_weights_host(ge)
    names = {"ids": ge.input_ids.name, "pos": ge.position_ids.name, "k": [t.name for t in ge.k_caches], "v": [t.name for t in ge.v_caches],
             "logits": ge.logits.name}
    sharded = X.parallel_model(exp.save(), world, rank, OracleHandler)
    ops = [n.op_type for n in sharded.graph.nodes]
    assert ops.count("ReduceSum") == 2 * cfg.layers, ops
    stub = X.OnnxStub(X.save_model(sharded), handler=OracleHandler())
    for li in range(cfg.layers):
        stub.inputs[names["k"][li]].copyin_numpy(G.llama_cache_values(cfg, li, "k", world, rank))
        stub.inputs[names["v"][li]].copyin_numpy(G.llama_cache_values(cfg, li, "v", world, rank))
    stub.inputs[names["ids"]].copyin_numpy(np.array([[1], [5], [7]], np.int64))
    stub.inputs[names["pos"]].copyin_numpy(np.full((3, 1), 9, np.int64))
    stub.run()
    np.save(os.environ["TP_OUT"] + f".{rank}.npy", stub.outputs[names["logits"]].f32())
    dist.barrier()


def test_parallel_model_rewrite_gloo_world2(tmp_path):
    """parallel_opt.py's job on the package-free ONNX classes: the unsharded Llama file rewritten per rank (column / row split,
    ReduceSum(communicator) -> AllReduceSum, head-sharded caches, Reshape targets) reproduces the single-rank logits."""
    import os
    import subprocess
    import sys
    from infinitensor_b200 import graphs as G
    from oracle.graph_oracle import OracleHandler
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = str(tmp_path / "tp_onnx")
    env = dict(os.environ, TP_OUT=out, MASTER_ADDR="127.0.0.1", OMP_NUM_THREADS="2")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", str(_free_port()), os.path.abspath(__file__), "--tp-onnx-worker"]
    subprocess.run(cmd, check=True, env=env, cwd=root, timeout=240)
    cfg = G.LlamaConfig.tiny(dtype=1, layers=2, batch=3)
    oh = OracleHandler()
    g = G.build_llama_decode(oh, cfg)
    G.fill_llama_weights_host(g)
    for li in range(cfg.layers):
        g.k_caches[li].copyin_numpy(G.llama_cache_values(cfg, li, "k"))
        g.v_caches[li].copyin_numpy(G.llama_cache_values(cfg, li, "v"))
    g.input_ids.copyin_numpy(np.array([[1], [5], [7]], np.int64))
    g.position_ids.copyin_numpy(np.full((3, 1), 9, np.int64))
    oh.run()
    for r in range(2):
        np.testing.assert_allclose(np.load(out + f".{r}.npy"), g.logits.f32(), rtol=1e-4, atol=1e-5)


def test_parallel_model_gpt2_packed_qkv_gloo_world2(tmp_path):
    """The generic rules on a second topology: packed q/k/v Gemm split per group before its Split, sharded Gemm biases,
    head-sharded activation x activation MatMuls, Softmax, the bias of a row-split Gemm added once after the all-reduce."""
    import subprocess
    from infinitensor_b200 import graphs as G
    from oracle.graph_oracle import OracleHandler
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = str(tmp_path / "tp_gpt2")
    env = dict(os.environ, TP_OUT=out, TP_MODEL="gpt2", MASTER_ADDR="127.0.0.1", OMP_NUM_THREADS="2")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", str(_free_port()), os.path.abspath(__file__), "--tp-onnx-worker"]
    subprocess.run(cmd, check=True, env=env, cwd=root, timeout=240)
    cfg = G.GPT2Config.tiny(1)
    oh = OracleHandler()
    g = G.build_gpt2(oh, cfg)
    oh.data_malloc()
    G.fill_gpt2_weights_host(g)
    ids, pos = _tp_gpt2_inputs(cfg)
    g.input_ids.copyin_numpy(ids)
    g.position_ids.copyin_numpy(pos)
    oh.run()
    for r in range(2):
        np.testing.assert_allclose(np.load(out + f".{r}.npy"), g.out.f32(), rtol=2e-4, atol=2e-5)


if __name__ == "__main__":
    import os
    import sys
    if "--tp-onnx-worker" in sys.argv:
        sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
        _tp_onnx_worker()


def test_parallel_model_schedule_matches_the_direct_tp_build():
    """On the C++ host: the rank-1-of-2 rewrite of the unsharded file plans and schedules exactly like the graph that
    graphs.build_llama_decode(world=2, rank=1) builds directly (AllReduce + Add + RMSNorm steps included)."""
    from infinitensor_b200 import backend as B, graphs as G
    rt = B.HostPlanRuntime()
    cfg = G.LlamaConfig(layers=2, d_model=512, heads=4, head_dim=128, ffn=1024, vocab=128, s_max=32, batch=16)
    direct = B.GraphHandler(rt)
    G.build_llama_decode(direct, cfg, 2, 1)
    exp = X.OnnxExporter(B.GraphHandler(rt))
    G.build_llama_decode(exp, cfg)
    sharded = X.parallel_model(exp.save(), 2, 1)
    h = B.GraphHandler(rt)
    X.OnnxStub(sharded, handler=h, upload=False)
    assert h.schedule() == direct.schedule()
    assert sum(s.startswith("AllReduceAddNorm") for s in h.schedule()) == 4
    h.data_malloc()
    direct.data_malloc()
    assert h.arena_bytes() == direct.arena_bytes()


def test_where_with_constant_mask_and_neg_inf_becomes_add():
    """reference onnx.py:1055-1081: Where(const mask, x, -inf) is lowered as x + (0 / -inf bias); other Wheres stay selects."""
    mask = np.tril(np.ones((1, 1, 4, 4), np.bool_))
    inits = [X.TensorProto("mask", [1, 1, 4, 4], 9, mask), X.TensorProto("ninf", [], F32, np.array(-np.inf, np.float32)),
             X.TensorProto("small", [], F32, np.array(-1e4, np.float32))]
    g = X.Graph([X.Node("Where", ["mask", "x", "ninf"], ["a"]), X.Node("Where", ["mask", "x", "small"], ["b"])], inits,
                [X.ValueInfo("x", F32, [2, 3, 4, 4])], [X.ValueInfo("a", F32, [2, 3, 4, 4]), X.ValueInfo("b", F32, [2, 3, 4, 4])])
    stub = X.OnnxStub(X.Model(g), handler=_oracle())
    assert "mask_alt" in stub._data and stub._data["mask_alt"].array.dtype == np.float32
    x = np.random.default_rng(0).standard_normal((2, 3, 4, 4)).astype(np.float32)
    stub.inputs["x"].copyin_numpy(x)
    stub.run()
    np.testing.assert_array_equal(stub.outputs["a"].copyout_numpy(), np.where(mask, x, -np.inf).astype(np.float32))
    np.testing.assert_array_equal(stub.outputs["b"].copyout_numpy(), np.where(mask, x, np.float32(-1e4)))
