#!/usr/bin/env python
"""Run an ONNX file on 1..N B200s -- the counterpart of the reference's examples/distributed/launch.py, without the
`onnx` package:

    python tools/launch_onnx.py model.onnx                                      # one GPU
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/launch_onnx.py model.onnx

Each rank rewrites the UNSHARDED file for itself with onnx_lite.parallel_model (column / row split + all-reduce), lowers
it with OnnxStub onto its CudaRuntime, sets up NCCL + the NVLink peer-memory all-reduce, feeds seeded random inputs and
replays the CUDA graph; rank 0 prints one JSON line (ms per run, device-timed, max over ranks).
`--export-llama out.onnx` writes a small Llama-decode file with graphs.build_llama_decode through OnnxExporter first."""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np


def export_llama(path, layers, d_model, heads, ffn, vocab, s_max, batch):
    from infinitensor_b200 import backend as B, graphs as G, onnx_lite as X
    cfg = G.LlamaConfig(layers=layers, d_model=d_model, heads=heads, head_dim=128, ffn=ffn, vocab=vocab, s_max=s_max, batch=batch)
    exp = X.OnnxExporter(B.GraphHandler(B.HostPlanRuntime()))
    g = G.build_llama_decode(exp, cfg)

    class _Capture:  # the planning runtime has no device: route the weight uploads into the exporter only
        pass
    for name, (t, shape, kind, shard) in g.weights.items():
        w = G.llama_weight_values(name, shape, kind)
        exp._weights[t.name] = G.to_storage(w, cfg.dtype)
    with open(path, "wb") as f:
        f.write(exp.save())
    return path


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("model", nargs="?")
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--export-llama", metavar="OUT")
    ap.add_argument("--llama", default="2,1024,8,2816,2048,128,16", help="layers,d_model,heads,ffn,vocab,s_max,batch for --export-llama")
    args = ap.parse_args()
    if args.export_llama:
        print(export_llama(args.export_llama, *[int(v) for v in args.llama.split(",")]))
        if not args.model:
            return
    import torch
    import torch.distributed as dist
    from infinitensor_b200 import backend as B, onnx_lite as X

    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    rt = B.CudaRuntime(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        box = [B.CudaRuntime.nccl_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(box, src=0)
        rt.init_comm_with_id(box[0], world, rank)
        if os.environ.get("ITB_NO_P2P", "0") != "1":
            handles = [None] * world
            dist.all_gather_object(handles, rt.p2p_export())
            rt.p2p_import(handles, world, rank)
    model = X.parallel_model(args.model, world, rank)
    stub = X.OnnxStub(model, rt)
    rng = np.random.default_rng(0)  # same seed on every rank: replicated inputs agree, sharded ones are synthetic anyway
    for name, t in stub.inputs.items():
        shp, dt = t.shape(), t.dtype()
        if dt in (X.I64, X.I32):
            v = rng.integers(0, 7, size=shp).astype(X._NP[dt])
        elif dt == X.BF16:
            v = (rng.standard_normal(shp).astype(np.float32) * 0.5).view(np.uint32) >> 16
            v = v.astype(np.uint16)
        else:
            v = (rng.standard_normal(shp) * 0.5).astype(X._NP[dt])
        t.copyin_numpy(v)
    h = stub.handler
    for _ in range(args.warmup):
        h.run_with_cudagraph()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    stream = torch.cuda.ExternalStream(rt.stream())
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(stream):
        e0.record()
        for _ in range(args.steps):
            h.launch_cudagraph_async()
        e1.record()
    e1.synchronize()
    ms = e0.elapsed_time(e1) / args.steps
    if world > 1:
        t = torch.tensor([ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    outs = {k: v.copyout_numpy() for k, v in stub.outputs.items()}
    if rank == 0:
        finite = all(np.isfinite(o.astype(np.float32) if o.dtype != np.uint16 else (o.astype(np.uint32) << 16).view(np.float32)).all()
                     for o in outs.values())
        print(json.dumps({"model": os.path.basename(args.model), "n_gpus": world, "ms_per_run": round(ms, 4), "steps": args.steps,
                          "scheduled_steps": len(h.schedule()), "ops": len(stub.lowered_nodes), "folded": len(stub.folded),
                          "outputs_finite": bool(finite)}))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
