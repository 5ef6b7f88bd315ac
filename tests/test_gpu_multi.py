"""Multi-GPU (tensor-parallel) parity on real GPUs: needs >= 2 devices (gpurun --gpus 2); skipped otherwise.
Both collective paths -- the fused NVLink one-shot all-reduce (+ residual + RMSNorm) and in-graph NCCL -- must
reproduce the unsharded graph executed by the CPU oracle."""
import os
import subprocess
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _ngpu():
    import torch
    return torch.cuda.device_count() if torch.cuda.is_available() else 0


@pytest.mark.parametrize("dtype", [16, 1])
def test_tensor_parallel_two_gpus(tmp_path, dtype):
    if _ngpu() < 2:
        pytest.skip("needs 2 GPUs")
    from infinitensor_b200 import graphs as G
    from oracle.graph_oracle import OracleHandler
    out = str(tmp_path / "tp")
    env = dict(os.environ, TP_OUT=out, TP_DTYPE=str(dtype), MASTER_ADDR="127.0.0.1")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29617", os.path.join(ROOT, "tools", "tp_worker.py")]
    subprocess.run(cmd, check=True, env=env, cwd=ROOT, timeout=600)
    cfg = G.LlamaConfig(layers=2, d_model=1024, heads=8, head_dim=128, ffn=2816, vocab=2048, s_max=64, batch=16, dtype=dtype)
    oh = OracleHandler()
    g = G.build_llama_decode(oh, cfg)
    G.fill_llama_weights_host(g)
    for li in range(cfg.layers):
        g.k_caches[li].copyin_numpy(G.to_storage(G.llama_cache_values(cfg, li, "k"), dtype))
        g.v_caches[li].copyin_numpy(G.to_storage(G.llama_cache_values(cfg, li, "v"), dtype))
    refs = []
    for step in range(3):
        g.input_ids.copyin_numpy((np.arange(cfg.batch, dtype=np.int64).reshape(-1, 1) * 7 + step) % cfg.vocab)
        g.position_ids.copyin_numpy(np.full((cfg.batch, 1), 9 + step, np.int64))
        oh.run()
        refs.append(g.logits.f32().copy())
    ref = np.stack(refs).astype(np.float64)
    tol = 1e-3 if dtype == 1 else 3e-2
    sched = open(out + ".p2p.sched").read().split("\n")
    assert sum(s.startswith("AllReduceAddNorm:AllReduceSum+Add+RMSNorm") for s in sched) == 4  # 2 per layer
    assert not any(s.startswith("AllReduceAddNorm") for s in open(out + ".nccl.sched").read().split("\n"))
    for mode in ("p2p", "nccl"):
        got = [np.load(f"{out}.{mode}.{r}.npy").astype(np.float64) for r in range(2)]
        assert np.array_equal(got[0], got[1]), f"{mode}: ranks disagree"
        err = np.abs(got[0] - ref).max() / np.abs(ref).max()
        assert err < tol, f"{mode}: rel-to-max error {err:.3e}"
