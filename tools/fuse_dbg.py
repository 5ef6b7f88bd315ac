import os, sys, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from infinitensor_b200 import backend as B, graphs as G
MODE = sys.argv[1] if len(sys.argv) > 1 else 'eager'
def run(dtype, mask, nofuse="0"):
    os.environ["ITB_NO_FUSION"]=nofuse; os.environ["ITB_FUSION_MASK"]=str(mask)
    cfg = G.LlamaConfig(layers=2, d_model=1024, heads=8, head_dim=128, ffn=2816, vocab=2048, s_max=64, batch=16, dtype=dtype)
    rt = B.CudaRuntime(0); h = B.GraphHandler(rt); g = G.build_llama_decode(h, cfg); h.data_malloc(); G.fill_llama_weights_host(g)
    for li in range(cfg.layers):
        g.k_caches[li].copyin_numpy(G.to_storage(G.llama_cache_values(cfg, li, "k"), dtype))
        g.v_caches[li].copyin_numpy(G.to_storage(G.llama_cache_values(cfg, li, "v"), dtype))
    g.input_ids.copyin_numpy(np.arange(16, dtype=np.int64).reshape(16, 1) * 7 % cfg.vocab)
    g.position_ids.copyin_numpy(np.full((16, 1), 33, np.int64))
    if MODE == 'graph':
        h.run_with_cudagraph(); h.run_with_cudagraph()
    else:
        h.run()
    return G.from_storage(g.logits.copyout_numpy(), dtype).astype(np.float64)
for dtype in (1, 16):
    base = run(dtype, 0, "1")
    base2 = run(dtype, 0, "1")
    print("dtype", dtype, "unfused repeat diff", np.abs(base-base2).max())
    for mask in (1,2,4,8,15):
        got = run(dtype, mask)
        print("  mask", mask, "maxdiff", np.abs(got-base).max(), "n diff", int((got!=base).sum()))
