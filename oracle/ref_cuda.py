"""ctypes loader for oracle/_ref/libit_ref_cuda.so -- the UNMODIFIED reference CUDA kernels (rms_norm.cu, rope.cu,
attention_kvcache.cu, softmax.cu, layer_norm.cu) compiled from /root/reference by `make -C oracle ref_cuda`.
TEST INFRASTRUCTURE: the `-m gpu` parity tests run them on the B200 beside this repo's kernels (same-box pin)."""
import ctypes
import os

_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref", "libit_ref_cuda.so")
_lib = None


def available() -> bool:
    return os.path.exists(_PATH)


def lib():
    global _lib
    if _lib is None:
        _lib = ctypes.CDLL(_PATH)  # RTLD_LOCAL: it carries its own copy of the reference's infini:: host symbols
        vp, ci, cf = ctypes.c_void_p, ctypes.c_int, ctypes.c_float
        _lib.ref_cuda_rmsnorm.argtypes = [ci, vp, vp, vp, ci, ci]
        _lib.ref_cuda_rope.argtypes = [ci, vp, vp, vp, ci, ci, ci, ci, ci]
        _lib.ref_cuda_attention_kvcache.argtypes = [vp] * 7 + [ci] * 4 + [vp, vp]
        _lib.ref_cuda_softmax_f32.argtypes = [vp, vp, ci, ci, ci]
        _lib.ref_cuda_softmax_f16.argtypes = [vp, vp, ci, ci, ci]
        _lib.ref_cuda_layernorm_f32.argtypes = [vp, vp, cf, ci, ci, ci, ci, vp, vp, ci]
    return _lib
