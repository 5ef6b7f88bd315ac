"""ctypes loader for oracle/_ref/libit_ref.so -- the UNMODIFIED reference (host core + native-CPU kernels)
compiled from /root/reference by `make -C oracle ref`.  TEST INFRASTRUCTURE: validates the restatement in
it_oracle.c against the reference's own kernels; may serve as the "reference" CPU baseline."""
import ctypes
import os

import numpy as np

_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref", "libit_ref.so")
_lib = None


def available() -> bool:
    return os.path.exists(_PATH)


def lib():
    global _lib
    if _lib is None:
        _lib = ctypes.CDLL(_PATH)
        _lib.ref_last_error.restype = ctypes.c_char_p
    return _lib


def run_op(name, inputs, iattrs=(), which_out=0, out_cap=None):
    """Run one operator of the reference on its NativeCpuRuntimeObj.  inputs: float32 arrays."""
    ins = [np.ascontiguousarray(a, dtype=np.float32) for a in inputs]
    ptrs = (ctypes.POINTER(ctypes.c_float) * len(ins))(*[a.ctypes.data_as(ctypes.POINTER(ctypes.c_float)) for a in ins])
    dims = [d for a in ins for d in a.shape]
    cd = (ctypes.c_int * max(len(dims), 1))(*dims)
    cr = (ctypes.c_int * len(ins))(*[a.ndim for a in ins])
    ia = (ctypes.c_int * max(len(iattrs), 1))(*[int(x) for x in iattrs])
    cap = out_cap or max(int(sum(a.size for a in ins)) * 64, 1 << 16)
    out = np.empty(cap, dtype=np.float32)
    od = (ctypes.c_int * 8)()
    orank = ctypes.c_int(0)
    r = lib().ref_run_op(name.encode(), len(ins), ptrs, cd, cr, ia, len(iattrs), which_out,
                         out.ctypes.data_as(ctypes.POINTER(ctypes.c_float)), ctypes.c_longlong(cap), od, ctypes.byref(orank))
    if r != 0:
        raise RuntimeError(lib().ref_last_error().decode())
    shape = [od[i] for i in range(orank.value)]
    return out[:int(np.prod(shape)) if shape else 1].reshape(shape).copy()
