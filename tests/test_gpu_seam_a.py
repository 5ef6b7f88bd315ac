"""SEAM A on the GPU: the reference's OWN gtests (test/kernels/cuda/test_cuda_*.cc, test/cuda/test_cudagraph.cc) run against
this repo's kernels.  oracle/Makefile `seam_a` builds the reference's host tree + its CudaRuntimeObj from /root/reference
with src/kernels/cuda REPLACED by infinitensor_b200/csrc/host/b200_kernels.cc (compiled against the reference's real headers)
and the kernel objects of infinitensor_b200/csrc/kernels; the binary travels to the GPU box prebuilt (oracle/_ref/)."""
import os
import re
import subprocess

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "oracle", "_ref", "seam_a", "seam_a_tests")

# every suite the binary links: the golden-vector / CPU-comparison tests of SURVEY 8(c) for rows a1-a15 and the CUDA-graph
# semantics of row a17
SUITES = ["cuBLAS_Matmul", "CUDA_SoftmaxFP32", "CUDA_SoftmaxFP16", "CUDA_LayernormFp32", "CUDA_LayernormFp16", "AttentionKVCache",
          "RoPE", "cuda_Transpose", "Concat", "ConcatToIdentity", "ConcatFp16", "Split", "SplitFp16", "Gather", "CUDA_WhereFp32",
          "CUDA_WhereFp16", "Expand", "CUDA_ReduceMean", "CUDA_ReduceSum", "cuDNN_MaxPool", "cuDNN_AvgPool", "CUDA_BatchNorm",
          "CUDA_Reshape", "CUDA_Flatten", "CUDA_Identity", "cuDNN_ElementWise", "LeakyRelu", "Elu", "cuDNN_Unary", "CUDA_Slice", "Pad",
          "cuDNN_Conv", "cuDNN_Conv_FP16", "TestCudaRuntime"]


def test_reference_gtests_pass_on_our_kernels():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no GPU")
    if not os.path.exists(BIN):
        pytest.skip("oracle/_ref/seam_a/seam_a_tests not built (needs /root/reference: `make -C oracle seam_a`)")
    r = subprocess.run([BIN, "--gtest_brief=1"], capture_output=True, text=True, timeout=600, cwd=ROOT)
    out = r.stdout + r.stderr
    m = re.search(r"\[==========\] (\d+) tests from (\d+) test suites ran", out)
    assert m, out[-2000:]
    failed = re.findall(r"\[  FAILED  \] (\S+)", out)
    passed = re.search(r"\[  PASSED  \] (\d+) tests", out)
    assert not failed, f"reference gtests failing on this repo's kernels: {sorted(set(failed))}"
    assert r.returncode == 0 and passed and int(passed.group(1)) == int(m.group(1)) >= 50
    listed = subprocess.run([BIN, "--gtest_list_tests"], capture_output=True, text=True, timeout=60).stdout
    for s in SUITES:
        assert re.search(rf"^{re.escape(s)}\.$", listed, flags=re.M), f"suite {s} missing from the binary"
