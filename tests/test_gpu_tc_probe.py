"""Pins the tcgen05 shared-memory descriptor conventions of csrc/tc05.cuh on hardware:
single-CTA bf16 GEMM D[128,N] = A*B for all K-major / MN-major operand combinations."""
import ctypes

import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("N,K", [(128, 32), (32, 128), (64, 64), (128, 128), (16, 16)])
@pytest.mark.parametrize("variant", [0, 1, 2, 3])
def test_probe(N, K, variant):
    import os
    lib = ctypes.CDLL(os.path.join(os.path.dirname(os.path.abspath(__file__)), "probe", "libgaot_tcprobe.so"))   # built by __graft_entry__.build()
    lib.gaot_tc_probe.restype = ctypes.c_int
    lib.gaot_tc_probe.argtypes = [ctypes.c_void_p] * 3 + [ctypes.c_int] * 3 + [ctypes.c_void_p]
    dev = torch.device("cuda:0")
    torch.manual_seed(variant * 100 + N + K)
    a_mn, b_mn = bool(variant & 2), bool(variant & 1)
    A = torch.randn(128, K, device=dev).bfloat16().float()       # logical A [M,K]
    B = torch.randn(N, K, device=dev).bfloat16().float()         # logical B [N,K]
    A_in = A.t().contiguous() if a_mn else A.contiguous()        # MN-major operand is handed over as [K, M]
    B_in = B.t().contiguous() if b_mn else B.contiguous()
    D = torch.full((128, N), float("nan"), device=dev)
    rc = lib.gaot_tc_probe(ctypes.c_void_p(A_in.data_ptr()), ctypes.c_void_p(B_in.data_ptr()), ctypes.c_void_p(D.data_ptr()),
                           N, K, variant, ctypes.c_void_p(torch.cuda.current_stream().cuda_stream))
    assert rc == 0
    torch.cuda.synchronize()
    ref = A.double() @ B.double().t()
    err = (D.double() - ref).abs().max().item()
    assert err < 1e-3 * max(1.0, ref.abs().max().item()), f"variant {variant} N={N} K={K}: max err {err}"
