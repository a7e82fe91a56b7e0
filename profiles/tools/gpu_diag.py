"""Diagnostic sweep run on the GPU box (not a pytest): prints per-component max errors and keeps
going after failures so that one gpurun call yields the whole picture."""
import ctypes
import sys
import time
import traceback

import torch

sys.path.insert(0, ".")


def probe_table():
    from gaot_3d_b200 import _lib
    lib = _lib.load()
    dev = torch.device("cuda:0")
    print("tcgen05 probe: rel err per variant (bit0 B mn-major, bit1 A mn-major, bit2 swap LBO/SBO)")
    for N, K in ((128, 32), (32, 128), (64, 64)):
        row = []
        for variant in range(4):
            a_mn, b_mn = bool(variant & 2), bool(variant & 1)
            torch.manual_seed(1)
            A = torch.randn(128, K, device=dev).bfloat16().float()
            B = torch.randn(N, K, device=dev).bfloat16().float()
            A_in = A.t().contiguous() if a_mn else A
            B_in = B.t().contiguous() if b_mn else B
            D = torch.zeros(128, N, device=dev)
            rc = lib.gaot_tc_probe(ctypes.c_void_p(A_in.data_ptr()), ctypes.c_void_p(B_in.data_ptr()),
                                   ctypes.c_void_p(D.data_ptr()), N, K, variant, None)
            try:
                torch.cuda.synchronize()
                ref = A @ B.t()
                row.append(f"{((D - ref).abs().max() / ref.abs().max()).item():.2e}")
            except Exception as e:
                row.append("ERR")
                print("   cuda error:", str(e)[:100])
                return
        print(f"  N={N:3d} K={K:3d}: " + "  ".join(row))


def section(name, fn):
    print(f"=== {name}", flush=True)
    t = time.time()
    try:
        fn()
        print(f"--- {name} ok ({time.time() - t:.1f}s)", flush=True)
    except Exception:
        traceback.print_exc()
        print(f"--- {name} FAILED", flush=True)


if __name__ == "__main__":
    print(torch.cuda.get_device_name(0), torch.version.cuda)
    section("probe", probe_table)
