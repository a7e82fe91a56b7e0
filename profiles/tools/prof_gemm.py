"""Dev helper (GPU box): time the bf16 dense kernels on the transformer-block shapes, against torch bf16 matmul (cuBLAS)."""
import sys
import torch
sys.path.insert(0, ".")
from gaot_3d_b200 import ops  # noqa: E402

dev = torch.device("cuda:0")
torch.manual_seed(0)
M = 16384


def timeit(fn, n=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3


for N, K in ((768, 256), (2048, 256), (256, 1024), (256, 256)):
    x = torch.randn(M, K, device=dev).bfloat16()
    w = (torch.randn(N, K, device=dev) / K ** 0.5).bfloat16()
    dy = torch.randn(M, N, device=dev).bfloat16()
    res = torch.randn(M, N, device=dev)
    t_f = timeit(lambda: ops._linear_fwd_raw(x, None, w, None, None, torch.bfloat16))
    t_fr = timeit(lambda: ops._linear_fwd_raw(x, None, w, None, res, torch.float32))
    t_x = timeit(lambda: ops._linear_bwd_x_raw(dy, w, torch.bfloat16))
    dw = torch.empty(N, K, device=dev)
    t_w = timeit(lambda: ops._linear_bwd_w_raw(dy, x, dw))
    t_t = timeit(lambda: x @ w.t())
    t_tx = timeit(lambda: dy @ w)
    t_tw = timeit(lambda: dy.t() @ x)
    fl = 2.0 * M * N * K
    y = ops._linear_fwd_raw(x, None, w, None, None, torch.float32)
    err = (y - (x.float() @ w.float().t())).abs().max().item()
    print(f"N={N} K={K}: fwd {t_f:.1f} us ({fl / t_f / 1e6:.0f} TF/s) fwd+res fp32 {t_fr:.1f} | dX {t_x:.1f} | dW {t_w:.1f} | "
          f"cuBLAS bf16 fwd {t_t:.1f} dX {t_tx:.1f} dW {t_tw:.1f} | max err {err:.2e}", flush=True)
