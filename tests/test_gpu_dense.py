"""GPU parity: the tcgen05 nn.Linear kernels (forward, d input, d weight with split-K) vs torch.
Oracle = the reference's nn.Linear arithmetic (F.linear, reference attn.py:104-106,:163,:223) evaluated in
fp64 on CPU.  Two bars: (i) against the fp64 product of the bf16-ROUNDED operands the kernel must be exact up
to fp32 accumulation order (1e-5 of the output scale) -- this pins tile indexing, majors and split-K;
(ii) against the un-rounded fp64 product, rtol 2e-2 of the output scale (north star, BF16 operands)."""
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def rb(t):  # bf16 rounding of an fp32 tensor, back in fp64
    return t.to(torch.bfloat16).double()


def close(a, ref, tol, what):
    a, ref = a.detach().double().cpu(), ref.detach().double().cpu()
    err = (a - ref).abs().max().item()
    scale = ref.abs().max().item()
    assert err <= tol * scale + 1e-30, f"{what}: max abs err {err:.3e} vs scale {scale:.3e} (tol {tol})"


@pytest.mark.parametrize("M,N,K", [(128, 128, 64), (300, 256, 256), (1000, 768, 256), (257, 136, 328), (4096, 256, 1024),
                                   (64, 1024, 512), (16384, 256, 256)])
@pytest.mark.parametrize("bias,res", [(False, False), (True, True)])
def test_linear_forward_backward(M, N, K, bias, res):
    from gaot_3d_b200 import ops
    torch.manual_seed(M + N + K)
    x = torch.randn(M, K)
    w = torch.randn(N, K) / K ** 0.5
    b = torch.randn(N) if bias else None
    r = torch.randn(M, N) if res else None
    go = torch.randn(M, N)
    xd, wd = x.to(DEV).requires_grad_(True), w.to(DEV).requires_grad_(True)
    bd = b.to(DEV).requires_grad_(True) if bias else None
    rd = r.to(DEV).requires_grad_(True) if res else None
    y = ops.linear(xd, wd, bd, rd)
    y.backward(go.to(DEV))
    extra = (b.double() if bias else 0) + (r.double() if res else 0)
    # (i) exact arithmetic on the rounded operands
    close(y, rb(x) @ rb(w).T + extra, 2e-5, "y (rounded operands)")
    close(xd.grad, rb(go) @ rb(w), 2e-5, "dx (rounded operands)")
    close(wd.grad, rb(go).T @ rb(x), 5e-5, "dw (rounded operands)")
    # (ii) the reference's fp32 nn.Linear within the BF16 tolerance
    close(y, x.double() @ w.double().T + extra, 2e-2, "y")
    close(xd.grad, go.double() @ w.double(), 2e-2, "dx")
    close(wd.grad, go.double().T @ x.double(), 2e-2, "dw")
    if bias:
        close(bd.grad, go.double().sum(0), 1e-5, "db")
    if res:
        close(rd.grad, go, 0.0, "dresidual")


def test_linear_concat_input():
    """skip_proj over cat[x, skip] (reference attn.py:222-224) without materialising the concatenation."""
    from gaot_3d_b200 import ops
    torch.manual_seed(3)
    M, K1, K2, N = 777, 256, 256, 256
    x, s = torch.randn(M, K1), torch.randn(M, K2)
    w, b = torch.randn(N, K1 + K2) / 16, torch.randn(N)
    go = torch.randn(M, N)
    xd, sd, wd, bd = (t.to(DEV).requires_grad_(True) for t in (x, s, w, b))
    y = ops.linear(xd, wd, bd, x2=sd)
    y.backward(go.to(DEV))
    cat = torch.cat([x, s], 1)
    close(y, rb(cat) @ rb(w).T + b.double(), 2e-5, "y")
    dcat = rb(go) @ rb(w)
    close(xd.grad, dcat[:, :K1], 2e-5, "dx")
    close(sd.grad, dcat[:, K1:], 2e-5, "dskip")
    close(wd.grad, rb(go).T @ rb(cat), 5e-5, "dw")


def test_linear_3d_input_and_determinism():
    from gaot_3d_b200 import ops
    torch.manual_seed(5)
    x = torch.randn(2, 500, 256, device=DEV, requires_grad=True)
    w = torch.randn(1024, 256, device=DEV, requires_grad=True)
    outs = []
    for _ in range(2):
        x.grad = w.grad = None
        y = ops.linear(x, w)
        assert y.shape == (2, 500, 1024)
        y.square().sum().backward()
        outs.append((y.detach().clone(), x.grad.clone(), w.grad.clone()))
    for a, b in zip(*outs):
        assert torch.equal(a, b), "dense layer kernels must be bit-reproducible (fixed-order split-K reduction)"


def test_linear_rejects_cpu_and_odd_shapes():
    from gaot_3d_b200 import ops
    with pytest.raises(RuntimeError):
        ops.linear(torch.randn(4, 64), torch.randn(64, 64))
    with pytest.raises(NotImplementedError):
        ops.linear(torch.randn(4, 6, device=DEV), torch.randn(32, 6, device=DEV))


@pytest.mark.parametrize("M,N,K", [(300, 256, 256), (1000, 768, 320), (4096, 2048, 256), (16384, 256, 1024)])
def test_gemm_bf16_operand_paths(M, N, K):
    """The cp.async (bf16 source) and mixed fp32/bf16 operand paths used inside the fused TransformerBlock."""
    from gaot_3d_b200 import ops
    torch.manual_seed(M + N)
    bf = torch.bfloat16
    x32 = torch.randn(M, K, device=DEV)
    xb = x32.to(bf)
    wb = (torch.randn(N, K, device=DEV) / K ** 0.5).to(bf)
    dy32 = torch.randn(M, N, device=DEV)
    dyb = dy32.to(bf)
    res = torch.randn(M, N, device=DEV)
    X, W, DY = xb.double().cpu(), wb.double().cpu(), dyb.double().cpu()
    # forward: bf16 x bf16 -> bf16 / fp32 (+ residual); fp32 x bf16
    y = ops._linear_fwd_raw(xb, None, wb, None, res)
    close(y, X @ W.T + res.double().cpu(), 2e-5, "y bf16/bf16 + residual")
    yb = ops._linear_fwd_raw(xb, None, wb, None, None, bf)
    close(yb, X @ W.T, 1e-2, "y bf16 output")
    y2 = ops._linear_fwd_raw(x32, None, wb, None, None)
    close(y2, rb(x32.cpu()) @ W.T, 2e-5, "y fp32 x / bf16 w")
    # d input: bf16 dy (cp.async) and fp32 dy (register path), bf16 MN-major weight
    close(ops._linear_bwd_x_raw(dyb, wb), DY @ W, 2e-5, "dx bf16 dy")
    close(ops._linear_bwd_x_raw(dy32, wb, bf), rb(dy32.cpu()) @ W, 1e-2, "dx fp32 dy -> bf16")
    # d weight: bf16/bf16, fp32/bf16
    dw = torch.empty(N, K, device=DEV)
    close(ops._linear_bwd_w_raw(dyb, xb, dw), DY.T @ X, 5e-5, "dw bf16/bf16")
    close(ops._linear_bwd_w_raw(dy32, xb, dw), rb(dy32.cpu()).T @ X, 5e-5, "dw fp32 dy / bf16 x")
