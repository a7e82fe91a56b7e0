"""GPU parity: the fused two-layer node MLP (decoder projection head, reference magno.py:640-644,:796-797) against the
reference's arithmetic -- Linear, exact-erf GELU, Linear (mlp.py:327-335) -- evaluated in fp64 on CPU.
Tolerance: rtol 2e-2 of the output / gradient scale (north star, f16/bf16 tensor-core operands)."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def rel(a, ref):
    a, ref = a.detach().double().cpu(), ref.detach().double().cpu()
    return ((a - ref).abs().max() / ref.abs().max().clamp(min=1e-30)).item()


@pytest.mark.parametrize("n,c_out", [(1, 4), (127, 4), (128, 4), (1000, 3), (40001, 4), (300000, 8)])
def test_node_mlp2_forward_backward(n, c_out):
    from gaot_3d_b200 import ops
    torch.manual_seed(n + c_out)
    x = torch.randn(n, 32)
    w1, b1 = torch.randn(256, 32) / 32 ** 0.5, torch.randn(256) * 0.2
    w2, b2 = torch.randn(c_out, 256) / 16, torch.randn(c_out) * 0.3
    go = torch.randn(n, c_out)
    ref_in = [t.double().requires_grad_(True) for t in (x, w1, b1, w2, b2)]
    yr = F.linear(F.gelu(F.linear(ref_in[0], ref_in[1], ref_in[2])), ref_in[3], ref_in[4])
    yr.backward(go.double())
    dev_in = [t.to(DEV).requires_grad_(True) for t in (x, w1, b1, w2, b2)]
    y = ops.node_mlp2(*dev_in)
    assert y.shape == (n, c_out)
    assert rel(y, yr) < 2e-2, ("y", rel(y, yr))
    y.backward(go.to(DEV))
    for name, a, r in zip(("dx", "dW1", "db1", "dW2", "db2"), dev_in, ref_in):
        assert rel(a.grad, r.grad) < 2e-2, (name, rel(a.grad, r.grad))
    # deterministic: fixed-order reduction of the per-CTA partials
    dev2 = [t.to(DEV).requires_grad_(True) for t in (x, w1, b1, w2, b2)]
    y2 = ops.node_mlp2(*dev2)
    y2.backward(go.to(DEV))
    assert torch.equal(y, y2) and all(torch.equal(a.grad, b.grad) for a, b in zip(dev_in, dev2))


def test_projection_module_uses_the_fused_kernel_when_asked():
    """MAGNODecoder.projection through _apply_node_mlp: 'fused' mode == torch mode within the mixed-precision tolerance,
    for both node-MLP flavours (Linear and kernel-size-1 Conv1d share the [out, in] weight layout)."""
    from gaot_3d_b200 import ops
    from gaot_3d_b200.layers.magno import _apply_node_mlp, _node_mlp
    torch.manual_seed(0)
    x = torch.randn(5000, 32, device=DEV)
    for mlp_type in ("linear", "channel"):
        mlp = _node_mlp(mlp_type, 32, 4, hidden=256).to(DEV)
        ops.set_node_mlp_mode("torch")
        ref = _apply_node_mlp(mlp, mlp_type, x)
        ops.reset_launch_count()
        ops.set_node_mlp_mode("fused")
        try:
            out = _apply_node_mlp(mlp, mlp_type, x)
            assert ops.launch_count() >= 1, "fused mode did not launch the library kernel"
        finally:
            ops.set_node_mlp_mode("torch")
            ops.set_node_mlp_tf32(False)
        assert rel(out, ref) < 2e-2, (mlp_type, rel(out, ref))
        # shapes outside the envelope keep the torch path
        other = _node_mlp(mlp_type, 64, 32).to(DEV)
        assert _apply_node_mlp(other, mlp_type, torch.randn(10, 64, device=DEV)).shape == (10, 32)


@pytest.mark.parametrize("n,k_in,c_out,bias", [(0, 6, 32, True), (1, 6, 32, True), (1000, 5, 16, True), (40001, 6, 32, False),
                                               (300000, 6, 32, True), (5000, 16, 64, True), (777, 3, 48, True), (2049, 1, 1, True)])
def test_node_linear_forward_backward(n, k_in, c_out, bias):
    """The lifting layer's streaming kernel (reference magno.py:421-424, :540-545; Linear forward mlp.py:327-335) against the
    same Linear in fp64: fp32 FMA in both directions -> the FP32 tier's rtol 1e-5 (of the output / gradient scale)."""
    from gaot_3d_b200 import ops
    torch.manual_seed(n + k_in)
    x = torch.randn(n, k_in)
    w, b = torch.randn(c_out, k_in) / k_in ** 0.5, (torch.randn(c_out) * 0.3 if bias else None)
    go = torch.randn(n, c_out)
    ref_in = [t.double().requires_grad_(True) for t in (x, w)] + ([b.double().requires_grad_(True)] if bias else [])
    yr = F.linear(ref_in[0], ref_in[1], ref_in[2] if bias else None)
    yr.backward(go.double())
    dev_in = [t.to(DEV).requires_grad_(True) for t in (x, w)] + ([b.to(DEV).requires_grad_(True)] if bias else [])
    y = ops.node_linear(dev_in[0], dev_in[1], dev_in[2] if bias else None)
    assert y.shape == (n, c_out)
    y.backward(go.to(DEV))
    if n == 0:
        assert all(float(a.grad.abs().sum()) == 0.0 for a in dev_in[1:])
        return
    assert rel(y, yr) < 1e-5, ("y", rel(y, yr))
    for name, a, r in zip(("dx", "dW", "db"), dev_in, ref_in):
        assert rel(a.grad, r.grad) < 2e-5, (name, rel(a.grad, r.grad))
    dev2 = [t.to(DEV).requires_grad_(True) for t in (x, w)] + ([b.to(DEV).requires_grad_(True)] if bias else [])
    y2 = ops.node_linear(dev2[0], dev2[1], dev2[2] if bias else None)
    y2.backward(go.to(DEV))
    assert torch.equal(y, y2) and all(torch.equal(a.grad, c.grad) for a, c in zip(dev_in, dev2)), "not deterministic"


def test_lifting_module_uses_the_streaming_kernel():
    """MAGNOEncoder.lifting through _apply_node_mlp: both node-MLP flavours, default (strict fp32) mode."""
    from gaot_3d_b200 import ops
    from gaot_3d_b200.layers.magno import _apply_node_mlp, _node_mlp
    torch.manual_seed(0)
    x = torch.randn(5000, 6, device=DEV)
    for mlp_type in ("linear", "channel"):
        mlp = _node_mlp(mlp_type, 6, 32).to(DEV)
        ref = mlp(x) if mlp_type == "linear" else mlp(x.transpose(0, 1)).transpose(0, 1)
        ops.reset_launch_count()
        out = _apply_node_mlp(mlp, mlp_type, x)
        assert ops.launch_count() >= 1, "the lifting layer did not launch the library kernel"
        assert rel(out, ref) < 1e-5, (mlp_type, rel(out, ref))
