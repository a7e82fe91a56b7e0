"""GPU parity: tcgen05 flash attention forward/backward vs the CPU oracle (fp64) and the golden
vectors of the reference's GroupQueryFlashAttention.  BF16 operands -> tolerance rtol 2e-2 (north star)."""
import os

import pytest
import torch

from oracle import attn as oattn

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")
DEV = "cuda:0"


def relerr(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return ((a - b).norm() / b.norm().clamp(min=1e-30)).item(), ((a - b).abs().max() / b.abs().max().clamp(min=1e-30)).item()


@pytest.mark.parametrize("B,S,H,Hkv,d,rope", [(1, 512, 8, 8, 32, True), (2, 200, 4, 2, 32, True), (1, 384, 2, 2, 64, False),
                                                (1, 1024, 8, 8, 32, False), (1, 129, 2, 1, 32, True), (1, 2300, 2, 1, 32, True)])
def test_attention_vs_oracle(B, S, H, Hkv, d, rope):
    from gaot_3d_b200 import ops
    from gaot_3d_b200.layers.attn import RotaryEmbedding
    torch.manual_seed(S + H)
    q = torch.randn(B, S, H * d)
    k = torch.randn(B, S, Hkv * d)
    v = torch.randn(B, S, Hkv * d)
    qr, kr, vr = (t.double().requires_grad_(True) for t in (q, k, v))
    ref = oattn.attention_core(qr, kr, vr, H, Hkv, rope, dtype=torch.float64)
    go = torch.randn_like(ref)
    ref.backward(go)
    qd, kd, vd = (t.to(DEV).requires_grad_(True) for t in (q, k, v))
    freqs = RotaryEmbedding(d).freqs.to(DEV) if rope else None
    out = ops.attention(qd, kd, vd, H, Hkv, rope_freqs=freqs)
    out.backward(go.float().to(DEV))
    for name, a, b in (("out", out, ref), ("dq", qd.grad, qr.grad), ("dk", kd.grad, kr.grad), ("dv", vd.grad, vr.grad)):
        l2, mx = relerr(a, b)
        assert l2 < 1e-2 and mx < 2e-2, f"{name}: rel l2 {l2:.3e}, rel max {mx:.3e}"


@pytest.mark.parametrize("tag", ["rope_gqa", "abs_mha"])
def test_attention_module_golden(tag):
    from gaot_3d_b200.layers.attn import GroupQueryFlashAttention
    g = torch.load(os.path.join(GOLD, "attn_golden.pt"))[tag]
    m = GroupQueryFlashAttention(g["hidden"], g["hidden"], hidden_size=g["hidden"], num_heads=g["num_heads"],
                                 num_kv_heads=g["num_kv_heads"], positional_embedding="rope" if g["rope"] else "absolute").to(DEV).eval()
    m.load_state_dict(g["state"])            # strict: same keys as the reference module (incl. rotary_emb.freqs)
    x = g["x"].to(DEV).requires_grad_(True)
    o = m(x, relative_positions=torch.zeros(1) if g["rope"] else None)
    o.backward(g["d_out"].to(DEV))
    l2, mx = relerr(o, g["out"])
    assert l2 < 1e-2 and mx < 2e-2, (l2, mx)
    l2, mx = relerr(x.grad, g["d_x"])
    assert l2 < 1.5e-2 and mx < 3e-2, (l2, mx)
    for n, p in m.named_parameters():
        if n in g["d_state"]:
            l2, mx = relerr(p.grad, g["d_state"][n])
            assert l2 < 1.5e-2, (n, l2, mx)


@pytest.mark.parametrize("B,S,H,Hkv,d,p", [(1, 384, 4, 4, 32, 0.1), (2, 200, 4, 2, 32, 0.3), (1, 1100, 2, 2, 32, 0.2)])
def test_attention_dropout_matches_oracle_mask(B, S, H, Hkv, d, p):
    """Training-mode dropout (reference attn.py:122-126): the kernels' counter-based mask is restated in
    oracle.attn.dropout_keep, so forward and backward are checked exactly like the p=0 case."""
    from gaot_3d_b200 import ops
    torch.manual_seed(3)
    q, k, v = torch.randn(B, S, H * d), torch.randn(B, S, Hkv * d), torch.randn(B, S, Hkv * d)
    seed = 0x1234_5678_9ABC_DEF1
    qr, kr, vr = (t.double().requires_grad_(True) for t in (q, k, v))
    ref = oattn.attention_core(qr, kr, vr, H, Hkv, False, dtype=torch.float64, dropout_p=p, seed=seed)
    go = torch.randn_like(ref)
    ref.backward(go)
    keep = oattn.dropout_keep(B, H, S, p, seed)
    assert abs(keep.float().mean().item() - (1 - p)) < 0.01
    qd, kd, vd = (t.to(DEV).requires_grad_(True) for t in (q, k, v))
    out = ops.attention(qd, kd, vd, H, Hkv, dropout_p=p, seed=seed)
    out.backward(go.float().to(DEV))
    for name, a, b in (("out", out, ref), ("dq", qd.grad, qr.grad), ("dk", kd.grad, kr.grad), ("dv", vd.grad, vr.grad)):
        l2, mx = relerr(a, b)
        assert l2 < 1e-2 and mx < 2.5e-2, f"{name}: rel l2 {l2:.3e}, rel max {mx:.3e}"
    # different seeds -> different masks; module in train() mode runs and differs from eval()
    out2 = ops.attention(qd, kd, vd, H, Hkv, dropout_p=p, seed=seed + 1)
    assert (out2 - out).abs().max().item() > 1e-3
