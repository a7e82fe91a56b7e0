"""Generates the committed golden fixtures by running the UNMODIFIED reference modules from
/root/reference on CPU (through oracle.ref_loader's stubs; the neighbour search / coalesce / RoPE
arithmetic that lives in un-vendored third-party wheels comes from the restatements in
oracle.graph / oracle.rope, checked against brute force).  Run in the dev container:

    python tests/golden/make_golden.py

Outputs (small, committed): graph_golden.npz, gno_golden.pt, geo_golden.pt, attn_golden.pt,
model_golden.pt.  The GPU box has no /root/reference: tests there read only these files.
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import graph as og, ref_loader  # noqa: E402
from tests import synth  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def main():
    ref = ref_loader.load_reference()
    torch.manual_seed(1234)
    # ------------------------------------------------------------------ graph (brute-force definition)
    N, G, r, k = 1500, (10, 8, 6), 0.22, 3
    phys, lat = synth.surface_cloud(N, seed=11), synth.latent_grid(G)
    g = {"phys": phys, "lat": lat, "r": np.float64(r), "k": np.int64(k)}
    g["enc_radius_raw"] = og.radius_bruteforce(phys, lat, r)       # [latent, phys]
    g["dec_radius_raw"] = og.radius_bruteforce(lat, phys, r)       # [phys, latent]
    g["knn_raw"] = og.knn_bruteforce(lat, phys, k)                 # [phys, latent]
    for name, dec in (("knn", False), ("radius", False), ("bidirectional", False), ("knn", True), ("radius", True),
                      ("bidirectional", True), ("reverse", True)):
        # composition through the reference's own get_neighbor_strategy (magno.py:116-295)
        e = ref.magno.get_neighbor_strategy(name, torch.from_numpy(phys), torch.zeros(N, dtype=torch.long),
                                            torch.from_numpy(lat), torch.zeros(len(lat), dtype=torch.long), r, k, dec)
        g[f"{'dec' if dec else 'enc'}_{name}"] = e.numpy()
    np.savez_compressed(os.path.join(OUT, "graph_golden.npz"), **g)

    # ------------------------------------------------------------------ IntegralTransform fwd + grads
    C = 32
    physt, latt = torch.from_numpy(phys), torch.from_numpy(lat)
    gno = {}
    for tag, layers, ypos, xpos, ei in (
            ("enc", [6, 64, 64, 64, C], physt, latt, torch.from_numpy(g["enc_bidirectional"])),
            ("dec", [6, 64, 64, C], latt, physt, torch.from_numpy(g["dec_radius"]))):
        it = ref.integral_transform.IntegralTransform(channel_mlp_layers=layers)
        f = torch.randn(ypos.shape[0], C, requires_grad=True)
        out = it(ypos, xpos, ei, f)
        gout = torch.randn_like(out)
        out.backward(gout)
        gno[tag] = dict(layers=layers, y_pos=ypos, x_pos=xpos, edge_index=ei, f_y=f.detach(), out=out.detach(), d_out=gout,
                        d_f=f.grad.clone(), weights=[l.weight.detach().clone() for l in it.channel_mlp.fcs],
                        biases=[l.bias.detach().clone() for l in it.channel_mlp.fcs],
                        d_weights=[l.weight.grad.clone() for l in it.channel_mlp.fcs],
                        d_biases=[l.bias.grad.clone() for l in it.channel_mlp.fcs])
    torch.save(gno, os.path.join(OUT, "gno_golden.pt"))

    # ------------------------------------------------------------------ GeometricEmbedding
    ge = ref.geoembed.GeometricEmbedding(3, C)
    ei = torch.from_numpy(g["enc_bidirectional"])
    feats = ge._compute_statistical_features_pyg(physt, latt, ei)
    emb = ge(physt, latt, ei)
    torch.save(dict(source_pos=physt, query_pos=latt, edge_index=ei, features=feats.detach(), embedding=emb.detach(),
                    state={k_: v.detach().clone() for k_, v in ge.state_dict().items()}), os.path.join(OUT, "geo_golden.pt"))

    # ------------------------------------------------------------------ attention module fwd + grads
    att = {}
    for tag, (S, hid, nh, nkv, rope) in {"rope_gqa": (320, 128, 4, 2, True), "abs_mha": (256, 64, 2, 2, False)}.items():
        a = ref.attn.GroupQueryFlashAttention(hid, hid, hidden_size=hid, num_heads=nh, num_kv_heads=nkv,
                                              positional_embedding="rope" if rope else "absolute").eval()
        x = torch.randn(1, S, hid, requires_grad=True)
        o = a(x, relative_positions=torch.zeros(1) if rope else None)
        go = torch.randn_like(o)
        o.backward(go)
        att[tag] = dict(S=S, hidden=hid, num_heads=nh, num_kv_heads=nkv, rope=rope, x=x.detach(), out=o.detach(), d_out=go,
                        d_x=x.grad.clone(), state={k_: v.detach().clone() for k_, v in a.state_dict().items()},
                        d_state={n: p.grad.clone() for n, p in a.named_parameters() if p.grad is not None})
    torch.save(att, os.path.join(OUT, "attn_golden.pt"))

    # ------------------------------------------------------------------ full model (config A family)
    models = {}
    for tag, strat, geo in (("radius_reverse", ["radius", "reverse"], [True, False]), ("knn", "knn", [False, False]),
                            ("bidirectional", "bidirectional", [True, False])):
        MC = ref.magno.MAGNOConfig(gno_coord_dim=3, lifting_channels=C, neighbor_strategy=strat, gno_radius=r, mlp_type="linear",
                                   precompute_edges=False, use_geoembed=geo, encoder_feature_attr=["pos", "c"], k_neighbors=k)
        TC = ref.attn.TransformerConfig(patch_size=2, hidden_size=128, num_layers=3, positional_embedding="rope")
        TC.attn_config.hidden_size = 128
        TC.attn_config.num_heads = 4
        TC.attn_config.num_kv_heads = 2 if tag == "knn" else 4
        TC.ffn_config.hidden_size = 128
        m = ref.gaot_3d.GAOT3D(6, 4, MC, TC, latent_tokens=G).eval()
        cn = torch.from_numpy(synth.unit_normals(N, seed=5))
        y = m(ref_loader.SimpleBatch(pos=physt, c=cn), tokens_pos=latt)
        models[tag] = dict(strategy=strat, use_geoembed=geo, latent_tokens=G, radius=r, k=k, pos=physt, c=cn, tokens_pos=latt,
                           out=y.detach(), state={k_: v.detach().clone().half() if v.dtype == torch.float32 and v.numel() > 4096
                                                  else v.detach().clone() for k_, v in m.state_dict().items()})
        # weights are stored in fp16 to keep the fixture small; recompute the output with the rounded weights
        sd = {k_: v.float() for k_, v in models[tag]["state"].items()}
        m.load_state_dict(sd)
        models[tag]["out"] = m(ref_loader.SimpleBatch(pos=physt, c=cn), tokens_pos=latt).detach()
    torch.save(models, os.path.join(OUT, "model_golden.pt"))
    for f in sorted(os.listdir(OUT)):
        print(f, os.path.getsize(os.path.join(OUT, f)))


if __name__ == "__main__":
    main()
