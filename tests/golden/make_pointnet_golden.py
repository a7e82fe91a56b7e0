"""Generates tests/golden/pointnet_golden.pt by running the UNMODIFIED reference GeometricEmbedding(method='pointnet')
(src/model/layers/geoembed.py:184-222) on CPU through oracle/ref_loader.py.  Dev-container only (needs /root/reference):
    python tests/golden/make_pointnet_golden.py
Stored per pooling: inputs, state_dict, forward output, and the parameter gradients for a fixed upstream gradient."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import graph as og, ref_loader  # noqa: E402
from tests import synth  # noqa: E402

if __name__ == "__main__":
    ref = ref_loader.load_reference()
    torch.manual_seed(11)
    phys, lat = synth.surface_cloud(3000, seed=9), synth.latent_grid((8, 8, 8))
    ei = torch.from_numpy(og.radius_np(phys, lat, 0.15)[::-1].copy())          # [phys, latent], some tokens stay empty
    P, L = torch.from_numpy(phys), torch.from_numpy(lat)
    out = {}
    for pooling in ("max", "mean"):
        ge = ref.geoembed.GeometricEmbedding(3, 16, method="pointnet", pooling=pooling)
        y = ge(P, L, ei)
        g = torch.randn_like(y)
        y.backward(g)
        out[pooling] = dict(source_pos=P, query_pos=L, edge_index=ei, state={k: v.detach().clone() for k, v in ge.state_dict().items()},
                            out=y.detach(), d_out=g, grads={k: v.grad.clone() for k, v in ge.named_parameters()})
    assert (torch.bincount(ei[1], minlength=L.shape[0]) == 0).any(), "fixture must contain empty queries"
    torch.save(out, os.path.join(ROOT, "tests", "golden", "pointnet_golden.pt"))
    print("wrote pointnet_golden.pt", ei.shape, os.path.getsize(os.path.join(ROOT, "tests", "golden", "pointnet_golden.pt")))
