"""TEST INFRASTRUCTURE ONLY -- import the *unmodified* reference modules from
/root/reference on CPU (dev container only; the GPU box has no /root/reference).

The reference depends on five packages that are absent from this image
(torch_geometric, torch_cluster, torch_scatter, rotary_embedding_torch,
omegaconf).  We inject minimal ``sys.modules`` stand-ins so that
``src.model.layers.{integral_transform,geoembed,magno,attn,mlp}`` and
``src.model.gaot_3d`` import unchanged:

* torch_scatter is deliberately NOT provided -> the reference falls back to its
  own ``scatter_native`` (reference src/model/layers/integral_transform.py:20-24).
* torch_geometric.nn.{radius,knn} and torch_geometric.utils.{coalesce,
  dropout_edge} are bound to the restatements in ``oracle.graph`` (SURVEY.md
  Appendix A1-A4 semantics; the arithmetic lives in the un-vendored, unpinned
  torch-cluster / torch-geometric wheels).
* rotary_embedding_torch.RotaryEmbedding is bound to ``oracle.rope`` (App. A6).

Used (a) by tests/golden/make_golden.py to generate the committed fixtures and
(b) by ``-m "not gpu"`` tests to pin the travelling torch restatements in
``oracle.gno`` / ``oracle.attn`` against the reference's own code.
"""
from __future__ import annotations

import os
import sys
import types
import dataclasses

REFERENCE_ROOT = os.environ.get("GAOT_REFERENCE_ROOT", "/root/reference")


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "src", "model", "layers"))


def _install_stubs() -> None:
    import torch
    from . import graph as _g
    from . import rope as _r

    if "omegaconf" not in sys.modules:
        oc = types.ModuleType("omegaconf")

        class DictConfig(dict):
            pass

        class OmegaConf:  # only the names the reference touches at import time
            @staticmethod
            def create(x=None):
                return DictConfig(x or {})

            @staticmethod
            def merge(*a):
                out = DictConfig()
                for d in a:
                    out.update(d)
                return out

            @staticmethod
            def structured(x):
                return x

            @staticmethod
            def to_object(x):
                return x

            @staticmethod
            def load(path):
                raise RuntimeError("omegaconf stub: load() unavailable")

        oc.DictConfig = DictConfig
        oc.OmegaConf = OmegaConf
        sys.modules["omegaconf"] = oc

    if "rotary_embedding_torch" not in sys.modules:
        rt = types.ModuleType("rotary_embedding_torch")
        rt.RotaryEmbedding = _r.RotaryEmbedding
        rt.apply_rotary_emb = _r.apply_rotary_emb
        sys.modules["rotary_embedding_torch"] = rt

    if "torch_geometric" not in sys.modules:
        pyg = types.ModuleType("torch_geometric")
        nn_m = types.ModuleType("torch_geometric.nn")
        ut_m = types.ModuleType("torch_geometric.utils")
        da_m = types.ModuleType("torch_geometric.data")

        def radius(x, y, r, batch_x=None, batch_y=None, max_num_neighbors=32,
                   num_workers=1, batch_size=None):
            return _g.radius_torch(x, y, r, batch_x, batch_y, max_num_neighbors)

        def knn(x, y, k, batch_x=None, batch_y=None, cosine=False,
                num_workers=1, batch_size=None):
            assert not cosine
            return _g.knn_torch(x, y, k, batch_x, batch_y)

        nn_m.radius = radius
        nn_m.knn = knn
        ut_m.coalesce = _g.coalesce_torch
        ut_m.dropout_edge = _g.dropout_edge_torch

        class Data:  # duck-typed container
            def __init__(self, **kw):
                for k, v in kw.items():
                    setattr(self, k, v)

        class Batch(Data):
            pass

        da_m.Data = Data
        da_m.Batch = Batch
        pyg.nn = nn_m
        pyg.utils = ut_m
        pyg.data = da_m
        sys.modules["torch_geometric"] = pyg
        sys.modules["torch_geometric.nn"] = nn_m
        sys.modules["torch_geometric.utils"] = ut_m
        sys.modules["torch_geometric.data"] = da_m


_REF = None


def load_reference():
    """Return a namespace with the reference's hot-path modules (unchanged code)."""
    global _REF
    if _REF is not None:
        return _REF
    if not reference_available():
        raise RuntimeError(f"reference tree not found at {REFERENCE_ROOT}")
    _install_stubs()
    # the reference is imported as the top-level package ``src`` (its own layout)
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    import importlib
    import io
    import contextlib

    with contextlib.redirect_stdout(io.StringIO()):
        magno = importlib.import_module("src.model.layers.magno")
        it = importlib.import_module("src.model.layers.integral_transform")
        geo = importlib.import_module("src.model.layers.geoembed")
        attn = importlib.import_module("src.model.layers.attn")
        mlp = importlib.import_module("src.model.layers.mlp")
        gaot = importlib.import_module("src.model.gaot_3d")
        scat = importlib.import_module("src.model.layers.utils.scatter_native")
    _REF = types.SimpleNamespace(magno=magno, integral_transform=it, geoembed=geo,
                                 attn=attn, mlp=mlp, gaot_3d=gaot, scatter_native=scat)
    return _REF


class SimpleBatch:
    """Duck-typed stand-in for torch_geometric.data.Batch (reference magno.py:480-499)."""

    def __init__(self, pos, batch=None, num_graphs=1, **attrs):
        import torch
        self.pos = pos
        self.batch = batch if batch is not None else torch.zeros(pos.shape[0], dtype=torch.long, device=pos.device)
        self.num_graphs = num_graphs
        for k, v in attrs.items():
            setattr(self, k, v)
