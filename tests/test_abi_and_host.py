"""CPU-side checks of the drop-in boundary: the C-ABI library loads and exports every symbol that
include/gaot_b200.h declares (no compute without a GPU), host-side logic and error behaviour."""
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    from gaot_3d_b200 import build, _lib
    build.build()
    return _lib.load()


def test_exports_match_header(lib):
    from gaot_3d_b200 import _lib
    hdr = open(os.path.join(ROOT, "include", "gaot_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(gaot_[a-z0-9_]+)\s*\(", hdr))
    assert declared, "no declarations parsed"
    assert declared == set(_lib.EXPORTED_SYMBOLS), declared ^ set(_lib.EXPORTED_SYMBOLS)
    for name in declared:
        assert hasattr(lib, name), name
    assert lib.gaot_abi_version() == 1
    # size queries are pure host arithmetic
    assert lib.gaot_radius_workspace_bytes(500000, 131072) > 0
    assert lib.gaot_attn_workspace_bytes(1, 16384, 8, 8, 32) > 0


def test_no_cpu_fallback():
    import gaot_3d_b200 as G
    p = torch.rand(10, 3)
    with pytest.raises(RuntimeError, match="GPU only|no CPU fallback"):
        G.get_neighbor_strategy("knn", p, None, p, None, 0.1)
    it = G.IntegralTransform(channel_mlp_layers=[6, 8, 4])
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        it(p, p, torch.zeros(2, 3, dtype=torch.long), torch.rand(10, 4))
    # the reference's zero-edge early-out needs no kernel and keeps working
    assert it(p, p, torch.empty(2, 0, dtype=torch.long), None).shape == (10, 4)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        G.ops.attention(torch.rand(1, 8, 64), torch.rand(1, 8, 64), torch.rand(1, 8, 64), 2, 2)


def test_product_does_not_import_oracle():
    """The oracle is test infrastructure: nothing under gaot_3d_b200/ may reference it."""
    for dp, _, fs in os.walk(os.path.join(ROOT, "gaot_3d_b200")):
        for f in fs:
            if f.endswith((".py", ".cu", ".cuh")):
                src = open(os.path.join(dp, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), os.path.join(dp, f)
                assert "/root/reference" not in src, os.path.join(dp, f)


def test_host_logic():
    import gaot_3d_b200 as G
    from gaot_3d_b200.graph import apply_neighbor_sampling, parse_geoembed_strategy, _example_slices
    assert G.parse_neighbor_strategy("knn") == ("knn", "knn")
    assert G.parse_neighbor_strategy(["radius", "reverse"]) == ("radius", "reverse")
    with pytest.raises(ValueError):
        G.parse_neighbor_strategy(["a", "b", "c"])
    assert parse_geoembed_strategy([True, False]) == (True, False)
    with pytest.raises(ValueError):
        parse_geoembed_strategy("yes")
    ei = torch.tensor([[0, 1, 2, 3, 4, 5], [0, 0, 0, 1, 1, 2]])
    assert apply_neighbor_sampling(ei, 3, None, None) is ei
    assert apply_neighbor_sampling(ei, 3, None, "ratio", sample_ratio=1.0) is ei
    assert apply_neighbor_sampling(ei, 3, None, "ratio", sample_ratio=0.5, training=False) is ei
    out = apply_neighbor_sampling(ei, 3, torch.device("cpu"), "max_neighbors", max_neighbors=2)
    assert torch.bincount(out[1], minlength=3).tolist() == [2, 2, 1]
    with pytest.raises(ValueError):
        apply_neighbor_sampling(ei, 3, None, "bogus")
    with pytest.raises(ValueError):
        apply_neighbor_sampling(ei, 3, None, "ratio")
    b = torch.tensor([0, 0, 1, 1, 1, 3])
    assert _example_slices(b, 6, 4) == [(0, 2), (2, 5), (5, 5), (5, 6)]
    # module shells: error behaviour of the reference
    mc = G.MAGNOConfig(gno_coord_dim=3, lifting_channels=8, mlp_type="linear", use_geoembed=False, precompute_edges=True)
    enc = G.MAGNOEncoder(3, 8, mc)
    bt = G.Batch(pos=torch.rand(5, 3), x=torch.rand(5, 3))
    with pytest.raises(AttributeError, match="encoder_edge_index_s0"):
        enc(bt, torch.rand(8, 3), torch.zeros(8, dtype=torch.long))
    mc2 = G.MAGNOConfig(gno_coord_dim=3, lifting_channels=8, mlp_type="linear", use_geoembed=False, encoder_feature_attr="nope")
    with pytest.raises(AttributeError, match="nope"):
        G.MAGNOEncoder(3, 8, mc2)(bt, torch.rand(8, 3), torch.zeros(8, dtype=torch.long))
    with pytest.raises(ValueError):
        G.init_model(3, 1, "other")


def test_head_parallel_column_bookkeeping():
    """Sharded mode splits the attention core by heads (tblock.head_slices / merge_head_slices): slicing every rank's heads
    out of [q | k | v] and merging the per-rank blocks back is the identity, with and without grouped kv heads; the
    eligibility rule needs both head counts to divide by the world size."""
    from gaot_3d_b200 import tblock
    torch.manual_seed(0)
    for H, Hkv, d, R in ((8, 8, 32, 2), (8, 8, 32, 8), (8, 4, 32, 4), (4, 2, 64, 2)):
        qkv = torch.randn(17, (H + 2 * Hkv) * d)
        parts = [tblock.head_slices(qkv, r, R, H, Hkv, d) for r in range(R)]
        assert all(p.shape == (17, (H // R + 2 * (Hkv // R)) * d) and p.is_contiguous() for p in parts)
        assert torch.equal(tblock.merge_head_slices(parts, H, Hkv, d), qkv)
        # rank r's block holds exactly its q heads and the kv heads those q heads attend to
        r = R - 1
        Ha, g = H // R, H // Hkv
        q_heads = range(r * Ha, (r + 1) * Ha)
        kv_needed = sorted({h // g for h in q_heads})
        assert kv_needed == list(range(r * (Hkv // R), (r + 1) * (Hkv // R)))
    assert tblock._head_shard(8, 8) is None                   # off unless the sharded forward switches it on
    import gaot_3d_b200 as G
    with pytest.raises(ValueError):
        G.set_node_mlp_mode("fast")


def test_search_cache_scope_logic():
    """graph._cached / sample_scope (host logic, no kernels): nothing is cached outside a scope; inside one the same tensors +
    parameters hit, different parameters or an in-place update miss; the LRU is bounded; the outermost exit drops everything."""
    from gaot_3d_b200 import graph
    calls = []

    def build():
        calls.append(1)
        return torch.zeros(2, 3, dtype=torch.long)

    x, y = torch.rand(5, 3), torch.rand(4, 3)
    graph._cached("knn", (x, y, None, None), (1,), build)
    graph._cached("knn", (x, y, None, None), (1,), build)
    assert len(calls) == 2 and not graph._CACHE["entries"], "no caching outside a sample scope"
    with graph.sample_scope():
        a = graph._cached("knn", (x, y, None, None), (1,), build)
        with graph.sample_scope():                                  # nests
            b = graph._cached("knn", (x, y, None, None), (1,), build)
        assert a is b and len(calls) == 3
        assert graph._CACHE["entries"], "the inner exit must not clear the outer scope's entries"
        graph._cached("knn", (x, y, None, None), (2,), build)        # other parameters
        graph._cached("radius", (x, y, None, None), (1,), build)      # other kind
        assert len(calls) == 5
        x.add_(1.0)                                                   # in-place update: version counter moves
        graph._cached("knn", (x, y, None, None), (1,), build)
        assert len(calls) == 6
        for i in range(3, 3 + 2 * graph._CACHE["max"]):
            graph._cached("knn", (x, y, None, None), (i,), build)
        assert len(graph._CACHE["entries"]) <= graph._CACHE["max"]
    assert not graph._CACHE["entries"] and graph._CACHE["depth"] == 0
    graph.set_graph_cache(False)
    try:
        with graph.sample_scope():
            n = len(calls)
            graph._cached("knn", (x, y, None, None), (1,), build)
            graph._cached("knn", (x, y, None, None), (1,), build)
            assert len(calls) == n + 2
    finally:
        graph.set_graph_cache(True)


def test_peer_collectives_decline_cpu_tensors():
    """p2p.* answer None / False for anything the peer path cannot carry (CPU tensors, odd sizes): callers then use
    torch.distributed.  No process group is needed to find that out."""
    from gaot_3d_b200 import p2p
    t = torch.rand(2, 8, 4)
    assert p2p.all_to_all(t) is None and p2p.all_gather(t) is None and p2p.reduce_scatter(t) is None
    assert p2p.all_reduce_(torch.rand(10)) is False
    assert p2p.backend(None) == "nccl"


def test_bench_reference_arm_contract():
    """`bench.py --impl reference` (the CPU arm the driver runs beside ours) prints ONE JSON line carrying the contract's keys;
    non-zero ranks of a torchrun launch print nothing and exit 0."""
    import json
    import subprocess
    import sys
    env = dict(os.environ, OMP_NUM_THREADS="1")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", "small", "--steps", "1",
                        "--warmup", "1"], capture_output=True, text=True, timeout=600, env=env)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, lines
    d = json.loads(lines[0])
    for k in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "dtype",
              "data", "config", "cpu_baseline", "e2e"):
        assert k in d, k
    assert d["impl"] == "reference" and d["unit"] == "samples/s" and d["higher_is_better"] is True and d["extrapolated"] is False
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and abs(d["ms_per_step"] * d["value"] - 1e3) < 1e-6 * 1e3
    assert d["cpu_baseline"]["cores"] == (len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else os.cpu_count()), \
        "the CPU arm must not inherit torchrun's OMP_NUM_THREADS=1"
    r1 = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", "small", "--steps", "1"],
                        capture_output=True, text=True, timeout=120, env=dict(env, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1"))
    assert r1.returncode == 0 and r1.stdout.strip() == ""
