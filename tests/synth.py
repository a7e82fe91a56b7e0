"""Synthetic inputs of the BASELINE shapes: thin-shell surface clouds on the
dataset's domain box, rescaled to [-1,1] with the global scalar min/max (reference
src/utils/scale.py:13-25), and the trainer's latent grid (src/trainer/stat.py:239-252)."""
import numpy as np

BOXES = {
    "drivaernet": ([-1.16, -1.20, 0.0], [4.21, 1.19, 1.77]),      # reference src/data/metadata.py:32
    "drivaerml": ([-0.943, -1.14, -0.318], [4.14, 1.14, 1.25]),   # metadata.py:117
    "crm": ([2.3495, -29.460142, 2.3101413], [66.744965, 29.460142, 8.833843]),   # NASA CRM, metadata.py:66
    "unit": ([-1.0, -1.0, -1.0], [1.0, 1.0, 1.0]),
}


def surface_cloud(n, box="drivaernet", seed=0):
    rng = np.random.default_rng(seed)
    lo, hi = (np.asarray(b, dtype=np.float64) for b in BOXES[box])
    p = rng.uniform(lo, hi, (n, 3))
    f = rng.integers(0, 6, n)
    ax, side = f // 2, f % 2
    p[np.arange(n), ax] = np.where(side == 0, lo[ax], hi[ax])
    return _rescale(p, box)


def latent_grid(shape, box="drivaernet"):
    lo, hi = BOXES[box]
    ax = [np.linspace(lo[a], hi[a], shape[a]) for a in range(3)]
    g = np.stack(np.meshgrid(*ax, indexing="ij"), -1).reshape(-1, 3)
    return _rescale(g, box)


def _rescale(p, box):
    lo, hi = BOXES[box]
    mn, mx = min(lo), max(hi)
    return ((p - mn) / (mx - mn) * 2.0 - 1.0).astype(np.float32)


def unit_normals(n, seed=0):
    rng = np.random.default_rng(seed + 1000)
    v = rng.normal(size=(n, 3))
    return (v / np.linalg.norm(v, axis=1, keepdims=True)).astype(np.float32)


def mach_aoa(n, seed=0):
    """NASA-CRM condition features: one (Mach, AOA) pair per sample, broadcast to every point (metadata.py:73)."""
    rng = np.random.default_rng(seed + 2000)
    return np.tile(np.array([[rng.uniform(0.7, 0.9), rng.uniform(0.0, 4.0)]], dtype=np.float32), (n, 1))
