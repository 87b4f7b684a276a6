"""Small C3-shaped sweeps for compute-sanitizer (racecheck / synccheck / memcheck):

    compute-sanitizer --tool racecheck python profiles/sanitize_small.py

Both draw engines, with and without free table slots, two sweeps each with the hyper step.  Sizes are kept
small because the sanitizer serialises the kernels; the shapes (tiles per CTA > 1, ragged last tile) are those
of the headline configuration."""
import os
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
for p in (ROOT, ROOT / "multiview-clustering_b200"):
    sys.path.insert(0, str(p))
import mvc_b200  # noqa: E402


def run(engine, k_true, n):
    dims, cap = [64, 64, 64], 64
    rng = np.random.default_rng(5)
    z = rng.integers(0, k_true, n)
    views = []
    for d in dims:
        mu = rng.normal(0, 2, (k_true, d))
        views.append((mu[z] + rng.normal(0, 1, (n, d))).astype(np.float32))
    tab = np.where(rng.random(n) < 0.1, rng.integers(0, k_true, n), z).astype(np.int32)
    dish = np.full((3, cap), -1, np.int32)
    dish[:, :k_true] = np.arange(k_true)
    s = mvc_b200.Sampler(n, dims, cap=cap, seed=11, engine=engine)
    for v, x in enumerate(views):
        s.upload_view(v, x)
    s.set_state(tab, dish, [1.0] * 3, [0.5] * 3, [1.0] * 3, 1.0, 0.6)
    s.sweep(2, do_hyper=True)
    st = s.get_state()
    assert int(st["n_t"].sum()) == n
    s.close()
    print(f"engine {engine} k_true {k_true} n {n}: ok, {int((st['n_t'] > 0).sum())} tables live", flush=True)


if __name__ == "__main__":
    n = int(os.environ.get("SAN_ROWS", 128 * 148 * 2 + 37))
    for engine in (2, 1):
        for k_true in (64, 60):
            run(engine, k_true, n if engine == 2 else min(n, 4096 + 37))
