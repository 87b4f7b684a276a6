"""A short C3-shaped run for ncu: N = 1M, 3 x 64 dims, cap 64; `python profiles/ncu_c3.py [k_true] [sweeps]`."""
import sys
from pathlib import Path
import numpy as np
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "multiview-clustering_b200"))
import bench, mvc_b200

k_true = int(sys.argv[1]) if len(sys.argv) > 1 else 64
sweeps = int(sys.argv[2]) if len(sys.argv) > 2 else 4
n = int(sys.argv[3]) if len(sys.argv) > 3 else bench.N_ROWS
mus = bench.planted_means(np.random.default_rng(bench.SEED))
views, z = bench.make_rows_numpy(0, n, mus, k_true=k_true)
tab, dish, hyp = bench.initial_state(z, k_true)
s = mvc_b200.Sampler(n, bench.DIMS, cap=64, seed=bench.SEED)
for v in range(3):
    s.upload_view(v, views[v])
s.set_state(tab, dish, hyp["alpha_v"], hyp["sigma_v"], hyp["tau_v"], hyp["alpha_g"], hyp["sigma_g"])
import os
os.environ["MVG_NO_GRAPHS"] = "1"
for _ in range(sweeps):
    s.sweep(1, True)
s.sync()
print("ok", s.get_state(with_rows=False)["n_t"].sum())
s.close()
