"""Where the host-buffer path spends its time: per-call wall clock of upload / set_state / run on a fresh handle."""
import sys, time
from pathlib import Path
import numpy as np, torch
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "multiview-clustering_b200"))
import mvc_b200, bench

mus = bench.planted_means(np.random.default_rng(bench.SEED))
views, z = bench.make_rows_numpy(0, 1_000_000, mus)
tab, dish, hyp = bench.initial_state(z)
pinned = [torch.from_numpy(v).pin_memory() for v in views]
torch.cuda.synchronize()
for rep in range(3):
    t = [time.perf_counter()]
    s = mvc_b200.Sampler(1_000_000, bench.DIMS, cap=64, seed=1999)
    t.append(time.perf_counter())
    for v in range(3):
        s.upload_view(v, pinned[v].numpy()); t.append(time.perf_counter())
    s.sync(); t.append(time.perf_counter())
    s.set_state(tab, dish, hyp["alpha_v"], hyp["sigma_v"], hyp["tau_v"], hyp["alpha_g"], hyp["sigma_g"]); t.append(time.perf_counter())
    s.sweep(200, True); s.sync(); t.append(time.perf_counter())
    st = s.get_state(); t.append(time.perf_counter())
    s.close(); t.append(time.perf_counter())
    names = ["create", "up0", "up1", "up2", "sync", "set_state", "200 sweeps", "get_state", "close"]
    print(rep, {n: round(1e3 * (b - a), 2) for n, a, b in zip(names, t, t[1:])}, flush=True)
