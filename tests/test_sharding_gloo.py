"""Row sharding (N > 1 ranks) on CPU: two gloo ranks each hold half of the customers.

What the multi-GPU path relies on, checked here without a GPU through the CPU oracle:
  * the synthetic data of bench.py is generated per row block, so a shard equals the slice of the whole;
  * the Philox stream is addressed by the GLOBAL row index (row_offset), so draws do not depend on sharding;
  * given the summed per-table statistics (the one exchange per sweep), a shard's draws are the
    corresponding slice of the single-process draws — counts exactly, float sums to rounding.
"""
import os
import socket
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parents[1]


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out_dir):
    import torch
    import torch.distributed as dist
    for p in (ROOT, ROOT / "oracle", ROOT / "tests"):
        sys.path.insert(0, str(p))
    import pyoracle as po
    from conftest import make_mixture
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    n, dims, cap, k_true = 640, [6, 4], 32, 7
    views, z = make_mixture(n, dims, k_true, seed=5)
    rng = np.random.default_rng(9)
    tab = np.where(rng.random(n) < 0.2, rng.integers(0, k_true + 2, n), z).astype(np.int32)
    dish = np.full((2, cap), -1, np.int32)
    dish[:, :k_true + 2] = rng.integers(0, 5, (2, k_true + 2))
    lo, hi = rank * n // world, (rank + 1) * n // world

    # the rank's shard, with global Philox addressing
    o = po.OracleState([v[lo:hi] for v in views], cap, seed=31, row_offset=lo, n_global=n)
    o.sweep = 4
    o.tau_v[:] = 0.8
    o.set_assignment(tab[lo:hi], dish)                      # statistics of the shard's rows only
    # the one exchange per sweep: element-wise sums of the per-table / per-dish statistics
    for a in [o.n_t, o.n_vk] + o.S1 + [o.S2]:
        t = torch.from_numpy(a)
        dist.all_reduce(t)
    live = o.n_t > 0
    for v in range(2):                                      # tables per dish from the global table counts
        o.l_vk[v] = np.bincount(dish[v][live & (dish[v] >= 0)], minlength=cap)[:cap]
    mine = o.draw_rows()

    gathered = [None] * world
    dist.all_gather_object(gathered, (lo, hi, mine))
    # an opaque 128-byte id travels from rank 0 to everyone exactly as bench.py ships the NCCL unique id
    uid = [bytes(range(128)) if rank == 0 else None]
    dist.broadcast_object_list(uid, src=0)
    assert uid[0] == bytes(range(128))
    if rank == 0:
        full = po.OracleState(views, cap, seed=31)
        full.sweep = 4
        full.tau_v[:] = 0.8
        full.set_assignment(tab, dish)
        want = full.draw_rows()
        got = np.concatenate([g[2] for g in sorted(gathered, key=lambda g: g[0])])
        np.testing.assert_array_equal(o.n_t, full.n_t)
        np.testing.assert_array_equal(o.n_vk, full.n_vk)
        np.testing.assert_array_equal(o.l_vk, full.l_vk)
        np.testing.assert_allclose(o.S2, full.S2, rtol=1e-12)
        agree = float((got == want).mean())
        Path(out_dir, "ok").write_text(f"{agree}")
        assert agree == 1.0, agree                          # FP64 oracle: sums differ by rounding only
    dist.barrier()
    dist.destroy_process_group()


def test_two_gloo_ranks_draw_the_slices_of_the_single_process_draws(tmp_path, oracle):
    import torch.multiprocessing as mp
    mp.spawn(_worker, args=(2, _free_port(), str(tmp_path)), nprocs=2, join=True)
    assert float((tmp_path / "ok").read_text()) == 1.0


def test_bench_rows_are_shard_invariant():
    sys.path.insert(0, str(ROOT))
    import bench
    mus = bench.planted_means(np.random.default_rng(bench.SEED))
    whole, z = bench.make_rows_numpy(0, 200_000, mus)
    for lo, hi in [(0, 70_000), (70_000, 131_073), (131_073, 200_000)]:
        part, zp = bench.make_rows_numpy(lo, hi, mus)
        np.testing.assert_array_equal(zp, z[lo:hi])
        for v in range(3):
            np.testing.assert_array_equal(part[v], whole[v][lo:hi])
