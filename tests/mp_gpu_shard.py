"""Launched by tests/test_gpu_multi.py under torchrun (one rank per GPU): a row-sharded chain over WORLD
GPUs against the same chain on one GPU.  Prints "SHARD_OK <agreement>" on rank 0."""
import os
import sys
from pathlib import Path

import numpy as np
import torch
import torch.distributed as dist

ROOT = Path(__file__).resolve().parents[1]
for p in (ROOT, ROOT / "multiview-clustering_b200", ROOT / "tests"):
    sys.path.insert(0, str(p))
import mvc_b200  # noqa: E402
from conftest import make_mixture  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    dist.init_process_group("gloo")
    torch.cuda.set_device(local)
    n, dims, cap, k_true = 128 * 300 + 19, [64, 64, 64], 64, 50
    views, z = make_mixture(n, dims, k_true, seed=17)
    rng = np.random.default_rng(3)
    tab = np.where(rng.random(n) < 0.1, rng.integers(0, k_true, n), z).astype(np.int32)
    dish = np.full((3, cap), -1, np.int32)
    dish[:, :k_true] = np.arange(k_true)
    hyp = ([1.0] * 3, [0.5] * 3, [1.0] * 3, 1.0, 0.6)
    lo, hi = rank * n // world, (rank + 1) * n // world
    s = mvc_b200.Sampler(hi - lo, dims, cap=cap, seed=77, device=local, engine=2, rank=rank, world=world,
                         row_offset=lo, n_rows_global=n)
    uid = [mvc_b200.Sampler.nccl_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(uid, src=0)
    s.comm_init_rank(uid[0])
    for v in range(3):
        s.upload_view(v, views[v][lo:hi])
    if os.environ.get("MVG_TEST_P2P") == "1":                 # packets over NVLink peer memory instead of ncclAllGather
        handles = [None] * world
        dist.all_gather_object(handles, s.p2p_export())
        s.p2p_attach(handles)
        try:                                                  # a second attach is refused: mappings and sequence number persist
            s.p2p_attach(handles)
            raise SystemExit("second p2p_attach was accepted")
        except mvc_b200.MvgError as e:
            assert e.code == -4, e
    if os.environ.get("MVG_TEST_INCR") == "1":                # incremental statistics on every shard
        s.set_stats_mode(True, 64)
    s.set_state(tab[lo:hi], dish, *hyp)
    s.sweep(1, do_hyper=True)
    st1 = s.get_state()
    s.sweep(3, do_hyper=True)
    st = s.get_state()
    parts = [None] * world
    dist.all_gather_object(parts, (lo, st1["table_of"], st["table_of"], st["n_t"], st["dish_of"], st["tau_v"], st["alpha_g"]))
    s.close()
    if rank == 0:
        parts.sort(key=lambda p: p[0])
        for p in parts[1:]:                                   # replicated state is identical on every rank
            np.testing.assert_array_equal(p[3], parts[0][3])
            np.testing.assert_array_equal(p[4], parts[0][4])
            np.testing.assert_array_equal(p[5], parts[0][5])
            assert p[6] == parts[0][6]
        tab1 = np.concatenate([p[1] for p in parts])
        tabN = np.concatenate([p[2] for p in parts])
        assert int(parts[0][3].sum()) == n and np.array_equal(np.bincount(tabN, minlength=cap), parts[0][3])
        one = mvc_b200.Sampler(n, dims, cap=cap, seed=77, device=local, engine=2)
        for v in range(3):
            one.upload_view(v, views[v])
        one.set_state(tab, dish, *hyp)
        one.sweep(1, do_hyper=True)
        a1 = float((one.get_state()["table_of"] == tab1).mean())
        one.sweep(3, do_hyper=True)
        ref = one.get_state()
        aN = float((ref["table_of"] == tabN).mean())
        one.close()
        # one sweep from the same state: the statistics are identical, the draws are addressed by global row
        assert a1 == 1.0, a1
        # later sweeps: FP32 partial sums are grouped differently per shard, so parameters may differ in the
        # last bit and a draw sitting on a CDF edge may flip
        assert aN > 0.999, aN
        np.testing.assert_allclose(ref["tau_v"], parts[0][5], rtol=1e-6)
        print(f"SHARD_OK world={world} first_sweep={a1} after4={aN} p2p={os.environ.get('MVG_TEST_P2P', '0')}")
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
