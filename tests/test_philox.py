"""Philox4x32-10: Random123 known answers, oracle restatement == product host mirror."""
import ctypes as C

import numpy as np

KAT = [   # Random123 kat_vectors, philox4x32 10 rounds
    ([0, 0, 0, 0], [0, 0], [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]),
    ([0xffffffff] * 4, [0xffffffff] * 2, [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]),
    ([0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344], [0xa4093822, 0x299f31d0],
     [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]),
]


def _raw(fn, ctr, key):
    c = (C.c_uint32 * 4)(*ctr)
    k = (C.c_uint32 * 2)(*key)
    o = (C.c_uint32 * 4)()
    fn(c, k, o)
    return list(o)


def test_oracle_known_answers(oracle):
    for ctr, key, want in KAT:
        assert _raw(oracle.lib().mvo_philox_raw, ctr, key) == want


def test_product_known_answers():
    import mvc_b200
    for ctr, key, want in KAT:
        assert _raw(mvc_b200.lib().mvg_philox4x32_10, ctr, key) == want


def test_product_mirror_equals_oracle_stream(oracle):
    import mvc_b200
    L, P = oracle.lib(), mvc_b200.lib()
    rng = np.random.default_rng(3)
    for _ in range(500):
        seed = int(rng.integers(0, 2**63))
        chain, dom, slot, sweep = (int(rng.integers(0, 2**31)), int(rng.integers(0, 7)), int(rng.integers(0, 16)),
                                   int(rng.integers(0, 2**31)))
        idx = int(rng.integers(0, 2**40))
        assert L.mvo_uf(seed, chain, dom, slot, sweep, idx) == P.mvg_philox_uniform_f32(seed, chain, dom, slot, sweep, idx)
        assert L.mvo_u53(seed, chain, dom, slot, sweep, idx) == P.mvg_philox_uniform_f64(seed, chain, dom, slot, sweep, idx)
        assert abs(L.mvo_z(seed, chain, dom, slot, sweep, idx) - P.mvg_philox_normal(seed, chain, dom, slot, sweep, idx)) < 1e-12


def test_uniform_open_interval_and_moments(oracle):
    L = oracle.lib()
    u = np.array([L.mvo_uf(1999, 0, 0, 0, 7, i) for i in range(20000)])
    assert u.min() > 0.0 and u.max() < 1.0
    assert abs(u.mean() - 0.5) < 0.01 and abs(u.var() - 1 / 12) < 0.005
    z = np.array([L.mvo_z(1999, 0, 2, 0, 7, i) for i in range(20000)])
    assert abs(z.mean()) < 0.03 and abs(z.std() - 1) < 0.03
    # different sweeps / rows / domains give different numbers
    assert L.mvo_uf(1, 0, 0, 0, 0, 5) != L.mvo_uf(1, 0, 0, 0, 1, 5) != L.mvo_uf(1, 0, 1, 0, 0, 5)
