"""Properties of the CPU oracle itself: FP32 mirror vs FP64 restatement, leave-one-out, reseating,
hyper-step closed forms, and that the synchronous sampler recovers planted clusters."""
import numpy as np
import pytest
from conftest import c1_data, make_mixture


def test_exp2m_log2m_accuracy(oracle):
    L = oracle.lib()
    d = np.concatenate([np.linspace(-124.5, 0, 4001), -np.logspace(-6, 1, 200)]).astype(np.float32)
    got = np.array([L.mvo_exp2m(float(x)) for x in d], np.float64)
    np.testing.assert_allclose(got, np.exp2(d.astype(np.float64)), rtol=4e-7)
    assert L.mvo_exp2m(0.0) == 1.0
    assert 0.0 < L.mvo_exp2m(-1e30) < 1e-37              # masked options get a negligible, positive weight
    s = np.concatenate([np.linspace(1, 70, 3001), np.logspace(-20, 20, 300)]).astype(np.float32)
    got = np.array([L.mvo_log2m(float(x)) for x in s], np.float64)
    np.testing.assert_allclose(got, np.log2(s.astype(np.float64)), rtol=3e-7, atol=3e-7)


@pytest.mark.parametrize("dims,cap", [([1, 1], 32), ([8, 8, 8], 32), ([64, 64, 64], 64), ([5, 12], 64)])
def test_fp32_mirror_agrees_with_fp64_restatement(oracle, dims, cap):
    views, z = make_mixture(600, dims, 5, seed=11)
    s = oracle.OracleState(views, cap, seed=42)
    dish = np.full((len(dims), cap), -1, np.int32)
    dish[:, :5] = np.arange(5)
    # a slightly scrambled seating so that plenty of rows want to move
    rng = np.random.default_rng(0)
    tab = np.where(rng.random(600) < 0.2, rng.integers(0, 5, 600), z)
    s.set_assignment(tab, dish)
    s.tau_v[:] = 0.8
    ch64 = s.draw_rows()
    ch32 = oracle.mirror_draw_rows(s)
    assert (ch64 == ch32).mean() > 0.995                  # draws differ only at CDF edges
    P = s.make_params()
    ps = oracle.params_struct(P)
    worst = 0.0
    for i in range(0, 600, 13):
        a = np.stack([oracle.stageA_f32(views[v][i], P["m"][v])[0] for v in range(len(dims))])
        q = np.array([oracle.stageA_f32(views[v][i], P["m"][v])[1] for v in range(len(dims))])
        _, lw32 = oracle.stageB_f32(ps, a, q, tab[i], 0.5, want_lw=True)
        lw64 = s.row_logweights(i) / np.log(2.0)
        ok = np.isfinite(lw64)
        assert np.all(lw32[~ok] < -1e29)
        worst = max(worst, np.max(np.abs(lw32[ok] - lw64[ok]) / np.maximum(1.0, np.abs(lw64[ok]))))
    assert worst < 1e-5                                    # the stated FP32 tolerance


@pytest.mark.parametrize("shared_dishes", [False, True])
def test_tensor_core_mirror_agrees_with_fp64_restatement(oracle, shared_dishes):
    """The CPU mirror of the tensor-core engine's epilogue (mvo_stageB_tc: pre-scaled means b = 2 A m, scalar
    leave-one-out corrections, new-table log-weight taken as an input) against the FP64 restatement, on the CPU alone:
    fed FP32 dot products x.b and the FP64 new-table log-weight it must reproduce the log-weights to the stated 1e-5 and
    the draws but for CDF edges.  With the FP64 restatement pinned to the compiled reference at D = 1
    (test_oracle_vs_reference.py) this closes the chain  GPU kernel == mirror (bit-exact, GPU tests)  ~  restatement
    (here)  ==  reference  without a GPU.  Free table slots (the new-table option has weight) and, in the second case,
    tables that share dishes (the leave-one-out correction reaches several tables)."""
    n, dims, cap, k_true = 400, [64, 64, 64], 64, 6
    views, z = make_mixture(n, dims, k_true, seed=17)
    rng = np.random.default_rng(3)
    n_tab = 12 if shared_dishes else k_true
    tab = np.where(rng.random(n) < 0.15, rng.integers(0, n_tab, n), z + (k_true * (rng.random(n) < 0.5) if shared_dishes else 0))
    tab = tab.astype(np.int32)
    dish = np.full((len(dims), cap), -1, np.int32)
    dish[:, :n_tab] = np.arange(n_tab) % k_true              # shared: tables t and t + 6 serve the same dish
    s = oracle.OracleState(views, cap, seed=7)
    s.set_assignment(tab, dish)
    s.tau_v[:] = 0.9
    P = s.make_params()
    ps = oracle.params_struct(P)
    ch64 = s.draw_rows()
    L = oracle.lib()
    agree, worst = 0, 0.0
    for i in range(n):
        acc = np.stack([(views[v][i].astype(np.float64) @ oracle.scaled_means(P["A"][v], P["m"][v]).astype(np.float64).T)
                        .astype(np.float32) for v in range(len(dims))])
        xx = np.array([float((views[v][i].astype(np.float64) ** 2).sum()) for v in range(len(dims))], np.float32)
        lw64 = s.row_logweights(i) / np.log(2.0)
        lnew = np.float32(lw64[cap]) if np.isfinite(lw64[cap]) else np.float32(-1e30)
        u = L.mvo_uf(7, 0, 0, 0, s.sweep, i)
        ch, lw32 = oracle.stageB_tc(ps, acc, xx, tab[i], u, lnew, want_lw=True)
        agree += int(ch == ch64[i])
        ok = np.isfinite(lw64)
        assert np.all(lw32[~ok] < -1e29)
        worst = max(worst, np.max(np.abs(lw32[ok] - lw64[ok]) / np.maximum(1.0, np.abs(lw64[ok]))))
    assert worst < 1e-5, worst
    assert agree / n > 0.99, agree


def test_dominant_option_shortcut_holds_in_the_mirror_arithmetic(oracle):
    """The tensor-core kernel skips exponentials, totals and the scan when one option dominates: every other table at
    least 31 below it in log2 and the new-table option at least 27 (csrc/mv_draw_tc.cu).  The claim behind it — the
    full FP32 computation then returns the dominant option for EVERY uniform the stream can produce — is checked here
    against the mirror that always runs the full computation, at the worst case the thresholds admit: all 63 other
    tables exactly 31 below, the new table exactly 27 below, the smallest and the largest uniform.  (Parameters chosen
    so that the log-weights are the inputs themselves: one view, A = C = 0, so lw[t] = acc[t] and lw[t0] = 0.)"""
    cap, t0 = 64, 17
    z = np.zeros((1, cap), np.float32)
    P = {"dish": np.arange(cap, dtype=np.int32)[None, :].copy(), "A": z.copy(), "C": z.copy(), "A1": z.copy(), "C1": z.copy(),
         "W": z.copy(), "W1": z.copy(), "lone": np.ones((1, cap), np.int32), "AN": np.zeros(1, np.float32),
         "CN": np.zeros(1, np.float32), "WN": np.zeros(2, np.float32), "LD": np.zeros(2, np.float32),
         "LM": np.zeros(cap, np.float32), "LM1": np.zeros(cap, np.float32), "single": np.zeros(cap, np.int32),
         "LMN": np.zeros(2, np.float32)}
    ps = oracle.params_struct(P)
    xx = np.zeros(1, np.float32)
    u_lo, u_hi = np.float32(2.0 ** -24), np.float32(1.0 - 2.0 ** -24)
    for winner in ("own", "other-low", "other-high", "new"):
        acc = np.full((1, cap), -31.0, np.float32)            # lw[t] = acc[t]; lw[t0] = 0 whatever acc[t0] is
        lnew, want, shift = np.float32(-27.0), t0, 0.0
        if winner.startswith("other"):
            want, shift = (3 if winner == "other-low" else 60), 40.0   # a table before / after the own one in the scan
            acc[:] = shift - 31.0
            acc[0, want] = shift
            lnew = np.float32(shift - 27.0)                    # (the own table, lw = 0, is 40 below: far)
        if winner == "new":
            want, lnew = -1, np.float32(40.0)
            acc[:] = 40.0 - 31.0                               # every table (the own one, lw = 0, too) at least 31 below
        for u in (u_lo, np.float32(0.5), u_hi):
            ch, lw = oracle.stageB_tc(ps, acc, xx, t0, u, lnew, want_lw=True)
            assert lw[t0] == 0.0
            assert ch == want, (winner, float(u), ch)
    # random admissible configurations: any winner, the others anywhere at or below the thresholds
    rng = np.random.default_rng(12)
    for _ in range(300):
        top = np.float32(rng.uniform(32.0, 60.0))             # (the own table sits at 0: keep it far as well)
        acc = (top - 31.0 - rng.exponential(8.0, (1, cap))).astype(np.float32)
        want = int(rng.integers(-1, cap))
        lnew = np.float32(top - 27.0 - rng.exponential(8.0))
        if want == -1:
            lnew = top
        elif want == t0:
            acc = (-31.0 - rng.exponential(8.0, (1, cap))).astype(np.float32)
            lnew = np.float32(-27.0 - rng.exponential(8.0))
        else:
            acc[0, want] = top
        for u in (u_lo, u_hi, np.float32(rng.uniform(0.0, 1.0))):
            assert oracle.stageB_tc(ps, acc, xx, t0, u, lnew) == want
    # control: with the other tables only 20 below, the smallest uniform does land on one of them
    acc = np.full((1, cap), -20.0, np.float32)
    assert oracle.stageB_tc(ps, acc, xx, t0, u_lo, np.float32(-27.0)) != t0


def test_leave_one_out_equals_explicit_removal(oracle):
    """Row weights must equal what one gets by really deleting the row and rebuilding the state."""
    views, z = make_mixture(80, [3, 2], 4, seed=2)
    cap = 16
    dish = np.full((2, cap), -1, np.int32)
    tab = z * 2
    tab[7] = 9                                             # a customer alone at its table ...
    for t in set(tab.tolist()):
        dish[:, t] = [t % 3, (t + 1) % 3]
    dish[1, 9] = 7                                         # ... eating a dish nobody else eats
    s = oracle.OracleState(views, cap)
    s.set_assignment(tab, dish)
    s.tau_v[:] = [0.7, 1.3]
    for i in (0, 7, 33, 79):
        keep = np.arange(80) != i
        s2 = oracle.OracleState([v[keep] for v in views] , cap)
        s2.set_assignment(tab[keep], dish)
        s2.tau_v[:] = s.tau_v
        # score row i against the reduced state: append it as a fresh singleton at a free slot
        free = [t for t in range(cap) if s2.n_t[t] == 0][0]
        s3 = oracle.OracleState([np.vstack([v[keep], v[i:i + 1]]) for v in views], cap)
        d3 = s2.dish_of.copy()
        d3[:, free] = [5, 5]                               # an otherwise unused dish: removed again by LOO
        s3.set_assignment(np.append(tab[keep], free), d3)
        s3.tau_v[:] = s.tau_v
        lw_a = s.row_logweights(i)
        lw_b = s3.row_logweights(79)
        own = tab[i]
        mask = np.ones(cap + 1, bool)
        mask[[free]] = False
        if s.n_t[own] == 1:
            mask[own] = False
        np.testing.assert_allclose(lw_a[mask], lw_b[mask], rtol=1e-10, atol=1e-10)


def test_reseat_births_overflow_and_deaths(oracle):
    views, z = make_mixture(64, [2], 3, seed=4)
    cap = 8
    dish = np.full((1, cap), -1, np.int32)
    dish[0, :6] = [0, 1, 2, 0, 1, 2]
    tab = np.arange(64) % 6
    s = oracle.OracleState(views, cap, seed=9)
    s.set_assignment(tab, dish)
    choice = tab.copy()
    choice[tab == 5] = 0                                   # table 5 is abandoned (its would-be births stay though)
    births = [3, 10, 11, 40]                               # four rows ask for a new table, two slots are free
    choice[births] = -1
    stay = [b for b in births if tab[b] == 5]
    ns, rows, w = s.reseat(choice, want_births=True)
    assert ns == 2 and list(rows) == births[:2]
    assert s.table_of[births[0]] == 6 and s.table_of[births[1]] == 7
    for b in births[2:]:
        assert s.table_of[b] == tab[b]                     # overflow: stays where it was
    assert s.n_t.sum() == 64 and np.all(s.n_t >= 0)
    assert (s.n_t[5] == len(stay)) and ((s.dish_of[0, 5] == -1) == (len(stay) == 0))
    assert np.all((s.dish_of[0] >= 0) == (s.n_t > 0))
    for k in range(cap):
        assert s.l_vk[0, k] == np.sum(s.dish_of[0] == k)
        assert s.n_vk[0, k] == np.sum(s.n_t[s.dish_of[0] == k])
    np.testing.assert_allclose(s.S1[0].sum(0), views[0].astype(np.float64).sum(0), rtol=1e-12)
    assert np.all(w >= 0) and np.all(w.max(axis=2) == 1.0)


def test_eppf_closed_form_equals_loops(oracle):
    views, z = make_mixture(300, [1, 1], 6, seed=8)
    s = oracle.OracleState(views, 32)
    s.init_reference()
    s.sweep_n(5, do_hyper=True)
    L = oracle.lib()
    for a, sg in [(1.0, 0.5), (0.3, 0.9), (5.0, 0.01), (2.0, 0.999)]:
        for v in range(2):
            np.testing.assert_allclose(L.mvo_log_EPPF_view(s.ref(), v, a, sg, 1), L.mvo_log_EPPF_view(s.ref(), v, a, sg, 0), rtol=1e-11)
        np.testing.assert_allclose(L.mvo_log_EPPF_global(s.ref(), a, sg, 1), L.mvo_log_EPPF_global(s.ref(), a, sg, 0), rtol=1e-11)
    assert L.mvo_log_EPPF_view(s.ref(), 0, 1.0, 1e-7, 1) == -np.inf        # guards of multiview_hyper.cpp:297-298
    assert L.mvo_log_EPPF_view(s.ref(), 0, -0.6, 0.5, 1) == -np.inf


def test_sync_sampler_recovers_planted_clusters(oracle):
    """Config-1-shaped data: two scalar views with 2 and 3 planted groups."""
    from sklearn.metrics import adjusted_rand_score as ari
    views, truth = c1_data(500)
    s = oracle.OracleState(views, 32, seed=1999)
    s.init_reference()
    s.sweep_n(700, threads=4, do_hyper=True)     # the synchronous kernel mixes slower than the sequential one
    lab = s.labels()
    assert ari(truth[0], lab[:, 0]) > 0.85
    assert ari(truth[1], lab[:, 1]) > 0.75
    # kernel variance learnt by the tau update: the compiled reference ends near (1.82, 1.55) on this data
    assert 1.3 < s.tau_v[0] < 2.4 and 1.1 < s.tau_v[1] < 2.1


def test_reference_init_matches_reference_shape(oracle):
    views, _ = c1_data(200)
    s = oracle.OracleState(views, 32, seed=3)
    s.init_reference()
    assert s.n_t[:4].sum() == 200 and np.all(s.n_t[4:] == 0)
    assert np.all(s.dish_of[:, :4] >= 0) and np.all(s.dish_of[:, :4] < 2)
    for v in range(2):
        var = np.var(views[v].astype(np.float64), ddof=1)
        np.testing.assert_allclose(s.tau_v[v], var * 0.25 * 0.01, rtol=1e-12)   # multiview_gibbs.cpp:94
    assert s.alpha_g == 1.0 and s.sigma_g == 0.6 and np.all(s.alpha_v == 1.0) and np.all(s.sigma_v == 0.5)


# ---------------------------------------------------------------------------------------------
# count views (no reference counterpart — the oracle is pinned to the written formula)
# ---------------------------------------------------------------------------------------------
def _count_state(oracle, n=120, W=40, cap=32, k_true=3, seed=5):
    from conftest import make_count_view
    rng = np.random.default_rng(seed)
    z = rng.integers(0, k_true, n)
    cv = make_count_view(n, W, z, k_true, seed=seed, mean_len=12)
    dense = (rng.normal(0, 2, (k_true, 2))[z] + rng.normal(0, 1, (n, 2))).astype(np.float32)
    o = oracle.OracleState([dense, cv], cap, seed=seed)
    o.init_reference()
    return o, cv, z


def test_count_view_loglik_matches_formula(oracle):
    """log f_vk(x) = sum_w x_w log((beta + c_kw - [own] x_w) / (W beta + C_k - [own] |x|)); log f_new = -|x| log W."""
    o, cv, _ = _count_state(oracle)
    W, beta, cap = cv["vocab"], o.count_beta, o.cap
    dense_rows = np.zeros((o.n, W))
    for i in range(o.n):
        j0, j1 = cv["rowptr"][i], cv["rowptr"][i + 1]
        dense_rows[i, cv["col"][j0:j1]] = cv["val"][j0:j1]
    # the rebuilt dish counts are the column sums of the rows of each dish
    lab = o.dish_of[1][o.table_of]
    for k in range(cap):
        np.testing.assert_array_equal(o.cd[1][k], dense_rows[lab == k].sum(0).astype(np.int64))
        assert o.ctot[1][k] == int(dense_rows[lab == k].sum())
    for i in (0, 7, 33, 119):
        _, L = o.row_logweights(i, want_L=True)
        k0 = lab[i]
        x = dense_rows[i]
        for t in range(cap):
            k = o.dish_of[1][t]
            if k < 0 or o.l_vk[1][k] == 0:
                continue
            own = (k == k0)
            c = o.cd[1][k] - (x if own else 0)
            den = W * beta + o.ctot[1][k] - (x.sum() if own else 0)
            want = float((x * np.log((beta + c) / den)).sum())
            np.testing.assert_allclose(L[1][t], want, rtol=1e-12, atol=1e-12)
        np.testing.assert_allclose(L[1][cap], -x.sum() * np.log(W), rtol=1e-12, atol=1e-12)


def test_count_view_chain_keeps_invariants(oracle):
    o, cv, _ = _count_state(oracle, n=200)
    total = int(cv["val"].sum())
    for _ in range(15):
        o.sweep_n(1, threads=2, do_hyper=True)
        assert o.n_t.sum() == o.n and (o.n_vk.sum(1) == o.n).all()
        assert int(o.ctot[1].sum()) == total and int(o.cd[1].sum()) == total
        assert o.tau_v[1] == 1.0                                   # a count view has no kernel variance


def test_count_view_fp32_mirror_tracks_fp64(oracle):
    """The FP32 stage-A mirror of a count row (fmaf chains over log2 theta) against the FP64 log f, and the mixed
    stage-B mirror equals the dense one when no view is a count view."""
    o, cv, _ = _count_state(oracle)
    W, beta, cap = cv["vocab"], o.count_beta, o.cap
    P = o.make_params()
    # tables as the device would build them: per table slot, the dish's counts and log2 theta
    cdt = np.zeros((W, cap), np.int32)
    l2t = np.zeros((W, cap), np.float32)
    for t in range(cap):
        k = o.dish_of[1][t]
        if k >= 0:
            cdt[:, t] = o.cd[1][k]
            l2t[:, t] = np.log2((beta + o.cd[1][k]) / (W * beta + o.ctot[1][k]))
    for i in (3, 50, 101):
        t0 = o.table_of[i]
        k0 = o.dish_of[1][t0]
        j0, j1 = cv["rowptr"][i], cv["rowptr"][i + 1]
        acc, loo, tot = oracle.stageA_counts_f32(cv["col"][j0:j1], cv["val"][j0:j1], l2t, cdt, t0, beta,
                                                 np.float32(W * beta + o.ctot[1][k0]))
        _, L = o.row_logweights(i, want_L=True)
        assert tot == cv["val"][j0:j1].sum()
        for t in range(cap):
            k = o.dish_of[1][t]
            if k < 0 or o.l_vk[1][k] == 0:
                continue
            want = L[1][t] / np.log(2.0)
            got = loo if k == k0 else acc[t]
            assert abs(got - want) <= 2e-5 * max(1.0, abs(want)), (i, t, got, want)
    ps = oracle.params_struct(P)
    rng = np.random.default_rng(0)
    acc = rng.normal(0, 1, (2, cap)).astype(np.float32)
    xx = np.abs(rng.normal(0, 1, 2)).astype(np.float32)
    a = oracle.stageB_f32(ps, acc, xx, int(o.table_of[0]), 0.37)
    b = oracle.stageB_f32_mixed(ps, np.zeros(2, np.int32), acc, xx, np.zeros(2, np.float32), int(o.table_of[0]), 0.37)
    assert a == b
