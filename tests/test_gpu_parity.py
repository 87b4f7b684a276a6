"""GPU parity tests (run with -m gpu on the B200 box).  Everything goes through the C ABI
(libmvg_b200.so via ctypes); the CPU oracle is the checker:

  * the FP32 parameter block equals the oracle's (computed independently in FP64 then rounded);
  * stage A (dot products) — CUDA-core engine: bit-exact against the fmaf-chain restatement;
    log-likelihoods within 1e-5 relative of the FP64 restatement (the tolerance north_star states);
  * stage B (draw): integer choices bit-exact against oracle/mv_oracle.c:mvo_stageB_f32 fed the
    device's own dot products and the same Philox uniforms;
  * births / deaths / counts (table_of, n_t, dish_of, n_vk, l_vk): bit-exact against mvo_reseat;
  * float statistics within 1e-5 relative; hyperparameters against mvo_hyper_step.
"""
import numpy as np
import pytest
from conftest import c1_data, make_mixture

pytestmark = pytest.mark.gpu
RTOL_LOGLIK = 1e-5     # north_star: "<= 1e-5 FP32"
RTOL_STATS = 1e-5


def _mk_sampler(views, cap, seed, engine=1, debug=True, **kw):
    import mvc_b200
    s = mvc_b200.Sampler(views[0].shape[0], [v.reshape(len(v), -1).shape[1] for v in views], cap=cap, seed=seed,
                         engine=engine, debug_export=debug, **kw)
    for v, x in enumerate(views):
        s.upload_view(v, x)
    return s


def _oracle_from_device(po, views, cap, seed, st, n_global=None, row_offset=0):
    o = po.OracleState(views, cap, seed=seed, row_offset=row_offset, n_global=n_global)
    o.alpha_v[:] = st["alpha_v"]; o.sigma_v[:] = st["sigma_v"]; o.tau_v[:] = st["tau_v"]
    o.alpha_g, o.sigma_g = st["alpha_g"], st["sigma_g"]
    o.sweep = st["sweep"]
    o.set_assignment(st["table_of"], st["dish_of"])
    # counts must agree exactly; then adopt the device's float statistics so that both sides start the
    # sweep from identical inputs (the device sums in FP32 tiles, the oracle in FP64: tolerance-level)
    np.testing.assert_array_equal(st["n_t"], o.n_t)
    np.testing.assert_array_equal(st["n_vk"], o.n_vk)
    np.testing.assert_array_equal(st["l_vk"], o.l_vk)
    for v in range(o.V):
        np.testing.assert_allclose(st["S1"][v], o.S1[v], rtol=RTOL_STATS, atol=1e-4)
        o.S1[v][:] = st["S1"][v]
    np.testing.assert_allclose(st["sum_y2"], o.S2, rtol=RTOL_STATS, atol=1e-4)
    o.S2[:] = st["sum_y2"]
    return o


def _check_params(po, o, P):
    Q = o.make_params()
    for k in ("dish", "lone", "single"):
        np.testing.assert_array_equal(P[k], Q[k], err_msg=k)
    for k in ("A", "C", "A1", "C1", "W", "W1", "AN", "CN", "WN", "LD", "LM", "LM1", "LMN"):
        a, b = P[k].astype(np.float64), Q[k].astype(np.float64)
        masked = b < -1e29
        np.testing.assert_array_equal(a[masked] < -1e29, True, err_msg=k)
        np.testing.assert_allclose(a[~masked], b[~masked], rtol=2e-5, atol=1e-6, err_msg=k)
    for v in range(o.V):
        np.testing.assert_allclose(P["m"][v], Q["m"][v], rtol=2e-5, atol=1e-6)


def _one_sweep_parity(po, s, views, cap, seed, do_hyper, simt_bit_exact=True, fast_weights=False, tc=False):
    """Compare ONE device sweep with the oracle started from the device's own pre-sweep state.
    tc: tensor-core engine — acc holds the dot products with the pre-scaled means b = 2 A m and the mirror is
    mvo_stageB_tc, fed the device's new-table log-weight as well (itself checked against FP64 below)."""
    pre = s.get_state()
    P = s.get_params()
    o = _oracle_from_device(po, views, cap, seed, pre)
    _check_params(po, o, P)
    s.sweep(1, do_hyper=do_hyper)
    acc, xx, raw = s.get_debug_rows()
    lnew_dev = s.get_debug_lnew() if tc else None
    n = o.n
    # ---- stage A
    ps = po.params_struct(P)
    L = po.lib()
    worst = 0.0
    flips = 0
    mm = [np.sum(P["m"][v].astype(np.float64) ** 2, axis=1) for v in range(o.V)]
    for i in range(n):
        if simt_bit_exact:
            for v in range(o.V):
                a, q = po.stageA_f32(o.views[v][i], P["m"][v])
                assert np.array_equal(a, acc[i, v]) and q == xx[i, v], (i, v)
        u = L.mvo_uf(seed, 0, 0, 0, pre["sweep"], i)
        if tc:
            ch, lw32 = po.stageB_tc(ps, acc[i], xx[i], pre["table_of"][i], u, lnew_dev[i], want_lw=True)
        else:
            ch, lw32 = po.stageB_f32(ps, acc[i], xx[i], pre["table_of"][i], u, want_lw=True)
        if not fast_weights:
            assert ch == raw[i], (i, ch, raw[i])                # integer draw: bit-exact
        elif ch != raw[i]:
            # MUFU weights (<= 2 ulp each): a draw may differ from the mirrored one only when u*total
            # fell within rounding distance of a CDF edge, and then only to the neighbouring option
            if tc:
                _, margin = po.stageB_tc(ps, acc[i], xx[i], pre["table_of"][i], u, lnew_dev[i], want_margin=True)
            else:
                _, margin = po.stageB_f32_margin(ps, acc[i], xx[i], pre["table_of"][i], u)
            assert margin < 1e-5, (i, ch, raw[i], margin)
            flips += 1
        if i % 7 == 0:
            lw64 = o.row_logweights(i) / np.log(2.0)
            ok = np.isfinite(lw64)
            assert np.all(lw32[~ok] < -1e29)
            # Relative tolerance of a cancelling sum: log f = C + A(2x.m - |x|^2) with C ~ -A|m|^2, so the
            # error is measured against the larger of the result and the terms that cancel, A(|x|^2+|m|^2).
            scale = np.zeros(cap + 1)
            for v in range(o.V):
                scale[:cap] += np.maximum(P["A"][v], P["A1"][v]) * (xx[i, v] + mm[v])
                scale[cap] += P["AN"][v] * xx[i, v]
            denom = np.maximum(np.maximum(1.0, np.abs(lw64[ok])), scale[ok])
            worst = max(worst, float(np.max(np.abs(lw32[ok] - lw64[ok]) / denom)))
    assert worst < RTOL_LOGLIK, worst
    assert flips <= max(2, n // 2000), flips
    # FP64 restatement draws agree except at CDF edges
    agree = float((o.draw_rows(threads=4) == raw).mean())
    assert agree > 0.99, agree
    # ---- births / deaths / counts from the device's raw draws
    ns, rows, w = o.reseat(raw, want_births=True)
    post = s.get_state()
    np.testing.assert_array_equal(post["table_of"], o.table_of)
    np.testing.assert_array_equal(post["n_t"], o.n_t)
    np.testing.assert_array_equal(post["dish_of"], o.dish_of)
    np.testing.assert_array_equal(post["n_vk"], o.n_vk)
    np.testing.assert_array_equal(post["l_vk"], o.l_vk)
    dns, drows, dw = s.get_debug_births()
    assert dns == ns and list(drows) == list(rows)
    if ns:
        np.testing.assert_allclose(dw, w, rtol=1e-9, atol=1e-12)
    for v in range(o.V):
        np.testing.assert_allclose(post["S1"][v], o.S1[v], rtol=RTOL_STATS, atol=1e-4)
    np.testing.assert_allclose(post["sum_y2"], o.S2, rtol=RTOL_STATS, atol=1e-4)
    assert post["sweep"] == pre["sweep"] + 1
    # ---- hyper step: feed the oracle the device's statistics so both see identical inputs
    if do_hyper:
        for v in range(o.V):
            o.S1[v][:] = post["S1"][v]
        o.S2[:] = post["sum_y2"]
        o.hyper_step(use_lgamma=True)
        np.testing.assert_allclose(post["alpha_v"], o.alpha_v, rtol=1e-9)
        np.testing.assert_allclose(post["sigma_v"], o.sigma_v, rtol=1e-9)
        np.testing.assert_allclose(post["tau_v"], o.tau_v, rtol=1e-9)
        np.testing.assert_allclose([post["alpha_g"], post["sigma_g"]], [o.alpha_g, o.sigma_g], rtol=1e-9)
    else:
        np.testing.assert_array_equal(post["tau_v"], pre["tau_v"])
    return ns


def test_scalar_views_config1_parity(oracle):
    """Config 1 shape: two scalar views (D = 1), reference init, 12 sweeps with births and deaths."""
    views, _ = c1_data(500)
    s = _mk_sampler(views, 32, seed=1999)
    s.init_state_reference()
    st = s.get_state()
    o = oracle.OracleState(views, 32, seed=1999)
    o.init_reference()
    np.testing.assert_array_equal(st["table_of"], o.table_of)        # same Philox init
    np.testing.assert_array_equal(st["dish_of"], o.dish_of)
    np.testing.assert_allclose(st["tau_v"], o.tau_v, rtol=1e-5)
    births = 0
    for it in range(12):
        births += _one_sweep_parity(oracle, s, views, 32, 1999, do_hyper=True)
    assert births > 0                                                # the run exercised the birth path
    s.close()


@pytest.mark.parametrize("dims,cap,n", [([8, 8, 8], 32, 700), ([64, 64, 64], 64, 900), ([5, 12], 64, 333), ([3], 32, 65)])
def test_dense_views_parity(oracle, dims, cap, n):
    views, z = make_mixture(n, dims, 6, seed=21)
    s = _mk_sampler(views, cap, seed=77)
    rng = np.random.default_rng(1)
    tab = np.where(rng.random(n) < 0.15, rng.integers(0, 6, n), z).astype(np.int32)
    tab[:3] = [7, 8, 9]                                              # three customers alone at their tables
    dish = np.full((len(dims), cap), -1, np.int32)
    for t in range(10):
        dish[:, t] = rng.integers(0, 5, len(dims))
    V = len(dims)
    s.set_state(tab, dish, np.full(V, 1.0), np.full(V, 0.5), np.full(V, 0.9), 1.0, 0.6, sweep=5)
    for it in range(6):
        _one_sweep_parity(oracle, s, views, cap, 77, do_hyper=(it % 2 == 0))
    s.close()


def test_full_capacity_masks_new_table(oracle):
    """All cap slots occupied: no birth may happen (capacity rule) and nothing is lost."""
    n, cap = 640, 32
    views, z = make_mixture(n, [4, 4], 32, seed=3)
    s = _mk_sampler(views, cap, seed=5)
    tab = (np.arange(n) % cap).astype(np.int32)
    dish = np.tile(np.arange(cap, dtype=np.int32), (2, 1))
    s.set_state(tab, dish, [1.0, 1.0], [0.5, 0.5], [1.0, 1.0], 5.0, 0.9)
    for _ in range(4):
        _one_sweep_parity(oracle, s, views, cap, 5, do_hyper=False)
        _, _, raw = s.get_debug_rows()
        assert (raw >= 0).all()
    assert s.get_state()["n_t"].sum() == n
    s.close()


def test_error_codes_and_ordering():
    import mvc_b200
    s = mvc_b200.Sampler(64, [2, 2], cap=32)
    with pytest.raises(mvc_b200.MvgError) as e:
        s.sweep(1)
    assert e.value.code == -4                                        # MVG_ESTATE: no state yet
    s.upload_view(0, np.zeros((64, 2), np.float32))
    with pytest.raises(mvc_b200.MvgError) as e:
        s.init_state_reference()
    assert e.value.code == -4 and "view 1" in str(e.value)           # second view missing
    s.upload_view(1, np.zeros((64, 2), np.float32))
    with pytest.raises(mvc_b200.MvgError) as e:
        s.set_state(np.full(64, 40), np.zeros((2, 32)), [1, 1], [.5, .5], [1, 1], 1, .6)
    assert e.value.code == -1                                        # MVG_EINVAL: table outside [0,cap)
    s.close()


def test_run_gibbs_posterior_summaries_match_reference(oracle):
    """north_star's third bullet: posterior summaries of a long GPU chain (synchronous sweeps)
    against the sequential reference chain on config-1 data: ARI vs truth, kernel variance, cluster count."""
    from sklearn.metrics import adjusted_rand_score as ari
    import mvc_b200
    views, truth = c1_data(500)
    res = mvc_b200.run_gibbs([v.astype(np.float64) for v in views], M=1500, burn_in=1200, thin=10, cap=32, seed=1999,
                             engine=1)
    assert len(res["table_of"]) == 30 and len(res["alpha_v"]) == 2
    assert len(res["loglik"]) == 30 and np.all(np.isfinite(res["loglik"])) and np.all(np.asarray(res["loglik"]) < 0)   # saved_loglik is filled
    aris = np.array([[ari(truth[v], np.asarray(res["dish_of"][s][v])[res["table_of"][s]]) for v in range(2)]
                     for s in range(30)])
    tau = np.array([res["tau_v"][v].mean() for v in range(2)])
    assert aris[:, 0].mean() > 0.85 and aris[:, 1].mean() > 0.75, aris.mean(0)
    if oracle.have_ref():
        y = np.stack([v.astype(np.float64) for v in views])
        tr = oracle.ref_run_gibbs(y, 1500, 1200, 10, seed=1999)
        ref_ari = np.array([[ari(truth[v], t["dish_of"][v][t["table_of"]]) for v in range(2)] for t in tr])
        ref_tau = np.array([[t["tau_v"][v] for v in range(2)] for t in tr]).mean(0)
        assert np.all(np.abs(aris.mean(0) - ref_ari.mean(0)) < 0.12), (aris.mean(0), ref_ari.mean(0))
        assert np.all(np.abs(tau / ref_tau - 1.0) < 0.25), (tau, ref_tau)


def _chain_summaries(table_of, dish_of, hyp_tau, truth, n):
    """Posterior summaries of a saved trace: per-view number of dishes, number of tables, ARI vs truth, pooled
    co-clustering matrix per view (posterior similarity), mean tau."""
    from sklearn.metrics import adjusted_rand_score as ari
    S, V = len(table_of), len(truth)
    n_dishes, n_tables, aris = np.zeros((S, V)), np.zeros(S), np.zeros((S, V))
    cocl = [np.zeros((n, n)) for _ in range(V)]
    for s in range(S):
        tab = np.asarray(table_of[s])
        n_tables[s] = len(np.unique(tab))
        for v in range(V):
            lab = np.asarray(dish_of[s][v])[tab]
            n_dishes[s, v] = len(np.unique(lab))
            aris[s, v] = ari(truth[v], lab)
            cocl[v] += lab[:, None] == lab[None, :]
    return {"n_dishes": n_dishes.mean(0), "n_tables": n_tables.mean(), "ari": aris.mean(0),
            "cocl": [c / S for c in cocl], "tau": np.array([np.mean(t) for t in hyp_tau])}


def test_long_chain_statistics_against_the_sequential_reference(oracle):
    """north_star's third correctness bullet, stated: over 4 seeds of config 1 (N = 500, two scalar views, 2000 sweeps,
    the last 500 kept every 10th) the GPU chain and the UNMODIFIED sequential reference (oracle/_ref, run_gibbs_cpp) agree
    on the per-view number of clusters (dishes), the pooled co-clustering matrix, the ARI against the truth and the
    kernel variances.  The number of TABLES is where a synchronous sweep differs from the sequential one (several
    customers open tables in the same sweep; tables sharing a dish do not change the clustering): it is printed, and the
    blocked sweep (mvg_set_sweep_blocks: statistics refreshed between row blocks) must move it towards the reference."""
    import mvc_b200
    if not oracle.have_ref():
        pytest.skip("compiled reference not available")
    views, truth = c1_data(500)
    y = np.stack([v.astype(np.float64) for v in views])
    n, seeds = 500, (1999, 7, 42, 2024)
    M, burn, thin = 2000, 1500, 10

    def pooled(runs):
        out = {k: np.mean([r[k] for r in runs], axis=0) for k in ("n_dishes", "n_tables", "ari", "tau")}
        out["cocl"] = [np.mean([r["cocl"][v] for r in runs], axis=0) for v in range(2)]
        return out

    ref_runs, gpu_runs, blk_runs = [], [], []
    for sd in seeds:
        tr = oracle.ref_run_gibbs(y, M, burn, thin, seed=sd)
        ref_runs.append(_chain_summaries([t["table_of"] for t in tr], [t["dish_of"] for t in tr],
                                         [[t["tau_v"][v] for t in tr] for v in range(2)], truth, n))
        for blocks, runs in ((1, gpu_runs), (32, blk_runs)):
            res = mvc_b200.run_gibbs([v.astype(np.float64) for v in views], M=M, burn_in=burn, thin=thin, cap=64, seed=sd,
                                     engine=1, blocks=blocks)
            runs.append(_chain_summaries(res["table_of"], res["dish_of"], res["tau_v"], truth, n))
    R, G, B = pooled(ref_runs), pooled(gpu_runs), pooled(blk_runs)
    print("\n[statistical parity, 4 seeds] dishes/view  ref %s  gpu %s  gpu-blocked32 %s" % (R["n_dishes"], G["n_dishes"], B["n_dishes"]))
    print("[statistical parity] tables  ref %.1f  gpu %.1f  gpu-blocked32 %.1f" % (R["n_tables"], G["n_tables"], B["n_tables"]))
    print("[statistical parity] ARI  ref %s  gpu %s  blocked %s;  tau  ref %s  gpu %s  blocked %s" % (R["ari"], G["ari"], B["ari"], R["tau"], G["tau"], B["tau"]))
    # The blocked sweep (32 passes per sweep) must sit on the reference.  The plain synchronous sweep is a visibly coarser
    # sampler at N = 500 — measured in round 2 over these 4 seeds: 4.2 / 11.5 dishes per view against the reference's
    # 2.9 / 4.1, ARI 0.82 against 0.93, mean |dP| 0.067 — because many customers re-seat against the same stale
    # statistics; it only has to stay a sensible clustering here (at N = 1M the fraction of customers whose neighbours
    # move in the same sweep is tiny).  DESIGN.md §2 states this and recommends blocks >= 16 for small N.
    for name, X in (("gpu", G), ("blocked", B)):
        d = [float(np.abs(X["cocl"][v] - R["cocl"][v]).mean()) for v in range(2)]
        print("[statistical parity] mean |dP| of the pooled co-clustering matrices, %s vs reference: %s" % (name, d))
        if name == "gpu":
            assert max(d) < 0.12 and np.all(X["ari"] > 0.7), (d, X["ari"])
            continue
        assert max(d) < 0.05, (name, d)                                   # posterior similarity matrices agree
        assert np.all(np.abs(X["ari"] - R["ari"]) < 0.05), (name, X["ari"], R["ari"])
        assert np.all(np.abs(X["n_dishes"] - R["n_dishes"]) < 1.5), (name, X["n_dishes"], R["n_dishes"])   # clusters per view
        assert np.all(np.abs(X["tau"] / R["tau"] - 1.0) < 0.15), (name, X["tau"], R["tau"])
    # the table count is the synchronous sweep's known difference; blocking must not make it worse and should reduce it
    assert abs(B["n_tables"] - R["n_tables"]) <= abs(G["n_tables"] - R["n_tables"]) + 0.5, (R["n_tables"], G["n_tables"], B["n_tables"])


# ---------------------------------------------------------------------------------------------
# tcgen05 engine (MVG_ENGINE_TCGEN05): stage A runs on the tensor cores (3-pass TF32 split), so the
# dot products are tolerance-level; everything downstream of them is bit-exact against the mirror.
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("n,k_true,engine,n_views,own_dishes", [
    (900, 6, 2, 3, False), (128 * 148 + 77, 40, 2, 3, False), (128 * 148 * 2 + 5, 40, 3, 3, False), (700, 5, 3, 3, False),
    (128 * 5 + 9, 6, 2, 1, False), (128 * 149 + 3, 12, 2, 2, False),
    # every table its own dish (the kernel's fast path: the leave-one-out correction touches the own table only),
    # with free slots (new-table marginal evaluated) and at full capacity (skipped)
    (128 * 148 + 77, 40, 2, 3, True), (128 * 20 + 5, 64, 2, 3, True), (128 * 7 + 1, 60, 3, 2, True)])
def test_tcgen05_engine_parity(oracle, n, k_true, engine, n_views, own_dishes):
    dims, cap = [64] * n_views, 64
    views, z = make_mixture(n, dims, k_true, seed=11)
    s = _mk_sampler(views, cap, seed=123, engine=engine)
    rng = np.random.default_rng(2)
    tab = np.where(rng.random(n) < 0.15, rng.integers(0, k_true, n), z).astype(np.int32)
    V = n_views
    dish = np.full((V, cap), -1, np.int32)
    if own_dishes:
        tab[:k_true] = np.arange(k_true)                             # every table occupied
        dish[:, :k_true] = np.arange(k_true)
    else:
        tab[:3] = [k_true + 1, k_true + 2, k_true + 3]               # three customers alone at their tables
        for t in range(k_true + 4):
            dish[:, t] = rng.integers(0, max(2, k_true - 1), 3)[:V]
    s.set_state(tab, dish, np.full(V, 1.0), np.full(V, 0.5), np.full(V, 0.9), 1.0, 0.6, sweep=3)
    for it in range(3):
        P = s.get_params()
        _one_sweep_parity(oracle, s, views, cap, 123, do_hyper=(it % 2 == 0), simt_bit_exact=False,
                          fast_weights=(engine == 3), tc=True)
        acc, xx, _ = s.get_debug_rows()
        for v in range(V):                                           # stage A against FP64: x . b with b = float32(2 A m)
            x64, b64 = views[v].astype(np.float64), oracle.scaled_means(P["A"][v], P["m"][v]).astype(np.float64)
            ref = x64 @ b64.T
            bound = np.abs(x64) @ np.abs(b64).T
            assert np.max(np.abs(acc[:, v, :] - ref) / (bound + 1e-30)) < 2.0 ** -18     # measured 1.9e-6: the tensor core truncates when it aligns addends
            np.testing.assert_allclose(xx[:, v], (x64 * x64).sum(1), rtol=1e-6)
    s.close()


def test_incremental_statistics_parity_and_agreement_with_rebuild(oracle):
    """MVG_STATS_INCREMENTAL: only moved rows are re-read (added to their new table, subtracted from their old one) into
    running FP64 sums.  (1) Every sweep still passes the one-sweep parity check against the oracle (counts exact, sums
    <= 1e-5, births seated identically); (2) a chain run incrementally stays with the chain that rebuilds every sweep:
    identical counts, sums to 1e-9, assignments identical but for draws on a CDF edge."""
    n, k_true, cap, V = 128 * 148 + 77, 40, 64, 3
    views, z = make_mixture(n, [64] * V, k_true, seed=21, spread=1.2)          # overlapping clusters: rows keep moving
    rng = np.random.default_rng(5)
    tab = np.where(rng.random(n) < 0.3, rng.integers(0, k_true, n), z).astype(np.int32)
    dish = np.full((V, cap), -1, np.int32)
    dish[:, :k_true] = np.arange(k_true)
    hyp = (np.full(V, 1.0), np.full(V, 0.5), np.full(V, 0.9), 1.0, 0.6)
    a = _mk_sampler(views, cap, seed=9, engine=2)
    b = _mk_sampler(views, cap, seed=9, engine=2)
    b.set_stats_mode(True, rebuild_every=1000)
    a.set_state(tab, dish, *hyp)
    b.set_state(tab, dish, *hyp)
    moved_total = 0
    for it in range(6):
        if it in (1, 4):
            _one_sweep_parity(oracle, b, views, cap, 9, do_hyper=True, simt_bit_exact=False, tc=True)
            a.sweep(1, do_hyper=True)
        else:
            a.sweep(1, do_hyper=True)
            b.sweep(1, do_hyper=True)
        sa, sb = a.get_state(), b.get_state()
        np.testing.assert_array_equal(sa["n_t"].sum(), n)
        agree = float((sa["table_of"] == sb["table_of"]).mean())
        assert agree > 0.9995, (it, agree)
        if agree == 1.0:
            np.testing.assert_array_equal(sa["n_t"], sb["n_t"])
            np.testing.assert_array_equal(sa["n_vk"], sb["n_vk"])
            for v in range(V):
                np.testing.assert_allclose(sb["S1"][v], sa["S1"][v], rtol=1e-6, atol=1e-3)
            np.testing.assert_allclose(sb["sum_y2"], sa["sum_y2"], rtol=1e-6)
        moved_total += int((sb["table_of"] != tab).sum())
        tab = sb["table_of"]
    assert moved_total > n // 20, moved_total            # the test did exercise moves
    # and the incremental sums equal a from-scratch FP64 rebuild of the final assignment
    o = oracle.OracleState(views, cap, seed=9)
    fin = b.get_state()
    o.set_assignment(fin["table_of"], fin["dish_of"])
    np.testing.assert_array_equal(fin["n_t"], o.n_t)
    for v in range(V):
        np.testing.assert_allclose(fin["S1"][v], o.S1[v], rtol=RTOL_STATS, atol=1e-4)
    a.close(); b.close()


@pytest.mark.parametrize("misplaced", [0.01, 0.6])
def test_incremental_statistics_row_and_tile_mode(oracle, misplaced):
    """The DELTA statistics kernel picks, per CTA, between fetching the 64-row tiles that hold a moved row (many rows
    moved) and fetching the moved rows alone (few): 1 % and 60 % misplaced customers move back in the first sweep; the
    running sums must equal a from-scratch FP64 rebuild of the final assignment, counts exactly."""
    n, k_true, cap, V = 64 * 148 * 2 + 41, 48, 64, 3
    views, z = make_mixture(n, [64] * V, k_true, seed=33)
    rng = np.random.default_rng(8)
    tab = np.where(rng.random(n) < misplaced, rng.integers(0, k_true, n), z).astype(np.int32)
    dish = np.full((V, cap), -1, np.int32)
    dish[:, :k_true] = np.arange(k_true)
    s = _mk_sampler(views, cap, seed=4, engine=2)
    s.set_stats_mode(True, rebuild_every=1000)
    s.set_state(tab, dish, np.full(V, 1.0), np.full(V, 0.5), np.full(V, 1.0), 1.0, 0.6)
    s.sweep(2, do_hyper=False)
    fin = s.get_state()
    moved = int((fin["table_of"] != tab).sum())
    assert moved > 0.5 * misplaced * n * (1 - 1 / k_true), moved
    o = oracle.OracleState(views, cap, seed=4)
    o.set_assignment(fin["table_of"], fin["dish_of"])
    np.testing.assert_array_equal(fin["n_t"], o.n_t)
    np.testing.assert_array_equal(fin["n_vk"], o.n_vk)
    for v in range(V):
        np.testing.assert_allclose(fin["S1"][v], o.S1[v], rtol=RTOL_STATS, atol=1e-3)
    s.close()


def test_tile_statistics_match_general_kernel(oracle, monkeypatch):
    """The register-accumulating statistics kernel of the C3 shape (mv_stats_tile.cu) against the general
    kernel (MVG_STATS_GENERIC=1) and the FP64 oracle: counts and assignments bit-exact, sums within 1e-5.
    Sizes: a partial last tile, fewer tiles than CTAs, and many tiles per CTA."""
    dims, cap = [64, 64, 64], 64
    for n, k_true in [(77, 5), (64 * 148 * 3 + 19, 50)]:
        views, z = make_mixture(n, dims, k_true, seed=4)
        rng = np.random.default_rng(9)
        tab = np.where(rng.random(n) < 0.2, rng.integers(0, k_true, n), z).astype(np.int32)
        dish = np.full((3, cap), -1, np.int32)
        dish[:, :k_true] = np.arange(k_true)
        out = []
        for generic in ("1", "0"):
            monkeypatch.setenv("MVG_STATS_GENERIC", generic)
            s = _mk_sampler(views, cap, seed=31, engine=1, debug=False)
            s.set_state(tab, dish, np.full(3, 1.0), np.full(3, 0.5), np.full(3, 0.9), 1.0, 0.6)
            s.sweep(2, do_hyper=False)
            out.append(s.get_state())
            s.close()
        g, t = out
        for k in ("table_of", "n_t", "dish_of", "n_vk", "l_vk"):
            np.testing.assert_array_equal(g[k], t[k], err_msg=k)
        for v in range(3):
            np.testing.assert_allclose(t["S1"][v], g["S1"][v], rtol=RTOL_STATS, atol=1e-4)
        np.testing.assert_allclose(t["sum_y2"], g["sum_y2"], rtol=RTOL_STATS, atol=1e-4)
        o = oracle.OracleState(views, cap, seed=31)
        o.set_assignment(t["table_of"], t["dish_of"])
        np.testing.assert_array_equal(t["n_vk"], o.n_vk)
        for v in range(3):
            np.testing.assert_allclose(t["S1"][v], o.S1[v], rtol=RTOL_STATS, atol=1e-4)
        np.testing.assert_allclose(t["sum_y2"], o.S2, rtol=RTOL_STATS, atol=1e-4)


def test_posterior_summaries(oracle):
    """cluster labels, adjusted Rand index + contingency table, co-clustering counts and the
    joint log-likelihood, computed on the device, against numpy / scikit-learn / the FP64 formula."""
    from sklearn.metrics import adjusted_rand_score
    n, dims, cap = 777, [2, 3], 32
    views, z = make_mixture(n, dims, 5, seed=8)
    s = _mk_sampler(views, cap, seed=17, debug=False)
    s.init_state_reference()
    s.sweep(25, do_hyper=True)
    st = s.get_state()
    labels = s.cluster_labels()
    for v in range(2):
        np.testing.assert_array_equal(labels[v], st["dish_of"][v][st["table_of"]])
        ari, tab = s.adjusted_rand_index(v, z)
        assert abs(ari - adjusted_rand_score(z, labels[v])) < 1e-12
        ref = np.zeros_like(tab)
        np.add.at(ref, (labels[v], z), 1)
        np.testing.assert_array_equal(tab, ref)
    ari_t, _ = s.adjusted_rand_index(-1, z)
    assert abs(ari_t - adjusted_rand_score(z, st["table_of"])) < 1e-12
    # co-clustering over three kept states
    s.coclustering_begin(0)
    want = np.zeros((n, n), np.uint32)
    for _ in range(3):
        s.sweep(2, do_hyper=True)
        s.coclustering_accumulate()
        lab = s.cluster_labels()[0]
        want += (lab[:, None] == lab[None, :]).astype(np.uint32)
    got, ns = s.coclustering_get()
    assert ns == 3
    np.testing.assert_array_equal(got, want)
    # log-likelihood: the reference's log p(y_S) per live dish (multiview_utils.cpp:316-320), per coordinate
    st = s.get_state()
    tot, pv = s.log_likelihood()
    ref_tot = 0.0
    for v in range(2):
        tau, D = st["tau_v"][v], dims[v]
        lv = 0.0
        for k in range(cap):
            nk = int(st["n_vk"][v][k])
            if nk == 0:
                continue
            S1, S2 = st["S1"][v][k], st["sum_y2"][v][k]
            lv += (-0.5 * nk * D * np.log(2 * np.pi * tau) - 0.5 * D * np.log(tau * (tau + nk)) - 0.5 * S2 / tau
                   + 0.5 * float(S1 @ S1) / (tau * (tau + nk)))
        np.testing.assert_allclose(pv[v], lv, rtol=1e-10)
        ref_tot += lv
    np.testing.assert_allclose(tot, ref_tot, rtol=1e-10)
    s.close()


def test_state_dump_and_resume_is_bit_identical():
    """a chain restarted from a dumped state (mvg_get_state -> mvg_set_state on a NEW handle)
    continues exactly as the uninterrupted one: draws are addressed by (seed, sweep, row), statistics are rebuilt
    in a fixed order."""
    n, dims, cap = 1500, [64, 64, 64], 64
    views, z = make_mixture(n, dims, 7, seed=5)
    a = _mk_sampler(views, cap, seed=41, engine=0, debug=False)
    a.init_state_reference()
    a.sweep(6, do_hyper=True)
    dump = a.get_state()
    a.sweep(5, do_hyper=True)
    ref = a.get_state()
    a.close()
    b = _mk_sampler(views, cap, seed=41, engine=0, debug=False)
    b.set_state(dump["table_of"], dump["dish_of"], dump["alpha_v"], dump["sigma_v"], dump["tau_v"], dump["alpha_g"],
                dump["sigma_g"], sweep=dump["sweep"])
    b.sweep(5, do_hyper=True)
    got = b.get_state()
    b.close()
    for k in ("table_of", "n_t", "dish_of", "n_vk", "l_vk"):
        np.testing.assert_array_equal(got[k], ref[k], err_msg=k)
    for k in ("alpha_v", "sigma_v", "tau_v"):
        np.testing.assert_array_equal(got[k], ref[k], err_msg=k)
    assert got["alpha_g"] == ref["alpha_g"] and got["sigma_g"] == ref["sigma_g"] and got["sweep"] == ref["sweep"]


def test_binary_checkpoint_file_resumes_bit_identically(tmp_path):
    """mvg_save_checkpoint / mvg_load_checkpoint: the chain continued from the FILE equals the uninterrupted chain; a file
    of another shape or a damaged file is refused."""
    import mvc_b200
    n, dims, cap = 1300, [64, 64], 64
    views, z = make_mixture(n, dims, 6, seed=8)
    a = _mk_sampler(views, cap, seed=3, engine=0, debug=False)
    a.init_state_reference()
    a.sweep(5, do_hyper=True)
    ck = tmp_path / "chain.mvg"
    a.save_checkpoint(ck)
    a.sweep(4, do_hyper=True)
    ref = a.get_state()
    a.close()
    b = _mk_sampler(views, cap, seed=3, engine=0, debug=False)
    b.load_checkpoint(ck)
    b.sweep(4, do_hyper=True)
    got = b.get_state()
    for k in ("table_of", "n_t", "dish_of", "n_vk", "l_vk", "alpha_v", "sigma_v", "tau_v"):
        np.testing.assert_array_equal(got[k], ref[k], err_msg=k)
    assert got["sweep"] == ref["sweep"] == 9 and got["alpha_g"] == ref["alpha_g"]
    raw = ck.read_bytes()
    (tmp_path / "cut.mvg").write_bytes(raw[: len(raw) // 2])
    with pytest.raises(mvc_b200.MvgError):
        b.load_checkpoint(tmp_path / "cut.mvg")
    b.close()
    c = _mk_sampler([views[0][:700], views[1][:700]], cap, seed=3, engine=0, debug=False)
    with pytest.raises(mvc_b200.MvgError) as e:
        c.load_checkpoint(ck)
    assert "another shape" in str(e.value)
    c.close()


# ---------------------------------------------------------------------------------------------
# Sparse count views (CSR; SURVEY.md A.3 — no reference counterpart: pinned to the FP64 restatement only)
# ---------------------------------------------------------------------------------------------
def _mk_mixed_sampler(views, cap, seed):
    import mvc_b200
    n = len(views[0]["rowptr"]) - 1 if isinstance(views[0], dict) else len(views[0])
    dims = [0 if isinstance(v, dict) else np.asarray(v).reshape(n, -1).shape[1] for v in views]
    s = mvc_b200.Sampler(n, dims, cap=cap, seed=seed, engine=0, debug_export=True)
    for v, x in enumerate(views):
        if isinstance(x, dict):
            s.upload_view_csr(v, x["rowptr"], x["col"], x["val"], x["vocab"])
        else:
            s.upload_view(v, x)
    return s


@pytest.mark.parametrize("layout", ["dense+counts", "counts+counts", "c4-mix"])
def test_count_views_parity(oracle, layout):
    from conftest import make_count_view
    n, cap, k_true, seed = 600, 32, 5, 91
    if layout == "c4-mix":
        n, cap = 260, 64                                             # BASELINE configs[3]'s view mix: 4 dense (D = 64) + 4 CSR
    rng = np.random.default_rng(3)
    z = rng.integers(0, k_true, n)
    cv1 = make_count_view(n, 200, z, k_true, seed=1)
    if layout == "dense+counts":
        mu = rng.normal(0, 2, (k_true, 4))
        views = [(mu[z] + rng.normal(0, 1, (n, 4))).astype(np.float32), cv1]
    elif layout == "c4-mix":
        views = []
        for v in range(4):
            mu = rng.normal(0, 2, (k_true, 64))
            views.append((mu[z] + rng.normal(0, 1, (n, 64))).astype(np.float32))
            views.append(make_count_view(n, 300 + 50 * v, z, k_true, seed=10 + v, mean_len=20))
    else:
        views = [cv1, make_count_view(n, 64, (z + 1) % k_true, k_true, seed=2, mean_len=8)]
    V = len(views)
    kind = np.array([1 if isinstance(v, dict) else 0 for v in views], np.int32)
    s = _mk_mixed_sampler(views, cap, seed)
    s.init_state_reference()
    o = oracle.OracleState(views, cap, seed=seed)
    o.init_reference()
    st = s.get_state()
    np.testing.assert_array_equal(st["table_of"], o.table_of)
    np.testing.assert_array_equal(st["dish_of"], o.dish_of)
    L = oracle.lib()
    births = 0
    for it in range(3 if layout == "c4-mix" else 8):
        pre, P = s.get_state(), s.get_params()
        # the oracle restarts from the device's state: integer statistics must agree exactly
        o = oracle.OracleState(views, cap, seed=seed)
        o.alpha_v[:] = pre["alpha_v"]; o.sigma_v[:] = pre["sigma_v"]; o.tau_v[:] = pre["tau_v"]
        o.alpha_g, o.sigma_g, o.sweep = pre["alpha_g"], pre["sigma_g"], pre["sweep"]
        o.set_assignment(pre["table_of"], pre["dish_of"])
        for k in ("n_t", "n_vk", "l_vk"):
            np.testing.assert_array_equal(pre[k], getattr(o, k), err_msg=k)
        tabs = {}
        for v in range(V):
            if not kind[v]:
                np.testing.assert_allclose(pre["S1"][v], o.S1[v], rtol=RTOL_STATS, atol=1e-4)
                o.S1[v][:] = pre["S1"][v]
                continue
            l2t, cd, ct = s.get_count_tables(v)
            tabs[v] = (l2t, cd)
            W = views[v]["vocab"]
            for t in range(cap):                                     # dish counts as seen from every table slot: exact
                k = pre["dish_of"][v][t]
                want = o.cd[v][k] if k >= 0 else np.zeros(W, np.int64)
                np.testing.assert_array_equal(cd[:, t], want, err_msg=f"dish counts view {v} table {t}")
                if k >= 0:
                    ref = np.log2((o.count_beta + o.cd[v][k]) / (W * o.count_beta + o.ctot[v][k]))
                    np.testing.assert_allclose(l2t[:, t], ref, rtol=2e-6, atol=1e-6)
            np.testing.assert_array_equal(pre["sum_y2"][v], o.S2[v])  # token totals per dish: exact integers
        o.S2[:] = pre["sum_y2"]
        Q = o.make_params()
        for k in ("dish", "lone", "single"):
            np.testing.assert_array_equal(P[k], Q[k], err_msg=k)
        for k in ("A", "C", "W", "W1", "AN", "CN", "WN", "LD", "LM", "LM1", "LMN"):
            a, b = P[k].astype(np.float64), Q[k].astype(np.float64)
            masked = b < -1e29
            np.testing.assert_array_equal(a[masked] < -1e29, True, err_msg=k)
            np.testing.assert_allclose(a[~masked], b[~masked], rtol=2e-5, atol=1e-6, err_msg=k)
        s.sweep(1, do_hyper=True)
        acc, xx, raw = s.get_debug_rows()
        loo = s.get_debug_loo()
        ps = oracle.params_struct(P)
        worst = 0.0
        for i in range(n):
            t0 = pre["table_of"][i]
            for v in range(V):
                if kind[v]:                                          # stage A of a count row: bit-exact fmaf chains
                    cv = views[v]
                    j0, j1 = cv["rowptr"][i], cv["rowptr"][i + 1]
                    a, l, tot = oracle.stageA_counts_f32(cv["col"][j0:j1], cv["val"][j0:j1], tabs[v][0], tabs[v][1], t0,
                                                         o.count_beta, P["C1"][v][t0])
                    assert np.array_equal(a, acc[i, v]) and tot == xx[i, v], (i, v)
                    assert l == loo[i, v], (i, v, l, loo[i, v])
                else:
                    a, q = oracle.stageA_f32(views[v][i], P["m"][v])
                    assert np.array_equal(a, acc[i, v]) and q == xx[i, v], (i, v)
            u = L.mvo_uf(seed, 0, 0, 0, pre["sweep"], i)
            ch, lw32 = oracle.stageB_f32_mixed(ps, kind, acc[i], xx[i], loo[i], t0, u, want_lw=True)
            assert ch == raw[i], (i, ch, raw[i])                      # integer draw: bit-exact
            if i % 5 == 0:                                           # log-weights against the FP64 restatement
                lw64 = o.row_logweights(i) / np.log(2.0)
                ok = np.isfinite(lw64)
                assert np.all(lw32[~ok] < -1e29)
                denom = np.maximum(1.0, np.abs(lw64[ok]))
                worst = max(worst, float(np.max(np.abs(lw32[ok] - lw64[ok]) / denom)))
        assert worst < 2e-5, worst                                   # count rows: sums of ~30 FP32 log2 terms
        ns, rows, w = o.reseat(raw, want_births=True)
        births += ns
        post = s.get_state()
        for k in ("table_of", "n_t", "dish_of", "n_vk", "l_vk"):
            np.testing.assert_array_equal(post[k], getattr(o, k), err_msg=k)
        dns, drows, dw = s.get_debug_births()
        assert dns == ns and list(drows) == list(rows)
        if ns:
            np.testing.assert_allclose(dw, w, rtol=1e-9, atol=1e-12)
        for v in range(V):
            if not kind[v]:
                o.S1[v][:] = post["S1"][v]
        o.S2[:] = post["sum_y2"]
        o.hyper_step(use_lgamma=True)
        np.testing.assert_allclose(post["alpha_v"], o.alpha_v, rtol=1e-9)
        np.testing.assert_allclose(post["sigma_v"], o.sigma_v, rtol=1e-9)
        np.testing.assert_allclose(post["tau_v"], o.tau_v, rtol=1e-9)
        np.testing.assert_allclose([post["alpha_g"], post["sigma_g"]], [o.alpha_g, o.sigma_g], rtol=1e-9)
    assert births > 0
    s.close()


def test_count_views_recover_planted_clusters():
    """A chain on a dense + a count view finds the planted partition (ARI), counts stay consistent."""
    from conftest import make_count_view
    n, cap, k_true = 3000, 32, 4
    rng = np.random.default_rng(12)
    z = rng.integers(0, k_true, n)
    mu = rng.normal(0, 3, (k_true, 3))
    views = [(mu[z] + rng.normal(0, 1, (n, 3))).astype(np.float32), make_count_view(n, 300, z, k_true, seed=4)]
    s = _mk_mixed_sampler(views, cap, seed=7)
    s.init_state_reference()
    s.sweep(120, do_hyper=True)
    st = s.get_state()
    assert st["n_t"].sum() == n
    a_dense, _ = s.adjusted_rand_index(0, z)
    a_counts, _ = s.adjusted_rand_index(1, z)
    # the dense view separates the clusters; the count view only re-labels through the dishes new tables pick (the
    # reference never re-samples the dish of an existing table), so its partition is coarser
    assert a_dense > 0.8 and a_counts > 0.2, (a_dense, a_counts)
    l2t, cd, ct = s.get_count_tables(1)
    assert ct.sum() == int(views[1]["val"].sum())
    s.close()


def test_independent_chains_coclustering_agrees_with_cpu_restatement(oracle):
    """BASELINE configs[4] in miniature: independent chains (chain id in the Philox key) of a count + dense data set,
    several per GPU on their own streams; the pooled co-clustering (posterior similarity) matrix of the GPU chains
    against the one of CPU chains run by the FP64 restatement from the same starts."""
    from conftest import make_count_view
    import mvc_b200
    n, cap, k_true, n_chains, burn, keep = 300, 32, 4, 4, 40, 30
    rng = np.random.default_rng(21)
    z = rng.integers(0, k_true, n)
    mu = rng.normal(0, 2.5, (k_true, 2))
    views = [make_count_view(n, 120, z, k_true, seed=3, mean_len=25), (mu[z] + rng.normal(0, 1, (n, 2))).astype(np.float32)]
    starts = []
    for ch in range(n_chains):
        r = np.random.default_rng([5, ch])
        tab = r.integers(0, cap // 2, n).astype(np.int32)
        dish = np.full((2, cap), -1, np.int32)
        dish[:, :cap // 2] = np.arange(cap // 2)
        starts.append((tab, dish))
    # GPU chains, interleaved launches
    chains = []
    for ch in range(n_chains):
        s = mvc_b200.Sampler(n, [0, 2], cap=cap, seed=77, chain=ch, engine=0)
        s.upload_view_csr(0, views[0]["rowptr"], views[0]["col"], views[0]["val"], views[0]["vocab"])
        s.upload_view(1, views[1])
        s.set_state(starts[ch][0], starts[ch][1], [1.0, 1.0], [0.5, 0.5], [1.0, 0.5], 1.0, 0.6)
        s.coclustering_begin(1)                 # dishes of the dense view
        chains.append(s)
    for s in chains:
        s.sweep(burn, True)
    for _ in range(keep):
        for s in chains:
            s.sweep(1, True)
            s.coclustering_accumulate()
    P_gpu = np.zeros((n, n))
    ari_gpu = []
    for s in chains:
        cnt, ns = s.coclustering_get()
        assert ns == keep
        P_gpu += cnt / float(keep)
        ari_gpu.append(s.adjusted_rand_index(1, z)[0])
        s.close()
    P_gpu /= n_chains
    # CPU chains (FP64 restatement of the same synchronous sampler, same Philox keys)
    P_cpu = np.zeros((n, n))
    ari_cpu = []
    from sklearn.metrics import adjusted_rand_score
    for ch in range(n_chains):
        o = oracle.OracleState(views, cap, seed=77, chain=ch)
        o.tau_v[:] = [1.0, 0.5]
        o.set_assignment(*starts[ch])
        o.sweep_n(burn, threads=4, do_hyper=True)
        for _ in range(keep):
            o.sweep_n(1, threads=4, do_hyper=True)
            lab = o.dish_of[1][o.table_of]
            P_cpu += (lab[:, None] == lab[None, :]) / float(keep)
        ari_cpu.append(adjusted_rand_score(z, o.dish_of[1][o.table_of]))
    P_cpu /= n_chains
    assert np.abs(P_gpu - P_cpu).mean() < 0.05, np.abs(P_gpu - P_cpu).mean()
    assert abs(np.mean(ari_gpu) - np.mean(ari_cpu)) < 0.15, (ari_gpu, ari_cpu)
    assert np.mean(ari_gpu) > 0.6, ari_gpu


def test_long_chain_from_reference_init_tensor_core_engine():
    """A chain on the C3 shape from the reference's T = 4 start: hundreds of births and deaths through the tensor-core
    draw, the tile statistics kernel's birth resolution and the CUDA-graph replay; invariants and the planted partition."""
    n, dims, cap, k_true = 20000 + 37, [64, 64, 64], 64, 9
    views, z = make_mixture(n, dims, k_true, seed=13)
    s = _mk_sampler(views, cap, seed=3, engine=0, debug=False)
    s.init_state_reference()
    live = []
    for _ in range(12):
        s.sweep(10, do_hyper=True)
        st = s.get_state()
        assert st["n_t"].sum() == n and (st["n_vk"].sum(1) == n).all()
        assert ((st["n_t"] > 0) == (st["dish_of"][0] >= 0)).all()
        for v in range(3):
            np.testing.assert_array_equal(np.bincount(st["dish_of"][v][st["n_t"] > 0], minlength=cap), st["l_vk"][v])
        live.append(int((st["n_t"] > 0).sum()))
    assert max(live) > 4                                     # tables were born
    # What this start does NOT give, here or in the FP64 restatement (checked on the CPU: 3 dishes per view after 60
    # sweeps): the planted partition.  A new table picks its dish between the existing (mixed) dishes and a new one at
    # the prior predictive N(0, tau); in 64 dimensions both are hopeless and the existing one wins on weight, and the
    # dish of an existing table is never re-sampled (multiview_utils.cpp:224-289).  That is the reference's algorithm,
    # restated; chains on such data are started from an over-split state instead (bench.py C3/C2).
    st = s.get_state()
    assert 1 <= int((st["l_vk"][0] > 0).sum()) <= cap
    tot, _ = s.log_likelihood()
    assert np.isfinite(tot)
    s.close()


def test_over_split_start_merges_to_planted_partition_tensor_core_engine():
    """The same data from an over-split start (cap/2 random tables with their own dishes): the sampler merges — tables
    die until the planted clusters remain — through the tensor-core draw and the tile statistics kernel."""
    n, dims, cap, k_true = 20000 + 37, [64, 64, 64], 64, 9
    views, z = make_mixture(n, dims, k_true, seed=13)
    rng = np.random.default_rng(4)
    tab = rng.integers(0, cap // 2, n).astype(np.int32)
    dish = np.full((3, cap), -1, np.int32)
    dish[:, :cap // 2] = np.arange(cap // 2)
    s = _mk_sampler(views, cap, seed=3, engine=0, debug=False)
    s.set_state(tab, dish, [1.0] * 3, [0.5] * 3, [1.0] * 3, 1.0, 0.6)
    s.sweep(80, do_hyper=True)
    st = s.get_state()
    assert st["n_t"].sum() == n
    aris = [s.adjusted_rand_index(v, z)[0] for v in range(3)]
    live = int((st["n_t"] > 0).sum())
    assert min(aris) > 0.8 and k_true - 2 <= live <= k_true + 4, (aris, live)     # at most a pair of clusters fused
    s.close()


def test_run_gibbs_accepts_sparse_count_views():
    """run_gibbs with a scipy.sparse document-term matrix and a dense view, from an over-split start."""
    import scipy.sparse as sp
    import mvc_b200
    from conftest import make_count_view
    n, cap, k_true = 1500, 32, 4
    rng = np.random.default_rng(8)
    z = rng.integers(0, k_true, n)
    cv = make_count_view(n, 150, z, k_true, seed=6)
    X = sp.csr_matrix((cv["val"], cv["col"], cv["rowptr"]), shape=(n, cv["vocab"]))
    mu = rng.normal(0, 3, (k_true, 2))
    dense = mu[z] + rng.normal(0, 1, (n, 2))
    tab = rng.integers(0, cap // 2, n).astype(np.int32)
    dish = np.full((2, cap), -1, np.int32)
    dish[:, :cap // 2] = np.arange(cap // 2)
    res = mvc_b200.run_gibbs([X, dense], M=120, burn_in=100, thin=5, cap=cap, seed=5, start=(tab, dish))
    assert len(res["table_of"]) == 4 and len(res["tau_v"]) == 2
    from sklearn.metrics import adjusted_rand_score as ari
    last_tab, last_dish = res["table_of"][-1], res["dish_of"][-1]
    assert ari(z, np.asarray(last_dish[0])[last_tab]) > 0.7 and ari(z, np.asarray(last_dish[1])[last_tab]) > 0.7


def test_real_reuters_collection_on_the_gpu(oracle):
    """BASELINE configs[1]: the three count views ingested from the Reuters-21578 .sgm files (mvc_b200/reuters.py restating
    dataset/reuters/data pre-process.R:9-108; cached as CSR in data_cache/, which travels with the working tree) through
    a GPU chain: 25 sweeps from an over-split start, then against the CPU restatement started from the device's state —
    customer counts and per-dish word counts exact, and the next sweep's draws agree except on CDF edges."""
    import mvc_b200
    from mvc_b200 import reuters
    cached = reuters.load_cached()
    if cached is None:
        pytest.skip("data_cache/reuters21578_csr.npz not present (python -m mvc_b200.reuters writes it where the .sgm files are)")
    views, ids = cached
    n, cap, seed = len(ids), 64, 1999
    assert n == 21578 and [v["vocab"] for v in views][2] == 445
    rng = np.random.default_rng(seed)
    tab = rng.integers(0, cap // 2, n).astype(np.int32)
    dish = np.full((3, cap), -1, np.int32)
    dish[:, :cap // 2] = np.arange(cap // 2)
    s = mvc_b200.Sampler(n, [0, 0, 0], cap=cap, seed=seed, engine=1, debug_export=True)
    for v, x in enumerate(views):
        s.upload_view_csr(v, x["rowptr"], x["col"], x["val"], x["vocab"])
    s.set_state(tab, dish, [1.0] * 3, [0.5] * 3, [1.0] * 3, 1.0, 0.6)
    s.sweep(25, do_hyper=True)
    pre = s.get_state()
    assert int(pre["n_t"].sum()) == n
    o = oracle.OracleState(views, cap, seed=seed)
    o.alpha_v[:] = pre["alpha_v"]; o.sigma_v[:] = pre["sigma_v"]; o.tau_v[:] = pre["tau_v"]
    o.alpha_g, o.sigma_g, o.sweep = pre["alpha_g"], pre["sigma_g"], pre["sweep"]
    o.set_assignment(pre["table_of"], pre["dish_of"])
    for k in ("n_t", "n_vk", "l_vk"):
        np.testing.assert_array_equal(pre[k], getattr(o, k), err_msg=k)
    for v in range(3):                                               # word counts of every dish, as seen from its tables: exact
        l2t, cd, ct = s.get_count_tables(v)
        for t in np.nonzero(pre["n_t"] > 0)[0]:
            np.testing.assert_array_equal(cd[:, t], o.cd[v][pre["dish_of"][v][t]])
    np.testing.assert_array_equal(pre["sum_y2"], o.S2)               # token totals per dish
    s.sweep(1, do_hyper=True)
    _, _, raw = s.get_debug_rows()
    agree = float((o.draw_rows(threads=8) == raw).mean())
    assert agree > 0.995, agree
    live = int((s.get_state(with_rows=False)["n_t"] > 0).sum())
    print(f"\n[reuters] 26 sweeps on the real collection: {live} tables live, draw agreement with the FP64 restatement {agree:.5f}")
    s.close()


def test_blocked_sweeps_and_statistics_modes_are_deterministic_and_conserve_customers():
    """mvg_set_sweep_blocks / mvg_set_stats_mode: a chain is a function of (data, seed, start, blocks) — two handles agree
    exactly; blocks = 1 is the plain sweep; customers are conserved; bad arguments are rejected."""
    import mvc_b200
    views, _ = c1_data(400)
    def chain(blocks, sweeps=40):
        s = mvc_b200.Sampler(400, [1, 1], cap=32, seed=5, engine=1)
        for v, x in enumerate(views):
            s.upload_view(v, x.reshape(-1, 1))
        if blocks is not None:
            s.set_sweep_blocks(blocks)
        s.init_state_reference()
        s.sweep(sweeps, do_hyper=True)
        st = s.get_state()
        s.close()
        return st
    a, b, c, d = chain(8), chain(8), chain(1), chain(None)
    for k in ("table_of", "n_t", "dish_of", "n_vk", "tau_v", "alpha_v"):
        np.testing.assert_array_equal(a[k], b[k], err_msg=k)
        np.testing.assert_array_equal(c[k], d[k], err_msg=k)
    assert int(a["n_t"].sum()) == 400 and int(c["n_t"].sum()) == 400
    assert not np.array_equal(a["table_of"], c["table_of"])          # blocking changes the chain (it sees fresher statistics)
    s = mvc_b200.Sampler(64, [1], cap=32, seed=1, engine=1)
    for bad in (lambda: s.set_sweep_blocks(0), lambda: s.set_stats_mode(True, 0)):
        with pytest.raises(mvc_b200.MvgError) as e:
            bad()
        assert e.value.code == -1
    s.close()
