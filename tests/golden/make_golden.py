"""tests/golden/make_golden.py — regenerate the golden vectors from the UNMODIFIED reference.

Run in the build container only (needs /root/reference; oracle/Makefile compiles it into
oracle/_ref/libmvref.so):

    python tests/golden/make_golden.py

Writes tests/golden/reference_d1.json.  Every number in that file was produced by the reference's
own functions (compute_f_vk, compute_f_vk_new, compute_marginal_likelihood_new_table,
compute_table_probs_with_cache, remove_customer, sample_dish_for_new_table, log_EPPF,
log_prior_*, log_posterior_given_tau, update_hyperparameters) through oracle/refshim/ref_shim.cpp.
The tests compare the plain-C restatement (oracle/mv_oracle.c) with these vectors, so they also
run where the reference tree is absent (the GPU box).
"""
import ctypes as C
import json
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT / "oracle"))
import pyoracle as po  # noqa: E402


def state_case(rng, n, d, cap, slots, n_dishes, singleton_rows, hyp):
    """A random seating on `slots` (table slot ids < cap), dishes < n_dishes; some rows alone."""
    y = rng.normal(0, 3, (d, n)).astype(np.float32).astype(np.float64)   # exactly FP32-representable
    table = np.asarray(slots)[rng.integers(0, len(slots), n)]
    free = [t for t in range(cap) if t not in slots]
    for r, t in zip(singleton_rows, free):
        table[r] = t
    used = sorted(set(table.tolist()))
    dish = np.full((d, cap), -1, np.int64)
    for t in used:
        dish[:, t] = rng.integers(0, n_dishes, d)
    return {"n": n, "d": d, "cap": cap, "y": y, "table_of": table, "dish_of": dish, **hyp}


def reference_rows(case):
    """For every row: remove_customer(i) on a fresh load, then the reference's table weights."""
    R = po.ref()
    n, d, cap = case["n"], case["d"], case["cap"]
    table, dish = case["table_of"], case["dish_of"]
    alive = sorted(set(table.tolist()))
    comp = {t: c for c, t in enumerate(alive)}
    rows = []
    for i in range(n):
        po.ref_load(case["y"], [comp[t] for t in table], dish[:, alive], [cap] * d, case["alpha_v"],
                    case["sigma_v"], case["tau_v"], case["alpha_g"], case["sigma_g"])
        slots = list(alive)
        c0 = comp[table[i]]
        emptied = int(np.sum(table == table[i])) == 1
        assert R.ref_remove_customer(i) == 0
        if emptied:                      # the reference swap-deletes the emptied table
            slots[c0] = slots[-1]
            slots.pop()
        pe, pn = po.ref_table_probs(i, len(slots))
        w = np.zeros(cap + 1)
        w[slots] = pe
        w[cap] = pn
        f = np.zeros((d, cap))           # f_vk for every dish slot that is live after the removal
        fnew = np.zeros(d)
        marg = np.zeros(d)
        lv = np.zeros((d, cap), np.int32)
        for v in range(d):
            nv = np.zeros(cap, np.int32); l = np.zeros(cap, np.int32); s1 = np.zeros(cap); s2 = np.zeros(cap)
            R.ref_get_view_stats(v, nv.ctypes.data_as(po._i32p), l.ctypes.data_as(po._i32p),
                                 s1.ctypes.data_as(po._f64p), s2.ctypes.data_as(po._f64p))
            lv[v] = l
            for k in range(cap):
                if l[k] > 0:
                    f[v, k] = R.ref_compute_f_vk(v, k, i)
            fnew[v] = R.ref_compute_f_vk_new(v, i)
            marg[v] = R.ref_marginal_new_table(v, i)
        rows.append({"weights": w.tolist(), "f": f.tolist(), "f_new": fnew.tolist(), "marg": marg.tolist(),
                     "l_after": lv.tolist()})
    return rows


def reference_hyper(case, seed):
    """update_hyperparameters() on the loaded state; z/u are the call-ordered stream of the shim."""
    R = po.ref()
    n, d, cap = case["n"], case["d"], case["cap"]
    table, dish = case["table_of"], case["dish_of"]
    alive = sorted(set(table.tolist()))
    comp = {t: c for c, t in enumerate(alive)}
    po.ref_load(case["y"], [comp[t] for t in table], dish[:, alive], [cap] * d, case["alpha_v"],
                case["sigma_v"], case["tau_v"], case["alpha_g"], case["sigma_g"])
    out = {"log_EPPF": [R.ref_log_EPPF(v, float(case["alpha_v"][v]), float(case["sigma_v"][v])) for v in range(d)],
           "log_posterior_tau": [R.ref_log_posterior_given_tau(v, float(case["tau_v"][v])) for v in range(d)],
           "log_prior_alpha": [R.ref_log_prior_alpha(float(a)) for a in case["alpha_v"]],
           "log_prior_sigma": [R.ref_log_prior_sigma(float(s)) for s in case["sigma_v"]]}
    R.ref_set_seed(seed)
    R.ref_update_hyperparameters()
    L = po.lib()
    k = 3 * d + 2
    out["z"] = [L.mvo_z(seed, 0, 6, 1, 0, 2 * j) for j in range(k)]
    out["u"] = [L.mvo_u53(seed, 0, 6, 0, 0, 2 * j + 1) for j in range(k)]
    a, s, t, g = np.empty(d), np.empty(d), np.empty(d), np.empty(2)
    R.ref_get_hypers(a.ctypes.data_as(po._f64p), s.ctypes.data_as(po._f64p), t.ctypes.data_as(po._f64p),
                     g.ctypes.data_as(po._f64p))
    out["after"] = {"alpha_v": a.tolist(), "sigma_v": s.tolist(), "tau_v": t.tolist(),
                    "alpha_g": float(g[0]), "sigma_g": float(g[1])}
    assert R.ref_uniform_calls() == k and R.ref_normal_calls() == k
    return out


def reference_dish_draws(case, rows_to_try, seed):
    """sample_dish_for_new_table(v, i) after remove_customer(i), with a scripted uniform."""
    R = po.ref()
    n, d, cap = case["n"], case["d"], case["cap"]
    table, dish = case["table_of"], case["dish_of"]
    alive = sorted(set(table.tolist()))
    comp = {t: c for c, t in enumerate(alive)}
    L = po.lib()
    out = []
    for i in rows_to_try:
        po.ref_load(case["y"], [comp[t] for t in table], dish[:, alive], [cap] * d, case["alpha_v"],
                    case["sigma_v"], case["tau_v"], case["alpha_g"], case["sigma_g"])
        assert R.ref_remove_customer(i) == 0
        for v in range(d):
            u = L.mvo_u53(seed, 0, 1, v, 0, i)
            arr = np.array([u])
            R.ref_set_seed(seed)
            R.ref_push_uniforms(arr.ctypes.data_as(po._f64p), 1)
            k = R.ref_sample_dish_for_new_table(v, i)
            is_new = k >= cap                    # the reference appends a brand-new slot
            out.append({"row": int(i), "view": v, "u": u, "dish": (-1 if is_new else int(k))})
    return out


def tolist(case):
    return {k: (v.tolist() if isinstance(v, np.ndarray) else v) for k, v in case.items()}


def main():
    po.build()
    rng = np.random.default_rng(20261018)
    hyp2 = {"alpha_v": np.array([1.0, 0.8]), "sigma_v": np.array([0.5, 0.3]), "tau_v": np.array([0.7, 1.9]),
            "alpha_g": 1.0, "sigma_g": 0.6}
    hyp3 = {"alpha_v": np.array([0.4, 2.5, 1.1]), "sigma_v": np.array([0.2, 0.7, 0.45]),
            "tau_v": np.array([0.05, 3.0, 0.9]), "alpha_g": 2.2, "sigma_g": 0.35}
    # Appendix B of SURVEY.md
    appB = {"n": 6, "d": 2, "cap": 3,
            "y": np.array([[1.2, -0.4, 3.1, 2.9, -0.1, 0.3], [10, 9.5, -3, -2.5, 9.8, -3.2]]),
            "table_of": np.array([0, 0, 1, 1, 2, -1]), "dish_of": np.array([[0, 1, 0], [0, 1, 0]]), **hyp2}
    R = po.ref()
    po.ref_load(appB["y"], appB["table_of"], appB["dish_of"], [3, 3], appB["alpha_v"], appB["sigma_v"],
                appB["tau_v"], 1.0, 0.6)
    pe, pn = po.ref_table_probs(5, 3)
    appB_out = {
        "f": [[R.ref_compute_f_vk(v, k, 5) for k in range(3)] for v in range(2)],
        "f_new": [R.ref_compute_f_vk_new(v, 5) for v in range(2)],
        "prob_existing": pe.tolist(), "prob_new": pn,
        "log_EPPF": [R.ref_log_EPPF(0, 1.0, 0.5), R.ref_log_EPPF(1, 0.8, 0.3)],
        "log_prior_alpha_1": R.ref_log_prior_alpha(1.0), "log_prior_sigma_05": R.ref_log_prior_sigma(0.5),
        "log_posterior_tau": [R.ref_log_posterior_given_tau(0, 0.7), R.ref_log_posterior_given_tau(1, 1.9)],
    }
    cases = [
        state_case(rng, 40, 2, 16, [0, 2, 3, 7, 9, 12], 4, [5, 17], hyp2),
        state_case(rng, 25, 3, 8, [1, 2, 5], 3, [0, 9, 24], hyp3),
        state_case(rng, 30, 2, 32, list(range(0, 20, 2)), 6, [3], hyp2),
    ]
    out = {"generator": "tests/golden/make_golden.py", "source": "compiled unmodified reference (oracle/_ref/libmvref.so)",
           "appendix_B": {"state": tolist(appB), "expect": appB_out}, "cases": []}
    for ci, case in enumerate(cases):
        out["cases"].append({"state": tolist(case), "rows": reference_rows(case),
                             "hyper": reference_hyper(case, 100 + ci),
                             "dish_draws": reference_dish_draws(case, list(range(0, case["n"], 3)), 200 + ci)})
    path = Path(__file__).with_name("reference_d1.json")
    path.write_text(json.dumps(out))
    print("wrote", path, path.stat().st_size, "bytes")


if __name__ == "__main__":
    main()
