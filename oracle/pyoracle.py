"""oracle/pyoracle.py — TEST INFRASTRUCTURE, not product code.

ctypes access to the two CPU checkers built by oracle/Makefile:

  * ``libmvoracle.so`` — the plain-C restatement (mv_oracle.c): FP64 synchronous sweep,
    hyper step, and the FP32 mirror of the device epilogue.
  * ``libmvref.so``    — the UNMODIFIED reference sampler compiled from
    /root/reference/Multiview against the Rcpp stand-in (only buildable where the
    reference tree exists; the prebuilt file travels to the GPU box).

Only tests/, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU legs import this module.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from pathlib import Path

import numpy as np

ORACLE_DIR = Path(__file__).resolve().parent
REF_DIR = ORACLE_DIR / "_ref"
MVO_NEW = -1
MASKED = np.float32(-1.0e30)

_i32p = C.POINTER(C.c_int32)
_f32p = C.POINTER(C.c_float)
_f64p = C.POINTER(C.c_double)


def build(quiet: bool = True) -> None:
    """Run oracle/Makefile (compiles the restatement; the reference too when its sources exist)."""
    subprocess.run(["make", "-C", str(ORACLE_DIR)], check=True,
                   stdout=subprocess.DEVNULL if quiet else None)


def _ptr(a: np.ndarray, typ):
    return a.ctypes.data_as(typ)


class _MvoState(C.Structure):
    _fields_ = [
        ("n", C.c_int32), ("V", C.c_int32), ("cap", C.c_int32), ("reserved0", C.c_int32),
        ("row_offset", C.c_int64), ("n_global", C.c_int64),
        ("D", _i32p), ("x", C.POINTER(_f32p)),
        ("table_of", _i32p), ("n_t", _i32p), ("dish_of", _i32p), ("n_vk", _i32p), ("l_vk", _i32p),
        ("S1", C.POINTER(_f64p)), ("S2", _f64p),
        ("alpha_v", _f64p), ("sigma_v", _f64p), ("tau_v", _f64p),
        ("alpha_g", C.c_double), ("sigma_g", C.c_double),
        ("seed", C.c_uint64), ("chain", C.c_uint32), ("sweep", C.c_uint32),
        ("csr", C.c_void_p), ("count_beta", C.c_double),
    ]


class _MvoCsr(C.Structure):
    _fields_ = [
        ("rowptr", _i32p), ("col", _i32p), ("val", _f32p), ("vocab", C.c_int32), ("pad", C.c_int32),
        ("cd", C.POINTER(C.c_int64)), ("ctot", C.POINTER(C.c_int64)),
    ]


class _MvoParams(C.Structure):
    _fields_ = [
        ("V", C.c_int32), ("cap", C.c_int32),
        ("dish", _i32p), ("A", _f32p), ("C", _f32p), ("A1", _f32p), ("C1", _f32p),
        ("W", _f32p), ("W1", _f32p), ("lone", _i32p),
        ("AN", _f32p), ("CN", _f32p), ("WN", _f32p), ("LD", _f32p),
        ("LM", _f32p), ("LM1", _f32p), ("single", _i32p), ("LMN", _f32p),
    ]


_lib = None
_ref = None


def lib():
    global _lib
    if _lib is None:
        path = REF_DIR / "libmvoracle.so"
        if not path.exists():
            build()
        L = C.CDLL(str(path))
        L.mvo_uf.restype = C.c_float
        L.mvo_uf.argtypes = [C.c_uint64, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint64]
        L.mvo_u53.restype = C.c_double
        L.mvo_u53.argtypes = L.mvo_uf.argtypes
        L.mvo_z.restype = C.c_double
        L.mvo_z.argtypes = L.mvo_uf.argtypes
        L.mvo_philox_raw.argtypes = [C.POINTER(C.c_uint32)] * 3
        L.mvo_log_f_vk.restype = C.c_double
        L.mvo_log_f_vk.argtypes = [C.POINTER(_MvoState), C.c_int, C.c_int, _f32p, C.c_int]
        L.mvo_log_f_new.restype = C.c_double
        L.mvo_log_f_new.argtypes = [C.POINTER(_MvoState), C.c_int, _f32p]
        L.mvo_row_logweights.argtypes = [C.POINTER(_MvoState), C.c_int, _f64p, _f64p]
        L.mvo_draw_from_logweights.argtypes = [C.POINTER(_MvoState), C.c_int, _f64p, C.c_double]
        L.mvo_draw_rows.argtypes = [C.POINTER(_MvoState), _i32p, C.c_int]
        L.mvo_reseat.argtypes = [C.POINTER(_MvoState), _i32p, _i32p, _i32p, _f64p]
        L.mvo_hyper_step.argtypes = [C.POINTER(_MvoState), _f64p, _f64p, C.c_int]
        L.mvo_log_EPPF_view.restype = C.c_double
        L.mvo_log_EPPF_view.argtypes = [C.POINTER(_MvoState), C.c_int, C.c_double, C.c_double, C.c_int]
        L.mvo_log_EPPF_global.restype = C.c_double
        L.mvo_log_EPPF_global.argtypes = [C.POINTER(_MvoState), C.c_double, C.c_double, C.c_int]
        L.mvo_log_posterior_tau.restype = C.c_double
        L.mvo_log_posterior_tau.argtypes = [C.POINTER(_MvoState), C.c_int, C.c_double]
        L.mvo_sweep.argtypes = [C.POINTER(_MvoState), C.c_int, C.c_int, C.c_int]
        L.mvo_rebuild_stats.argtypes = [C.POINTER(_MvoState)]
        L.mvo_init_reference.argtypes = [C.POINTER(_MvoState)]
        L.mvo_exp2m.restype = C.c_float
        L.mvo_exp2m.argtypes = [C.c_float]
        L.mvo_log2m.restype = C.c_float
        L.mvo_log2m.argtypes = [C.c_float]
        L.mvo_stageA_f32.argtypes = [_f32p, C.c_int, _f32p, C.c_int, _f32p, _f32p]
        L.mvo_stageB_f32.argtypes = [C.POINTER(_MvoParams), _f32p, _f32p, C.c_int, C.c_float, _f32p]
        L.mvo_stageB_f32_ex.argtypes = [C.POINTER(_MvoParams), _f32p, _f32p, C.c_int, C.c_float, _f32p, _f32p]
        L.mvo_stageB_f32_mixed.argtypes = [C.POINTER(_MvoParams), _i32p, _f32p, _f32p, _f32p, C.c_int, C.c_float, _f32p, _f32p]
        L.mvo_stageB_tc.argtypes = [C.POINTER(_MvoParams), _f32p, _f32p, C.c_int, C.c_float, C.c_float, _f32p, _f32p]
        L.mvo_stageA_counts_f32.argtypes = [_i32p, _f32p, C.c_int, _f32p, _i32p, C.c_int, C.c_int, C.c_float, C.c_float,
                                            _f32p, _f32p, _f32p]
        L.mvo_make_params.argtypes = ([C.POINTER(_MvoState), _i32p] + [_f32p] * 6 + [_i32p] + [_f32p] * 6
                                      + [_i32p, _f32p, C.POINTER(_f32p)])
        _lib = L
    return _lib


def have_ref() -> bool:
    return (REF_DIR / "libmvref.so").exists()


def ref():
    """The compiled, unmodified reference (raises FileNotFoundError if it was never built)."""
    global _ref
    if _ref is None:
        path = REF_DIR / "libmvref.so"
        if not path.exists():
            if Path("/root/reference/Multiview/multiview_gibbs.cpp").exists():
                build()
            if not path.exists():
                raise FileNotFoundError(str(path))
        R = C.CDLL(str(path))
        dbl = C.c_double
        R.ref_last_error.restype = C.c_char_p
        R.ref_set_seed.argtypes = [C.c_uint64]
        R.ref_push_uniforms.argtypes = [_f64p, C.c_int]
        R.ref_uniform_calls.restype = C.c_uint64
        R.ref_normal_calls.restype = C.c_uint64
        R.ref_load_state.argtypes = [C.c_int, C.c_int, _f64p, _i32p, C.c_int, _i32p, _i32p, _f64p, _f64p, _f64p, dbl, dbl]
        R.ref_get_dims.argtypes = [_i32p, _i32p, _i32p]
        R.ref_get_tables.argtypes = [_i32p, _i32p]
        R.ref_get_dish_of.argtypes = [C.c_int, _i32p]
        R.ref_get_view_stats.argtypes = [C.c_int, _i32p, _i32p, _f64p, _f64p]
        R.ref_get_hypers.argtypes = [_f64p, _f64p, _f64p, _f64p]
        for name in ("ref_compute_f_vk",):
            getattr(R, name).restype = dbl
            getattr(R, name).argtypes = [C.c_int, C.c_int, C.c_int]
        for name in ("ref_compute_f_vk_new", "ref_marginal_new_table"):
            getattr(R, name).restype = dbl
            getattr(R, name).argtypes = [C.c_int, C.c_int]
        R.ref_table_probs.argtypes = [C.c_int, _f64p, _f64p]
        R.ref_log_EPPF.restype = dbl
        R.ref_log_EPPF.argtypes = [C.c_int, dbl, dbl]
        R.ref_log_prior_alpha.restype = dbl
        R.ref_log_prior_alpha.argtypes = [dbl]
        R.ref_log_prior_sigma.restype = dbl
        R.ref_log_prior_sigma.argtypes = [dbl]
        R.ref_log_posterior_given_tau.restype = dbl
        R.ref_log_posterior_given_tau.argtypes = [C.c_int, dbl]
        R.ref_run_gibbs.argtypes = [C.c_int, C.c_int, _f64p, C.c_int, C.c_int, C.c_int]
        R.ref_saved_table_of.argtypes = [C.c_int, _i32p]
        R.ref_saved_dish_of.argtypes = [C.c_int, C.c_int, _i32p]
        R.ref_saved_hypers.argtypes = [C.c_int, _f64p, _f64p, _f64p, _f64p]
        _ref = R
    return _ref


class OracleState:
    """Owns the numpy arrays behind one ``mvo_state``."""

    def __init__(self, views, cap, seed=1999, chain=0, row_offset=0, n_global=None, count_beta=0.5):
        """views: dense arrays [n, D], or CSR count views given as dicts
        {"rowptr": int32[n+1], "col": int32[nnz], "val": float32[nnz], "vocab": W}."""
        self.csr = [v if isinstance(v, dict) else None for v in views]
        n_of = lambda v: len(v["rowptr"]) - 1 if isinstance(v, dict) else len(v)
        self.n = int(n_of(views[0]))
        self.views = [np.zeros((self.n, 0), np.float32) if isinstance(v, dict)
                      else np.ascontiguousarray(v, dtype=np.float32).reshape(len(v), -1) for v in views]
        self.V = len(self.views)
        self.cap = int(cap)
        self.D = np.array([v.shape[1] for v in self.views], dtype=np.int32)
        self.table_of = np.zeros(self.n, dtype=np.int32)
        self.n_t = np.zeros(self.cap, dtype=np.int32)
        self.dish_of = np.full((self.V, self.cap), -1, dtype=np.int32)
        self.n_vk = np.zeros((self.V, self.cap), dtype=np.int32)
        self.l_vk = np.zeros((self.V, self.cap), dtype=np.int32)
        self.S1 = [np.zeros((self.cap, int(d)), dtype=np.float64) for d in self.D]
        self.S2 = np.zeros((self.V, self.cap), dtype=np.float64)
        self.alpha_v = np.ones(self.V, dtype=np.float64)
        self.sigma_v = np.full(self.V, 0.5, dtype=np.float64)
        self.tau_v = np.ones(self.V, dtype=np.float64)
        self._xptrs = (_f32p * self.V)(*[_ptr(v, _f32p) for v in self.views])
        self._s1ptrs = (_f64p * self.V)(*[_ptr(a, _f64p) for a in self.S1])
        s = _MvoState()
        s.n, s.V, s.cap = self.n, self.V, self.cap
        s.row_offset = row_offset
        s.n_global = self.n if n_global is None else n_global
        s.D = _ptr(self.D, _i32p)
        s.x = C.cast(self._xptrs, C.POINTER(_f32p))
        s.table_of = _ptr(self.table_of, _i32p)
        s.n_t = _ptr(self.n_t, _i32p)
        s.dish_of = _ptr(self.dish_of, _i32p)
        s.n_vk = _ptr(self.n_vk, _i32p)
        s.l_vk = _ptr(self.l_vk, _i32p)
        s.S1 = C.cast(self._s1ptrs, C.POINTER(_f64p))
        s.S2 = _ptr(self.S2, _f64p)
        s.alpha_v = _ptr(self.alpha_v, _f64p)
        s.sigma_v = _ptr(self.sigma_v, _f64p)
        s.tau_v = _ptr(self.tau_v, _f64p)
        s.alpha_g, s.sigma_g = 1.0, 0.6
        s.seed, s.chain, s.sweep = seed, chain, 0
        s.count_beta = float(count_beta)
        self.count_beta = float(count_beta)
        if any(c is not None for c in self.csr):
            self._csr_arr = (_MvoCsr * self.V)()
            self._csr_keep = []
            self.cd, self.ctot = [None] * self.V, [None] * self.V
            for v, cv in enumerate(self.csr):
                if cv is None:
                    continue
                rp = np.ascontiguousarray(cv["rowptr"], np.int32)
                col = np.ascontiguousarray(cv["col"], np.int32)
                val = np.ascontiguousarray(cv["val"], np.float32)
                W = int(cv["vocab"])
                self.cd[v] = np.zeros((self.cap, W), np.int64)
                self.ctot[v] = np.zeros(self.cap, np.int64)
                self.csr[v] = {"rowptr": rp, "col": col, "val": val, "vocab": W}
                e = self._csr_arr[v]
                e.rowptr, e.col, e.val, e.vocab = _ptr(rp, _i32p), _ptr(col, _i32p), _ptr(val, _f32p), W
                e.cd = self.cd[v].ctypes.data_as(C.POINTER(C.c_int64))
                e.ctot = self.ctot[v].ctypes.data_as(C.POINTER(C.c_int64))
            s.csr = C.cast(self._csr_arr, C.c_void_p)
        self.kind = np.array([0 if c is None else 1 for c in self.csr], np.int32)
        self.c = s

    # -- scalar fields live in the struct -------------------------------------------------
    alpha_g = property(lambda self: self.c.alpha_g, lambda self, v: setattr(self.c, "alpha_g", float(v)))
    sigma_g = property(lambda self: self.c.sigma_g, lambda self, v: setattr(self.c, "sigma_g", float(v)))
    sweep = property(lambda self: self.c.sweep, lambda self, v: setattr(self.c, "sweep", int(v)))
    seed = property(lambda self: self.c.seed, lambda self, v: setattr(self.c, "seed", int(v)))

    def ref(self):
        return C.byref(self.c)

    def set_assignment(self, table_of, dish_of):
        self.table_of[:] = np.asarray(table_of, dtype=np.int32)
        self.dish_of[:, :] = np.asarray(dish_of, dtype=np.int32)
        rc = lib().mvo_rebuild_stats(self.ref())
        if rc:
            raise ValueError(f"mvo_rebuild_stats rc={rc}")

    def init_reference(self):
        rc = lib().mvo_init_reference(self.ref())
        if rc:
            raise ValueError(f"mvo_init_reference rc={rc}")

    def row_logweights(self, i, want_L=False):
        lw = np.empty(self.cap + 1, dtype=np.float64)
        L = np.empty((self.V, self.cap + 1), dtype=np.float64) if want_L else None
        lib().mvo_row_logweights(self.ref(), int(i), _ptr(lw, _f64p), _ptr(L, _f64p) if want_L else None)
        return (lw, L) if want_L else lw

    def draw_rows(self, threads=1):
        choice = np.empty(self.n, dtype=np.int32)
        lib().mvo_draw_rows(self.ref(), _ptr(choice, _i32p), threads)
        return choice

    def reseat(self, choice, want_births=False):
        choice = np.ascontiguousarray(choice, dtype=np.int32)
        n_seated = C.c_int32(0)
        rows = np.full(self.cap, -1, dtype=np.int32)
        w = np.zeros((self.cap, self.V, self.cap + 1), dtype=np.float64)
        rc = lib().mvo_reseat(self.ref(), _ptr(choice, _i32p), C.byref(n_seated), _ptr(rows, _i32p), _ptr(w, _f64p))
        if rc:
            raise ValueError(f"mvo_reseat rc={rc}")
        ns = n_seated.value
        return (ns, rows[:ns].copy(), w[:ns].copy()) if want_births else ns

    def hyper_step(self, z=None, u=None, use_lgamma=True):
        zp = _ptr(np.ascontiguousarray(z, dtype=np.float64), _f64p) if z is not None else None
        up = _ptr(np.ascontiguousarray(u, dtype=np.float64), _f64p) if u is not None else None
        lib().mvo_hyper_step(self.ref(), zp, up, int(use_lgamma))

    def sweep_n(self, n_sweeps=1, threads=1, do_hyper=True):
        rc = lib().mvo_sweep(self.ref(), n_sweeps, threads, int(do_hyper))
        if rc:
            raise ValueError(f"mvo_sweep rc={rc}")

    def make_params(self):
        """FP32 parameter block + per-table means, as mvo_make_params defines them."""
        V, cap = self.V, self.cap
        P = {
            "dish": np.zeros((V, cap), np.int32), "A": np.zeros((V, cap), np.float32),
            "C": np.zeros((V, cap), np.float32), "A1": np.zeros((V, cap), np.float32),
            "C1": np.zeros((V, cap), np.float32), "W": np.zeros((V, cap), np.float32),
            "W1": np.zeros((V, cap), np.float32), "lone": np.zeros((V, cap), np.int32),
            "AN": np.zeros(V, np.float32), "CN": np.zeros(V, np.float32),
            "WN": np.zeros((V, 2), np.float32), "LD": np.zeros((V, 2), np.float32),
            "LM": np.zeros(cap, np.float32), "LM1": np.zeros(cap, np.float32),
            "single": np.zeros(cap, np.int32), "LMN": np.zeros(2, np.float32),
        }
        m = [np.zeros((cap, int(d)), np.float32) for d in self.D]
        mptrs = (_f32p * V)(*[_ptr(a, _f32p) for a in m])
        lib().mvo_make_params(
            self.ref(), _ptr(P["dish"], _i32p), _ptr(P["A"], _f32p), _ptr(P["C"], _f32p), _ptr(P["A1"], _f32p),
            _ptr(P["C1"], _f32p), _ptr(P["W"], _f32p), _ptr(P["W1"], _f32p), _ptr(P["lone"], _i32p),
            _ptr(P["AN"], _f32p), _ptr(P["CN"], _f32p), _ptr(P["WN"], _f32p), _ptr(P["LD"], _f32p),
            _ptr(P["LM"], _f32p), _ptr(P["LM1"], _f32p), _ptr(P["single"], _i32p), _ptr(P["LMN"], _f32p),
            C.cast(mptrs, C.POINTER(_f32p)))
        P["m"] = m
        return P

    def labels(self):
        """N x V matrix of dish labels, the clustering New_Simulation.R:135-149 extracts."""
        return np.stack([self.dish_of[v][self.table_of] for v in range(self.V)], axis=1)


def params_struct(P):
    """Wrap a dict of parameter arrays (from OracleState.make_params or the device export)."""
    keep = {k: np.ascontiguousarray(P[k]) for k in
            ("dish", "A", "C", "A1", "C1", "W", "W1", "lone", "AN", "CN", "WN", "LD", "LM", "LM1", "single", "LMN")}
    s = _MvoParams()
    s.V, s.cap = keep["dish"].shape
    for k, a in keep.items():
        setattr(s, k, _ptr(a, _i32p if a.dtype == np.int32 else _f32p))
    s._keep = keep
    return s


def stageA_f32(x_row, m):
    x_row = np.ascontiguousarray(x_row, np.float32)
    m = np.ascontiguousarray(m, np.float32)
    acc = np.empty(m.shape[0], np.float32)
    xx = C.c_float()
    lib().mvo_stageA_f32(_ptr(x_row, _f32p), x_row.shape[0], _ptr(m, _f32p), m.shape[0], _ptr(acc, _f32p), C.byref(xx))
    return acc, np.float32(xx.value)


def stageB_f32(pstruct, acc, xx, t0, uf, want_lw=False):
    acc = np.ascontiguousarray(acc, np.float32)
    xx = np.ascontiguousarray(xx, np.float32)
    lw = np.empty(pstruct.cap + 1, np.float32) if want_lw else None
    ch = lib().mvo_stageB_f32(C.byref(pstruct), _ptr(acc, _f32p), _ptr(xx, _f32p), int(t0), C.c_float(float(uf)),
                              _ptr(lw, _f32p) if want_lw else None)
    return (ch, lw) if want_lw else ch


def stageB_tc(pstruct, acc, xx, t0, uf, lnew_dev, want_lw=False, want_margin=False):
    """Draw stage of the tensor-core engine: acc = dot products with the pre-scaled means, lnew_dev = the device's
    log2 weight of a new table."""
    acc = np.ascontiguousarray(acc, np.float32)
    xx = np.ascontiguousarray(xx, np.float32)
    lw = np.empty(pstruct.cap + 1, np.float32) if want_lw else None
    mg = C.c_float()
    ch = lib().mvo_stageB_tc(C.byref(pstruct), _ptr(acc, _f32p), _ptr(xx, _f32p), int(t0), C.c_float(float(uf)),
                             C.c_float(float(lnew_dev)), _ptr(lw, _f32p) if want_lw else None, C.byref(mg))
    out = (ch,)
    if want_lw:
        out += (lw,)
    if want_margin:
        out += (mg.value,)
    return out if len(out) > 1 else ch


def scaled_means(A, m):
    """b[t] = float32(2 * A[t] * m[t]) for one view: the tensor-core engine's B operand."""
    return (2.0 * A.astype(np.float64)[:, None] * m.astype(np.float64)).astype(np.float32)


def stageB_f32_mixed(pstruct, kind, acc, xx, acc_loo, t0, uf, want_lw=False):
    """Stage B for a mix of dense and count views (kind[v] = 1: count view)."""
    kind = np.ascontiguousarray(kind, np.int32)
    acc = np.ascontiguousarray(acc, np.float32)
    xx = np.ascontiguousarray(xx, np.float32)
    acc_loo = np.ascontiguousarray(acc_loo, np.float32)
    lw = np.empty(pstruct.cap + 1, np.float32) if want_lw else None
    ch = lib().mvo_stageB_f32_mixed(C.byref(pstruct), _ptr(kind, _i32p), _ptr(acc, _f32p), _ptr(xx, _f32p),
                                    _ptr(acc_loo, _f32p), int(t0), C.c_float(float(uf)),
                                    _ptr(lw, _f32p) if want_lw else None, None)
    return (ch, lw) if want_lw else ch


def stageA_counts_f32(col, val, l2t, cdt, t0, beta, wbeta_plus_ctot):
    """Stage A of one CSR row against the device's tables l2t / cdt ([vocab, cap])."""
    col = np.ascontiguousarray(col, np.int32)
    val = np.ascontiguousarray(val, np.float32)
    l2t = np.ascontiguousarray(l2t, np.float32)
    cdt = np.ascontiguousarray(cdt, np.int32)
    cap = l2t.shape[1]
    acc = np.empty(cap, np.float32)
    loo, tot = C.c_float(), C.c_float()
    lib().mvo_stageA_counts_f32(_ptr(col, _i32p), _ptr(val, _f32p), len(col), _ptr(l2t, _f32p), _ptr(cdt, _i32p), cap,
                                int(t0), C.c_float(float(beta)), C.c_float(float(wbeta_plus_ctot)), _ptr(acc, _f32p),
                                C.byref(loo), C.byref(tot))
    return acc, np.float32(loo.value), np.float32(tot.value)


def stageB_f32_margin(pstruct, acc, xx, t0, uf):
    """Stage B plus the relative distance of the draw to the nearest CDF edge."""
    acc = np.ascontiguousarray(acc, np.float32)
    xx = np.ascontiguousarray(xx, np.float32)
    mg = C.c_float()
    ch = lib().mvo_stageB_f32_ex(C.byref(pstruct), _ptr(acc, _f32p), _ptr(xx, _f32p), int(t0), C.c_float(float(uf)),
                                 None, C.byref(mg))
    return ch, mg.value


def mirror_draw_rows(state: OracleState, P=None, acc=None, xx=None):
    """Full FP32 mirror over all rows.  With acc/xx (device stage-A export, [n][V][cap] and [n][V])
    only stage B is restated; without them stage A is restated too (CUDA-core engine)."""
    if P is None:
        P = state.make_params()
    ps = params_struct(P)
    out = np.empty(state.n, np.int32)
    L = lib()
    for i in range(state.n):
        if acc is None:
            a = np.empty((state.V, state.cap), np.float32)
            q = np.empty(state.V, np.float32)
            for v in range(state.V):
                a[v], q[v] = stageA_f32(state.views[v][i], P["m"][v])
        else:
            a, q = acc[i], xx[i]
        u = L.mvo_uf(state.c.seed, state.c.chain, 0, 0, state.c.sweep, state.c.row_offset + i)
        out[i] = stageB_f32(ps, a, q, state.table_of[i], u)
    return out


# ---------------------------------------------------------------------------------------------
# helpers for driving the compiled reference
# ---------------------------------------------------------------------------------------------
def ref_load(y, table_of, dish_of, K, alpha_v, sigma_v, tau_v, alpha_g, sigma_g):
    R = ref()
    y = np.ascontiguousarray(y, np.float64)
    d, n = y.shape
    table_of = np.ascontiguousarray(table_of, np.int32)
    dish_of = np.ascontiguousarray(dish_of, np.int32)
    T = dish_of.shape[1]
    K = np.ascontiguousarray(K, np.int32)
    rc = R.ref_load_state(n, d, _ptr(y, _f64p), _ptr(table_of, _i32p), T, _ptr(dish_of, _i32p), _ptr(K, _i32p),
                          _ptr(np.ascontiguousarray(alpha_v, np.float64), _f64p),
                          _ptr(np.ascontiguousarray(sigma_v, np.float64), _f64p),
                          _ptr(np.ascontiguousarray(tau_v, np.float64), _f64p), float(alpha_g), float(sigma_g))
    if rc:
        raise ValueError(R.ref_last_error().decode())


def ref_table_probs(i, T):
    R = ref()
    pe = np.zeros(T, np.float64)
    pn = C.c_double()
    if R.ref_table_probs(int(i), _ptr(pe, _f64p), C.byref(pn)):
        raise ValueError(R.ref_last_error().decode())
    return pe, pn.value


def ref_run_gibbs(y, M, burn_in, thin, seed=1999):
    """run_gibbs_cpp through the compiled reference; returns the saved states as numpy."""
    R = ref()
    y = np.ascontiguousarray(y, np.float64)
    d, n = y.shape
    R.ref_set_seed(seed)
    S = R.ref_run_gibbs(n, d, _ptr(y, _f64p), M, burn_in, thin)
    if S < 0:
        raise ValueError(R.ref_last_error().decode())
    out = []
    for s in range(S):
        T = R.ref_saved_T(s)
        tab = np.empty(n, np.int32)
        R.ref_saved_table_of(s, _ptr(tab, _i32p))
        dish = np.empty((d, T), np.int32)
        for v in range(d):
            row = np.empty(T, np.int32)
            R.ref_saved_dish_of(s, v, _ptr(row, _i32p))
            dish[v] = row
        a, sg, tau, g = np.empty(d), np.empty(d), np.empty(d), np.empty(2)
        R.ref_saved_hypers(s, _ptr(a, _f64p), _ptr(sg, _f64p), _ptr(tau, _f64p), _ptr(g, _f64p))
        out.append({"table_of": tab, "dish_of": dish, "alpha_v": a, "sigma_v": sg, "tau_v": tau,
                    "alpha_global": g[0], "sigma_global": g[1]})
    return out
