"""The C-ABI library loads without a GPU and exports every symbol include/mvg.h declares."""
import ctypes as C
import re
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parents[1]


def declared_symbols():
    text = (ROOT / "include" / "mvg.h").read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(mvg_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    import mvc_b200
    L = mvc_b200.lib()
    names = declared_symbols()
    assert len(names) >= 25
    for n in names:
        assert hasattr(L, n), f"{n} declared in include/mvg.h but not exported by libmvg_b200.so"
    assert L.mvg_abi_version() == 2


def test_no_cpu_fallback_without_gpu():
    """Without a CUDA device, creation must fail loudly (MVG_ECUDA), never fall back to the CPU."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    import mvc_b200
    with pytest.raises(mvc_b200.MvgError) as e:
        mvc_b200.Sampler(100, [1, 1], cap=32)
    assert e.value.code == -2 and "no CPU path" in str(e.value)


def test_product_never_imports_oracle():
    """The package must not reference oracle/ (the checker) anywhere."""
    pkg = ROOT / "multiview-clustering_b200"
    for path in list(pkg.rglob("*.py")) + list(pkg.rglob("*.cu")) + list(pkg.rglob("*.cuh")) + list(pkg.rglob("*.h")) + list(pkg.rglob("*.cpp")):
        if "build" in path.parts:
            continue
        text = path.read_text()
        assert "pyoracle" not in text and "mv_oracle.h" not in text.replace("oracle/mv_oracle.h:", "").replace("oracle/mv_oracle.c:", ""), path


def test_config_validation_is_host_side():
    import mvc_b200
    L = mvc_b200.lib()
    cfg = mvc_b200._Config()
    h = C.c_void_p()
    cfg.abi_version = 99
    assert L.mvg_create(C.byref(cfg), C.byref(h)) == -1
    cfg.abi_version = 2
    cfg.n_rows = 10
    cfg.n_rows_global = 10
    cfg.n_views = 1
    cfg.cap = 48
    cfg.world = 1
    assert L.mvg_create(C.byref(cfg), C.byref(h)) == -6       # MVG_EUNSUPPORTED: cap must be 32 or 64
    assert b"cap" in L.mvg_last_error(None)
