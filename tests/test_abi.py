"""CPU-only checks of the drop-in boundary: the C-ABI library loads and exports every symbol that
include/rxgauss.h declares, the ctypes table matches the header, and compute fails loudly
without a GPU (no CPU fallback)."""
import ctypes
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "rxgauss.h")


def header_functions():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(rxg_[a-z0-9_]+)\s*\(", src)))


def test_header_declares_expected_surface():
    fns = header_functions()
    for must in ("rxg_create", "rxg_lgssm_smooth_f32", "rxg_lgssm_filter_f32", "rxg_rule_mul_out_f32",
                 "rxg_rule_mul_in_f32", "rxg_prod_gaussian_f32", "rxg_hgf_filter_f32", "rxg_allgather_posteriors"):
        assert must in fns


def test_library_exports_every_declared_symbol(rx):
    lib = rx._lib.load()
    assert rx._lib.MISSING == [], f"the built library is older than the ctypes table: {rx._lib.MISSING}"
    for name in header_functions():
        assert hasattr(lib, name), f"{name} declared in rxgauss.h but not exported"


def test_ctypes_table_matches_header(rx):
    assert sorted(rx._lib.SIGNATURES) == header_functions()


def test_exports_are_plain_c(rx):
    out = subprocess.run(["nm", "-D", "--defined-only", rx._lib.LIB_PATH], capture_output=True, text=True).stdout
    syms = {l.split()[-1] for l in out.splitlines() if " T " in l}
    for name in header_functions():
        assert name in syms
    import torch  # noqa: F401  (the library itself must not depend on torch / python)
    ldd = subprocess.run(["ldd", rx._lib.LIB_PATH], capture_output=True, text=True).stdout
    assert "torch" not in ldd and "python" not in ldd


def test_version_and_supports(rx):
    lib = rx._lib.load()
    assert lib.rxg_version() == 100
    assert lib.rxg_supports(4, 4) == 1 and lib.rxg_supports(2, 2) == 1 and lib.rxg_supports(64, 64) == 1 and lib.rxg_supports(5, 5) == 1 and lib.rxg_supports(64, 32) == 1 \
        and lib.rxg_supports(65, 4) == 0 and lib.rxg_supports(4, 0) == 0


def test_fails_loudly_without_gpu(rx):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    lib = rx._lib.load()
    h = ctypes.c_void_p()
    assert lib.rxg_create(ctypes.byref(h), 0, 0) == rx._lib.RXG_ERR_NO_DEVICE
    assert not h
    with pytest.raises(rx.RxGaussError):
        rx.Context()
    # every compute entry refuses a null context instead of computing on the CPU
    null = ctypes.c_void_p(None)
    fpn = ctypes.cast(null, rx._lib.fp)
    rc = lib.rxg_lgssm_smooth_f32(null, 4, 4, 1, 1, fpn, fpn, fpn, fpn, fpn, fpn, fpn, fpn,
                                  ctypes.cast(null, rx._lib.u8p), fpn, fpn, fpn, ctypes.cast(null, rx._lib.i32p), 0)
    assert rc == rx._lib.RXG_ERR_BAD_ARG


def test_product_package_does_not_import_oracle():
    pkg = os.path.join(ROOT, "rxinfer.jl_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".jl")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in txt and "from oracle" not in txt and "oracle/" not in txt, f
