"""CPU-side checks of the drop-in boundary: the C-ABI library loads and exports every symbol that
include/zerocaf_b200.h declares; without a CUDA device the product fails loudly (no CPU fallback)."""
import ctypes
import os
import subprocess

import pytest

import dusk_zerocaf_b200 as zc
from dusk_zerocaf_b200 import _lib


@pytest.fixture(scope="module")
def so():
    if not os.path.exists(_lib.SO_PATH):
        _lib.build()
    return ctypes.CDLL(_lib.SO_PATH)


def test_header_declares_expected_surface():
    syms = _lib.header_symbols()
    for must in ("zc_ctx_create", "zc_fe_mul_batch", "zc_fe_mul_square_batch_dev", "zc_scalar_mul_batch",
                 "zc_point_add_batch", "zc_point_double_batch", "zc_point_scalar_mul_batch", "zc_msm",
                 "zc_msm_sharded_dev", "zc_ristretto_eq_batch"):
        assert must in syms
    assert len(syms) >= 50


def test_library_exports_every_declared_symbol(so):
    missing = [s for s in _lib.header_symbols() if not hasattr(so, s)]
    assert missing == []


def test_no_undeclared_public_symbols():
    out = subprocess.check_output(["nm", "-D", "--defined-only", _lib.SO_PATH], text=True)
    exported = {l.split()[-1] for l in out.splitlines() if " T " in l and l.split()[-1].startswith("zc_")}
    declared = set(_lib.header_symbols())
    # C++-mangled internals are not `zc_`-prefixed plain symbols; everything plain must be in the header
    assert exported - declared == set()


def test_product_does_not_link_or_import_the_oracle():
    out = subprocess.check_output(["ldd", _lib.SO_PATH], text=True)
    assert "oracle" not in out
    pkg = os.path.dirname(_lib.SO_PATH)
    for root, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                with open(os.path.join(root, f)) as fh:
                    text = fh.read()
                assert "zerocaf_oracle" not in text and "import oracle" not in text and "from oracle" not in text, f


def test_version_string(so):
    so.zc_version.restype = ctypes.c_char_p
    assert b"sm_100a" in so.zc_version()


def test_fails_loudly_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(zc.ZerocafError):
        zc.Context(0)


def test_scalar_from_bytes_range_check():
    # reference panics on > L-1 (scalar.rs:465); host mirror raises
    L = 2**249 + 14490550575682688738086195780655237219
    zc.Scalar.from_bytes((L - 1).to_bytes(32, "little"))
    with pytest.raises(ValueError):
        zc.Scalar.from_bytes(L.to_bytes(32, "little"))


def test_bytes_roundtrip_matches_oracle(oracle):
    import numpy as np
    rng = np.random.default_rng(7)
    for _ in range(64):
        b = bytearray(rng.integers(0, 256, 32, dtype=np.uint8).tobytes())
        b[31] &= 0x07
        fe = zc.FieldElement.from_bytes(b)
        assert np.array_equal(fe.limbs, oracle.fe_from_bytes(bytes(b)))
        assert fe.to_bytes() == oracle.fe_to_bytes(fe.limbs) == bytes(b)


def test_ctypes_binding_matches_header_prototypes(so):
    """Every int32_t-returning prototype in the header has ctypes argtypes of the same arity in _lib.py (ABI drift guard:
    a missing or mis-sized argtypes entry silently truncates 64-bit arguments)."""
    import re
    with open(os.path.join(os.path.dirname(_lib.SO_PATH), "..", "include", "zerocaf_b200.h")) as f:
        text = re.sub(r"/\*.*?\*/", "", f.read(), flags=re.S)
    protos = re.findall(r"int32_t\s+(zc_\w+)\s*\(([^)]*)\)\s*;", text)
    assert len(protos) >= 90
    L = _lib.lib()
    bad = []
    for name, args in protos:
        n_args = 0 if args.strip() in ("", "void") else len(args.split(","))
        at = getattr(getattr(L, name), "argtypes", None)
        if at is None or len(at) != n_args:
            bad.append((name, n_args, None if at is None else len(at)))
    assert bad == []


def test_header_is_plain_c(tmp_path):
    """The boundary header compiles as C99 (what a cgo / Rust bindgen / JNI consumer would feed it to)."""
    src = tmp_path / "t.c"
    src.write_text('#include "zerocaf_b200.h"\nint main(void) { return (int)sizeof(zc_ctx *) * 0; }\n')
    inc = os.path.join(os.path.dirname(_lib.SO_PATH), "..", "include")
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Werror", "-fsyntax-only", "-I", inc, str(src)])
