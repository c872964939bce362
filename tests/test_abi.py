"""CPU-side checks of the drop-in boundary: the C-ABI library builds, loads, and exports every symbol that
include/gl_commit.h declares; without a GPU it refuses to create a context (no CPU fallback)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_symbols():
    text = open(os.path.join(ROOT, "include", "gl_commit.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(gl_[a-z0-9_]+)\s*\(", text)))


def test_library_builds_and_exports_every_declared_symbol():
    import plonky25_b200 as g
    from plonky25_b200 import _lib
    path = g.build()
    assert os.path.exists(path)
    lib = ctypes.CDLL(path)
    declared = _header_symbols()
    assert len(declared) >= 25
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/gl_commit.h but not exported"
    # and the Python binding covers exactly the header
    assert sorted(_lib.SIGNATURES) == declared
    assert _lib.load().gl_abi_version() == 2


def test_sass_is_sm_100a_only():
    import plonky25_b200 as g
    import subprocess
    out = subprocess.run(["cuobjdump", "-lelf", g.build()], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_\d+a?", out))
    assert archs == {"sm_100a"}, archs


def test_no_cpu_fallback_without_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import plonky25_b200 as g
    with pytest.raises(g.GlError):
        g.Context(0)
    lib = g._lib.load()
    h = ctypes.c_void_p()
    assert lib.gl_ctx_create(ctypes.byref(h), 0) == g._lib.GL_ERR_CUDA
    assert lib.gl_strerror(g._lib.GL_ERR_INVALID) == b"invalid argument"
    # NULL context is rejected, not dereferenced
    assert lib.gl_tree_free(None, 1) == g._lib.GL_ERR_INVALID


def test_host_mirror_validates_like_upstream_asserts():
    import numpy as np
    import plonky25_b200 as g

    class _NoCtx:   # argument validation happens before any device work
        pass
    with pytest.raises(ValueError, match="inconsistent"):
        g.PolynomialBatch._cols([np.zeros(4, dtype=np.uint64), np.zeros(8, dtype=np.uint64)])
    with pytest.raises(ValueError, match="power of two"):
        g.PolynomialBatch._cols([np.zeros(6, dtype=np.uint64)])
    with pytest.raises(ValueError, match="empty"):
        g.PolynomialBatch._cols([])
    with pytest.raises(ValueError, match="blinding"):
        g.PolynomialBatch.from_values([np.zeros(4, dtype=np.uint64)], 1, True, 0, ctx=_NoCtx())


def test_product_never_imports_oracle():
    """The product path must not route through oracle/ (which is test infrastructure)."""
    pkg = os.path.join(ROOT, "plonky2.5_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                text = open(os.path.join(dirpath, f)).read()
                assert "gl_oracle" not in text and "libgl_oracle" not in text, f
                assert not re.search(r"^\s*(from|import)\s+oracle", text, flags=re.M), f


def test_rust_ffi_matches_the_header():
    """rust/gl-commit/src/ffi.rs is exactly what tools/gen_rust_ffi.py produces from include/gl_commit.h (one declaration per
    prototype), and the safe wrappers in src/lib.rs only call symbols that exist (no Rust toolchain here: a textual check)."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("gen_rust_ffi", os.path.join(ROOT, "tools", "gen_rust_ffi.py"))
    gen = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(gen)
    ffi = open(os.path.join(ROOT, "rust", "gl-commit", "src", "ffi.rs")).read()
    assert ffi == gen.generate(), "run `python tools/gen_rust_ffi.py` after changing include/gl_commit.h"
    declared = set(re.findall(r"pub fn (gl_[a-z0-9_]+)\(", ffi))
    assert sorted(declared) == _header_symbols()
    # argument counts agree with the ctypes table (an independent transcription of the same header)
    from plonky25_b200 import _lib
    for name, (_, argtypes) in _lib.SIGNATURES.items():
        m = re.search(r"pub fn %s\((.*?)\)" % name, ffi, flags=re.S)
        n_args = len([a for a in m.group(1).split(",") if a.strip()])
        assert n_args == len(argtypes), name
    lib_rs = open(os.path.join(ROOT, "rust", "gl-commit", "src", "lib.rs")).read()
    used = set(re.findall(r"ffi::(gl_[a-z0-9_]+)\(", lib_rs))
    assert used and used <= declared, used - declared
    consts = set(re.findall(r"ffi::(GL_[A-Z_0-9]+)", lib_rs))
    assert consts <= set(re.findall(r"pub const (GL_[A-Z_0-9]+)", ffi)), consts


def test_commit_multi_rejects_null_without_touching_a_device():
    from plonky25_b200 import _lib
    lib = _lib.load()
    assert lib.gl_commit_multi(None, 2, None, 4, 3, 1, 1, 0, None, None) == _lib.GL_ERR_INVALID


def test_reduction_arity_bits_mirrors_match_the_oracle():
    """FriReductionStrategy::ConstantArityBits in the Python mirror == the oracle's restatement ([4, 4, 4] for the wrapper)"""
    import plonky25_b200 as g
    import gl_oracle as o
    assert g.reduction_arity_bits(16, 3, 4) == [4, 4, 4]
    for degree_bits in range(0, 24):
        for rate_bits in (1, 3):
            for cap_height in (0, 4, 14):
                for ab, fp in ((4, 5), (3, 2), (1, 0)):
                    assert g.reduction_arity_bits(degree_bits, rate_bits, cap_height, ab, fp) == \
                        o.reduction_arity_bits_constant(ab, fp, degree_bits, rate_bits, cap_height)
