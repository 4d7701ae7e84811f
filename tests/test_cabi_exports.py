"""CPU (`-m "not gpu"`): the C-ABI library builds for sm_100a, loads, and exports every symbol include/pka_b200.h
declares (no compute calls here -- there is no GPU in the build container)."""
import ctypes
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "pka_b200.h")


def declared_symbols():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(pka_[a-z0-9_]+)\s*\(", text)))


@pytest.fixture(scope="module")
def library():
    import __graft_entry__
    __graft_entry__.build()
    from pytorch_kaldi_asr_b200 import _lib
    return _lib


def test_header_declares_the_expected_surface():
    syms = declared_symbols()
    for must in ("pka_frontend_fwd", "pka_gemm_f32", "pka_gemm_tc", "pka_attn_fwd", "pka_attn_bwd",
                 "pka_add_layernorm_fwd", "pka_add_layernorm_bwd", "pka_ce_fwd", "pka_ce_bwd", "pka_beam_advance",
                 "pka_tree_attn", "pka_adam_step", "pka_last_error"):
        assert must in syms


def test_library_exports_every_declared_symbol(library):
    lib = ctypes.CDLL(library.LIB_PATH)
    missing = [s for s in declared_symbols() if not hasattr(lib, s)]
    assert not missing, "declared in pka_b200.h but not exported: %s" % missing


def test_library_is_sm100a_only_and_has_no_torch_dependency(library):
    out = subprocess.run(["cuobjdump", "--list-elf", library.LIB_PATH], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_(\d+a?)", out))
    assert archs == {"100a"}, archs
    ldd = subprocess.run(["ldd", library.LIB_PATH], capture_output=True, text=True).stdout
    assert "torch" not in ldd and "c10" not in ldd


def test_error_reporting_without_a_gpu(library):
    lib = library.lib()
    assert lib.pka_version() >= 100
    rc = lib.pka_gemm_f32(None, None)                     # argument validation happens before any CUDA call
    assert rc == 1
    assert b"null" in lib.pka_last_error()
    with pytest.raises(RuntimeError):
        library.check(rc, "gemm_f32")


def test_struct_layouts_match_the_header(library):
    """ctypes mirrors of the descriptor structs must have the C sizes (guards against silent ABI drift)."""
    src = r'''
    #include <stdio.h>
    #include "pka_b200.h"
    int main(void){ printf("%zu %zu %zu %zu %zu %zu %zu\n", sizeof(pka_dropout), sizeof(pka_gemm_desc), sizeof(pka_attn_desc), sizeof(pka_beam_desc), sizeof(pka_tc_desc), sizeof(pka_reduce_job), sizeof(pka_relayout_job)); return 0; }
    '''
    import tempfile
    with tempfile.TemporaryDirectory() as d:
        c = os.path.join(d, "s.c")
        open(c, "w").write(src)
        exe = os.path.join(d, "s")
        subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), c, "-o", exe], check=True)
        sizes = [int(x) for x in subprocess.run([exe], capture_output=True, text=True).stdout.split()]
    assert sizes == [ctypes.sizeof(library.Dropout), ctypes.sizeof(library.GemmDesc), ctypes.sizeof(library.AttnDesc),
                     ctypes.sizeof(library.BeamDesc), ctypes.sizeof(library.TcDesc), ctypes.sizeof(library.ReduceJob),
                     ctypes.sizeof(library.RelayoutJob)]


def test_product_package_never_imports_the_oracle():
    """oracle/ is test infrastructure: only tests/, __graft_entry__.smoke() and bench.py's CPU arm may use it."""
    import re
    pkg = os.path.join(ROOT, "pytorch-kaldi-asr_b200")
    pattern = re.compile(r"^\s*(from\s+oracle\b|import\s+oracle\b)", re.M)
    offenders = []
    for base, _, files in os.walk(pkg):
        for name in files:
            if name.endswith(".py") and pattern.search(open(os.path.join(base, name), encoding="utf-8").read()):
                offenders.append(os.path.relpath(os.path.join(base, name), ROOT))
    assert offenders == []


def test_layernorm_backward_partial_row_contract_is_host_arithmetic(library):
    """pka_ln_bwd_blocks(rows) is pure host code (no device call): it is both the grid of the LayerNorm backward kernel and
    the number of partial rows the caller sizes its workspace / reduction job with (include/pka_b200.h).  Small tensors:
    one row per warp, at most 256 partial rows; large ones: a persistent grid that is a whole number of waves for 3, 2 and
    1 resident CTAs per SM (148 * 6), mid-sized ones one CTA pair per SM (148 * 2)."""
    lib = ctypes.CDLL(library.LIB_PATH)
    lib.pka_ln_bwd_blocks.restype = ctypes.c_int
    lib.pka_ln_bwd_blocks.argtypes = [ctypes.c_int]
    f = lib.pka_ln_bwd_blocks
    assert f(1) == 1 and f(8) == 1 and f(9) == 2
    assert f(2016) == 252                      # the timed step's decoder rows: one row per warp, 8 warps per CTA
    assert f(8192) == 256 and f(4096) == 256
    assert f(8193) == 296 and f(12792) == 296  # config 5's encoder rows
    assert f(148 * 6 * 8 * 4) == 888 and f(223552) == 888
    for rows in (1, 7, 2016, 8192, 8193, 30000, 1 << 20):
        assert 1 <= f(rows) <= 888
