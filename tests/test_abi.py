"""The C-ABI library loads, exports every symbol include/savont_b200.h declares, and refuses to work
without a CUDA device (no CPU fallback).  CPU only; no compute calls."""
import ctypes as C
import os
import re

import pytest


def test_library_exports_every_declared_symbol():
    from savont_b200 import _lib
    assert os.path.exists(_lib.LIB_PATH), "run python -m savont_b200.build"
    names = _lib.declared_symbols()
    assert len(names) >= 20 and "savont_b200_map_reads" in names and "savont_b200_align_pairs" in names
    lib = C.CDLL(_lib.LIB_PATH)
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, missing
    lib.savont_b200_abi_version.restype = C.c_int
    hdr = open(_lib.HEADER_PATH).read()
    assert lib.savont_b200_abi_version() == int(re.search(r"#define SAVONT_B200_ABI_VERSION (\d+)", hdr).group(1))


def test_header_cites_reference_interfaces():
    from savont_b200 import _lib
    hdr = open(_lib.HEADER_PATH).read()
    for cite in ("src/seeding.rs:317-327", "src/asv_cluster.rs:356-383", "src/alignment.rs:1848-1863", "src/alignment.rs:1786-1925",
                 "src/kmer_comp.rs:39-66", "src/kmer_comp.rs:41-48"):
        assert cite in hdr


def test_default_params_and_struct_layout():
    from savont_b200 import _lib
    p = _lib.default_params()
    assert (p.k, p.c, p.l, p.min_bq, p.band, p.strand_mode) == (17, 11, 3, 25, 64, 0)
    assert p.ratio_thresh == 0.005 and p.min_mini_frac_base == 0.95
    assert C.sizeof(_lib.Params) == 48 and C.sizeof(_lib.Stats) == 16 * 8
    assert p.candidate_mode == 0


def test_no_cpu_fallback_without_a_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    from savont_b200 import _lib
    with pytest.raises(_lib.SavontB200Error) as e:
        _lib.Context(0)
    assert e.value.code == _lib.ERR_CUDA and "no CPU fallback" in str(e.value)


def test_product_does_not_import_the_oracle():
    root = os.path.join(os.path.dirname(__file__), "..", "savont_b200")
    for dp, _, files in os.walk(root):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".hpp")):
                txt = open(os.path.join(dp, f), errors="ignore").read()
                assert "savont_oracle" not in txt and "from oracle" not in txt and "import oracle" not in txt, f


def test_integration_md_names_every_entry_point():
    """INTEGRATION.md is where a maintainer looks for the binding of an entry point: every function declared in include/savont_b200.h
    is named there (the kmer_shard family by its shorthand)"""
    import os
    from savont_b200 import _lib
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    text = open(os.path.join(root, "INTEGRATION.md")).read()
    text = text.replace("`savont_b200_kmer_shard_hist` / `_partitions` / `_alloc` / `_scatter` / `_count`",
                        " ".join("savont_b200_kmer_shard_" + x for x in ("hist", "partitions", "alloc", "scatter", "count")))
    missing = [sym for sym in _lib.declared_symbols() if sym not in text]
    assert not missing, missing
