"""CPU tests of the host side: the C-ABI library loads and exports every symbol of include/skf.h,
host-only queries work without a GPU, and the product-side initialiser equals the oracle's."""
import ctypes as C
import os
import re

import pytest
import torch

from oracle import sketchformer_oracle as O
from oracle.make_golden import CFG1, TINY
from sketchformer_b200 import _lib
from sketchformer_b200.engine import make_cfg
from sketchformer_b200.init import init_params

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    hdr = open(os.path.join(ROOT, "include", "skf.h")).read()
    return sorted(set(re.findall(r"\b(skf_[a-z0-9_]+)\s*\(", hdr)))


def test_library_exports_every_declared_symbol():
    lib = _lib.load()
    syms = declared_symbols()
    assert len(syms) >= 30
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in include/skf.h but not exported by libskf.so"
        assert s in _lib.SIGNATURES, f"{s} has no ctypes signature"
    assert lib.skf_version() == 1


def test_ctypes_struct_layouts_match_the_library():
    """include/skf.h structs as compiled vs their ctypes mirrors (also enforced by _lib.load())."""
    lib = _lib.load()
    assert lib.skf_abi_sizeof(0) == C.sizeof(_lib.SkfCfg)
    assert lib.skf_abi_sizeof(1) == C.sizeof(_lib.SkfGemmEpilogue)
    assert lib.skf_abi_sizeof(2) == C.sizeof(_lib.SkfParamEntry)
    assert lib.skf_abi_sizeof(99) == 0


def test_param_table_matches_oracle_param_count():
    lib = _lib.load()
    for kw, B, want in ((CFG1, 8, 2272630), (dict(num_layers=4, d_model=256, dff=1024, num_heads=8, lowerdim=256,
                                                   max_seq_len=200, vocab_size=1004, n_classes=345), 128, 8299989)):
        c = make_cfg(B, kw["max_seq_len"], kw["d_model"], kw["dff"], kw["num_heads"], kw["num_layers"],
                     kw["vocab_size"], kw["n_classes"], kw["lowerdim"])
        ent = (_lib.SkfParamEntry * 512)()
        n = lib.skf_param_table(C.byref(c), ent, 512)
        assert n > 0
        total = sum(ent[i].rows * ent[i].cols for i in range(n))
        assert total == want                      # SURVEY §8(d) parameter counts
        arena = lib.skf_param_count(C.byref(c))
        assert want <= arena <= want + 8 * n      # 32-byte alignment padding only (16 bytes in the bf16 images)
        offs = [(ent[i].offset, ent[i].rows * ent[i].cols) for i in range(n)]
        for (o0, s0), (o1, _) in zip(offs, offs[1:]):
            assert o0 % 8 == 0 and o0 + s0 <= o1
        assert lib.skf_workspace_bytes(C.byref(c)) > 0


def test_bad_configs_are_rejected_with_a_message():
    lib = _lib.load()
    c = make_cfg(8, 64, 130, 512, 8, 4, 1004, 10, 256)     # d_model % heads != 0
    assert lib.skf_param_count(C.byref(c)) < 0
    assert b"num_heads" in lib.skf_last_error()
    c = make_cfg(8, 64, 128, 512, 8, 4, 1002, 10, 256)    # vocab not a multiple of 4
    assert lib.skf_workspace_bytes(C.byref(c)) == 0
    assert b"vocab" in lib.skf_last_error()
    c = make_cfg(8, 64, 128, 512, 8, 4, 0, 10, 256, continuous=True)   # stroke-5 input needs no vocabulary
    assert lib.skf_workspace_bytes(C.byref(c)) > 0
    with pytest.raises(_lib.SkfError):
        _lib.check(lib.skf_set_option(b"nope", 1), "set_option")


def test_product_initialiser_equals_oracle():
    for kw in (TINY, CFG1, dict(TINY, use_continuous_data=True), dict(TINY, attn_version=2, class_buffer_layers=2)):
        a = init_params(kw, 0)
        b = O.init_params(O.OracleConfig(**kw), 0)
        assert list(a) == list(b)
        for k in a:
            assert torch.equal(a[k], b[k]), k


def test_engine_refuses_to_run_without_cuda():
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    from sketchformer_b200.engine import Engine
    with pytest.raises(_lib.SkfError):
        Engine(make_cfg(4, 24, 64, 128, 4, 2, 68, 10, 32))
