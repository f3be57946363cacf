"""The C-ABI library: loads, exports every symbol include/heroam_gpu.h declares, and
refuses to work without a GPU (no CPU fallback).  No compute calls here."""
from __future__ import annotations

import ctypes as C
import os
import re

import numpy as np
import pytest

from sphere_roaming_b200 import engine
from sphere_roaming_b200.types import make_config

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_functions():
    text = open(os.path.join(ROOT, "include", "heroam_gpu.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(heroam_gpu_\w+)\s*\(", text)))


def test_header_and_binding_agree():
    assert declared_functions() == sorted(engine.ABI_SYMBOLS)


def test_library_exports_every_declared_symbol():
    lib = engine.load_library()
    for name in declared_functions():
        assert hasattr(lib, name), f"libheroam_gpu.so does not export {name}"
    assert lib.heroam_gpu_abi_version() == 2


def test_struct_sizes_match_header():
    # heroam_options: int, uint, ull, int, uint, int, int; heroam_traj_result: 16 bytes
    assert C.sizeof(engine.Options) == 32
    assert engine.RESULT_DTYPE.itemsize == 16
    assert C.sizeof(engine.RunInfo) == 16 * 8
    assert C.sizeof(engine.Stats) == 8 * (2 + 256 + 4096 + 4096 + 8 + 1 + 1) + C.sizeof(engine.RunInfo)


def test_no_cpu_fallback(table):
    """Without a CUDA device the engine must fail loudly, not compute on the CPU."""
    if engine.device_count() > 0:
        pytest.skip("a GPU is present")
    with pytest.raises(engine.HeroamError, match="CUDA|device"):
        engine.Engine(make_config(), table)


def test_argument_validation_needs_no_gpu(table):
    lib = engine.load_library()
    h = C.c_void_p()
    assert lib.heroam_gpu_create(None, None, 0, None, C.byref(h)) == 2  # HEROAM_E_ARG
    assert b"null" in lib.heroam_gpu_last_error()
    assert lib.heroam_gpu_run(None) == 2
    info = engine.RunInfo()
    assert lib.heroam_gpu_last_run_info(None, C.byref(info)) == 2
