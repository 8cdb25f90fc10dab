"""The C-ABI library loads and exports every symbol include/b200pt.h declares; argument checking
and the no-fallback rule work without a GPU (no compute calls here)."""
import ctypes
import os
import re

import pytest

from cpuperformanceraytracer_b200 import api

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _has_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def test_header_symbols_exported():
    hdr = open(os.path.join(ROOT, "include", "b200pt.h")).read()
    declared = sorted(set(re.findall(r"\b(b200pt_[a-z_0-9]+)\s*\(", hdr)))
    assert declared, "no declarations found"
    lib = api.load_library()
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in b200pt.h but not exported"
    assert sorted(api.ABI_SYMBOLS) == declared
    assert lib.b200pt_api_version() == 2


def test_built_in_scenes_match_the_specialised_kernels_tables():
    """the scene-specialised kernels' compile-time tables (sphere data as immediates, zero components of the v4 quad
    vectors) are what the host builds from the reference's scene expressions -- otherwise the library would silently
    run the slower generic kernels"""
    lib = api.load_library()
    for profile in (api.PROFILE_V2, api.PROFILE_SIMT_TEXTURED, api.PROFILE_OPT_V4, api.PROFILE_V3_REDO, api.PROFILE_V3_REDO_SCENE0):
        assert lib.b200pt_static_tables_match(profile) == 1
    assert lib.b200pt_static_tables_match(99) == -1


def test_default_params_follow_reference_flags():
    p = api.default_params(api.PROFILE_OPT_V4)  # global_preprocessor_flags.h:56-66
    assert (p.env_kind, p.env_sampler, p.accum_mode, p.math_mode) == (api.ENV_EQUIRECT, api.SAMPLER_RANDOM,
                                                                      api.ACCUM_RUNNING_AVERAGE, api.MATH_PARITY)
    assert p.struct_size == ctypes.sizeof(api.Params)
    p = api.default_params(api.PROFILE_SIMT_TEXTURED)
    assert (p.env_kind, p.env_sampler) == (api.ENV_EQUIRECT, api.SAMPLER_POINT)
    with pytest.raises(api.B200PTError):
        api.default_params(7)


def test_invalid_arguments_rejected():
    lib = api.load_library()
    ctx = ctypes.c_void_p()
    assert lib.b200pt_create(None, ctypes.byref(ctx)) == 1
    p = api.default_params(api.PROFILE_V2)
    p.struct_size = 3
    assert lib.b200pt_create(ctypes.byref(p), ctypes.byref(ctx)) == 1
    p = api.default_params(api.PROFILE_V2)
    p.math_mode = 9
    assert lib.b200pt_create(ctypes.byref(p), ctypes.byref(ctx)) == 1
    assert lib.b200pt_render_frames(None, 1) == 1
    assert lib.b200pt_error_string(2).decode().startswith("CUDA")


@pytest.mark.skipif(_has_gpu(), reason="checks the no-GPU failure mode")
def test_no_cpu_fallback_without_gpu():
    # the product path must fail loudly, never render on the CPU
    with pytest.raises(api.B200PTError, match="no CPU fallback"):
        api.Renderer(profile=api.PROFILE_V2)


def test_product_does_not_reference_oracle():
    """Nothing under cpuperformanceraytracer_b200/ or include/ may import, include or link oracle/."""
    bad = []
    for base in ("cpuperformanceraytracer_b200", "include"):
        for dp, _, files in os.walk(os.path.join(ROOT, base)):
            if os.path.basename(dp) in ("build", "__pycache__"):
                continue
            for f in files:
                if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                    txt = open(os.path.join(dp, f), errors="ignore").read()
                    if re.search(r'#include\s+"[^"]*oracle|import\s+oracle|from\s+oracle|liboracle|pyoracle', txt):
                        bad.append(os.path.join(dp, f))
    assert not bad, bad
