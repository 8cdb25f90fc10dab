"""Oracle restatement vs the reference's own code run live (oracle/_ref binaries).  Skipped where
the binaries are absent; tests/golden covers those machines."""
import os

import numpy as np
import pytest

from oracle import pyoracle as po

needs_ref = pytest.mark.skipif(po.ref_binary("ref_v2_exact") is None, reason="oracle/_ref not built (needs /root/reference)")
TEX_DIR = "/root/reference/Textures"


@needs_ref
@pytest.mark.parametrize("bounces,frames,tiles", [(4, 5, (2, 4)), (8, 3, (1, 8)), (0, 2, (4, 2))])
def test_v2(oracle, bounces, frames, tiles):
    W, H = 160, 96
    o, _ = oracle.render(po.PROFILE_V2, W, H, tiles[0], tiles[1], bounces, frames)
    r = po.run_ref("ref_v2_exact", W, H, tiles[0], tiles[1], frames, bounces=bounces)["buffer"]
    assert np.array_equal(o, r)


@needs_ref
def test_simt_textured(oracle):
    W, H = 160, 96
    env = po.synthetic_env(256, 128)
    o, _ = oracle.render(po.PROFILE_SIMT_TEXTURED, W, H, 2, 4, 4, 4, env=env, env_kind=po.ENV_EQUIRECT)
    r = po.run_ref("ref_simt_textured_exact", W, H, 2, 4, 4, bounces=4, env=env)["buffer"]
    assert np.array_equal(o, r)


@needs_ref
@pytest.mark.parametrize("name,kind,sampler,envshape", [
    ("ref_v4_equirect_random_exact", po.ENV_EQUIRECT, po.SAMPLER_RANDOM, (256, 128)),
    ("ref_v4_equirect_bilinear_exact", po.ENV_EQUIRECT, po.SAMPLER_BILINEAR, (256, 128)),
    ("ref_v4_cubemap_random_exact", po.ENV_CUBEMAP, po.SAMPLER_RANDOM, (64, 384)),
    ("ref_v4_cubemap_bilinear_exact", po.ENV_CUBEMAP, po.SAMPLER_BILINEAR, (64, 384)),
])
def test_v4(oracle, name, kind, sampler, envshape):
    W, H = 320, 180
    env = po.synthetic_env(*envshape)
    o, _ = oracle.render(po.PROFILE_V4, W, H, 10, 15, 8, 4, env=env, env_kind=kind, env_sampler=sampler)
    r = po.run_ref(name, W, H, 10, 15, 4, bounces=8, env=env)["buffer"]
    assert np.array_equal(o, r)


@needs_ref
@pytest.mark.parametrize("name,flags", [("ref_v4_equirect_random_expexact_exact", po.V4_EXACT_EXP),
                                        ("ref_v4_equirect_random_sincos_exact", po.V4_SINCOS_UNIT_VECTORS),
                                        ("ref_v4_equirect_random_allexact_exact", po.V4_EXACT_EXP | po.V4_SINCOS_UNIT_VECTORS)])
def test_v4_non_default_switches(oracle, name, flags):
    """reference builds with USE_FAST_APPROXIMATE_EXP / USE_UNIT_VECTOR_REJECTION_SAMPLING (/ ..._ACES_TONEMAP) set to 0
    (global_preprocessor_flags.h:63-65; oracle/ref_build/build_ref.sh) against the oracle's v4_flags"""
    if po.ref_binary(name) is None:
        pytest.skip(name + " not built")
    W, H = 320, 180
    env = po.synthetic_env(256, 128)
    o, _ = oracle.render(po.PROFILE_V4, W, H, 10, 15, 8, 4, env=env, env_kind=po.ENV_EQUIRECT, env_sampler=po.SAMPLER_RANDOM, v4_flags=flags)
    r = po.run_ref(name, W, H, 10, 15, 4, bounces=8, env=env)["buffer"]
    assert np.array_equal(o, r)
    if flags == 3:  # that build's CopyOutputToFile uses the exact ACES curve
        res = po.run_ref(name, W, H, 10, 15, 4, bounces=8, env=env, threads=1, ldr=True)
        assert np.array_equal(res["ldr"], oracle.resolve_ldr(res["buffer"], W, H, 10, 15, mode=2))


@needs_ref
@pytest.mark.parametrize("name,mode", [("ref_v4_equirect_random_gammaexact_exact", 4), ("ref_v4_equirect_random_ldrexact_exact", 6)])
def test_v4_exact_gamma_tonemap(oracle, name, mode):
    """reference builds with USE_FAST_APPROXIMATE_GAMMA 0 (and ..._ACES_TONEMAP 0): CopyOutputToFile against the oracle's resolve"""
    if po.ref_binary(name) is None:
        pytest.skip(name + " not built")
    W, H = 320, 180
    env = po.synthetic_env(256, 128)
    res = po.run_ref(name, W, H, 10, 15, 4, bounces=8, env=env, threads=1, ldr=True)
    o, _ = oracle.render(po.PROFILE_V4, W, H, 10, 15, 8, 4, env=env, env_kind=po.ENV_EQUIRECT, env_sampler=po.SAMPLER_RANDOM)
    assert np.array_equal(res["buffer"], o)
    assert np.array_equal(res["ldr"], oracle.resolve_ldr(o, W, H, 10, 15, mode=mode))


@needs_ref
@pytest.mark.parametrize("bounces,frames,tiles", [(8, 4, (2, 4)), (16, 2, (4, 2))])
def test_v3_redo(oracle, bounces, frames, tiles):
    W, H = 256, 144
    env = po.synthetic_env(256, 128)
    o, _ = oracle.render(po.PROFILE_V3REDO, W, H, tiles[0], tiles[1], bounces, frames, env=env, env_kind=po.ENV_EQUIRECT,
                         env_sampler=po.SAMPLER_BILINEAR)
    r = po.run_ref("ref_v3redo_exact", W, H, tiles[0], tiles[1], frames, bounces=bounces, env=env)["buffer"]
    assert np.array_equal(o, r)


@needs_ref
@pytest.mark.parametrize("bounces,frames,tiles", [(8, 4, (2, 4)), (3, 3, (4, 3))])
def test_v3_redo_scene0(oracle, bounces, frames, tiles):
    """demofox_path_tracing_v3_redo.cpp compiled with `#define SCENE 0` (the Cornell box variant, :392-479, :530-580)"""
    if not po.ref_binary("ref_v3redo_scene0_exact"):
        pytest.skip("ref_v3redo_scene0_exact not built")
    W, H = 256, 144
    env = po.synthetic_env(256, 128)
    o, _ = oracle.render(po.PROFILE_V3REDO_SCENE0, W, H, tiles[0], tiles[1], bounces, frames, env=env, env_kind=po.ENV_EQUIRECT,
                         env_sampler=po.SAMPLER_BILINEAR)
    r = po.run_ref("ref_v3redo_scene0_exact", W, H, tiles[0], tiles[1], frames, bounces=bounces, env=env)["buffer"]
    assert np.array_equal(o, r)
    # a different picture than SCENE 1
    o1, _ = oracle.render(po.PROFILE_V3REDO, W, H, tiles[0], tiles[1], bounces, frames, env=env, env_kind=po.ENV_EQUIRECT,
                          env_sampler=po.SAMPLER_BILINEAR)
    assert not np.array_equal(o, o1)


@needs_ref
@pytest.mark.skipif(not os.path.exists(os.path.join(TEX_DIR, "HDR_040_Field_Env.hdr")), reason="reference textures absent")
def test_v4_real_textures(oracle, tmp_path):
    """The reference's shipped HDR env maps, decoded by the reference's own loader (stb_image)."""
    import subprocess
    tool = po.ref_binary("ref_asset_tool")
    out = tmp_path / "eq.f32"
    w, h, c = map(int, subprocess.run([tool, "equirect", os.path.join(TEX_DIR, "HDR_040_Field_Env.hdr"), str(out)],
                                      check=True, capture_output=True, text=True).stdout.split())
    env = np.fromfile(out, dtype=np.float32).reshape(h, w, 3)
    W, H = 256, 144
    o, _ = oracle.render(po.PROFILE_V4, W, H, 4, 6, 8, 3, env=env, env_kind=po.ENV_EQUIRECT, env_sampler=po.SAMPLER_RANDOM)
    r = po.run_ref("ref_v4_equirect_random_exact", W, H, 4, 6, 3, bounces=8, env=env)["buffer"]
    assert np.array_equal(o, r)
    o, _ = oracle.render(po.PROFILE_SIMT_TEXTURED, W, H, 2, 4, 4, 3, env=env, env_kind=po.ENV_EQUIRECT)
    r = po.run_ref("ref_simt_textured_exact", W, H, 2, 4, 3, bounces=4, env=env)["buffer"]
    assert np.array_equal(o, r)
    faces = [os.path.join(TEX_DIR, f + ".hdr") for f in ("px", "nx", "py", "ny", "pz", "nz")]
    out = tmp_path / "cube.f32"
    w, h, c = map(int, subprocess.run([tool, "cubemap"] + faces + [str(out)], check=True, capture_output=True,
                                      text=True).stdout.split())
    cube = np.fromfile(out, dtype=np.float32).reshape(h, w, 3)
    o, _ = oracle.render(po.PROFILE_V4, W, H, 4, 6, 8, 3, env=cube, env_kind=po.ENV_CUBEMAP, env_sampler=po.SAMPLER_RANDOM)
    r = po.run_ref("ref_v4_cubemap_random_exact", W, H, 4, 6, 3, bounces=8, env=cube)["buffer"]
    assert np.array_equal(o, r)


@needs_ref
def test_reference_is_thread_count_invariant():
    a = po.run_ref("ref_v2_exact", 64, 48, 2, 4, 2, bounces=4, threads=1)["buffer"]
    b = po.run_ref("ref_v2_exact", 64, 48, 2, 4, 2, bounces=4, threads=6)["buffer"]
    assert np.array_equal(a, b)
