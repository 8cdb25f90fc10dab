"""The C++ host mirror (cpuperformanceraytracer_b200/host/demofox_render.* -- the reference's own entry-point names
and signatures -- and the render_offline CLI, a mirror of ApplicationState::RenderOffline, Application.cpp:400-458)
run on the GPU and compared with the oracle BIT FOR BIT: the f32 dump after 2 warm-up + N frames is the oracle's
buffer after N + 2 frames, for every renderer variant, with .hdr assets read by the product's own loader."""
import os
import subprocess

import numpy as np
import pytest

from test_host_io import HOSTLIB, load_cubemap, load_hdr, rgbe_encode, write_hdr  # noqa: F401
from test_host_io import io  # noqa: F401  (fixture)

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CLI = os.path.join(ROOT, "cpuperformanceraytracer_b200", "render_offline")
W, H, NTX, NTY, FRAMES = 256, 128, 2, 4, 6
TEX_DIR = "/root/reference/Textures"


def _ngpus():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


def run_cli(tmp_path, *args, frames=FRAMES, name="dump"):
    dump = tmp_path / (name + ".f32")
    cmd = [CLI, "--width", str(W), "--height", str(H), "--tiles-x", str(NTX), "--tiles-y", str(NTY), "--frames", str(frames),
           "--dump-f32", str(dump), "--out", str(tmp_path / (name + ".bmp"))] + [str(a) for a in args]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert res.returncode == 0, res.stdout + res.stderr
    assert "Total render time" in res.stdout
    return np.fromfile(dump, np.float32), res.stdout


def make_equirect(tmp_path, oracle, io, w=128, h=64):
    path = tmp_path / "env.hdr"
    img = oracle.synthetic_env(w, h)[::-1] * 4.0 + 0.01  # top-down rows for the file
    write_hdr(path, rgbe_encode(img.astype(np.float32)), rle=True)
    tex = load_hdr(io, path)  # what LoadTexture hands to the renderer
    assert tex is not None and tex.shape == (h, w, 3)
    return path, tex


def make_cubemap(tmp_path, oracle, io, face=32):
    paths = []
    for k, nm in enumerate(("px", "nx", "py", "ny", "pz", "nz")):
        p = tmp_path / (nm + ".hdr")
        img = (oracle.synthetic_env(face, face) * (1.0 + k) + 0.02).astype(np.float32)
        write_hdr(p, rgbe_encode(img), rle=(k % 2 == 0))
        paths.append(p)
    atlas = load_cubemap(io, paths)
    assert atlas is not None and atlas.shape == (6 * face, face, 3)
    return paths, atlas


def test_render_offline_v2_is_the_oracle(oracle, tmp_path):
    g, _ = run_cli(tmp_path, "--variant", "v2", "--bounces", 8)
    o, _ = oracle.render(oracle.PROFILE_V2, W, H, NTX, NTY, 8, FRAMES + 2)
    assert np.array_equal(g, o)
    # one entry-point call per frame (DemofoxRenderV2 x N, the reference's own loop) == the batched form
    g1, _ = run_cli(tmp_path, "--variant", "v2", "--bounces", 8, "--per-frame-calls", name="perframe")
    assert np.array_equal(g1, o)
    # the reference's default bounce count (c_numBounces = 4, v2.cpp:22)
    g4, _ = run_cli(tmp_path, "--variant", "v2", name="b4")
    o4, _ = oracle.render(oracle.PROFILE_V2, W, H, NTX, NTY, 4, FRAMES + 2)
    assert np.array_equal(g4, o4)


def test_render_offline_v4_equirect(oracle, io, tmp_path):
    path, tex = make_equirect(tmp_path, oracle, io)
    for flags, sampler in ((["--env", path], oracle.SAMPLER_RANDOM), (["--env", path, "--bilinear"], oracle.SAMPLER_BILINEAR)):
        g, _ = run_cli(tmp_path, "--variant", "v4", *flags)
        o, _ = oracle.render(oracle.PROFILE_V4, W, H, NTX, NTY, 8, FRAMES + 2, env=tex, env_kind=oracle.ENV_EQUIRECT, env_sampler=sampler)
        assert np.array_equal(g, o)
    g, _ = run_cli(tmp_path, "--variant", "v4", "--env", path, "--per-frame-calls", name="perframe")
    o, _ = oracle.render(oracle.PROFILE_V4, W, H, NTX, NTY, 8, FRAMES + 2, env=tex, env_kind=oracle.ENV_EQUIRECT, env_sampler=oracle.SAMPLER_RANDOM)
    assert np.array_equal(g, o)
    # no texture: USE_ENV_MAP 0 (constant ambient)
    g, _ = run_cli(tmp_path, "--variant", "v4", name="noenv")
    o, _ = oracle.render(oracle.PROFILE_V4, W, H, NTX, NTY, 8, FRAMES + 2)
    assert np.array_equal(g, o)


def read_bmp24(path, w, h):
    """rows of the 24-bit BMP the CLI writes (bottom row first, B G R bytes) -> (h, w) of R | G<<8 | B<<16, top row first"""
    b = np.frombuffer(open(path, "rb").read(), np.uint8)
    stride = (w * 3 + 3) & ~3
    rows = b[54:54 + stride * h].reshape(h, stride)[::-1, :w * 3].reshape(h, w, 3).astype(np.uint32)
    return rows[:, :, 2] | (rows[:, :, 1] << 8) | (rows[:, :, 0] << 16)


def test_render_offline_v4_non_default_switches(oracle, io, tmp_path):
    """--exact-exp / --sincos-unit-vectors / --exact-aces / --exact-gamma = USE_FAST_APPROXIMATE_EXP, USE_UNIT_VECTOR_REJECTION_SAMPLING,
    USE_FAST_APPROXIMATE_ACES_TONEMAP, USE_FAST_APPROXIMATE_GAMMA set to 0 (global_preprocessor_flags.h:62-65): f32 dump and the .bmp"""
    path, tex = make_equirect(tmp_path, oracle, io)
    for args, flags, aces in ((["--exact-exp"], oracle.V4_EXACT_EXP, 0), (["--sincos-unit-vectors"], oracle.V4_SINCOS_UNIT_VECTORS, 0),
                              (["--exact-exp", "--sincos-unit-vectors", "--exact-aces"], 3, 2), (["--exact-gamma"], 0, 4),
                              (["--exact-aces", "--exact-gamma"], 0, 6)):
        g, _ = run_cli(tmp_path, "--variant", "v4", "--env", path, *args, name="sw")
        o, _ = oracle.render(oracle.PROFILE_V4, W, H, NTX, NTY, 8, FRAMES + 2, env=tex, env_kind=oracle.ENV_EQUIRECT,
                             env_sampler=oracle.SAMPLER_RANDOM, v4_flags=flags)
        assert np.array_equal(g, o)
        ldr = oracle.resolve_ldr(o, W, H, NTX, NTY, mode=aces).reshape(H, W) & 0xFFFFFF
        assert np.array_equal(read_bmp24(tmp_path / "sw.bmp", W, H), ldr)


def test_render_offline_v4_cubemap(oracle, io, tmp_path):
    paths, atlas = make_cubemap(tmp_path, oracle, io)
    for extra, sampler in (([], oracle.SAMPLER_RANDOM), (["--bilinear"], oracle.SAMPLER_BILINEAR)):
        g, _ = run_cli(tmp_path, "--variant", "v4", "--cubemap", *paths, *extra)
        o, _ = oracle.render(oracle.PROFILE_V4, W, H, NTX, NTY, 8, FRAMES + 2, env=atlas, env_kind=oracle.ENV_CUBEMAP, env_sampler=sampler)
        assert np.array_equal(g, o)


def test_render_offline_simt_and_v3redo(oracle, io, tmp_path):
    path, tex = make_equirect(tmp_path, oracle, io)
    g, _ = run_cli(tmp_path, "--variant", "simt", "--env", path, "--bounces", 4)
    o, _ = oracle.render(oracle.PROFILE_SIMT_TEXTURED, W, H, NTX, NTY, 4, FRAMES + 2, env=tex, env_kind=oracle.ENV_EQUIRECT)
    assert np.array_equal(g, o)
    g, _ = run_cli(tmp_path, "--variant", "v3redo", "--env", path)
    o, _ = oracle.render(oracle.PROFILE_V3REDO, W, H, NTX, NTY, 8, FRAMES + 2, env=tex, env_kind=oracle.ENV_EQUIRECT,
                         env_sampler=oracle.SAMPLER_BILINEAR)
    assert np.array_equal(g, o)
    g, _ = run_cli(tmp_path, "--variant", "v3redo0", "--env", path, name="scene0")   # `#define SCENE 0`
    o, _ = oracle.render(oracle.PROFILE_V3REDO_SCENE0, W, H, NTX, NTY, 8, FRAMES + 2, env=tex, env_kind=oracle.ENV_EQUIRECT,
                         env_sampler=oracle.SAMPLER_BILINEAR)
    assert np.array_equal(g, o)


@pytest.mark.skipif(not os.path.isdir(TEX_DIR), reason="the reference's shipped textures are only present in the build container")
def test_render_offline_with_the_references_own_textures(oracle, io, tmp_path):
    path = os.path.join(TEX_DIR, "HDR_040_Field_Env.hdr")
    tex = load_hdr(io, path)
    g, _ = run_cli(tmp_path, "--variant", "v4", "--env", path)
    o, _ = oracle.render(oracle.PROFILE_V4, W, H, NTX, NTY, 8, FRAMES + 2, env=tex, env_kind=oracle.ENV_EQUIRECT, env_sampler=oracle.SAMPLER_RANDOM)
    assert np.array_equal(g, o)


@pytest.mark.skipif(_ngpus() < 2, reason="needs at least 2 GPUs")
def test_render_offline_multi_gpu(oracle, tmp_path):
    """--gpus N: the C++ entry points shard a render call over the GPUs of the process (b200pt_group_*)"""
    n = min(_ngpus(), 8)
    o, _ = oracle.render(oracle.PROFILE_V2, W, H, NTX, NTY, 8, 24 + 2)
    g, out = run_cli(tmp_path, "--variant", "v2", "--bounces", 8, "--gpus", n, "--shard", "tiles", frames=24, name="tiles")
    assert np.array_equal(g, o) and "Cross-GPU combine step" in out
    for combine in ("nccl", "peer", "fused"):
        g, _ = run_cli(tmp_path, "--variant", "v2", "--bounces", 8, "--gpus", n, "--shard", "spp", "--combine", combine, frames=24,
                       name="spp_" + combine)
        assert np.allclose(g, o, rtol=3e-6, atol=3e-6)
