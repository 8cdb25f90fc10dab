"""pyoracle.py -- TEST INFRASTRUCTURE: ctypes access to oracle/liboracle.so (the C restatement,
pt_oracle.c) and subprocess access to the reference's own binaries under oracle/_ref/
(ref_build/build_ref.sh).  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs import this module; the product never does.
"""
import ctypes
import json
import os
import subprocess
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "liboracle.so")
REF_DIR = os.path.join(HERE, "_ref")

PROFILE_V2, PROFILE_SIMT_TEXTURED, PROFILE_V4, PROFILE_V3REDO, PROFILE_V3REDO_SCENE0 = 0, 1, 2, 3, 4
ENV_NONE, ENV_EQUIRECT, ENV_CUBEMAP = 0, 1, 2
SAMPLER_POINT, SAMPLER_BILINEAR, SAMPLER_RANDOM = 0, 1, 2


class OracleSceneV4(ctypes.Structure):
    _fields_ = [("num_quads", ctypes.c_int), ("num_spheres", ctypes.c_int), ("quad_vertices", ctypes.POINTER(ctypes.c_float)),
                ("spheres", ctypes.POINTER(ctypes.c_float)), ("materials", ctypes.POINTER(ctypes.c_float)),
                ("camera_position", ctypes.c_float * 3), ("camera_distance", ctypes.c_float)]


class OracleSceneCornell(ctypes.Structure):
    _fields_ = [("quad_vertices", ctypes.POINTER(ctypes.c_float)), ("spheres", ctypes.POINTER(ctypes.c_float)),
                ("materials", ctypes.POINTER(ctypes.c_float))]


class OracleParams(ctypes.Structure):
    _fields_ = [
        ("profile", ctypes.c_int),
        ("width", ctypes.c_int),
        ("height", ctypes.c_int),
        ("num_tiles_x", ctypes.c_int),
        ("num_tiles_y", ctypes.c_int),
        ("num_bounces", ctypes.c_int),
        ("env_kind", ctypes.c_int),
        ("env_sampler", ctypes.c_int),
        ("env", ctypes.POINTER(ctypes.c_float)),
        ("env_width", ctypes.c_int),
        ("env_height", ctypes.c_int),
        ("scene_v4", ctypes.POINTER(OracleSceneV4)),
        ("scene_cornell", ctypes.POINTER(OracleSceneCornell)),
        ("v4_flags", ctypes.c_int),
    ]
V4_EXACT_EXP, V4_SINCOS_UNIT_VECTORS = 1, 2


class OracleCounters(ctypes.Structure):
    _fields_ = [("paths", ctypes.c_uint64), ("segments", ctypes.c_uint64), ("escapes", ctypes.c_uint64)]


_lib = None


def build():
    subprocess.run(["make", "-C", HERE, "liboracle.so"], check=True, capture_output=True)


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            build()
        L = ctypes.CDLL(LIB_PATH)
        L.oracle_render.argtypes = [ctypes.POINTER(OracleParams), ctypes.POINTER(ctypes.c_float), ctypes.c_int,
                                    ctypes.c_int, ctypes.c_int, ctypes.POINTER(OracleCounters)]
        L.oracle_render.restype = ctypes.c_int
        L.oracle_path_radiance.argtypes = [ctypes.POINTER(OracleParams), ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                           ctypes.POINTER(ctypes.c_float)]
        L.oracle_wang_hash.argtypes = [ctypes.POINTER(ctypes.c_uint32)]
        L.oracle_wang_hash.restype = ctypes.c_uint32
        L.oracle_random01.argtypes = [ctypes.POINTER(ctypes.c_uint32)]
        L.oracle_random01.restype = ctypes.c_float
        L.oracle_seed.argtypes = [ctypes.c_int] * 3
        L.oracle_seed.restype = ctypes.c_uint32
        L.oracle_final_rng_state.argtypes = [ctypes.POINTER(OracleParams), ctypes.c_int, ctypes.c_int, ctypes.c_int]
        L.oracle_final_rng_state.restype = ctypes.c_uint32
        L.oracle_buffer_index.argtypes = [ctypes.c_int] * 7
        L.oracle_buffer_index.restype = ctypes.c_int64
        L.oracle_resolve_ldr.argtypes = [ctypes.POINTER(ctypes.c_float), ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                         ctypes.c_int, ctypes.POINTER(ctypes.c_uint32), ctypes.c_int]
        L.oracle_camera_distance.restype = ctypes.c_float
        _lib = L
    return _lib


def make_scene_v4(quads, spheres, materials, camera_position=(0.0, 0.0, 40.0), camera_distance=1.0):
    """Returns (OracleSceneV4, keepalive) from (nq,4,3) / (ns,4) / (nq+ns,17) arrays."""
    q = np.ascontiguousarray(quads, dtype=np.float32).reshape(-1, 12)
    s = np.ascontiguousarray(spheres, dtype=np.float32).reshape(-1, 4)
    m = np.ascontiguousarray(materials, dtype=np.float32).reshape(-1, 17)
    sc = OracleSceneV4()
    sc.num_quads, sc.num_spheres = q.shape[0], s.shape[0]
    fp = ctypes.POINTER(ctypes.c_float)
    sc.quad_vertices, sc.spheres, sc.materials = q.ctypes.data_as(fp), s.ctypes.data_as(fp), m.ctypes.data_as(fp)
    sc.camera_position = (ctypes.c_float * 3)(*camera_position)
    sc.camera_distance = camera_distance
    return sc, (q, s, m)


def make_scene_cornell(quads, spheres, materials):
    """(6, 4, 3) vertices, (3, 4) xyz + radius, (9, 11) legacy materials -> (OracleSceneCornell, keep-alive arrays)"""
    q = np.ascontiguousarray(quads, dtype=np.float32).reshape(6, 12)
    s = np.ascontiguousarray(spheres, dtype=np.float32).reshape(3, 4)
    m = np.ascontiguousarray(materials, dtype=np.float32).reshape(9, 11)
    fp = ctypes.POINTER(ctypes.c_float)
    sc = OracleSceneCornell(q.ctypes.data_as(fp), s.ctypes.data_as(fp), m.ctypes.data_as(fp))
    return sc, (q, s, m)


def make_params(profile, width, height, ntx, nty, bounces, env=None, env_kind=ENV_NONE, env_sampler=SAMPLER_POINT, v4_flags=0):
    p = OracleParams()
    p.v4_flags = int(v4_flags)
    p.profile, p.width, p.height = profile, width, height
    p.num_tiles_x, p.num_tiles_y, p.num_bounces = ntx, nty, bounces
    p.env_kind, p.env_sampler = env_kind, env_sampler
    keep = None
    if env is not None:
        keep = np.ascontiguousarray(env, dtype=np.float32)
        assert keep.ndim == 3 and keep.shape[2] == 3
        p.env = keep.ctypes.data_as(ctypes.POINTER(ctypes.c_float))
        p.env_height, p.env_width = keep.shape[0], keep.shape[1]
    return p, keep


def render(profile, width, height, ntx, nty, bounces, nframes, first_frame=1, env=None, env_kind=ENV_NONE,
           env_sampler=SAMPLER_POINT, target=None, nthreads=0, scene_v4=None, scene_cornell=None, v4_flags=0):
    """Returns (tile-major f32 buffer, counters dict).  scene_v4: result of make_scene_v4 (V4 profile);
    scene_cornell: result of make_scene_cornell (V2 / SIMT_TEXTURED profiles)."""
    p, keep = make_params(profile, width, height, ntx, nty, bounces, env, env_kind, env_sampler)
    if scene_v4 is not None:
        p.scene_v4 = ctypes.pointer(scene_v4[0])
    if scene_cornell is not None:
        p.scene_cornell = ctypes.pointer(scene_cornell[0])
    p.v4_flags = int(v4_flags)
    if target is None:
        target = np.zeros(width * height * 3, dtype=np.float32)
    else:
        target = np.ascontiguousarray(target, dtype=np.float32).copy()
    cnt = OracleCounters()
    rc = lib().oracle_render(ctypes.byref(p), target.ctypes.data_as(ctypes.POINTER(ctypes.c_float)), first_frame,
                             nframes, nthreads, ctypes.byref(cnt))
    if rc != 0:
        raise ValueError("oracle_render: invalid parameters")
    del keep
    return target, {"paths": cnt.paths, "segments": cnt.segments, "escapes": cnt.escapes}


def max_segments(profile, width, height, bounces, nframes, first_frame=1, env=None, env_kind=ENV_NONE,
                 env_sampler=SAMPLER_POINT, scene_v4=None, scene_cornell=None):
    """(H, W) uint32: max traced segments per pixel over the frame range."""
    p, keep = make_params(profile, width, height, 1, 1, bounces, env, env_kind, env_sampler)
    if scene_v4 is not None:
        p.scene_v4 = ctypes.pointer(scene_v4[0])
    if scene_cornell is not None:
        p.scene_cornell = ctypes.pointer(scene_cornell[0])
    out = np.zeros(width * height, dtype=np.uint32)
    L = lib()
    L.oracle_max_segments.argtypes = [ctypes.POINTER(OracleParams), ctypes.c_int, ctypes.c_int, ctypes.POINTER(ctypes.c_uint32)]
    if L.oracle_max_segments(ctypes.byref(p), first_frame, nframes, out.ctypes.data_as(ctypes.POINTER(ctypes.c_uint32))) != 0:
        raise ValueError("oracle_max_segments: invalid parameters")
    return out.reshape(height, width)


def resolve_ldr(target, width, height, ntx, nty, mode=0):
    t = np.ascontiguousarray(target, dtype=np.float32)
    out = np.zeros(width * height, dtype=np.uint32)
    rc = lib().oracle_resolve_ldr(t.ctypes.data_as(ctypes.POINTER(ctypes.c_float)), width, height, ntx, nty,
                                  out.ctypes.data_as(ctypes.POINTER(ctypes.c_uint32)), mode)
    if rc != 0:
        raise ValueError("oracle_resolve_ldr: invalid parameters")
    return out.reshape(height, width)


def detile(buf, width, height, ntx, nty):
    """tile-major SoA8 accumulation buffer -> (H, W, 3) row-major image (row 0 = top)."""
    tw, th = width // ntx, height // nty
    a = np.asarray(buf, dtype=np.float32).reshape(nty, ntx, th, tw // 8, 3, 8)
    return np.ascontiguousarray(a.transpose(0, 2, 1, 3, 5, 4).reshape(height, width, 3))


def tile(img, ntx, nty):
    """(H, W, 3) row-major image -> tile-major SoA8 buffer."""
    h, w, _ = img.shape
    tw, th = w // ntx, h // nty
    a = np.asarray(img, dtype=np.float32).reshape(nty, th, ntx, tw // 8, 8, 3)
    return np.ascontiguousarray(a.transpose(0, 2, 1, 3, 5, 4)).reshape(-1)


# ---- reference binaries ---------------------------------------------------------------------
def ref_binary(name):
    path = os.path.join(REF_DIR, name)
    return path if os.path.exists(path) else None


def run_ref(name, width, height, ntx, nty, frames, bounces=None, env=None, threads=None, start_frame=0,
            target=None, time_it=False, warmup=0, ldr=False, timeout=3600):
    """Runs oracle/_ref/<name>; returns dict(buffer=..., timing=..., ldr=...)."""
    exe = ref_binary(name)
    if exe is None:
        raise FileNotFoundError(f"{name} not built (oracle/ref_build/build_ref.sh needs /root/reference)")
    with tempfile.TemporaryDirectory() as td:
        out = os.path.join(td, "out.f32")
        cmd = [exe, "--w", str(width), "--h", str(height), "--ntx", str(ntx), "--nty", str(nty), "--frames",
               str(frames), "--out", out, "--start-frame", str(start_frame), "--warmup", str(warmup)]
        if bounces is not None:
            cmd += ["--bounces", str(bounces)]
        if threads is not None:
            cmd += ["--threads", str(threads)]
        if env is not None:
            e = np.ascontiguousarray(env, dtype=np.float32)
            ep = os.path.join(td, "env.f32")
            e.tofile(ep)
            cmd += ["--env", ep, "--envw", str(e.shape[1]), "--envh", str(e.shape[0])]
        if target is not None:
            ip = os.path.join(td, "in.f32")
            np.ascontiguousarray(target, dtype=np.float32).tofile(ip)
            cmd += ["--in", ip]
        if time_it:
            cmd += ["--time"]
        lp = os.path.join(td, "ldr.u32")
        if ldr:
            cmd += ["--ldr", lp]
        r = subprocess.run(cmd, capture_output=True, text=True, timeout=timeout)
        if r.returncode != 0:
            raise RuntimeError(f"{name} failed: {r.stderr}")
        res = {"buffer": np.fromfile(out, dtype=np.float32)}
        if time_it:
            res["timing"] = json.loads(r.stdout.strip().splitlines()[-1])
        if ldr:
            res["ldr"] = np.fromfile(lp, dtype=np.uint32).reshape(height, width)
        return res


def synthetic_env(width, height):
    """Deterministic synthetic equirect/cubemap-atlas texture:
    texel(r,c,ch) = 0.05 + 4*u^2 with u = Randomf3201(wang_hash stream seeded 1|((r*W+c)*3+ch)),
    plus a 'sun' disc of value 50 at uv (0.25, 0.75), radius 0.02 (in uv units)."""
    idx = (np.arange(width * height * 3, dtype=np.uint64)).astype(np.uint32) | np.uint32(1)
    s = idx.copy()
    s = (s ^ np.uint32(61)) ^ (s >> np.uint32(16))
    s = s * np.uint32(9)
    s = s ^ (s >> np.uint32(4))
    s = s * np.uint32(0x27D4EB2D)
    s = s ^ (s >> np.uint32(15))
    u = (s & np.uint32(0x7FFFFFFF)).astype(np.int32).astype(np.float32) / np.float32(2147483648.0)
    tex = (np.float32(0.05) + np.float32(4.0) * u * u).astype(np.float32).reshape(height, width, 3)
    rr, cc = np.meshgrid(np.arange(height), np.arange(width), indexing="ij")
    du = (cc + 0.5) / width - 0.25
    dv = (rr + 0.5) / height - 0.75
    tex[(du * du + dv * dv) < 0.02 * 0.02] = np.float32(50.0)
    return tex
