"""api.py -- Python binding (ctypes) of the C ABI in include/b200pt.h.

The product is libb200pt.so (hand-written sm_100a kernels behind a C ABI); this module is only
the thin host-side handle used by the tests, bench.py and the multi-GPU driver.  It never renders
on the CPU: importing it without the built library, or creating a Renderer without a B200,
raises.
"""
import ctypes
import os

import numpy as np

PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(PKG, "libb200pt.so")

PROFILE_V2, PROFILE_SIMT_TEXTURED, PROFILE_OPT_V4, PROFILE_V3_REDO, PROFILE_V3_REDO_SCENE0 = 0, 1, 2, 3, 4
MATH_PARITY, MATH_FAST = 0, 1
ENV_NONE, ENV_EQUIRECT, ENV_CUBEMAP = 0, 1, 2
SAMPLER_POINT, SAMPLER_BILINEAR, SAMPLER_RANDOM = 0, 1, 2
ACCUM_RUNNING_AVERAGE, ACCUM_SUM = 0, 1
LDR_FILE_RGBA, LDR_SCREEN_BGRA, LDR_EXACT_ACES, LDR_EXACT_GAMMA = 0, 1, 2, 4
TONEMAP_EXACT_ACES, TONEMAP_EXACT_GAMMA = 1, 2
SCHED_DEFAULT, SCHED_LANE, SCHED_SORTED = 0, 1, 2

# every symbol include/b200pt.h declares (tests/test_abi.py checks the library exports them all)
ABI_SYMBOLS = [
    "b200pt_api_version", "b200pt_error_string", "b200pt_last_error", "b200pt_default_params", "b200pt_create",
    "b200pt_destroy", "b200pt_set_env", "b200pt_resize", "b200pt_reset", "b200pt_set_frame_counter",
    "b200pt_get_frame_counter", "b200pt_render_frames", "b200pt_synchronize", "b200pt_upload_target",
    "b200pt_download_target", "b200pt_render_host", "b200pt_resolve_ldr", "b200pt_bind_device_target",
    "b200pt_get_device_target", "b200pt_set_stream", "b200pt_finalize_sum", "b200pt_download_rng_state",
    "b200pt_get_counters", "b200pt_compute_cull_rects", "b200pt_set_tile_row_range", "b200pt_set_tile_range", "b200pt_set_tile_stride", "b200pt_present_submit", "b200pt_present_acquire", "b200pt_present_blocking", "b200pt_set_scene_v4", "b200pt_compute_cull_rects_scene_v4", "b200pt_set_scene_cornell", "b200pt_compute_cull_rects_scene_cornell",
    "b200pt_eval_portable", "b200pt_check_portable_tiers", "b200pt_static_tables_match", "b200pt_measure_fp32_peak", "b200pt_scale_target", "b200pt_scale_target_span",
    "b200pt_group_create", "b200pt_group_destroy", "b200pt_group_size", "b200pt_group_context", "b200pt_group_set_env",
    "b200pt_group_resize", "b200pt_group_reset", "b200pt_group_set_bands", "b200pt_group_set_frame_counter", "b200pt_group_get_frame_counter",
    "b200pt_group_render_frames", "b200pt_group_synchronize", "b200pt_group_upload_target", "b200pt_group_download_target",
    "b200pt_group_render_host", "b200pt_group_resolve_ldr", "b200pt_group_get_counters", "b200pt_group_last_error",
]
SHARD_SPP, SHARD_TILES = 0, 1
COMBINE_NCCL, COMBINE_PEER, COMBINE_FUSED = 0, 1, 2
FN_SIN, FN_COS, FN_ATAN2, FN_ASIN, FN_EXP, FN_SQRT, FN_RCP, FN_DIV, FN_EQUIRECT_TEXEL, FN_POW = 0, 1, 2, 3, 4, 5, 6, 7, 8, 9


class Texture(ctypes.Structure):
    """struct texture, texture.h:6-12"""
    _fields_ = [("Data", ctypes.POINTER(ctypes.c_float)), ("Width", ctypes.c_int32), ("Height", ctypes.c_int32),
                ("Components", ctypes.c_int32)]


class Params(ctypes.Structure):
    _fields_ = [("struct_size", ctypes.c_int32), ("device", ctypes.c_int32), ("profile", ctypes.c_int32),
                ("math_mode", ctypes.c_int32), ("num_bounces", ctypes.c_int32), ("env_kind", ctypes.c_int32),
                ("env_sampler", ctypes.c_int32), ("accum_mode", ctypes.c_int32), ("output_to_screen", ctypes.c_int32),
                ("disable_camera_culling", ctypes.c_int32), ("generic_scene_tables", ctypes.c_int32),
                ("scheduler", ctypes.c_int32), ("disable_item_order", ctypes.c_int32), ("exact_exp", ctypes.c_int32),
                ("sincos_unit_vectors", ctypes.c_int32), ("exact_tonemap", ctypes.c_int32)]


class Counters(ctypes.Structure):
    _fields_ = [("paths", ctypes.c_uint64), ("segments", ctypes.c_uint64), ("escapes", ctypes.c_uint64),
                ("launches", ctypes.c_uint64), ("last_render_ms", ctypes.c_double), ("culled_segments", ctypes.c_uint64)]


class B200PTError(RuntimeError):
    pass


_lib = None


def load_library():
    """Loads libb200pt.so; raises if it has not been built (there is no fallback path)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise B200PTError(f"{LIB_PATH} is missing: run `python -m cpuperformanceraytracer_b200.build` "
                          "(the CUDA extension is the only implementation)")
    L = ctypes.CDLL(LIB_PATH)
    vp, i32 = ctypes.c_void_p, ctypes.c_int32
    L.b200pt_api_version.restype = ctypes.c_int
    L.b200pt_error_string.restype = ctypes.c_char_p
    L.b200pt_error_string.argtypes = [ctypes.c_int]
    L.b200pt_last_error.restype = ctypes.c_char_p
    L.b200pt_last_error.argtypes = [vp]
    L.b200pt_default_params.argtypes = [ctypes.c_int, ctypes.POINTER(Params)]
    L.b200pt_create.argtypes = [ctypes.POINTER(Params), ctypes.POINTER(vp)]
    L.b200pt_destroy.argtypes = [vp]
    L.b200pt_set_env.argtypes = [vp, Texture]
    L.b200pt_resize.argtypes = [vp, i32, i32, i32, i32]
    L.b200pt_reset.argtypes = [vp]
    L.b200pt_set_frame_counter.argtypes = [vp, i32]
    L.b200pt_get_frame_counter.argtypes = [vp, ctypes.POINTER(i32)]
    L.b200pt_render_frames.argtypes = [vp, i32]
    L.b200pt_synchronize.argtypes = [vp]
    L.b200pt_upload_target.argtypes = [vp, ctypes.POINTER(ctypes.c_float)]
    L.b200pt_download_target.argtypes = [vp, ctypes.POINTER(ctypes.c_float)]
    L.b200pt_render_host.argtypes = [vp, ctypes.POINTER(ctypes.c_float), i32, i32, i32, i32, i32, i32, i32, Texture, vp, i32]
    L.b200pt_resolve_ldr.argtypes = [vp, ctypes.POINTER(ctypes.c_uint32), i32, i32]
    L.b200pt_bind_device_target.argtypes = [vp, vp]
    L.b200pt_get_device_target.argtypes = [vp, ctypes.POINTER(vp), ctypes.POINTER(ctypes.c_size_t)]
    L.b200pt_set_stream.argtypes = [vp, vp]
    L.b200pt_finalize_sum.argtypes = [vp, i32]
    L.b200pt_set_tile_row_range.argtypes = [vp, i32, i32]
    L.b200pt_set_tile_range.argtypes = [vp, i32, i32]
    L.b200pt_set_tile_stride.argtypes = [vp, i32, i32]
    L.b200pt_set_scene_v4.argtypes = [vp, vp, i32, vp, i32, vp, vp]
    L.b200pt_set_scene_cornell.argtypes = [vp, vp, vp, vp]
    L.b200pt_compute_cull_rects_scene_cornell.argtypes = [vp, vp, i32, i32, ctypes.POINTER(ctypes.c_float), ctypes.POINTER(i32)]
    L.b200pt_present_submit.argtypes = [vp, i32]
    L.b200pt_present_acquire.argtypes = [vp, ctypes.POINTER(ctypes.POINTER(ctypes.c_uint32)), ctypes.POINTER(i32)]
    L.b200pt_present_blocking.argtypes = [vp, i32, ctypes.POINTER(ctypes.c_uint32), i32]
    L.b200pt_download_rng_state.argtypes = [vp, ctypes.POINTER(ctypes.c_uint32)]
    L.b200pt_get_counters.argtypes = [vp, ctypes.POINTER(Counters)]
    fpp = ctypes.POINTER(ctypes.c_float)
    L.b200pt_eval_portable.argtypes = [vp, ctypes.c_int, fpp, fpp, fpp, ctypes.c_size_t]
    L.b200pt_check_portable_tiers.argtypes = [vp, ctypes.c_int, ctypes.c_uint64, ctypes.c_uint64, ctypes.POINTER(ctypes.c_uint64),
                                              ctypes.POINTER(ctypes.c_uint64)]
    L.b200pt_compute_cull_rects.argtypes = [ctypes.c_int, i32, i32, ctypes.POINTER(ctypes.c_float), ctypes.POINTER(i32)]
    L.b200pt_scale_target.argtypes = [vp, ctypes.c_float]
    L.b200pt_measure_fp32_peak.argtypes = [vp, ctypes.POINTER(ctypes.c_double)]
    L.b200pt_scale_target_span.argtypes = [vp, ctypes.c_size_t, ctypes.c_size_t, ctypes.c_float, vp]
    # several GPUs of one process
    L.b200pt_group_create.argtypes = [ctypes.POINTER(Params), ctypes.POINTER(i32), i32, i32, i32, ctypes.POINTER(vp)]
    L.b200pt_group_destroy.argtypes = [vp]
    L.b200pt_group_size.argtypes = [vp]
    L.b200pt_group_context.argtypes = [vp, i32]
    L.b200pt_group_context.restype = vp
    L.b200pt_group_set_env.argtypes = [vp, Texture]
    L.b200pt_group_resize.argtypes = [vp, i32, i32, i32, i32]
    L.b200pt_group_reset.argtypes = [vp]
    L.b200pt_group_set_bands.argtypes = [vp, i32]
    L.b200pt_group_set_frame_counter.argtypes = [vp, i32]
    L.b200pt_group_get_frame_counter.argtypes = [vp, ctypes.POINTER(i32)]
    L.b200pt_group_render_frames.argtypes = [vp, i32]
    L.b200pt_group_synchronize.argtypes = [vp]
    L.b200pt_group_upload_target.argtypes = [vp, ctypes.POINTER(ctypes.c_float)]
    L.b200pt_group_download_target.argtypes = [vp, ctypes.POINTER(ctypes.c_float)]
    L.b200pt_group_render_host.argtypes = [vp, ctypes.POINTER(ctypes.c_float), i32, i32, i32, i32, i32, i32, i32, Texture, vp, i32]
    L.b200pt_group_resolve_ldr.argtypes = [vp, ctypes.POINTER(ctypes.c_uint32), i32, i32]
    L.b200pt_group_get_counters.argtypes = [vp, ctypes.POINTER(Counters), ctypes.POINTER(ctypes.c_double)]
    L.b200pt_group_last_error.argtypes = [vp]
    L.b200pt_group_last_error.restype = ctypes.c_char_p
    _lib = L
    return L


def default_params(profile):
    p = Params()
    rc = load_library().b200pt_default_params(profile, ctypes.byref(p))
    if rc != 0:
        raise B200PTError(f"b200pt_default_params: {load_library().b200pt_error_string(rc).decode()}")
    return p


def _fptr(a):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_float))


class Renderer:
    """One rendering context on one GPU = the reference's static render state (frame counter, scene,
    tile table) plus the HBM-resident accumulation buffer."""

    def __init__(self, profile=PROFILE_V2, math_mode=MATH_PARITY, num_bounces=-1, device=0, env_kind=None,
                 env_sampler=None, accum_mode=ACCUM_RUNNING_AVERAGE, output_to_screen=False,
                 disable_camera_culling=False, generic_scene_tables=False, scheduler=SCHED_DEFAULT, disable_item_order=False,
                 exact_exp=False, sincos_unit_vectors=False, exact_aces_tonemap=False, exact_gamma=False):
        self._lib = load_library()
        self._ctx = ctypes.c_void_p()
        p = default_params(profile)
        p.math_mode, p.num_bounces, p.device = math_mode, num_bounces, device
        p.accum_mode, p.output_to_screen = accum_mode, int(bool(output_to_screen))
        p.disable_camera_culling = int(bool(disable_camera_culling))
        p.generic_scene_tables = int(bool(generic_scene_tables))
        p.scheduler = int(scheduler)
        p.disable_item_order = int(bool(disable_item_order))
        # global_preprocessor_flags.h:63-65, non-default side (0 = the reference's checked-in "fast" variants)
        p.exact_exp, p.sincos_unit_vectors = int(bool(exact_exp)), int(bool(sincos_unit_vectors))
        p.exact_tonemap = (TONEMAP_EXACT_ACES if exact_aces_tonemap else 0) | (TONEMAP_EXACT_GAMMA if exact_gamma else 0)
        if env_kind is not None:
            p.env_kind = env_kind
        if env_sampler is not None:
            p.env_sampler = env_sampler
        self.params = p
        rc = self._lib.b200pt_create(ctypes.byref(p), ctypes.byref(self._ctx))
        if rc != 0:
            self._ctx = ctypes.c_void_p()
            raise B200PTError(f"b200pt_create failed: {self._lib.b200pt_error_string(rc).decode()} "
                              "(a B200 is required; there is no CPU fallback)")
        self.width = self.height = self.ntx = self.nty = 0
        self._env_keep = None

    # -- plumbing ------------------------------------------------------------------------------
    def _check(self, rc, what):
        if rc != 0:
            raise B200PTError(f"{what}: {self._lib.b200pt_error_string(rc).decode()}: "
                              f"{self._lib.b200pt_last_error(self._ctx).decode()}")

    def close(self):
        if self._ctx:
            self._lib.b200pt_destroy(self._ctx)
            self._ctx = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    # -- reference-shaped operations ---------------------------------------------------------------
    def set_env(self, env):
        """env: (H, W, 3) float32, row 0 = bottom (what LoadTexture / LoadCubemapTexture return)."""
        e = np.ascontiguousarray(env, dtype=np.float32)
        if e.ndim != 3 or e.shape[2] != 3:
            raise ValueError("env must be (H, W, 3)")
        self._env_keep = e
        t = Texture(_fptr(e), e.shape[1], e.shape[0], 3)
        self._check(self._lib.b200pt_set_env(self._ctx, t), "b200pt_set_env")

    def set_scene_v4(self, quads=None, spheres=None, materials=None, camera_position=(0.0, 0.0, 40.0), camera_distance=1.0):
        """quads: (nq, 4, 3) vertices; spheres: (ns, 4) xyz+radius; materials: (nq+ns, 17) in SceneMaterial order
        (albedo3, emissive3, specularChance, specularRoughness, specularColor3, IOR, refractionChance,
        refractionRoughness, refractionColor3).  No arguments: back to the built-in scene."""
        q = np.ascontiguousarray(quads if quads is not None else np.zeros((0, 4, 3)), dtype=np.float32).reshape(-1, 12)
        s = np.ascontiguousarray(spheres if spheres is not None else np.zeros((0, 4)), dtype=np.float32).reshape(-1, 4)
        m = np.ascontiguousarray(materials if materials is not None else np.zeros((0, 17)), dtype=np.float32).reshape(-1, 17)
        cam = np.array(list(camera_position) + [camera_distance], dtype=np.float32)
        if m.shape[0] != q.shape[0] + s.shape[0]:
            raise ValueError("one material per object")
        rc = self._lib.b200pt_set_scene_v4(self._ctx, q.ctypes.data_as(ctypes.c_void_p), q.shape[0], s.ctypes.data_as(ctypes.c_void_p),
                                           s.shape[0], m.ctypes.data_as(ctypes.c_void_p), cam.ctypes.data_as(ctypes.c_void_p))
        self._check(rc, "b200pt_set_scene_v4")

    def set_scene_cornell(self, quads=None, spheres=None, materials=None):
        """Cornell-family profiles: quads (6, 4, 3) vertices, spheres (3, 4) xyz + radius, materials (9, 11) in
        b200pt_material_legacy order.  No arguments: back to the reference's box."""
        if quads is None:
            self._check(self._lib.b200pt_set_scene_cornell(self._ctx, None, None, None), "b200pt_set_scene_cornell")
            return
        q = np.ascontiguousarray(quads, dtype=np.float32).reshape(6, 12)
        s = np.ascontiguousarray(spheres, dtype=np.float32).reshape(3, 4)
        m = np.ascontiguousarray(materials, dtype=np.float32).reshape(9, 11)
        vp = ctypes.c_void_p
        rc = self._lib.b200pt_set_scene_cornell(self._ctx, q.ctypes.data_as(vp), s.ctypes.data_as(vp), m.ctypes.data_as(vp))
        self._check(rc, "b200pt_set_scene_cornell")

    def resize(self, width, height, ntx, nty):
        self._check(self._lib.b200pt_resize(self._ctx, width, height, ntx, nty), "b200pt_resize")
        self.width, self.height, self.ntx, self.nty = width, height, ntx, nty

    def reset(self):
        self._check(self._lib.b200pt_reset(self._ctx), "b200pt_reset")

    @property
    def frame_counter(self):
        v = ctypes.c_int32()
        self._check(self._lib.b200pt_get_frame_counter(self._ctx, ctypes.byref(v)), "b200pt_get_frame_counter")
        return v.value

    @frame_counter.setter
    def frame_counter(self, v):
        self._check(self._lib.b200pt_set_frame_counter(self._ctx, int(v)), "b200pt_set_frame_counter")

    def render_frames(self, nframes, sync=True):
        self._check(self._lib.b200pt_render_frames(self._ctx, int(nframes)), "b200pt_render_frames")
        if sync:
            self.synchronize()

    def synchronize(self):
        self._check(self._lib.b200pt_synchronize(self._ctx), "b200pt_synchronize")

    def upload_target(self, buf):
        b = np.ascontiguousarray(buf, dtype=np.float32).reshape(-1)
        assert b.size == self.width * self.height * 3
        self._check(self._lib.b200pt_upload_target(self._ctx, _fptr(b)), "b200pt_upload_target")

    def download_target(self):
        out = np.empty(self.width * self.height * 3, dtype=np.float32)
        self._check(self._lib.b200pt_download_target(self._ctx, _fptr(out)), "b200pt_download_target")
        return out

    def render_host(self, buffer_out, width, height, ntx, nty, nframes, env=None, screen=None):
        """The reference-facing call (DemofoxRenderOptV4 signature + frame count) on HOST buffers."""
        assert buffer_out.dtype == np.float32 and buffer_out.flags["C_CONTIGUOUS"]
        if env is not None:
            e = env if (env.dtype == np.float32 and env.flags["C_CONTIGUOUS"]) else np.ascontiguousarray(env, np.float32)
            self._env_keep = e
            t = Texture(_fptr(e), e.shape[1], e.shape[0], 3)
        else:
            t = Texture(None, 0, 0, 3)
        sp = screen.ctypes.data_as(ctypes.c_void_p) if screen is not None else None
        rc = self._lib.b200pt_render_host(self._ctx, _fptr(buffer_out), width, height, ntx, nty, width // ntx,
                                          height // nty, 3, t, sp, int(nframes))
        self._check(rc, "b200pt_render_host")
        self.width, self.height, self.ntx, self.nty = width, height, ntx, nty

    def resolve_ldr(self, mode=LDR_FILE_RGBA, bump_frame_counter=False, out=None):
        """tone-mapped u32 frame; `out` = a caller-owned (H, W) uint32 array to fill (page-locked memory is copied
        into directly at PCIe speed), default a fresh array"""
        if out is None:
            out = np.empty((self.height, self.width), dtype=np.uint32)
        assert out.dtype == np.uint32 and out.flags["C_CONTIGUOUS"] and out.size == self.height * self.width
        rc = self._lib.b200pt_resolve_ldr(self._ctx, out.ctypes.data_as(ctypes.POINTER(ctypes.c_uint32)), mode,
                                          int(bump_frame_counter))
        self._check(rc, "b200pt_resolve_ldr")
        return out

    def present_submit(self, nframes=1):
        self._check(self._lib.b200pt_present_submit(self._ctx, int(nframes)), "b200pt_present_submit")

    def present_acquire(self, copy=True):
        """Oldest in-flight frame as an (H, W) uint32 array (a view of pinned memory unless copy) + its iFrame."""
        ptr, fr = ctypes.POINTER(ctypes.c_uint32)(), ctypes.c_int32()
        self._check(self._lib.b200pt_present_acquire(self._ctx, ctypes.byref(ptr), ctypes.byref(fr)), "b200pt_present_acquire")
        a = np.ctypeslib.as_array(ptr, shape=(self.height, self.width))
        return (a.copy() if copy else a), fr.value

    def present_blocking(self, out, nframes=1, bands=0):
        """nframes render calls + fused tone map + band-pipelined copy: `out` ((H, W) uint32, ideally page-locked)
        holds the screen frame when the call returns"""
        assert out.dtype == np.uint32 and out.flags["C_CONTIGUOUS"] and out.size == self.height * self.width
        rc = self._lib.b200pt_present_blocking(self._ctx, int(nframes), out.ctypes.data_as(ctypes.POINTER(ctypes.c_uint32)), int(bands))
        self._check(rc, "b200pt_present_blocking")
        return out

    def rng_state(self):
        out = np.empty((self.height, self.width), dtype=np.uint32)
        rc = self._lib.b200pt_download_rng_state(self._ctx, out.ctypes.data_as(ctypes.POINTER(ctypes.c_uint32)))
        self._check(rc, "b200pt_download_rng_state")
        return out

    def eval_portable(self, fn, a, b=None):
        """parity-mode transcendental `fn` evaluated on the device (b200pt_eval_portable)"""
        a = np.ascontiguousarray(a, dtype=np.float32)
        out = np.empty_like(a)
        fp = ctypes.POINTER(ctypes.c_float)
        bb = None if b is None else np.ascontiguousarray(b, dtype=np.float32)
        rc = self._lib.b200pt_eval_portable(self._ctx, int(fn), a.ctypes.data_as(fp), None if bb is None else bb.ctypes.data_as(fp),
                                            out.ctypes.data_as(fp), ctypes.c_size_t(a.size))
        self._check(rc, "b200pt_eval_portable")
        return out

    def check_portable_tiers(self, fn, first, count):
        """(mismatches, inputs that took the literal path) of b200pt_check_portable_tiers"""
        bad, lit = ctypes.c_uint64(0), ctypes.c_uint64(0)
        rc = self._lib.b200pt_check_portable_tiers(self._ctx, int(fn), ctypes.c_uint64(first), ctypes.c_uint64(count),
                                                   ctypes.byref(bad), ctypes.byref(lit))
        self._check(rc, "b200pt_check_portable_tiers")
        return bad.value, lit.value

    def measure_fp32_peak(self):
        """FFMA micro-benchmark on this GPU: TFLOP/s (FFMA = 2 flop)"""
        t = ctypes.c_double()
        self._check(self._lib.b200pt_measure_fp32_peak(self._ctx, ctypes.byref(t)), "b200pt_measure_fp32_peak")
        return t.value

    def counters(self):
        c = Counters()
        self._check(self._lib.b200pt_get_counters(self._ctx, ctypes.byref(c)), "b200pt_get_counters")
        return {"paths": c.paths, "segments": c.segments, "escapes": c.escapes, "launches": c.launches,
                "last_render_ms": c.last_render_ms, "culled_segments": c.culled_segments}

    # -- multi-GPU plumbing ------------------------------------------------------------------------
    def bind_device_target(self, device_ptr):
        self._check(self._lib.b200pt_bind_device_target(self._ctx, ctypes.c_void_p(device_ptr)), "b200pt_bind_device_target")

    def device_target(self):
        p, n = ctypes.c_void_p(), ctypes.c_size_t()
        self._check(self._lib.b200pt_get_device_target(self._ctx, ctypes.byref(p), ctypes.byref(n)), "b200pt_get_device_target")
        return p.value, n.value

    def set_stream(self, cuda_stream_ptr):
        self._check(self._lib.b200pt_set_stream(self._ctx, ctypes.c_void_p(cuda_stream_ptr)), "b200pt_set_stream")

    def set_tile_row_range(self, first_tile_row, num_tile_rows):
        self._check(self._lib.b200pt_set_tile_row_range(self._ctx, int(first_tile_row), int(num_tile_rows)),
                    "b200pt_set_tile_row_range")

    def set_tile_range(self, first_flat_tile, num_tiles):
        self._check(self._lib.b200pt_set_tile_range(self._ctx, int(first_flat_tile), int(num_tiles)), "b200pt_set_tile_range")

    def set_tile_stride(self, remainder, modulus):
        self._check(self._lib.b200pt_set_tile_stride(self._ctx, int(remainder), int(modulus)), "b200pt_set_tile_stride")

    def finalize_sum(self, total_frames):
        self._check(self._lib.b200pt_finalize_sum(self._ctx, int(total_frames)), "b200pt_finalize_sum")

    def scale_target(self, factor):
        self._check(self._lib.b200pt_scale_target(self._ctx, ctypes.c_float(factor)), "b200pt_scale_target")

    def scale_target_span(self, float_offset, float_count, factor, cuda_stream_ptr=None):
        rc = self._lib.b200pt_scale_target_span(self._ctx, int(float_offset), int(float_count), ctypes.c_float(factor),
                                                ctypes.c_void_p(cuda_stream_ptr))
        self._check(rc, "b200pt_scale_target_span")


class Group:
    """Several GPUs of this process behind one set of render entry points (b200pt_group_*): frames (SHARD_SPP) or
    tiles (SHARD_TILES) of every render call are sharded over `devices`; the image lives on devices[0]."""

    def __init__(self, devices, sharding=SHARD_SPP, combine=COMBINE_NCCL, profile=PROFILE_V2, math_mode=MATH_PARITY,
                 num_bounces=-1, env_kind=None, env_sampler=None, output_to_screen=False, scheduler=SCHED_DEFAULT,
                 exact_exp=False, sincos_unit_vectors=False, exact_aces_tonemap=False, exact_gamma=False):
        self._lib = load_library()
        self._g = ctypes.c_void_p()
        p = default_params(profile)
        p.math_mode, p.num_bounces = math_mode, num_bounces
        p.exact_exp, p.sincos_unit_vectors = int(bool(exact_exp)), int(bool(sincos_unit_vectors))
        p.exact_tonemap = (TONEMAP_EXACT_ACES if exact_aces_tonemap else 0) | (TONEMAP_EXACT_GAMMA if exact_gamma else 0)
        p.scheduler = int(scheduler)
        p.output_to_screen = int(bool(output_to_screen))
        if env_kind is not None:
            p.env_kind = env_kind
        if env_sampler is not None:
            p.env_sampler = env_sampler
        devs = (ctypes.c_int32 * len(devices))(*devices)
        rc = self._lib.b200pt_group_create(ctypes.byref(p), devs, len(devices), sharding, combine, ctypes.byref(self._g))
        if rc != 0:
            self._g = ctypes.c_void_p()
            raise B200PTError(f"b200pt_group_create failed: {self._lib.b200pt_error_string(rc).decode()} "
                              "(B200s are required; there is no CPU fallback)")
        self.size = len(devices)
        self.width = self.height = self.ntx = self.nty = 0
        self._env_keep = None

    def _check(self, rc, what):
        if rc != 0:
            raise B200PTError(f"{what}: {self._lib.b200pt_error_string(rc).decode()}: "
                              f"{self._lib.b200pt_group_last_error(self._g).decode()}")

    def close(self):
        if self._g:
            self._lib.b200pt_group_destroy(self._g)
            self._g = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def set_env(self, env):
        e = np.ascontiguousarray(env, dtype=np.float32)
        self._env_keep = e
        self._check(self._lib.b200pt_group_set_env(self._g, Texture(_fptr(e), e.shape[1], e.shape[0], 3)), "b200pt_group_set_env")

    def resize(self, width, height, ntx, nty):
        self._check(self._lib.b200pt_group_resize(self._g, width, height, ntx, nty), "b200pt_group_resize")
        self.width, self.height, self.ntx, self.nty = width, height, ntx, nty

    def reset(self):
        self._check(self._lib.b200pt_group_reset(self._g), "b200pt_group_reset")

    def set_bands(self, bands):
        self._check(self._lib.b200pt_group_set_bands(self._g, int(bands)), "b200pt_group_set_bands")

    @property
    def frame_counter(self):
        v = ctypes.c_int32()
        self._check(self._lib.b200pt_group_get_frame_counter(self._g, ctypes.byref(v)), "b200pt_group_get_frame_counter")
        return v.value

    @frame_counter.setter
    def frame_counter(self, v):
        self._check(self._lib.b200pt_group_set_frame_counter(self._g, int(v)), "b200pt_group_set_frame_counter")

    def render_frames(self, nframes, sync=True):
        self._check(self._lib.b200pt_group_render_frames(self._g, int(nframes)), "b200pt_group_render_frames")
        if sync:
            self.synchronize()

    def synchronize(self):
        self._check(self._lib.b200pt_group_synchronize(self._g), "b200pt_group_synchronize")

    def upload_target(self, buf):
        b = np.ascontiguousarray(buf, dtype=np.float32).reshape(-1)
        assert b.size == self.width * self.height * 3
        self._check(self._lib.b200pt_group_upload_target(self._g, _fptr(b)), "b200pt_group_upload_target")

    def download_target(self):
        out = np.empty(self.width * self.height * 3, dtype=np.float32)
        self._check(self._lib.b200pt_group_download_target(self._g, _fptr(out)), "b200pt_group_download_target")
        return out

    def render_host(self, buffer_out, width, height, ntx, nty, nframes, env=None, screen=None):
        assert buffer_out.dtype == np.float32 and buffer_out.flags["C_CONTIGUOUS"]
        if env is not None:
            e = env if (env.dtype == np.float32 and env.flags["C_CONTIGUOUS"]) else np.ascontiguousarray(env, np.float32)
            self._env_keep = e
            t = Texture(_fptr(e), e.shape[1], e.shape[0], 3)
        else:
            t = Texture(None, 0, 0, 3)
        sp = screen.ctypes.data_as(ctypes.c_void_p) if screen is not None else None
        rc = self._lib.b200pt_group_render_host(self._g, _fptr(buffer_out), width, height, ntx, nty, width // ntx,
                                                height // nty, 3, t, sp, int(nframes))
        self._check(rc, "b200pt_group_render_host")
        self.width, self.height, self.ntx, self.nty = width, height, ntx, nty

    def resolve_ldr(self, mode=LDR_FILE_RGBA, bump_frame_counter=False):
        out = np.empty((self.height, self.width), dtype=np.uint32)
        rc = self._lib.b200pt_group_resolve_ldr(self._g, out.ctypes.data_as(ctypes.POINTER(ctypes.c_uint32)), mode,
                                                int(bump_frame_counter))
        self._check(rc, "b200pt_group_resolve_ldr")
        return out

    def counters(self):
        c, ms = Counters(), ctypes.c_double()
        self._check(self._lib.b200pt_group_get_counters(self._g, ctypes.byref(c), ctypes.byref(ms)), "b200pt_group_get_counters")
        return {"paths": c.paths, "segments": c.segments, "escapes": c.escapes, "launches": c.launches,
                "last_render_ms": c.last_render_ms, "culled_segments": c.culled_segments, "combine_ms": ms.value}


def cull_rects(profile, width, height):
    """Host-only: the conservative fragCoord-space rectangles used for camera-ray culling, or None."""
    r = (ctypes.c_float * 48)()
    n = ctypes.c_int32()
    rc = load_library().b200pt_compute_cull_rects(profile, width, height, r, ctypes.byref(n))
    if rc != 0:
        raise B200PTError("b200pt_compute_cull_rects: invalid argument")
    if n.value < 0:
        return None
    return np.array(r[:4 * n.value], dtype=np.float32).reshape(n.value, 4)


def cull_rects_scene_v4(quads, spheres, camera_position, camera_distance, width, height):
    """Host-only: culling rectangles of a run-time OPT_V4 scene, or None when culling is impossible."""
    L = load_library()
    vp, i32 = ctypes.c_void_p, ctypes.c_int32
    L.b200pt_compute_cull_rects_scene_v4.argtypes = [vp, i32, vp, i32, vp, i32, i32, ctypes.POINTER(ctypes.c_float), ctypes.POINTER(i32)]
    q = np.ascontiguousarray(quads, dtype=np.float32).reshape(-1, 12)
    s = np.ascontiguousarray(spheres, dtype=np.float32).reshape(-1, 4)
    cam = np.array(list(camera_position) + [camera_distance], dtype=np.float32)
    r = (ctypes.c_float * 48)()
    n = ctypes.c_int32()
    rc = L.b200pt_compute_cull_rects_scene_v4(q.ctypes.data_as(vp), q.shape[0], s.ctypes.data_as(vp), s.shape[0],
                                              cam.ctypes.data_as(vp), width, height, r, ctypes.byref(n))
    if rc != 0:
        raise B200PTError("b200pt_compute_cull_rects_scene_v4: invalid argument")
    return None if n.value < 0 else np.array(r[:4 * n.value], dtype=np.float32).reshape(n.value, 4)


def cull_rects_scene_cornell(quads, spheres, width, height):
    """Host-only: culling rectangles of a run-time Cornell-family scene, or None when culling is impossible."""
    L = load_library()
    q = np.ascontiguousarray(quads, dtype=np.float32).reshape(6, 12)
    s = np.ascontiguousarray(spheres, dtype=np.float32).reshape(3, 4)
    r = (ctypes.c_float * 48)()
    n = ctypes.c_int32()
    vp = ctypes.c_void_p
    rc = L.b200pt_compute_cull_rects_scene_cornell(q.ctypes.data_as(vp), s.ctypes.data_as(vp), width, height, r, ctypes.byref(n))
    if rc != 0:
        raise B200PTError("b200pt_compute_cull_rects_scene_cornell: invalid argument")
    return None if n.value < 0 else np.array(r[:4 * n.value], dtype=np.float32).reshape(n.value, 4)


def detile(buf, width, height, ntx, nty):
    """tile-major SoA8 accumulation buffer (RenderTile, v4.cpp:1189-1252) -> (H, W, 3) image."""
    tw, th = width // ntx, height // nty
    a = np.asarray(buf, dtype=np.float32).reshape(nty, ntx, th, tw // 8, 3, 8)
    return np.ascontiguousarray(a.transpose(0, 2, 1, 3, 5, 4).reshape(height, width, 3))
