#!/usr/bin/env python
"""full_job_parity.py -- BASELINE.json configs[1] at its FULL size (Cornell P_v2, 1920x1080, 8 bounces, 1024 spp = 2.1e9 paths):
the GPU's f32 accumulation buffer against (a) the reference's own code, `oracle/_ref/ref_v2_exact` (reference sources compiled
in place, exact reciprocals; its v2 renderer takes at most 8 tiles = 8 threads, tiles 2x4), and (b) the C restatement
`oracle/pt_oracle.c`, both run on the host cores of the GPU box while the GPU result waits.  Bit for bit, whole image.

The parity tests in tests/ compare 1080p images at 6-16 frames and smaller images at more frames so that the suite runs in
minutes; this script is the same comparison on the headline job itself (several CPU-minutes).  One JSON line.
--profile: any renderer / sampler the reference build exists for (v2, simt, v3redo, v3redo0, v4, v4_bilinear, v4_cubemap,
v4_cubemap_bilinear); the v4 ones run the reference on all cores, the others on the 8 threads their tile table allows.
usage: full_job_parity.py [--profile P] [--spp 1024] [--width W --height H --bounces B] [--skip-oracle]"""
import argparse
import hashlib
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from cpuperformanceraytracer_b200 import api  # noqa: E402
from oracle import pyoracle as po  # noqa: E402  (the checker)

ap = argparse.ArgumentParser()
ap.add_argument("--spp", type=int, default=1024)
ap.add_argument("--skip-oracle", action="store_true")
# name -> (GPU profile kwargs, reference binary, its thread limit (None = all cores), oracle profile, oracle env kwargs, env shape, tiles)
EQ, CUBE = (2048, 1024), (512, 3072)
PROFILES = {
    "v2": (dict(profile=api.PROFILE_V2), "ref_v2_exact", 8, po.PROFILE_V2, {}, None, (2, 4)),
    "simt": (dict(profile=api.PROFILE_SIMT_TEXTURED), "ref_simt_textured_exact", 8, po.PROFILE_SIMT_TEXTURED, dict(env_kind=po.ENV_EQUIRECT), EQ, (2, 4)),
    "v3redo": (dict(profile=api.PROFILE_V3_REDO), "ref_v3redo_exact", 8, po.PROFILE_V3REDO,
               dict(env_kind=po.ENV_EQUIRECT, env_sampler=po.SAMPLER_BILINEAR), EQ, (2, 4)),
    "v3redo0": (dict(profile=api.PROFILE_V3_REDO_SCENE0), "ref_v3redo_scene0_exact", 8, po.PROFILE_V3REDO_SCENE0,
                dict(env_kind=po.ENV_EQUIRECT, env_sampler=po.SAMPLER_BILINEAR), EQ, (2, 4)),
    "v4": (dict(profile=api.PROFILE_OPT_V4, env_kind=api.ENV_EQUIRECT, env_sampler=api.SAMPLER_RANDOM), "ref_v4_equirect_random_exact", None,
           po.PROFILE_V4, dict(env_kind=po.ENV_EQUIRECT, env_sampler=po.SAMPLER_RANDOM), EQ, (10, 15)),
    "v4_bilinear": (dict(profile=api.PROFILE_OPT_V4, env_kind=api.ENV_EQUIRECT, env_sampler=api.SAMPLER_BILINEAR), "ref_v4_equirect_bilinear_exact", None,
                    po.PROFILE_V4, dict(env_kind=po.ENV_EQUIRECT, env_sampler=po.SAMPLER_BILINEAR), EQ, (10, 15)),
    "v4_cubemap": (dict(profile=api.PROFILE_OPT_V4, env_kind=api.ENV_CUBEMAP, env_sampler=api.SAMPLER_RANDOM), "ref_v4_cubemap_random_exact", None,
                   po.PROFILE_V4, dict(env_kind=po.ENV_CUBEMAP, env_sampler=po.SAMPLER_RANDOM), CUBE, (10, 15)),
    "v4_cubemap_bilinear": (dict(profile=api.PROFILE_OPT_V4, env_kind=api.ENV_CUBEMAP, env_sampler=api.SAMPLER_BILINEAR), "ref_v4_cubemap_bilinear_exact",
                            None, po.PROFILE_V4, dict(env_kind=po.ENV_CUBEMAP, env_sampler=po.SAMPLER_BILINEAR), CUBE, (10, 15)),
}
ap.add_argument("--profile", default="v2", choices=sorted(PROFILES))
ap.add_argument("--width", type=int, default=1920)   # BASELINE configs[4]: --width 8192 --height 8192 --bounces 16 (bounded --spp)
ap.add_argument("--height", type=int, default=1080)
ap.add_argument("--bounces", type=int, default=8)
a = ap.parse_args()
GKW, REF, REF_LIMIT, OPROFILE, OENV, ENV_SHAPE, (NTX, NTY) = PROFILES[a.profile]
V4 = ENV_SHAPE is not None  # "has an env map"
W, H, BOUNCES, SPP = a.width, a.height, a.bounces, a.spp
if H % NTY or (W // NTX) % 8 or W % NTX:  # sizes the 10x15 grid does not divide (8192^2, 3840x2160)
    NTX, NTY = 2, 4
ENV = po.synthetic_env(*ENV_SHAPE) if ENV_SHAPE else None
KW = dict(num_bounces=BOUNCES, **GKW)
REF_THREADS = REF_LIMIT or (os.cpu_count() or 16)
OKW = dict(env=ENV, **OENV) if ENV_SHAPE else {}

with api.Renderer(**KW) as r:
    if V4:
        r.set_env(ENV)
    r.resize(W, H, NTX, NTY)
    r.render_frames(SPP)
    gpu = r.download_target()
    c = r.counters()
    gpu_ms = c["last_render_ms"]
ALT = (2, 4) if (NTX, NTY) != (2, 4) else ((10, 15) if H % 15 == 0 and W % 10 == 0 and (W // 10) % 8 == 0 else (4, 8))
with api.Renderer(**KW) as r:  # another tiling: same pixels, other layout
    if V4:
        r.set_env(ENV)
    r.resize(W, H, *ALT)
    r.render_frames(SPP)
    gpu_alt = r.download_target()

res = {}


def run_reference():
    t0 = time.time()
    if po.ref_binary(REF) is None:
        res["ref"] = None
        return
    res["ref"] = po.run_ref(REF, W, H, NTX, NTY, SPP, bounces=BOUNCES, env=ENV, threads=REF_THREADS, timeout=7200)["buffer"]
    res["ref_s"] = time.time() - t0


def run_oracle():
    t0 = time.time()
    res["oracle"], res["oracle_counters"] = po.render(OPROFILE, W, H, NTX, NTY, BOUNCES, SPP, nthreads=max(1, (os.cpu_count() or 16) - 8), **OKW)
    res["oracle_s"] = time.time() - t0


threads = [threading.Thread(target=run_reference)]
if not a.skip_oracle:
    threads.append(threading.Thread(target=run_oracle))
for t in threads:
    t.start()
for t in threads:
    t.join()


def sha(b):
    return hashlib.sha256(np.ascontiguousarray(b).tobytes()).hexdigest()[:16]


JOB = f"profile {a.profile}, {W}x{H}, {BOUNCES} bounces, {SPP} spp, tiles {NTX}x{NTY}" + (
    f", synthetic {ENV_SHAPE[0]}x{ENV_SHAPE[1]} env" if ENV_SHAPE else "") + (" (the renderer's 8-tile limit)" if REF_LIMIT else "")
out = {"job": JOB, "paths": W * H * SPP, "gpu_kernel_ms": gpu_ms, "gpu_sha256_16": sha(gpu), "host_cores": os.cpu_count(),
       "gpu_tiles_%dx%d_same_pixels" % ALT: bool(np.array_equal(po.detile(gpu, W, H, NTX, NTY), po.detile(gpu_alt, W, H, *ALT)))}
if res.get("ref") is not None:
    out.update(reference="oracle/_ref/%s (%d threads)" % (REF, REF_THREADS), reference_seconds=res["ref_s"], reference_sha256_16=sha(res["ref"]),
               gpu_equals_reference_bit_for_bit=bool(np.array_equal(gpu, res["ref"])),
               max_abs_diff_vs_reference=float(np.abs(gpu.astype(np.float64) - res["ref"]).max()))
if "oracle" in res:
    oc = res["oracle_counters"]
    out.update(oracle_seconds=res["oracle_s"], oracle_sha256_16=sha(res["oracle"]), gpu_equals_oracle_bit_for_bit=bool(np.array_equal(gpu, res["oracle"])),
               counters_equal=bool((c["segments"], c["escapes"]) == (oc["segments"], oc["escapes"])),
               segments=oc["segments"], escapes=oc["escapes"])
print(json.dumps(out), flush=True)
