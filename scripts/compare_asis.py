#!/usr/bin/env python
"""compare_asis.py -- the CUDA path (parity mode) against the UNTOUCHED reference build (`asis`: hardware
rcpps / rsqrtps as mathlib.h:417,444 use them, glibc libm in place of MSVC SVML), statistically.

The bit-exact parity anchor is the `exact` build of the reference (1/x and 1/sqrt(x) exactly defined); the `asis`
build differs from it at approximation level (~2^-12 relative in every reciprocal), which decorrelates individual
paths: per-pixel differences are at Monte-Carlo noise level, while image statistics must agree.  This script
measures both (SURVEY.md section 7 hard part 1(a), BASELINE.md section 2 "sensitivity"):

  stats        512x512, 1024 spp, same seeds: RMSE, max-abs and the relative mean-brightness shift of the GPU image
               against ref_v2_asis and ref_v4_equirect_random_asis
  convergence  BASELINE config 4 (P_v4 + cubemap, progressive 1 spp per frame, Application.cpp:306-375): RMSE of the
               image after k = 1 .. 600 frames against a converged 16384-spp image, for the GPU and for the reference

One JSON line per measurement.  Needs a B200 and oracle/_ref (built where /root/reference exists)."""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from cpuperformanceraytracer_b200 import api  # noqa: E402
from oracle import pyoracle as po  # noqa: E402  (the checker: reference binaries + synthetic textures)


def image_stats(gpu, ref):
    g, r = gpu.astype(np.float64), ref.astype(np.float64)
    d = g - r
    return {"rmse": float(np.sqrt((d * d).mean())), "max_abs": float(np.abs(d).max()),
            "mean_gpu": float(g.mean()), "mean_ref": float(r.mean()), "rel_mean_shift": float((r.mean() - g.mean()) / g.mean())}


def stats(size=512, spp=1024):
    out = []
    threads = os.cpu_count() or 8
    # ---- Cornell P_v2, 8 bounces
    with api.Renderer(profile=api.PROFILE_V2, num_bounces=8) as r:
        r.resize(size, size, 2, 4)
        r.render_frames(spp)
        g = r.download_target()
        r.reset()
        r.render_frames(spp // 16)
        g_low = r.download_target()
    t0 = time.time()
    ref = po.run_ref("ref_v2_asis", size, size, 2, 4, spp, bounces=8, threads=8)["buffer"]
    d = image_stats(g, ref)
    d.update(what="stats", profile="v2", size=size, spp=spp, reference="ref_v2_asis (hardware rcpps/rsqrtps, libm)", ref_seconds=time.time() - t0,
             noise_floor_rmse_spp_over_16=image_stats(g_low, g)["rmse"])
    out.append(d)
    # ---- P_v4, equirect env, random-jitter sampler (the reference's checked-in flags)
    env = po.synthetic_env(1024, 512)
    with api.Renderer(profile=api.PROFILE_OPT_V4, num_bounces=8, env_kind=api.ENV_EQUIRECT, env_sampler=api.SAMPLER_RANDOM) as r:
        r.set_env(env)
        r.resize(size, size, 8, 8)
        r.render_frames(spp)
        g = r.download_target()
        r.reset()
        r.render_frames(spp // 16)
        g_low = r.download_target()
    t0 = time.time()
    ref = po.run_ref("ref_v4_equirect_random_asis", size, size, 8, 8, spp, bounces=8, env=env, threads=threads)["buffer"]
    d = image_stats(g, ref)
    d.update(what="stats", profile="v4_equirect_random", size=size, spp=spp, reference="ref_v4_equirect_random_asis", ref_seconds=time.time() - t0,
             noise_floor_rmse_spp_over_16=image_stats(g_low, g)["rmse"])
    out.append(d)
    return out


def convergence(size=512, frames=600, converged_spp=16384):
    """RMSE vs frame count, GPU and reference side by side, against the GPU's converged image"""
    cube = po.synthetic_env(256, 1536)
    ntx = nty = 8
    checkpoints = [k for k in (1, 2, 4, 8, 16, 32, 64, 128, 256, 512, frames) if k <= frames]
    checkpoints = sorted(set(checkpoints))
    kw = dict(profile=api.PROFILE_OPT_V4, num_bounces=8, env_kind=api.ENV_CUBEMAP, env_sampler=api.SAMPLER_RANDOM)
    # the converged image uses frames the progressive run never sees (the seeds depend on iFrame)
    with api.Renderer(accum_mode=api.ACCUM_SUM, **kw) as r:  # the plain mean of the same samples, for the comparison target
        r.set_env(cube)
        r.resize(size, size, ntx, nty)
        r.frame_counter = 100000
        r.render_frames(converged_spp)
        conv = r.download_target().astype(np.float64) / converged_spp
    curve_gpu, curve_ref = [], []
    with api.Renderer(**kw) as r:
        r.set_env(cube)
        r.resize(size, size, ntx, nty)
        done = 0
        for k in checkpoints:
            r.render_frames(k - done)
            done = k
            img = r.download_target().astype(np.float64) * (k + 1) / k  # undo the reference's 1/(N+1) bias for the RMSE
            curve_gpu.append(float(np.sqrt(((img - conv) ** 2).mean())))
    threads = os.cpu_count() or 8
    buf, done = None, 0
    t0 = time.time()
    for k in checkpoints:
        res = po.run_ref("ref_v4_cubemap_random_asis", size, size, ntx, nty, k - done, bounces=8, env=cube, threads=threads,
                         start_frame=done, target=buf)
        buf, done = res["buffer"], k
        img = buf.astype(np.float64) * (k + 1) / k
        curve_ref.append(float(np.sqrt(((img - conv) ** 2).mean())))
    return {"what": "convergence", "workload": f"P_v4 + cubemap 256x1536 atlas, {size}x{size}, 1 spp per frame, {frames} frames (BASELINE config 4 geometry reduced)",
            "converged": f"{converged_spp} spp (GPU, frames 100001..), plain mean", "frames": checkpoints, "rmse_gpu": curve_gpu,
            "rmse_reference_asis": curve_ref, "ref_seconds": time.time() - t0,
            "ratio_ref_over_gpu": [b / a for a, b in zip(curve_gpu, curve_ref)]}


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--size", type=int, default=512)
    ap.add_argument("--spp", type=int, default=1024)
    ap.add_argument("--frames", type=int, default=600)
    ap.add_argument("--skip-convergence", action="store_true")
    a = ap.parse_args()
    for line in stats(a.size, a.spp):
        print(json.dumps(line), flush=True)
    if not a.skip_convergence:
        print(json.dumps(convergence(a.size, a.frames)), flush=True)
