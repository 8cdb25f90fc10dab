#!/usr/bin/env python
"""fast_math_rmse.py -- B200PT_MATH_FAST against the oracle for every profile: RMSE / max-abs / mean shift at 256x192,
64 and 1024 spp, on the per-texel-noise synthetic env (worst case for point / jitter samplers: a 1-ulp direction
change can pick a neighbouring texel of unrelated value) and on a smooth env (what an HDR photograph looks like
at texel scale).  One JSON line per case; feeds the tolerances in tests/test_gpu_parity.py."""
import json, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from cpuperformanceraytracer_b200 import api
from oracle import pyoracle as po


def smooth_env(w, h):
    y, x = np.meshgrid(np.linspace(0, 1, h, dtype=np.float32), np.linspace(0, 1, w, dtype=np.float32), indexing="ij")
    e = np.stack([0.6 + 0.5 * np.sin(6.283 * x) * y, 0.5 + 0.4 * np.cos(6.283 * 2 * x), 0.3 + 1.5 * y * y], axis=2)
    e[(x - 0.25) ** 2 + (y - 0.75) ** 2 < 0.02 ** 2] = 50.0  # a sun
    return e.astype(np.float32)


W, H, NTX, NTY = 256, 192, 4, 6
CASES = [("v2", 0, None, 0, 0, 8), ("simt_textured", 1, (256, 128), 1, 0, 4), ("v4_equirect_random", 2, (256, 128), 1, 2, 8),
         ("v4_equirect_bilinear", 2, (256, 128), 1, 1, 8), ("v4_cubemap_random", 2, (64, 384), 2, 2, 8),
         ("v4_cubemap_bilinear", 2, (64, 384), 2, 1, 8), ("v4_no_env", 2, None, 0, 0, 8), ("v3_redo", 3, (256, 128), 1, 1, 8)]
GPU_PROFILE = {0: api.PROFILE_V2, 1: api.PROFILE_SIMT_TEXTURED, 2: api.PROFILE_OPT_V4, 3: api.PROFILE_V3_REDO}
for name, prof, shape, ek, es, bounces in CASES:
    for envname in (("noise", "smooth") if shape else ("none",)):
        env = None if not shape else (po.synthetic_env(*shape) if envname == "noise" else smooth_env(*shape))
        for spp in (64, 1024):
            o, _ = po.render(prof, W, H, NTX, NTY, bounces, spp, env=env, env_kind=ek, env_sampler=es, nthreads=os.cpu_count() or 8)
            kw = dict(env_kind=ek, env_sampler=es) if prof == 2 else {}
            with api.Renderer(profile=GPU_PROFILE[prof], math_mode=api.MATH_FAST, num_bounces=bounces, **kw) as r:
                if env is not None:
                    r.set_env(env)
                r.resize(W, H, NTX, NTY)
                r.render_frames(spp)
                g = r.download_target()
            d = g.astype(np.float64) - o
            print(json.dumps({"profile": name, "env": envname, "spp": spp, "rmse": float(np.sqrt((d * d).mean())), "max_abs": float(np.abs(d).max()),
                              "rel_mean_shift": float((g.astype(np.float64).mean() - o.mean()) / o.mean()),
                              "floats_identical": float((g == o).mean()), "image_mean": float(o.mean())}), flush=True)
