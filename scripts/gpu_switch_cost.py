#!/usr/bin/env python
"""Cost of the reference's non-default compile-time switches (global_preprocessor_flags.h:64-65, b200pt_params.exact_exp /
sincos_unit_vectors) on BASELINE config 3 (P_v4 + equirect, 1080p, 8 bounces, 256 spp): device ms per launch, both math modes.
Scene-specialised kernels carry the switches as compile-time values (pt_kernels_*_v4sw.cu); the generic (table-driven) kernel
reads them at run time."""
import json, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from cpuperformanceraytracer_b200 import api
from oracle import pyoracle as po   # synthetic env generator only

W, H, NTX, NTY, SPP = 1920, 1080, 10, 15, int(os.environ.get("SPP", "256"))
env = po.synthetic_env(2048, 1024)
CASES = [("default (static kernel)", {}), ("exact_exp", dict(exact_exp=True)), ("sincos_unit_vectors", dict(sincos_unit_vectors=True)),
         ("both", dict(exact_exp=True, sincos_unit_vectors=True)), ("generic tables", dict(generic_scene_tables=True)),
         ("generic tables + both", dict(generic_scene_tables=True, exact_exp=True, sincos_unit_vectors=True))]
for mm, mname in ((api.MATH_PARITY, "parity"), (api.MATH_FAST, "fast")):
    for name, kw in CASES:
        with api.Renderer(profile=api.PROFILE_OPT_V4, math_mode=mm, num_bounces=8, env_kind=api.ENV_EQUIRECT, env_sampler=api.SAMPLER_RANDOM, **kw) as r:
            r.set_env(env)
            r.resize(W, H, NTX, NTY)
            r.render_frames(8); r.render_frames(8)
            ms = []
            for _ in range(3):
                r.reset()
                r.render_frames(SPP)
                ms.append(r.counters()["last_render_ms"])
            c = r.counters()
        print(json.dumps({"math": mname, "case": name, "spp": SPP, "ms_min": min(ms), "ms_median": float(np.median(ms)),
                          "gpaths_per_s": W * H * SPP / min(ms) * 1e-6, "mean_segments_per_path": c["segments"] / max(c["paths"], 1)}), flush=True)
