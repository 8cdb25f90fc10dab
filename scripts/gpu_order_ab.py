#!/usr/bin/env python
"""A/B of the work-item pull order (scene-first / sky-last vs buffer order) on the 1080p jobs: device ms per launch."""
import json, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from cpuperformanceraytracer_b200 import api
from oracle import pyoracle as po   # synthetic env generator only

W, H, NTX, NTY = 1920, 1080, 10, 15
reps = int(os.environ.get("REPS", "3"))
spps = [int(x) for x in os.environ.get("SPPS", "1,128,1024").split(",")]
which = os.environ.get("PROFILES", "v2,v4_equirect").split(",")
PROFILES = {
    "v2": (dict(profile=api.PROFILE_V2, num_bounces=8), None),
    "v4_equirect": (dict(profile=api.PROFILE_OPT_V4, num_bounces=8, env_kind=api.ENV_EQUIRECT, env_sampler=api.SAMPLER_RANDOM), (2048, 1024)),
    "v4_cubemap": (dict(profile=api.PROFILE_OPT_V4, num_bounces=8, env_kind=api.ENV_CUBEMAP, env_sampler=api.SAMPLER_RANDOM), (512, 3072)),
}
for name in which:
    kw, env = PROFILES[name]
    for spp in spps:
        for off in (True, False):
            with api.Renderer(disable_item_order=off, **kw) as r:
                if env:
                    r.set_env(po.synthetic_env(*env))
                r.resize(W, H, NTX, NTY)
                r.render_frames(2); r.render_frames(2)
                ms = []
                for _ in range(reps if spp >= 16 else 30):
                    r.reset()
                    r.render_frames(spp)
                    ms.append(r.counters()["last_render_ms"])
            print(json.dumps({"profile": name, "spp": spp, "item_order": "buffer" if off else "scene-first", "ms_min": min(ms),
                              "ms_median": float(np.median(ms)), "gpaths_per_s": W * H * spp / min(ms) * 1e-6}), flush=True)
