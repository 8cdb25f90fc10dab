#!/usr/bin/env python
"""A/B of the two schedulers (B200PT_SCHED_LANE vs B200PT_SCHED_SORTED) on the 1080p jobs: device time per launch
(the library's own CUDA events), best of `reps`, one JSON line per (profile, spp, scheduler)."""
import json, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from cpuperformanceraytracer_b200 import api
from oracle import pyoracle as po   # synthetic env generator only

W, H, NTX, NTY = 1920, 1080, 10, 15
reps = int(os.environ.get("REPS", "3"))
spps = [int(x) for x in os.environ.get("SPPS", "128,1024").split(",")]
which = os.environ.get("PROFILES", "v2,v4_equirect,v4_cubemap,simt,v3redo").split(",")
equi = po.synthetic_env(2048, 1024)
cube = po.synthetic_env(512, 3072)
PROFILES = {
    "v2": (dict(profile=api.PROFILE_V2, num_bounces=8), None),
    "v4_equirect": (dict(profile=api.PROFILE_OPT_V4, num_bounces=8, env_kind=api.ENV_EQUIRECT, env_sampler=api.SAMPLER_RANDOM), equi),
    "v4_cubemap": (dict(profile=api.PROFILE_OPT_V4, num_bounces=8, env_kind=api.ENV_CUBEMAP, env_sampler=api.SAMPLER_RANDOM), cube),
    "simt": (dict(profile=api.PROFILE_SIMT_TEXTURED, num_bounces=4), equi),
    "v3redo": (dict(profile=api.PROFILE_V3_REDO, num_bounces=8), equi),
}
math = api.MATH_FAST if os.environ.get("MATH") == "fast" else api.MATH_PARITY
for name in which:
    kw, env = PROFILES[name]
    ref = {}
    for spp in spps:
        for sched, sname in ((api.SCHED_LANE, "lane"), (api.SCHED_SORTED, "sorted")):
            with api.Renderer(scheduler=sched, math_mode=math, **kw) as r:
                if env is not None:
                    r.set_env(env)
                r.resize(W, H, NTX, NTY)
                r.render_frames(8)
                best = 1e30
                for _ in range(reps):
                    r.reset()
                    r.render_frames(spp)
                    c = r.counters()
                    best = min(best, c["last_render_ms"])
                img = r.download_target()
            key = (name, spp)
            same = None
            if sname == "lane":
                ref[key] = img
            else:
                same = bool(np.array_equal(ref[key], img))
            print(json.dumps({"profile": name, "spp": spp, "scheduler": sname, "ms": best,
                              "gpaths_per_s": W * H * spp / best * 1e-6, "identical_to_lane": same}), flush=True)
