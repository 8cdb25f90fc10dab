#!/usr/bin/env python
"""primitive_sweep.py -- SURVEY.md 8(f) rank 4: the run-time-scene kernel of the OPT_V4 profile (b200pt_set_scene_v4, the
generic kernel that loops over the scene table) as the primitive count grows from 4 to 12 objects (MAX_OBJECTS,
demofox_path_tracing_optimization_v4.cpp:327): the reference's 4 quads plus 0..8 of its spheres ((-18 + 6 i, -8, 10),
radius 2.8, its Fresnel / refraction materials).  1920x1080, tiles 10x15, 8 bounces, equirect env with the random-jitter
sampler, 128 spp per launch.  Per point: device ms, Gpaths/s, traced segments per path, and the fraction of the FP32
roofline with the algorithmic flop figure of SURVEY.md 8(d) scaled to the scene:
    F_seg = 50 nq + 38 ns + 219,  + 25 + 9 per path, + 80 per escaped path;   peak = 148 x 128 x 2 x 1965 MHz.
The first line is the built-in scene on the scene-specialised kernel (compile-time tables) for comparison."""
import json, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from cpuperformanceraytracer_b200 import api
from oracle import pyoracle as po   # synthetic env + the reference's camera distance
from scene_fixtures import default_v4_scene

W, H, NTX, NTY, SPP = 1920, 1080, 10, 15, int(os.environ.get("SPP", "128"))
PEAK = 148 * 128 * 2 * 1965e6
env = po.synthetic_env(2048, 1024)
q, s7, m = default_v4_scene()
cam_dist = float(po.lib().oracle_camera_distance())
f = np.float32


def measure(r, nq, ns, label):
    r.render_frames(4); r.render_frames(4)
    best, c0 = 1e30, None
    for _ in range(3):
        r.reset()
        c0 = r.counters()
        r.render_frames(SPP)
        c1 = r.counters()
        best = min(best, c1["last_render_ms"])
    paths = c1["paths"] - c0["paths"]
    segs, esc, cull = c1["segments"] - c0["segments"], c1["escapes"] - c0["escapes"], c1["culled_segments"] - c0["culled_segments"]
    fseg = 50 * nq + 38 * ns + 219
    flops = segs * fseg + paths * 34 + esc * 80
    flops_traced = (segs - cull) * fseg + paths * 34 + esc * 80
    print(json.dumps({"scene": label, "quads": nq, "spheres": ns, "objects": nq + ns, "spp": SPP, "ms": best,
                      "gpaths_per_s": paths / best * 1e-6, "segments_per_path": segs / paths, "culled_segment_share": cull / segs,
                      "F_seg": fseg, "roofline_frac": flops / (best * 1e-3) / PEAK, "roofline_frac_traced_only": flops_traced / (best * 1e-3) / PEAK}),
          flush=True)


kw = dict(profile=api.PROFILE_OPT_V4, num_bounces=8, env_kind=api.ENV_EQUIRECT, env_sampler=api.SAMPLER_RANDOM)
with api.Renderer(**kw) as r:
    r.set_env(env)
    r.resize(W, H, NTX, NTY)
    measure(r, 4, 7, "built-in scene, scene-specialised kernel")
for ns in range(0, 9):
    spheres = np.array([[f(-18.0) + f(6.0) * f(i), -8.0, 10.0, 2.8] for i in range(ns)], dtype=np.float32).reshape(-1, 4)
    mats = np.zeros((4 + ns, 17), dtype=np.float32)
    mats[:4] = m[:4]
    for i in range(ns):
        mats[4 + i] = m[4 + min(i, 6)]
    with api.Renderer(**kw) as r:
        r.set_env(env)
        r.set_scene_v4(q, spheres, mats, (0.0, 0.0, 40.0), cam_dist)
        r.resize(W, H, NTX, NTY)
        measure(r, 4, ns, "run-time scene table, generic kernel")
