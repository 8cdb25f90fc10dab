"""First GPU contact: parity of every profile against the oracle + a timing probe."""
import sys, time, json
import numpy as np
sys.path.insert(0, '.')
from cpuperformanceraytracer_b200 import api
from oracle import pyoracle as po

def cmp(name, g, o):
    d = np.abs(g.astype(np.float64) - o.astype(np.float64))
    print(f"{name}: identical={np.array_equal(g, o)} frac_eq={(g == o).mean():.6f} maxabs={d.max():.3e} rmse={np.sqrt((d**2).mean()):.3e} mean={o.mean():.5f}", flush=True)

env = po.synthetic_env(256, 128)
cube = po.synthetic_env(64, 64 * 6)
cases = [
    ("v2", api.PROFILE_V2, po.PROFILE_V2, None, api.ENV_NONE, api.SAMPLER_POINT, 8),
    ("simt", api.PROFILE_SIMT_TEXTURED, po.PROFILE_SIMT_TEXTURED, env, api.ENV_EQUIRECT, api.SAMPLER_POINT, 4),
    ("v4_eq_rand", api.PROFILE_OPT_V4, po.PROFILE_V4, env, api.ENV_EQUIRECT, api.SAMPLER_RANDOM, 8),
    ("v4_eq_bil", api.PROFILE_OPT_V4, po.PROFILE_V4, env, api.ENV_EQUIRECT, api.SAMPLER_BILINEAR, 8),
    ("v4_cube_rand", api.PROFILE_OPT_V4, po.PROFILE_V4, cube, api.ENV_CUBEMAP, api.SAMPLER_RANDOM, 8),
    ("v4_cube_bil", api.PROFILE_OPT_V4, po.PROFILE_V4, cube, api.ENV_CUBEMAP, api.SAMPLER_BILINEAR, 8),
]
W, H, NTX, NTY, F = 256, 256, 2, 4, 16
for name, gp, op, e, ek, es, b in cases:
    o, oc = po.render(op, W, H, NTX, NTY, b, F, env=e, env_kind=ek, env_sampler=es)
    for mode, mname in ((api.MATH_PARITY, "parity"), (api.MATH_FAST, "fast")):
        r = api.Renderer(profile=gp, math_mode=mode, num_bounces=b, env_kind=ek if gp == api.PROFILE_OPT_V4 else None,
                         env_sampler=es if gp == api.PROFILE_OPT_V4 else None)
        if e is not None:
            r.set_env(e)
        r.resize(W, H, NTX, NTY)
        r.render_frames(F)
        g = r.download_target()
        c = r.counters()
        cmp(f"{name}/{mname}", g, o)
        if mode == api.MATH_PARITY:
            print("   counters gpu", c["segments"], c["escapes"], "oracle", oc["segments"], oc["escapes"], flush=True)
        r.close()

# timing probe: 1080p Cornell, 8 bounces
for mode, mname in ((api.MATH_PARITY, "parity"), (api.MATH_FAST, "fast")):
    r = api.Renderer(profile=api.PROFILE_V2, math_mode=mode, num_bounces=8)
    r.resize(1920, 1080, 10, 15)
    r.render_frames(16)
    for n in (64, 256):
        r.render_frames(n)
        c = r.counters()
        print(f"v2 1080p {mname} nframes={n}: {c['last_render_ms']:.2f} ms -> {1920*1080*n/c['last_render_ms']/1e3:.1f} Mpaths/s; seg/path={c['segments']/c['paths']:.3f}", flush=True)
    r.close()
r = api.Renderer(profile=api.PROFILE_OPT_V4, math_mode=api.MATH_FAST, num_bounces=8)
r.set_env(env); r.resize(1920, 1080, 10, 15); r.render_frames(16); r.render_frames(256)
c = r.counters(); print(f"v4 1080p fast nframes=256: {c['last_render_ms']:.2f} ms -> {1920*1080*256/c['last_render_ms']/1e3:.1f} Mpaths/s; seg/path={c['segments']/c['paths']:.3f}")
