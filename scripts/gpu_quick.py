"""Quick GPU check: bit-exactness of the v2/simt/v4 parity kernels vs the oracle + 1080p timing."""
import sys
import numpy as np
sys.path.insert(0, '.')
from cpuperformanceraytracer_b200 import api
from oracle import pyoracle as po

env = po.synthetic_env(256, 128)
for name, gp, op, e, ek, es, b in [("v2", api.PROFILE_V2, po.PROFILE_V2, None, 0, 0, 8),
                                   ("simt", api.PROFILE_SIMT_TEXTURED, po.PROFILE_SIMT_TEXTURED, env, 1, 0, 4),
                                   ("v4", api.PROFILE_OPT_V4, po.PROFILE_V4, env, 1, 2, 8)]:
    W, H, F = 256, 192, 12
    o, oc = po.render(op, W, H, 4, 6, b, F, env=e, env_kind=ek, env_sampler=es)
    for mode in (api.MATH_PARITY, api.MATH_FAST):
        kw = dict(env_kind=ek, env_sampler=es) if gp == api.PROFILE_OPT_V4 else {}
        r = api.Renderer(profile=gp, math_mode=mode, num_bounces=b, **kw)
        if e is not None: r.set_env(e)
        r.resize(W, H, 4, 6); r.render_frames(F); g = r.download_target(); c = r.counters()
        d = np.abs(g.astype(np.float64) - o)
        print(f"{name} mode={mode}: identical={np.array_equal(g, o)} rmse={np.sqrt((d**2).mean()):.3e} seg {c['segments']} vs {oc['segments']}", flush=True)
        r.close()
for prof, pname in ((api.PROFILE_V2, "v2"), (api.PROFILE_OPT_V4, "v4")):
    for mode, mname in ((api.MATH_PARITY, "parity"), (api.MATH_FAST, "fast")):
        r = api.Renderer(profile=prof, math_mode=mode, num_bounces=8)
        if prof == api.PROFILE_OPT_V4: r.set_env(env)
        r.resize(1920, 1080, 10, 15); r.render_frames(16)
        for n in (1, 256):
            r.reset(); r.render_frames(n); c = r.counters()
            print(f"{pname} 1080p {mname} nframes={n}: {c['last_render_ms']:.3f} ms -> {1920*1080*n/c['last_render_ms']/1e3:.1f} Mpaths/s", flush=True)
        r.close()
