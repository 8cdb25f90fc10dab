"""1080p x 256 spp timing of the four parity kernels (tuning probe)"""
import sys
sys.path.insert(0, '.')
from cpuperformanceraytracer_b200 import api
from oracle import pyoracle as po
env = po.synthetic_env(2048, 1024)
for name, prof, kw in (("v2", api.PROFILE_V2, {}), ("simt_textured", api.PROFILE_SIMT_TEXTURED, {}), ("v3_redo", api.PROFILE_V3_REDO, {}),
                       ("v4_equirect_random", api.PROFILE_OPT_V4, dict(env_kind=api.ENV_EQUIRECT, env_sampler=api.SAMPLER_RANDOM))):
    r = api.Renderer(profile=prof, math_mode=api.MATH_PARITY, num_bounces=8, **kw)
    if prof != api.PROFILE_V2: r.set_env(env)
    r.resize(1920, 1080, 10, 15); r.render_frames(8)
    best = 1e9
    for i in range(3):
        r.reset(); r.render_frames(256); best = min(best, r.counters()['last_render_ms'])
    print(f"{name} 1080p 256 spp: {best:.3f} ms -> {1920*1080*256/best/1e3:.1f} Mpaths/s", flush=True)
    r.close()
