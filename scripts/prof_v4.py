"""prof_v4.py -- P_v4 workload for ncu: 1080p, tiles 10x15, 8 bounces, synthetic 2048x1024 equirect env
(random-jitter sampler, the reference's default flags) or cubemap.  usage: prof_v4.py [equirect|cubemap] [spp]"""
import sys
sys.path.insert(0, '.')
from cpuperformanceraytracer_b200 import api
from oracle import pyoracle as po

kind = sys.argv[1] if len(sys.argv) > 1 else "equirect"
spp = int(sys.argv[2]) if len(sys.argv) > 2 else 128
if kind == "cubemap":
    env, ek = po.synthetic_env(512, 3072), api.ENV_CUBEMAP
else:
    env, ek = po.synthetic_env(2048, 1024), api.ENV_EQUIRECT
r = api.Renderer(profile=api.PROFILE_OPT_V4, math_mode=api.MATH_PARITY, num_bounces=8, env_kind=ek,
                 env_sampler=api.SAMPLER_RANDOM)
r.set_env(env)
r.resize(1920, 1080, 10, 15)
for i in range(5):
    r.reset(); r.render_frames(spp); c = r.counters()
    print(f"v4 {kind} 1080p spp={spp}: {c['last_render_ms']:.3f} ms -> {1920*1080*spp/c['last_render_ms']/1e3:.1f} Mpaths/s "
          f"counters {c}", flush=True)
r.close()
