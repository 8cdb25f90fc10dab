"""prof_any.py -- one 1080p workload for ncu.  usage: prof_any.py <v2|v4_equirect|v4_cubemap|simt|v3redo> [spp] [reps]
B200PT_SCHEDULER=lane|sorted selects the kernel."""
import sys
sys.path.insert(0, '.')
from cpuperformanceraytracer_b200 import api
from oracle import pyoracle as po   # synthetic env generator only

name = sys.argv[1] if len(sys.argv) > 1 else "v2"
spp = int(sys.argv[2]) if len(sys.argv) > 2 else 128
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 5
kw, env = {
    "v2": (dict(profile=api.PROFILE_V2, num_bounces=8), None),
    "v4_equirect": (dict(profile=api.PROFILE_OPT_V4, num_bounces=8, env_kind=api.ENV_EQUIRECT, env_sampler=api.SAMPLER_RANDOM), (2048, 1024)),
    "v4_cubemap": (dict(profile=api.PROFILE_OPT_V4, num_bounces=8, env_kind=api.ENV_CUBEMAP, env_sampler=api.SAMPLER_RANDOM), (512, 3072)),
    "simt": (dict(profile=api.PROFILE_SIMT_TEXTURED, num_bounces=4), (2048, 1024)),
    "v3redo": (dict(profile=api.PROFILE_V3_REDO, num_bounces=8), (2048, 1024)),
}[name]
r = api.Renderer(math_mode=api.MATH_PARITY, **kw)
if env:
    r.set_env(po.synthetic_env(*env))
r.resize(1920, 1080, 10, 15)
for i in range(reps):
    r.reset(); r.render_frames(spp); c = r.counters()
    print(f"{name} 1080p spp={spp}: {c['last_render_ms']:.3f} ms -> {1920*1080*spp/c['last_render_ms']/1e3:.1f} Mpaths/s", flush=True)
r.close()
