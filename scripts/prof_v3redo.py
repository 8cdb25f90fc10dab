"""prof_v3redo.py -- P_v3redo workload for ncu: 1080p, tiles 10x15, 8 bounces, synthetic 2048x1024 equirect env"""
import sys
sys.path.insert(0, '.')
from cpuperformanceraytracer_b200 import api
from oracle import pyoracle as po
spp = int(sys.argv[1]) if len(sys.argv) > 1 else 128
r = api.Renderer(profile=api.PROFILE_V3_REDO, math_mode=api.MATH_PARITY, num_bounces=8)
r.set_env(po.synthetic_env(2048, 1024))
r.resize(1920, 1080, 10, 15)
for i in range(5):
    r.reset(); r.render_frames(spp); c = r.counters()
    print(f"v3_redo 1080p spp={spp}: {c['last_render_ms']:.3f} ms -> {1920*1080*spp/c['last_render_ms']/1e3:.1f} Mpaths/s", flush=True)
r.close()
