"""tail/load-balance probe: same scene and aspect, growing image -> more work items per warp"""
import sys
sys.path.insert(0, '.')
from cpuperformanceraytracer_b200 import api
for (W, H, ntx, nty, spp) in ((960, 540, 10, 15, 512), (1920, 1080, 10, 15, 128), (1920, 1080, 10, 15, 512), (3840, 2160, 10, 15, 128), (7680, 4320, 20, 30, 32)):
    r = api.Renderer(profile=api.PROFILE_V2, math_mode=api.MATH_PARITY, num_bounces=8)
    r.resize(W, H, ntx, nty)
    r.render_frames(8)
    best = 1e9
    for i in range(3):
        r.reset(); r.render_frames(spp); c = r.counters(); best = min(best, c['last_render_ms'])
    print(f"{W}x{H} spp={spp}: {best:.3f} ms -> {W*H*spp/best/1e3:.1f} Mpaths/s", flush=True)
    r.close()
