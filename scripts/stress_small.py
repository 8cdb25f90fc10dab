"""small stress workload (compute-sanitizer is closed on the GPU pool, so this runs plain): both schedulers, all profiles, order table, tile stride,
groups with every combine mode, present paths"""
import sys
import numpy as np
sys.path.insert(0, '.')
from cpuperformanceraytracer_b200 import api
from oracle import pyoracle as po

W, H, NTX, NTY = 512, 288, 4, 6   # 4608 items: enough for the pull-order table
env = po.synthetic_env(128, 64)
cube = po.synthetic_env(32, 192)
cases = [dict(profile=api.PROFILE_V2, num_bounces=8), dict(profile=api.PROFILE_SIMT_TEXTURED, num_bounces=4),
         dict(profile=api.PROFILE_OPT_V4, num_bounces=8, env_kind=api.ENV_EQUIRECT, env_sampler=api.SAMPLER_RANDOM),
         dict(profile=api.PROFILE_OPT_V4, num_bounces=8, env_kind=api.ENV_CUBEMAP, env_sampler=api.SAMPLER_BILINEAR),
         dict(profile=api.PROFILE_V3_REDO, num_bounces=8), dict(profile=api.PROFILE_V3_REDO_SCENE0, num_bounces=8)]
for kw in cases:
    ref = None
    for sched in (api.SCHED_LANE, api.SCHED_SORTED):
        with api.Renderer(scheduler=sched, **kw) as r:
            if kw["profile"] != api.PROFILE_V2:
                r.set_env(cube if kw.get("env_kind") == api.ENV_CUBEMAP else env)
            r.resize(W, H, NTX, NTY)
            r.render_frames(2)
            r.render_frames(8)            # builds the order table
            r.set_tile_stride(1, 3)
            r.render_frames(8)
            r.set_tile_stride(0, 0)
            img = r.download_target()
        if ref is None:
            ref = img
        assert np.array_equal(ref, img)
    print("ok", kw["profile"], kw.get("env_kind"), flush=True)
for sharding, combine in ((api.SHARD_TILES, api.COMBINE_PEER), (api.SHARD_SPP, api.COMBINE_PEER), (api.SHARD_SPP, api.COMBINE_FUSED)):
    for sched in (api.SCHED_LANE, api.SCHED_SORTED):
        with api.Group([0, 0, 0], sharding=sharding, combine=combine, profile=api.PROFILE_V2, num_bounces=8, scheduler=sched) as g:
            g.resize(W, H, NTX, NTY)
            g.render_frames(9)
            g.set_bands(2)
            g.render_frames(9)
            g.download_target()
    print("ok group", sharding, combine, flush=True)
with api.Renderer(profile=api.PROFILE_V2, num_bounces=8, output_to_screen=True) as r:
    r.resize(W, H, NTX, NTY)
    frame = np.zeros((H, W), dtype=np.uint32)
    r.present_blocking(frame, 1, 2)
    r.present_submit(1); r.present_submit(1); r.present_acquire(); r.present_acquire()
    r.resolve_ldr()
    print("peak", r.measure_fp32_peak() > 1.0)
print("sanitize workload done")
