#!/usr/bin/env python
"""bench_configs.py -- the BASELINE.json configurations other than the headline one (bounded samples),
one JSON line each.  Run under torchrun for N > 1:  python -m torch.distributed.run --nproc-per-node N ...
    --config 1   Cornell P_v2 512x512, 64 spp, 8 bounces: full-size bit-exact check vs the oracle + time
    --config 3   simt_textured (and P_v4 equirect) 3840x2160, spp-sharded, synthetic 2048x1024 env
    --config 4   P_v4 + cubemap, 1080p progressive 1 spp/frame: per-frame latency (render + tone map + D2H)
    --config 5   8192x8192, 16 bounces, strong scaling: spp-shard reduce vs tile-shard gather
"""
import argparse, json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from cpuperformanceraytracer_b200 import api, dist as ptdist

ap = argparse.ArgumentParser()
ap.add_argument("--config", type=int, required=True)
ap.add_argument("--spp", type=int, default=0)
ap.add_argument("--math", default="parity")
args = ap.parse_args()
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    import torch.distributed as tdist
    tdist.init_process_group("nccl", device_id=dev)
MATH = api.MATH_PARITY if args.math == "parity" else api.MATH_FAST


def out(d):
    if rank == 0:
        d.update(n_gpus=world, math=args.math)
        print(json.dumps(d), flush=True)


def sync():
    if world > 1:
        tdist.barrier()
    torch.cuda.synchronize(dev)


if args.config == 1:
    from oracle import pyoracle as po
    W = H = 512
    t0 = time.time(); o, oc = po.render(po.PROFILE_V2, W, H, 2, 4, 8, 64); tcpu = time.time() - t0
    with api.Renderer(profile=api.PROFILE_V2, math_mode=MATH, num_bounces=8) as r:
        r.resize(W, H, 2, 4); r.render_frames(64); r.reset(); r.render_frames(64)
        g = r.download_target(); c = r.counters(); rs = r.rng_state()
    d = np.abs(g.astype(np.float64) - o)
    out({"config": 1, "workload": "Cornell P_v2 512x512 tiles 2x4, 64 spp, 8 bounces", "gpu_ms": c["last_render_ms"],
         "mpaths_per_s": W * H * 64 / c["last_render_ms"] * 1e-3, "bit_exact_vs_oracle": bool(np.array_equal(g, o)),
         "rmse": float(np.sqrt((d ** 2).mean())), "max_abs": float(d.max()), "oracle_port_seconds": tcpu,
         "segments_match": bool(c["segments"] % (2 ** 64) >= oc["segments"])})
elif args.config == 3:
    from oracle import pyoracle as po
    W, H, ntx, nty = 3840, 2160, 10, 15
    spp = args.spp or 256
    env = po.synthetic_env(2048, 1024)
    for name, prof, kw in (("simt_textured", api.PROFILE_SIMT_TEXTURED, {}),
                           ("v4_equirect_random", api.PROFILE_OPT_V4, dict(env_kind=api.ENV_EQUIRECT, env_sampler=api.SAMPLER_RANDOM))):
        bounces = 4 if prof == api.PROFILE_SIMT_TEXTURED else 8
        def factory(accum_mode=api.ACCUM_RUNNING_AVERAGE, device=local):
            r = api.Renderer(profile=prof, math_mode=MATH, num_bounces=bounces, device=device, accum_mode=accum_mode, **kw)
            r.set_env(env)
            return r
        sr = ptdist.SppShardedRenderer(factory, W, H, ntx, nty, rank, world, local)
        sr.render(spp * world); sr.stream.synchronize(); sync()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with torch.cuda.stream(sr.stream):
            e0.record(sr.stream); sr.render(spp * world); e1.record(sr.stream)
        sync()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1: tdist.all_reduce(ms, op=tdist.ReduceOp.MAX)
        out({"config": 3, "profile": name, "workload": f"{W}x{H}, {spp} spp per GPU (bounded sample of 4096), synthetic 2048x1024 equirect env, spp-shard + all-reduce",
             "ms": float(ms), "mpaths_per_s": W * H * spp * world / float(ms) * 1e-3})
        sr.close()
elif args.config == 4:
    from oracle import pyoracle as po
    W, H, ntx, nty, frames = 1920, 1080, 10, 15, 600
    cube = po.synthetic_env(512, 3072)
    with api.Renderer(profile=api.PROFILE_OPT_V4, math_mode=MATH, num_bounces=8, env_kind=api.ENV_CUBEMAP,
                      env_sampler=api.SAMPLER_RANDOM) as r:
        r.set_env(cube); r.resize(W, H, ntx, nty)
        lat = []
        screen = torch.empty((H, W), dtype=torch.int32).pin_memory().numpy().view(np.uint32)  # BackBuffer.Memory, page-locked
        for f in range(frames + 10):
            t0 = time.perf_counter()
            r.render_frames(1, sync=False)            # NUM_SAMPLES_PER_FRAME 1
            ldr = r.resolve_ldr(api.LDR_SCREEN_BGRA, out=screen)  # tone map + D2H of the u32 frame (blocks)
            lat.append((time.perf_counter() - t0) * 1e3)
        lat = np.array(lat[10:])
        c = r.counters()
        # the same iteration as one call: fused tone map, the copy of a finished band of tile rows overlaps the next band's render
        blocking = {}
        for bands in (-1, 1, 2, 3, 4):
            r.reset()
            l2 = []
            for f in range(frames // 2 + 10):
                t0 = time.perf_counter()
                r.present_blocking(screen, nframes=1, bands=bands)
                l2.append((time.perf_counter() - t0) * 1e3)
            l2 = np.array(l2[10:])
            blocking[str(bands)] = {"p50": float(np.percentile(l2, 50)), "p95": float(np.percentile(l2, 95))}
    # pipelined present ring: fused tone map, async D2H overlapping the next frame's render
    with api.Renderer(profile=api.PROFILE_OPT_V4, math_mode=MATH, num_bounces=8, env_kind=api.ENV_CUBEMAP,
                      env_sampler=api.SAMPLER_RANDOM, output_to_screen=True) as r:
        r.set_env(cube); r.resize(W, H, ntx, nty)
        r.present_submit(1)
        for f in range(10):
            r.present_submit(1); r.present_acquire(copy=False)
        t0 = time.perf_counter()
        for f in range(frames):
            r.present_submit(1); r.present_acquire(copy=False)
        ring_ms = (time.perf_counter() - t0) * 1e3 / frames
    out({"config": 4, "workload": "P_v4 + cubemap 512x3072 atlas, 1920x1080 progressive, 1 spp/frame, 600 frames",
         "latency_ms_p50": float(np.percentile(lat, 50)), "latency_ms_p95": float(np.percentile(lat, 95)),
         "latency_ms_mean": float(lat.mean()), "kernel_ms_last": c["last_render_ms"], "present_ring_ms_per_frame": ring_ms,
         "present_ring_fps": 1e3 / ring_ms, "present_blocking_latency_ms_by_bands": blocking,
         "definition": "host call b200pt_render_frames(1) -> b200pt_resolve_ldr returns with the u32 frame in (page-locked) host memory"})
elif args.config == 5:
    W = H = 8192
    ntx, nty = 16, 64
    total = args.spp or 64  # bounded sample of the 16384-spp job; strong scaling: total fixed
    def factory(accum_mode=api.ACCUM_RUNNING_AVERAGE, device=local):
        return api.Renderer(profile=api.PROFILE_V2, math_mode=MATH, num_bounces=16, device=device, accum_mode=accum_mode)
    for mode in ("spp-shard reduce", "tile-shard gather"):
        R = ptdist.SppShardedRenderer(factory, W, H, ntx, nty, rank, world, local) if mode.startswith("spp") else \
            ptdist.TileShardedRenderer(factory, W, H, ntx, nty, rank, world, local)
        def step():
            if mode.startswith("spp"):
                R.render(total)
            else:
                with torch.cuda.stream(R.stream):
                    R.buf.zero_()
                R.r.frame_counter = 0
                R.render(total)
        step(); R.stream.synchronize(); sync()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with torch.cuda.stream(R.stream):
            e0.record(R.stream); step(); e1.record(R.stream)
        sync()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1: tdist.all_reduce(ms, op=tdist.ReduceOp.MAX)
        out({"config": 5, "sharding": mode, "workload": f"Cornell P_v2 {W}x{H}, {total} spp total (bounded sample of 16384), 16 bounces, strong scaling",
             "ms": float(ms), "mpaths_per_s": W * H * total / float(ms) * 1e-3})
        R.close()
if world > 1:
    tdist.barrier(); tdist.destroy_process_group()
