#!/usr/bin/env python
"""bench_configs.py -- the BASELINE.json configurations other than the headline one (bounded samples),
one JSON line each.  Run under torchrun for N > 1:  python -m torch.distributed.run --nproc-per-node N ...
    --config 1   Cornell P_v2 512x512, 64 spp, 8 bounces: full-size bit-exact check vs the oracle + time
    --config 3   simt_textured (and P_v4 equirect) 3840x2160, spp-sharded, synthetic 2048x1024 env
    --config 4   P_v4 + cubemap, 1080p progressive 1 spp/frame: per-frame latency (render + tone map + D2H)
    --config 5   8192x8192, 16 bounces, strong scaling: spp-shard reduce vs tile-shard gather
Every line carries cpu_baseline: the reference's own renderer of that configuration (oracle/_ref/ref_*_asis) on all host
cores, on a bounded sample (--no-cpu-baseline skips it).
    --group N    configs 3 / 5 in ONE process through the C ABI's b200pt_group_* (no torch.distributed): N GPUs
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
ap.add_argument("--group", type=int, default=0)
ap.add_argument("--no-cpu-baseline", action="store_true")
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
        d.update(n_gpus=d.pop("n_gpus_group", world), math=args.math)
        print(json.dumps(d), flush=True)


def sync():
    if world > 1:
        tdist.barrier()
    torch.cuda.synchronize(dev)


def cpu_baseline(binary, W, H, ntx, nty, frames, bounces, env=None, threads=None):
    """the reference's own renderer (asis build) on the host, bounded sample; rank 0 only"""
    if rank != 0 or args.no_cpu_baseline:
        return None
    from oracle import pyoracle as po
    ncores = os.cpu_count() or 1
    if not po.ref_binary(binary):
        return {"kind": "unavailable", "sample": binary + " not built"}
    th = threads or ncores
    try:
        t = po.run_ref(binary, W, H, ntx, nty, frames, bounces=bounces, env=env, threads=th, time_it=True, warmup=1)["timing"]
    except Exception as e:
        return {"kind": "unavailable", "sample": str(e)[:200]}
    return {"value": t["mpaths_per_s"], "unit": "Mpaths/s", "cores": min(th, ncores), "host_cores": ncores, "kind": "reference",
            "sample": f"{binary}, {W}x{H}, tiles {ntx}x{nty}, {th} threads, 1 warm-up + {frames} timed frames, {t['seconds']:.2f} s"}


def timed_group(G, fn):
    """device time of fn() on a b200pt group: wall clock around fn + synchronize, after a warm-up call"""
    fn(); G.synchronize()
    t0 = time.perf_counter(); fn(); G.synchronize()
    return (time.perf_counter() - t0) * 1e3


if args.config == 1:
    from oracle import pyoracle as po
    W = H = 512
    t0 = time.time(); o, oc = po.render(po.PROFILE_V2, W, H, 2, 4, 8, 64); tcpu = time.time() - t0
    with api.Renderer(profile=api.PROFILE_V2, math_mode=MATH, num_bounces=8) as r:
        r.resize(W, H, 2, 4); r.render_frames(64); r.reset(); r.render_frames(64)
        g = r.download_target(); c = r.counters(); rs = r.rng_state()
    d = np.abs(g.astype(np.float64) - o)
    out({"config": 1, "workload": "Cornell P_v2 512x512 tiles 2x4, 64 spp, 8 bounces", "gpu_ms": c["last_render_ms"],
         "cpu_baseline": cpu_baseline("ref_v2_asis", W, H, 2, 4, 64, 8, threads=8),
         "mpaths_per_s": W * H * 64 / c["last_render_ms"] * 1e-3, "bit_exact_vs_oracle": bool(np.array_equal(g, o)),
         "rmse": float(np.sqrt((d ** 2).mean())), "max_abs": float(d.max()), "oracle_port_seconds": tcpu,
         "segments_match": bool(c["segments"] % (2 ** 64) >= oc["segments"])})
elif args.config == 3:
    from oracle import pyoracle as po
    W, H, ntx, nty = 3840, 2160, 10, 15
    spp = args.spp or 256
    env = po.synthetic_env(2048, 1024)
    for name, prof, kw, refbin in (("simt_textured", api.PROFILE_SIMT_TEXTURED, {}, "ref_simt_textured_asis"),
                                   ("v4_equirect_random", api.PROFILE_OPT_V4, dict(env_kind=api.ENV_EQUIRECT, env_sampler=api.SAMPLER_RANDOM),
                                    "ref_v4_equirect_random_asis")):
        bounces = 4 if prof == api.PROFILE_SIMT_TEXTURED else 8
        # simt_textured keeps the v2 renderer's 8-tile table (simt_textured.cpp WorkData[8]): tiles 2x4, 8 threads
        base = cpu_baseline(refbin, W, H, 2 if refbin.startswith("ref_simt") else ntx, 4 if refbin.startswith("ref_simt") else nty,
                            4 if refbin.startswith("ref_simt") else 16, bounces, env=env, threads=8 if refbin.startswith("ref_simt") else None)
        if args.group:
            n = args.group
            for sharding, combine, label in ((api.SHARD_SPP, api.COMBINE_NCCL, "spp-shard, NCCL reduce"),
                                             (api.SHARD_SPP, api.COMBINE_PEER, "spp-shard, peer-memory combine kernel"),
                                             (api.SHARD_SPP, api.COMBINE_FUSED, "spp-shard, fused render + reduce-scatter"),
                                             (api.SHARD_TILES, api.COMBINE_PEER, "tile-shard, interleaved tiles + gather kernel")):
                with api.Group(list(range(n)), sharding=sharding, combine=combine, profile=prof, math_mode=MATH, num_bounces=bounces, **kw) as G:
                    G.set_env(env); G.resize(W, H, ntx, nty)
                    def job():
                        G.reset(); G.render_frames(spp * n, sync=False)
                    ms = timed_group(G, job)
                    c = G.counters()
                out({"config": 3, "profile": name, "api": "b200pt_group_* (one process, C ABI)", "sharding": label, "n_gpus_group": n,
                     "workload": f"{W}x{H}, {spp} spp per GPU (bounded sample of 4096), synthetic 2048x1024 equirect env",
                     "ms": ms, "mpaths_per_s": W * H * spp * n / ms * 1e-3, "exposed_combine_ms": c["combine_ms"], "cpu_baseline": base})
            continue
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
             "ms": float(ms), "mpaths_per_s": W * H * spp * world / float(ms) * 1e-3, "cpu_baseline": base})
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
    base = cpu_baseline("ref_v4_cubemap_random_asis", W, H, ntx, nty, 16, 8, env=cube)
    out({"config": 4, "workload": "P_v4 + cubemap 512x3072 atlas, 1920x1080 progressive, 1 spp/frame, 600 frames", "cpu_baseline": base,
         "cpu_ms_per_frame": (W * H / (base["value"] * 1e3)) if base and base.get("value") else None,
         "latency_ms_p50": float(np.percentile(lat, 50)), "latency_ms_p95": float(np.percentile(lat, 95)),
         "latency_ms_mean": float(lat.mean()), "kernel_ms_last": c["last_render_ms"], "present_ring_ms_per_frame": ring_ms,
         "present_ring_fps": 1e3 / ring_ms, "present_blocking_latency_ms_by_bands": blocking,
         "definition": "host call b200pt_render_frames(1) -> b200pt_resolve_ldr returns with the u32 frame in (page-locked) host memory"})
elif args.config == 5:
    W = H = 8192
    ntx, nty = 16, 64
    total = args.spp or 512  # bounded sample of the 16384-spp job; strong scaling: total fixed whatever the GPU count
    base = cpu_baseline("ref_v2_asis", 2048, 2048, 2, 4, 2, 16, threads=8)  # a 2048^2 crop-sized image: throughput is size-independent
    wl = f"Cornell P_v2 {W}x{H}, {total} spp total (bounded sample of 16384), 16 bounces, strong scaling"
    if args.group:
        n = args.group
        for sharding, combine, bands, label in ((api.SHARD_SPP, api.COMBINE_NCCL, 1, "spp-shard, NCCL reduce after the render"),
                                                (api.SHARD_SPP, api.COMBINE_NCCL, 8, "spp-shard, NCCL reduce per band behind the render"),
                                                (api.SHARD_SPP, api.COMBINE_PEER, 1, "spp-shard, peer-memory combine kernel after the render"),
                                                (api.SHARD_SPP, api.COMBINE_PEER, 8, "spp-shard, peer-memory combine kernel per band behind the render"),
                                                (api.SHARD_SPP, api.COMBINE_FUSED, 1, "spp-shard, fused render + reduce-scatter"),
                                                (api.SHARD_TILES, api.COMBINE_PEER, 1, "tile-shard, interleaved tiles + gather kernel")):
            with api.Group(list(range(n)), sharding=sharding, combine=combine, profile=api.PROFILE_V2, math_mode=MATH, num_bounces=16) as G:
                G.resize(W, H, ntx, nty)
                G.set_bands(bands)
                def job():
                    G.reset(); G.render_frames(total, sync=False)
                ms = timed_group(G, job)
                c = G.counters()
            out({"config": 5, "api": "b200pt_group_* (one process, C ABI)", "sharding": label, "n_gpus_group": n, "workload": wl,
                 "ms": ms, "mpaths_per_s": W * H * total / ms * 1e-3, "exposed_combine_ms": c["combine_ms"], "cpu_baseline": base})
    else:
        def factory(accum_mode=api.ACCUM_RUNNING_AVERAGE, device=local):
            return api.Renderer(profile=api.PROFILE_V2, math_mode=MATH, num_bounces=16, device=device, accum_mode=accum_mode)
        for mode, bands in (("spp-shard, all-reduce after the render", 1), ("spp-shard, all-reduce per band behind the render", 8),
                            ("tile-shard gather", 1)):
            R = ptdist.SppShardedRenderer(factory, W, H, ntx, nty, rank, world, local) if mode.startswith("spp") else \
                ptdist.TileShardedRenderer(factory, W, H, ntx, nty, rank, world, local)
            def step():
                if mode.startswith("spp"):
                    R.render(total, bands=bands)
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
            out({"config": 5, "sharding": mode, "workload": wl, "ms": float(ms), "mpaths_per_s": W * H * total / float(ms) * 1e-3,
                 "cpu_baseline": base})
            R.close()
if world > 1:
    tdist.barrier(); tdist.destroy_process_group()
