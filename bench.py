#!/usr/bin/env python
"""bench.py -- headline benchmark of the path-tracing hot path (BASELINE.json).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--math parity|fast]

Metric: Mpaths/s = pixels x spp / second; 1 path = 1 (pixel, frame) sample.
Workload (N=1): BASELINE.json configs[1] -- Cornell-box scene (reference renderer
demofox_path_tracing_v2.cpp, bounces patched to 8), 1920x1080, tiles 10x15, 1024 spp.
One "step" = one full pass of the hot path over that job: 1024 render calls of the reference
(DemofoxRenderV2 x 1024) folded into the f32 accumulation buffer = ONE launch of the persistent
megakernel.  N>1 (STRONG scaling): the SAME 1024-spp job, its frame range spp-sharded N ways
(1024/N frames per GPU), SUM buffers all-reduced over NCCL and scaled -- zeroing, reduce and scale
are inside the step; the weak-scaling figure (1024 spp per GPU, an N*1024-spp image) is reported
as a sub-object.  At every N the line carries `parity_check`: a bounded job (1080p x 16 frames)
rendered spp-sharded and tile-sharded, compared with the sequential single-GPU render outside the
timed region.

Default math policy is PARITY: the kernel whose output is bit-identical to the oracle
(tests/test_gpu_parity.py).  --math fast reports the FMA-contracted / MUFU variant.

The reference arm (--impl reference) times the reference's own AVX2 multithreaded renderer
(oracle/_ref/ref_v2_asis, built in place from /root/reference by oracle/ref_build/build_ref.sh)
on this box's host cores.  One process of that renderer can use 8 threads at most (its tile table,
demofox_path_tracing_v2.cpp:641-643), so ALL cores are used by running host_cores/8 such processes
side by side, each on its own block of the job's frame range (frames are independent: the RNG
re-seeds per (pixel, frame)); `value` is their aggregate, the stock one-process figure is reported
next to it.  If the binaries are absent the oracle port is timed instead.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

WIDTH, HEIGHT, NTX, NTY = 1920, 1080, 10, 15
SPP = 1024
BOUNCES = 8
METRIC = "Mpaths/s (px*spp/s)"
WORKLOAD = "Cornell P_v2 1920x1080 tiles 10x15, 1024 spp, 8 bounces (BASELINE.json configs[1])"

# algorithmic flops per unit, SURVEY.md section 8(d) (hand count from the reference source;
# add/mul/div/sqrt = 1, fma = 2): P_v2 F_seg = 6 quads x 151 + 3 spheres x 43 + 127 shading
F_SEG_V2 = 6 * 151 + 3 * 43 + 127   # = 1162 flop per traced segment of a live path
F_CAM_V2 = 40                       # per path: seed, jitter, camera ray
F_ACC = 9                           # per path: running-average blend (3 x sub, mul, add)
ACC_BYTES_PER_PIXEL_PER_LAUNCH = 24  # f32 target read + write once per launch


def measured_dram_traffic():
    """dram__bytes_read.sum + dram__bytes_write.sum of pt_render_kernel from the newest committed ncu capture
    (profiles/*_metrics.csv, written by scripts/ncu_summary.py from an `ncu --set full` report); bytes per launch."""
    import csv
    import glob
    best = None
    for path in sorted(glob.glob(os.path.join(ROOT, "profiles", "r*_parity_v2_*_metrics.csv"))):
        best = path
    if not best:
        return None, None
    unit = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    tot = 0.0
    with open(best) as f:
        for row in csv.reader(f):
            if len(row) == 3 and row[0] in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
                tot += float(row[2]) * unit.get(row[1], 1.0)
    return (tot if tot > 0 else None), os.path.relpath(best, ROOT)


def measured_ncu_utilisation():
    """What the newest committed `ncu --set full` capture of the headline kernel (profiles/*_parity_v2_*_metrics.csv)
    says about pipe utilisation -- read from the file, not measured by this run: issue slots busy, FMA-pipe
    instruction share, active lanes per instruction, and executed FP32 flops (fadd + fmul + 2 ffma, thread level,
    per SM cycle) as a fraction of the 128 lanes x 2 flop x SM-cycle peak."""
    import csv
    import glob
    paths = sorted(glob.glob(os.path.join(ROOT, "profiles", "r*_parity_v2_*_metrics.csv")))
    if not paths:
        return None
    m = {}
    with open(paths[-1]) as f:
        for row in csv.reader(f):
            if len(row) == 3:
                try:
                    m[row[0]] = float(row[2])
                except ValueError:
                    pass
    out = {"source": os.path.relpath(paths[-1], ROOT)}
    out["issue_active_pct"] = m.get("smsp__issue_active.avg.pct_of_peak_sustained_active")
    out["fma_pipe_pct"] = m.get("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active")
    out["alu_pipe_pct"] = m.get("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active")
    out["xu_pipe_pct"] = m.get("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active")
    out["active_lanes"] = m.get("smsp__thread_inst_executed_per_inst_executed.ratio")
    # thread-level FP32 instructions per elapsed cycle, summed over the chip
    fadd = m.get("smsp__sass_thread_inst_executed_op_fadd_pred_on.sum.per_cycle_elapsed")
    fmul = m.get("smsp__sass_thread_inst_executed_op_fmul_pred_on.sum.per_cycle_elapsed")
    ffma = m.get("smsp__sass_thread_inst_executed_op_ffma_pred_on.sum.per_cycle_elapsed")
    if None not in (fadd, fmul, ffma):
        out["executed_fp32_frac"] = (fadd + fmul + 2.0 * ffma) / (148 * 128 * 2)
    else:
        out["executed_fp32_frac"] = None
    return out


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            p = json.load(f)
        return {"hbm_gbs": p.get("hbm_gbs"), "sm_max_mhz": p.get("sm_max_mhz", 1965.0), "source": "measured"}
    return {"hbm_gbs": 6650.0, "sm_max_mhz": 1965.0, "source": "fallback"}


class ClockSampler:
    """nvidia-smi clocks + throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.tmp = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.proc = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(gpu_index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=self.tmp,
                                         stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if self.proc is None:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        self.tmp.flush()
        self.tmp.seek(0)
        sm, mx, power, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in self.tmp.read().splitlines():
            f = [x.strip() for x in line.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0]))
                mx.append(float(f[1]))
                power.append(float(f[2]))
            except ValueError:
                continue
            for n, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        self.tmp.close()
        os.unlink(self.tmp.name)
        if sm:
            # median over the samples under load (power above half of the max seen)
            pm = max(power)
            load = sorted(s for s, p in zip(sm, power) if p >= 0.5 * pm) or sorted(sm)
            out["sm_mhz"] = load[len(load) // 2]
            out["sm_max_mhz"] = max(mx)
            out["power_w_max"] = pm
            out["samples"] = len(sm)
        out["reasons"] = sorted(reasons)
        return out


def _ref_v2_process(exe, frames, warmup, start_frame, threads):
    """one stock reference process: `warmup` untimed + `frames` timed render calls of DemofoxRenderV2"""
    cmd = [exe, "--w", str(WIDTH), "--h", str(HEIGHT), "--ntx", "2", "--nty", "4", "--frames", str(frames), "--warmup", str(warmup),
           "--start-frame", str(start_frame), "--bounces", str(BOUNCES), "--threads", str(threads), "--time"]
    return subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)


def cpu_reference_run(frames, warmup, timeout=3000):
    """Times the reference's own CPU renderer on a bounded sample of the 1080p job: `frames` timed frames per process.
    Returns the all-cores aggregate as `value` and the stock one-process figure as `stock_single_process`."""
    from oracle import pyoracle as po
    ncores = os.cpu_count() or 1
    exe = po.ref_binary("ref_v2_asis")
    if exe:
        # DemofoxRenderV2 sizes WorkData[8] (demofox_path_tracing_v2.cpp:641-643): at most 8 tiles, so at most 8 threads
        # of one process ever have work.
        th = min(ncores, 8)
        build = "g++ -O2 -mavx2 -mfma, hardware rcpps/rsqrtps, libm sin/cos"
        p = _ref_v2_process(exe, frames, warmup, 0, th)
        out, err = p.communicate(timeout=timeout)
        if p.returncode != 0:
            raise RuntimeError("ref_v2_asis failed: " + err[-300:])
        t1 = json.loads(out.strip().splitlines()[-1])
        single = {"value": t1["mpaths_per_s"], "cores": th, "seconds": t1["seconds"],
                  "sample": f"one process, tiles 2x4, {th} threads, {warmup} warm-up + {frames} timed frames"}
        # all cores: k processes side by side, process j on frames [j*frames, (j+1)*frames) of the job
        k = max(1, ncores // 8)
        if k == 1:
            agg, cores, sec, detail = single["value"], th, single["seconds"], "host has <= 15 cores: one process is all of it"
        else:
            t0 = time.perf_counter()
            procs = [_ref_v2_process(exe, frames, warmup, j * (frames + warmup), 8) for j in range(k)]
            res = []
            for q in procs:
                o, e = q.communicate(timeout=timeout)
                if q.returncode != 0:
                    raise RuntimeError("ref_v2_asis failed: " + e[-300:])
                res.append(json.loads(o.strip().splitlines()[-1]))
            wall = time.perf_counter() - t0
            sec = max(r["seconds"] for r in res)
            agg = WIDTH * HEIGHT * frames * k / sec * 1e-6  # all k blocks done when the slowest process is
            cores = min(ncores, 8 * k)
            detail = f"{k} concurrent processes x 8 threads, each {warmup} warm-up + {frames} timed frames of its own frame block; slowest {sec:.2f} s (wall incl. start-up {wall:.2f} s)"
        return {"value": agg, "unit": "Mpaths/s", "cores": cores, "kind": "reference", "host_cores": ncores,
                "sample": f"{WIDTH}x{HEIGHT}, tiles 2x4 (the v2 renderer's 8-tile limit), DemofoxRenderV2 ({build}); {detail}",
                "seconds": sec, "stock_single_process": single}
    # oracle port: scalar C restatement, pthreads over rows, all cores
    th = ncores
    po.render(po.PROFILE_V2, WIDTH, HEIGHT, NTX, NTY, BOUNCES, max(1, warmup), nthreads=th)
    t0 = time.perf_counter()
    po.render(po.PROFILE_V2, WIDTH, HEIGHT, NTX, NTY, BOUNCES, frames, nthreads=th)
    sec = time.perf_counter() - t0
    return {"value": WIDTH * HEIGHT * frames / sec * 1e-6, "unit": "Mpaths/s", "cores": th, "kind": "port",
            "sample": f"{WIDTH}x{HEIGHT}, {frames} frames of the scalar oracle port (exact rcp/rsqrt), {sec:.2f} s",
            "seconds": sec, "host_cores": ncores}


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    frames_per_step = 8  # ~1.4 s of the host's cores per step
    base = cpu_reference_run(frames=frames_per_step * args.steps, warmup=frames_per_step * args.warmup)
    k = max(1, (os.cpu_count() or 1) // 8) if base.get("kind") == "reference" else 1
    paths = WIDTH * HEIGHT * frames_per_step * args.steps * k
    line = {
        "impl": "reference", "metric": METRIC, "value": base["value"], "unit": "Mpaths/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": base["seconds"] * 1e3 / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "step": f"{frames_per_step} frames of the {SPP}-spp job per process (bounded sample; "
                                                 "throughput is spp-independent)", "paths_timed": paths},
        "cpu_baseline": base,
        "e2e": {"value": base["value"], "unit": "Mpaths/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)
    return 0


_JSON_FD = None


def claim_stdout():
    """stdout carries exactly ONE JSON line: whatever libraries print there (NCCL's version banner, ...) is sent to
    stderr by pointing fd 1 at fd 2; the JSON line goes to the saved original descriptor."""
    global _JSON_FD
    if _JSON_FD is None:
        sys.stdout.flush()
        _JSON_FD = os.dup(1)
        os.dup2(2, 1)


def emit(line):
    data = (json.dumps(line) + "\n").encode()
    if _JSON_FD is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_JSON_FD, data)


def main():
    claim_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--math", default="parity", choices=["parity", "fast"])
    ap.add_argument("--spp", type=int, default=SPP, help="frames per step per GPU (default: the 1024-spp job)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = 3  # timing rule: at least 3 warm-up steps
    if args.impl == "reference":
        return run_reference_arm(args)

    import numpy as np
    import torch
    from cpuperformanceraytracer_b200 import api, dist as ptdist

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("launch N>1 with torch.distributed.run (one rank per GPU)")
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a B200: the engine has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        import torch.distributed as tdist
        tdist.init_process_group("nccl", device_id=dev)

    math_mode = api.MATH_PARITY if args.math == "parity" else api.MATH_FAST
    spp = args.spp
    peaks = load_peaks()

    def factory(accum_mode=api.ACCUM_RUNNING_AVERAGE, device=local_rank):
        return api.Renderer(profile=api.PROFILE_V2, math_mode=math_mode, num_bounces=BOUNCES, device=device,
                            accum_mode=accum_mode)

    stream = torch.cuda.Stream(dev)
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)  # > 126 MB L2

    # `spp` is the job's total frame count: at N > 1 it is split N ways (strong scaling)
    if world == 1:
        r = factory()
        r.resize(WIDTH, HEIGHT, NTX, NTY)
        r.set_stream(stream.cuda_stream)
        sr = None

        def step():
            # one full job: zeroed buffer (Resize), frame counter 0, 1024 render calls
            r.reset()
            r.render_frames(spp, sync=False)
    else:
        sr = ptdist.SppShardedRenderer(factory, WIDTH, HEIGHT, NTX, NTY, rank, world, local_rank)
        stream = sr.stream
        r = sr.r

        def step():
            # the same job: rank r zeroes its SUM buffer, renders its block of the 1024 frames, all-reduce, scale
            sr.render(total_frames=spp)

    def barrier():
        if world > 1:
            tdist.barrier()
        torch.cuda.synchronize(dev)

    def timed(fn, steps, warmup):
        """`steps` calls of fn on `stream`, L2 flushed before each, CUDA events on that stream, max over ranks"""
        with torch.cuda.stream(stream):
            for _ in range(warmup):
                flush.zero_()
                fn()
        barrier()
        c_before = r.counters()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        with torch.cuda.stream(stream):
            e0.record(stream)
            for _ in range(steps):
                flush.zero_()  # L2 flush between timed iterations (inside the timed region)
                fn()
            e1.record(stream)
        barrier()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], dtype=torch.float64, device=dev)
            tdist.all_reduce(t, op=tdist.ReduceOp.MAX)
            ms = float(t.item())
        return ms, c_before, r.counters()

    sampler = ClockSampler(local_rank) if rank == 0 else None
    total_ms, c0, c1 = timed(step, args.steps, args.warmup)
    clocks = sampler.stop() if sampler else None
    # per-launch duration of the dominant kernel, CUDA events on the launching stream (the library
    # brackets every render launch with its own event pair; this is the last timed launch)
    last_kernel_ms = c1["last_render_ms"]

    def delta(key):
        v = float(c1[key] - c0[key])
        if world > 1:
            t = torch.tensor([v], dtype=torch.float64, device=dev)
            tdist.all_reduce(t, op=tdist.ReduceOp.SUM)
            v = float(t.item())
        return v

    segs, escs, culled = delta("segments"), delta("escapes"), delta("culled_segments")
    ffma_peak = r.measure_fp32_peak() if rank == 0 else None  # FFMA micro-benchmark, outside the timed region
    paths_per_step = WIDTH * HEIGHT * spp
    total_paths = paths_per_step * args.steps
    value = total_paths / (total_ms * 1e-3) * 1e-6
    launches = int(c1["launches"] - c0["launches"])

    # the image of the timed job, read back on rank 0: the same job must give the same picture at every N
    # (identical samples; only the order of the f32 additions depends on N)
    if world == 1:
        img = r.download_target()
    else:
        sr.stream.synchronize()
        img = sr.buf.cpu().numpy()
    image_check = {"mean": float(img.astype(np.float64).mean()), "sum_sq": float((img.astype(np.float64) ** 2).sum()),
                   "finite": bool(np.isfinite(img).all())}

    # ---- weak scaling as a sub-object (N > 1): every GPU renders the full frame count --------------------------
    weak = None
    if world > 1:
        wsteps = max(1, min(args.steps, 2))
        wms, _, _ = timed(lambda: sr.render(total_frames=spp * world), wsteps, 1)
        weak = {"value": WIDTH * HEIGHT * spp * world * wsteps / (wms * 1e-3) * 1e-6, "unit": "Mpaths/s", "scaling": "weak",
                "ms_per_step": wms / wsteps, "steps": wsteps, "workload": f"{spp} spp per GPU = a {spp * world}-spp image per step"}

    # ---- parity_check: sharded renders against the sequential one, outside the timed region --------------------
    PF = 16
    if world == 1:
        with factory() as q:
            q.resize(WIDTH, HEIGHT, NTX, NTY)
            q.render_frames(PF)
            seq = q.download_target()
        # one GPU: the sharding logic of b200pt_group_* with two ranks placed on this device
        with api.Group([local_rank, local_rank], sharding=api.SHARD_SPP, combine=api.COMBINE_PEER, profile=api.PROFILE_V2,
                       math_mode=math_mode, num_bounces=BOUNCES) as g:
            g.resize(WIDTH, HEIGHT, NTX, NTY)
            g.render_frames(PF)
            a = g.download_target()
        with api.Group([local_rank, local_rank], sharding=api.SHARD_TILES, profile=api.PROFILE_V2, math_mode=math_mode,
                       num_bounces=BOUNCES) as g:
            g.resize(WIDTH, HEIGHT, NTX, NTY)
            g.render_frames(PF)
            t_img = g.download_target()
        how = "b200pt_group with 2 ranks on this GPU (peer-memory combine) vs one context"
    else:
        sr.render(total_frames=PF)
        sr.stream.synchronize()
        a = sr.buf.cpu().numpy()
        sr.close()
        tr = ptdist.TileShardedRenderer(factory, WIDTH, HEIGHT, NTX, NTY, rank, world, local_rank)
        tr.render(PF)
        tr.stream.synchronize()
        t_img = tr.buf.cpu().numpy()
        tr.close()
        seq = None
        if rank == 0:
            with factory() as q:
                q.resize(WIDTH, HEIGHT, NTX, NTY)
                q.render_frames(PF)
                seq = q.download_target()
        how = f"SppShardedRenderer / TileShardedRenderer over {world} GPUs (NCCL) vs rank 0 alone"
    parity_check = None
    if rank == 0:
        rel = np.abs(a.astype(np.float64) - seq) / np.maximum(np.abs(seq.astype(np.float64)), 1e-3)
        parity_check = {"job": f"{WIDTH}x{HEIGHT}, {PF} frames", "how": how,
                        "spp_shard_max_rel": float(rel.max()), "spp_shard_ok": bool(rel.max() <= 3e-6),
                        "tile_shard_bit_exact": bool(np.array_equal(t_img, seq))}

    # ---- e2e: the reference-facing call on HOST buffers -----------------------------------------------------------
    nbytes = WIDTH * HEIGHT * 3 * 4
    if world == 1:
        # the caller's accumulation buffer, page-locked as the contract asks (render_host copies pinned buffers
        # directly and stages pageable ones)
        host_t = torch.zeros(WIDTH * HEIGHT * 3, dtype=torch.float32).pin_memory()
        host = host_t.numpy()
        r.set_stream(None)
        with api.Renderer(profile=api.PROFILE_V2, math_mode=math_mode, num_bounces=BOUNCES, device=local_rank) as rh:
            rh.render_host(host, WIDTH, HEIGHT, NTX, NTY, 8)  # warm-up: allocations, pinned staging
            e2e_steps = max(1, min(args.steps, 3))
            sec = 0.0
            for _ in range(e2e_steps):
                host[:] = 0.0  # a fresh accumulation state (not part of the call)
                rh.frame_counter = 0
                t0 = time.perf_counter()
                rh.render_host(host, WIDTH, HEIGHT, NTX, NTY, spp)  # H2D of the state, render, D2H of the result
                checksum = float(host[::4097].sum())  # device->host result is read on the host
                sec += time.perf_counter() - t0
        e2e = {"value": WIDTH * HEIGHT * spp * e2e_steps / sec * 1e-6, "unit": "Mpaths/s",
               "h2d_bytes_per_step": nbytes, "d2h_bytes_per_step": nbytes,
               "steps": e2e_steps, "api": "b200pt_render_host (DemofoxRenderV2 signature + frame count) on a page-locked host buffer, wall clock",
               "checksum": checksum}
    else:
        # N > 1: the caller's accumulation buffer lives in host memory every rank can address (POSIX shared memory,
        # page-locked in each process).  Rank r moves 1/N of it over its own PCIe link in both directions, renders its
        # frame block, and the SUM buffers are all-reduced in between (SppShardedRenderer.render_host_slices).
        import ctypes as _ct
        sr = ptdist.SppShardedRenderer(factory, WIDTH, HEIGHT, NTX, NTY, rank, world, local_rank)
        shm_path = "/dev/shm/b200pt_bench_%s.f32" % os.environ.get("MASTER_PORT", "0")
        nfl = WIDTH * HEIGHT * 3
        if rank == 0:
            np.zeros(nfl, dtype=np.float32).tofile(shm_path)
        barrier()
        host_np = np.memmap(shm_path, dtype=np.float32, mode="r+", shape=(nfl,))
        rc = torch.cuda.cudart().cudaHostRegister(host_np.ctypes.data, nbytes, 0)
        registered = (int(rc) == 0) if not isinstance(rc, tuple) else (int(rc[0]) == 0)
        host_t = torch.from_numpy(host_np)
        e2e_steps = max(1, min(args.steps, 3))
        sec, step_secs = 0.0, []
        for it in range(e2e_steps + 1):  # the first pass is an untimed warm-up of the whole call
            barrier()
            if rank == 0:
                host_np[:] = 0.0  # a fresh accumulation state (not part of the call)
            barrier()
            t0 = time.perf_counter()
            sr.render_host_slices(host_t, spp)
            barrier()  # every slice is in the caller's buffer
            if it > 0:
                step_secs.append(time.perf_counter() - t0)
        sec = sum(step_secs)
        checksum = float(host_np[::4097].sum()) if rank == 0 else 0.0
        e2e_mean = float(np.asarray(host_np, dtype=np.float64).mean()) if rank == 0 else 0.0
        tt = torch.tensor([sec], dtype=torch.float64, device=dev)
        tdist.all_reduce(tt, op=tdist.ReduceOp.MAX)
        sec = float(tt.item())
        e2e = {"value": paths_per_step * e2e_steps / sec * 1e-6, "unit": "Mpaths/s", "h2d_bytes_per_step": nbytes,
               "d2h_bytes_per_step": nbytes, "steps": e2e_steps, "checksum": checksum, "image_mean": e2e_mean,
               "host_buffer_page_locked": registered, "step_seconds": step_secs,
               "api": "the caller's f32 buffer in POSIX shared memory, page-locked in every rank's process: rank r copies 1/N of it in, "
                      "renders its frame block, NCCL all-reduce + scale, copies 1/N of the image out over its own PCIe link "
                      "(SppShardedRenderer.render_host_slices); wall clock between barriers, max over ranks"}
        try:
            torch.cuda.cudart().cudaHostUnregister(host_np.ctypes.data)
        except Exception:
            pass
        del host_t, host_np
        sr.close()
        barrier()
        if rank == 0:
            try:
                os.unlink(shm_path)
            except OSError:
                pass

    if rank == 0:
        # ---- roofline of the dominant kernel (pt_render_kernel): FP32 pipe, not HBM, not tensor ----
        step_ms = total_ms / args.steps
        fp32_peak = 148 * 128 * 2 * peaks["sm_max_mhz"] * 1e6 * 1e-12 * world  # TFLOP/s, FFMA = 2 flop
        flops_per_step = (segs / args.steps) * F_SEG_V2 + paths_per_step * (F_CAM_V2 + F_ACC)
        # the same without credit for segments whose scene trace the kernel provably skips (camera-culled pixels)
        flops_traced = ((segs - culled) / args.steps) * F_SEG_V2 + paths_per_step * (F_CAM_V2 + F_ACC)
        achieved = flops_per_step / (step_ms * 1e-3) * 1e-12
        hbm_bytes = WIDTH * HEIGHT * ACC_BYTES_PER_PIXEL_PER_LAUNCH * world
        traffic, traffic_src = measured_dram_traffic()
        roofline = {
            "bound": "fp32",
            "achieved": achieved, "peak": fp32_peak, "unit": "TFLOP/s", "frac": achieved / fp32_peak,
            "frac_traced_only": flops_traced / (step_ms * 1e-3) * 1e-12 / fp32_peak,
            "frac_definition": "frac: reference flop count x segments the reference traces / time / FFMA peak; "
                               "frac_traced_only: no flop credit for the segments of camera-culled pixels (counted by the "
                               "kernel), whose trace the kernel skips; executed_fp32_frac etc. under `ncu`: what the hardware "
                               "executed, from the committed ncu capture",
            "culled_segment_share": culled / segs if segs else None,
            "measured_ffma_peak_tflops_per_gpu": ffma_peak,
            "frac_vs_measured_ffma_peak": (achieved / (ffma_peak * world)) if ffma_peak else None,
            "ncu": measured_ncu_utilisation(),
            "traffic": traffic, "traffic_source": traffic_src,
            "peak_source": f"148 SMs x 128 FP32 lanes x 2 x sm_max_mhz {peaks['sm_max_mhz']:.0f} MHz "
                           f"(MEASURED_PEAKS.json clock, {peaks['source']}); the file's hbm/bf16 peaks do not bound this kernel",
            "algorithmic_flops_per_segment": F_SEG_V2, "segments_per_path": segs / total_paths,
            "kernel_ms_last_launch": last_kernel_ms,
            "hbm": {"achieved_gbs": hbm_bytes / (step_ms * 1e-3) * 1e-9, "peak_gbs": peaks["hbm_gbs"],
                    "algorithmic_bytes_per_launch": hbm_bytes},
        }
        cpu_base = None
        if world == 1 and not args.no_cpu_baseline:
            try:
                cpu_base = cpu_reference_run(frames=48, warmup=2)  # ~9 s one process + ~9 s all cores
            except Exception as e:  # the reported baseline must not take the bench down
                cpu_base = {"value": None, "unit": "Mpaths/s", "cores": 0, "kind": "unavailable", "sample": str(e)[:200]}
        line = {
            "metric": METRIC, "value": value, "unit": "Mpaths/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": step_ms, "higher_is_better": True,
            "scaling": "strong" if world > 1 else "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD if spp == SPP else WORKLOAD.replace("1024 spp", f"{spp} spp"), "math": args.math,
                       "spp_per_gpu_per_step": spp / world,
                       "rendered_per_step": f"one {WIDTH}x{HEIGHT} image of {spp} spp" + ("" if world == 1 else f", {spp}/{world} frames on each GPU"),
                       "sharding": "single GPU" if world == 1 else f"spp-shard x{world} of the fixed job: zero + render + NCCL all-reduce + 1/(N+1) scale inside the step",
                       "l2": "flushed between timed steps by a 256 MiB memset inside the timed region",
                       "parity": ("bit-exact vs oracle (math=parity); this very job against the reference build: profiles/r02_n_full_job_parity.json"
                                  if args.math == "parity" else "RMSE-bounded (math=fast)")},
            "roofline": roofline,
            "e2e": e2e,
            "gpu_launches": launches,
            "clocks": clocks,
            "parity_check": parity_check,
            "image_check": image_check,
        }
        if weak is not None:
            line["weak"] = weak
        if cpu_base is not None:
            line["cpu_baseline"] = cpu_base
        emit(line)
    if world > 1:
        tdist.barrier()
        tdist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
