#!/usr/bin/env python
"""bench.py -- headline benchmark of the path-tracing hot path (BASELINE.json).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--math parity|fast]

Metric: Mpaths/s = pixels x spp / second; 1 path = 1 (pixel, frame) sample.
Workload (N=1): BASELINE.json configs[1] -- Cornell-box scene (reference renderer
demofox_path_tracing_v2.cpp, bounces patched to 8), 1920x1080, tiles 10x15, 1024 spp.
One "step" = one full pass of the hot path over that job: 1024 render calls of the reference
(DemofoxRenderV2 x 1024) folded into the f32 accumulation buffer = ONE launch of the persistent
megakernel.  N>1: the frame range is spp-sharded, 1024 spp per GPU (weak scaling: an N*1024-spp
image), SUM buffers all-reduced over NCCL and scaled -- reduce and scale are inside the step.

Default math policy is PARITY: the kernel whose output is bit-identical to the oracle
(tests/test_gpu_parity.py).  --math fast reports the FMA-contracted / MUFU variant.

The reference arm (--impl reference) times the reference's own AVX2 multithreaded renderer
(oracle/_ref/ref_v2_asis, built in place from /root/reference by oracle/ref_build/build_ref.sh)
on this box's host cores; if those binaries are absent it times the oracle port instead.
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


def cpu_reference_run(frames, warmup, threads=None, timeout=3000):
    """Times the reference's own CPU renderer (or the oracle port) on `frames` 1080p frames."""
    from oracle import pyoracle as po
    ncores = os.cpu_count() or 1
    if po.ref_binary("ref_v2_asis"):
        # DemofoxRenderV2 sizes WorkData[8] (demofox_path_tracing_v2.cpp:641-643): at most 8 tiles,
        # so at most 8 threads ever have work; give it min(cores, 8) queue workers + the caller.
        th = threads or min(ncores, 8)
        res = po.run_ref("ref_v2_asis", WIDTH, HEIGHT, 2, 4, frames, bounces=BOUNCES, threads=th, time_it=True,
                         warmup=warmup, timeout=timeout)
        t = res["timing"]
        return {"value": t["mpaths_per_s"], "unit": "Mpaths/s", "cores": min(th + 1, 8, ncores), "kind": "reference",
                "sample": f"{WIDTH}x{HEIGHT}, tiles 2x4 (the v2 renderer's 8-tile limit), {warmup} warm-up + {frames} timed "
                          f"frames of DemofoxRenderV2 (g++ -O2 -mavx2 -mfma, hardware rcpps/rsqrtps, libm sin/cos), "
                          f"{t['seconds']:.2f} s", "seconds": t["seconds"], "host_cores": ncores}
    # oracle port: scalar C restatement, pthreads over rows, all cores
    th = threads or ncores
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
    paths = WIDTH * HEIGHT * frames_per_step * args.steps
    line = {
        "impl": "reference", "metric": METRIC, "value": base["value"], "unit": "Mpaths/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": base["seconds"] * 1e3 / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "step": f"{frames_per_step} frames of the {SPP}-spp job (bounded sample; "
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

    if world == 1:
        r = factory()
        r.resize(WIDTH, HEIGHT, NTX, NTY)
        r.set_stream(stream.cuda_stream)

        def step():
            # one full job: zeroed buffer (Resize), frame counter 0, 1024 render calls
            r.reset()
            r.render_frames(spp, sync=False)
    else:
        sr = ptdist.SppShardedRenderer(factory, WIDTH, HEIGHT, NTX, NTY, rank, world, local_rank)
        stream = sr.stream
        r = sr.r

        def step():
            sr.render(total_frames=spp * world)

    def barrier():
        if world > 1:
            tdist.barrier()
        torch.cuda.synchronize(dev)

    with torch.cuda.stream(stream):
        for _ in range(args.warmup):
            flush.zero_()
            step()
    barrier()
    c0 = r.counters()

    sampler = ClockSampler(local_rank) if rank == 0 else None
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    with torch.cuda.stream(stream):
        ev0.record(stream)
        for _ in range(args.steps):
            flush.zero_()  # L2 flush between timed iterations (inside the timed region)
            step()
        ev1.record(stream)
    barrier()
    total_ms = ev0.elapsed_time(ev1)
    clocks = sampler.stop() if sampler else None
    c1 = r.counters()
    # per-launch duration of the dominant kernel, CUDA events on the launching stream (the library
    # brackets every render launch with its own event pair; this is the last timed launch)
    last_kernel_ms = c1["last_render_ms"]

    if world > 1:
        t = torch.tensor([total_ms], dtype=torch.float64, device=dev)
        tdist.all_reduce(t, op=tdist.ReduceOp.MAX)
        total_ms = float(t.item())
        seg = torch.tensor([c1["segments"] - c0["segments"], c1["escapes"] - c0["escapes"]], dtype=torch.float64, device=dev)
        tdist.all_reduce(seg, op=tdist.ReduceOp.SUM)
        segs, escs = float(seg[0].item()), float(seg[1].item())
    else:
        segs, escs = float(c1["segments"] - c0["segments"]), float(c1["escapes"] - c0["escapes"])

    paths_per_step = WIDTH * HEIGHT * spp * world
    total_paths = paths_per_step * args.steps
    value = total_paths / (total_ms * 1e-3) * 1e-6
    launches = int(c1["launches"] - c0["launches"])

    # ---- e2e: the reference-facing call on HOST buffers (rank 0's GPU share of the job) ----------
    e2e = None
    cpu_base = None
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
               "h2d_bytes_per_step": WIDTH * HEIGHT * 3 * 4, "d2h_bytes_per_step": WIDTH * HEIGHT * 3 * 4,
               "steps": e2e_steps, "api": "b200pt_render_host (DemofoxRenderV2 signature + frame count) on a page-locked host buffer, wall clock",
               "checksum": checksum}
    else:
        # N>1: every rank's result tensor is read back to pinned host memory inside the step
        pinned = torch.empty(WIDTH * HEIGHT * 3, dtype=torch.float32).pin_memory()
        barrier()
        t0 = time.perf_counter()
        e2e_steps = max(1, min(args.steps, 3))
        for _ in range(e2e_steps):
            buf = sr.render(total_frames=spp * world)
            with torch.cuda.stream(sr.stream):
                pinned.copy_(buf, non_blocking=True)
            sr.stream.synchronize()
        barrier()
        sec = time.perf_counter() - t0
        tt = torch.tensor([sec], dtype=torch.float64, device=dev)
        tdist.all_reduce(tt, op=tdist.ReduceOp.MAX)
        sec = float(tt.item())
        e2e = {"value": paths_per_step * e2e_steps / sec * 1e-6, "unit": "Mpaths/s", "h2d_bytes_per_step": 0,
               "d2h_bytes_per_step": WIDTH * HEIGHT * 3 * 4 * world, "steps": e2e_steps,
               "api": "SppShardedRenderer.render + D2H of the reduced buffer on every rank, wall clock"}

    if rank == 0:
        # ---- roofline of the dominant kernel (pt_render_kernel): FP32 pipe, not HBM, not tensor ----
        flops_per_step = (segs / args.steps) * F_SEG_V2 + paths_per_step * (F_CAM_V2 + F_ACC)
        step_ms = total_ms / args.steps
        fp32_peak = 148 * 128 * 2 * peaks["sm_max_mhz"] * 1e6 * 1e-12 * world  # TFLOP/s, FFMA = 2 flop
        achieved = flops_per_step / (step_ms * 1e-3) * 1e-12
        hbm_bytes = WIDTH * HEIGHT * ACC_BYTES_PER_PIXEL_PER_LAUNCH * world
        traffic, traffic_src = measured_dram_traffic()
        roofline = {
            "bound": "fp32",
            "achieved": achieved, "peak": fp32_peak, "unit": "TFLOP/s", "frac": achieved / fp32_peak,
            "traffic": traffic, "traffic_source": traffic_src,
            "peak_source": f"148 SMs x 128 FP32 lanes x 2 x sm_max_mhz {peaks['sm_max_mhz']:.0f} MHz "
                           f"(MEASURED_PEAKS.json clock, {peaks['source']}); the file's hbm/bf16 peaks do not bound this kernel",
            "algorithmic_flops_per_segment": F_SEG_V2, "segments_per_path": segs / total_paths,
            "kernel_ms_last_launch": last_kernel_ms,
            "hbm": {"achieved_gbs": hbm_bytes / (step_ms * 1e-3) * 1e-9, "peak_gbs": peaks["hbm_gbs"],
                    "algorithmic_bytes_per_launch": hbm_bytes},
        }
        if world == 1 and not args.no_cpu_baseline:
            try:
                cpu_base = cpu_reference_run(frames=64, warmup=2)  # ~11 s of CPU work
            except Exception as e:  # the reported baseline must not take the bench down
                cpu_base = {"value": None, "unit": "Mpaths/s", "cores": 0, "kind": "unavailable", "sample": str(e)[:200]}
        line = {
            "metric": METRIC, "value": value, "unit": "Mpaths/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": step_ms, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "math": args.math, "spp_per_gpu_per_step": spp,
                       "sharding": "single GPU" if world == 1 else f"spp-shard x{world}: SUM buffers, NCCL all-reduce, 1/(N+1) scale",
                       "l2": "flushed between timed steps by a 256 MiB memset inside the timed region",
                       "parity": "bit-exact vs oracle (math=parity)" if args.math == "parity" else "RMSE-bounded (math=fast)"},
            "roofline": roofline,
            "e2e": e2e,
            "gpu_launches": launches,
            "clocks": clocks,
        }
        if cpu_base is not None:
            line["cpu_baseline"] = cpu_base
        emit(line)
    if world > 1:
        tdist.barrier()
        tdist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
