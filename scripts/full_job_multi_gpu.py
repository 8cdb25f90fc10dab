#!/usr/bin/env python
"""full_job_multi_gpu.py -- the headline job (Cornell P_v2, 1920x1080, 8 bounces, 1024 spp, tiles 2x4) through the C group API
(b200pt_group_*, one process, no torch) on every GPU of the box:
  tile sharding: the SHA-256 of the buffer must be the one recorded in profiles/r02_n_full_job_parity.json -- i.e. the
                 reference build's own buffer, bit for bit;
  spp sharding (NCCL, peer kernel, fused): another summation order, so the difference to that buffer is reported
                 (bar: <= 3e-6 relative, include/b200pt.h).
One JSON line.  usage: full_job_multi_gpu.py [--gpus N]"""
import argparse
import hashlib
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from cpuperformanceraytracer_b200 import api  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--gpus", type=int, default=0)
a = ap.parse_args()
import torch  # noqa: E402  (device count only)
n = a.gpus or torch.cuda.device_count()
W, H, NTX, NTY, BOUNCES, SPP = 1920, 1080, 2, 4, 8, 1024
want = json.load(open(os.path.join(ROOT, "profiles", "r02_n_full_job_parity.json")))["reference_sha256_16"]


def sha(b):
    return hashlib.sha256(np.ascontiguousarray(b).tobytes()).hexdigest()[:16]


out = {"job": f"Cornell P_v2 {W}x{H}, {BOUNCES} bounces, {SPP} spp, tiles {NTX}x{NTY}", "gpus": n, "reference_sha256_16": want}
with api.Group(list(range(n)), sharding=api.SHARD_TILES, profile=api.PROFILE_V2, num_bounces=BOUNCES) as g:
    g.resize(W, H, NTX, NTY)
    g.render_frames(SPP)
    exact = g.download_target()
    c = g.counters()
out.update(tile_shard_sha256_16=sha(exact), tile_shard_equals_reference_bit_for_bit=sha(exact) == want,
           tile_shard_ms=c.get("last_render_ms"), segments=c["segments"], escapes=c["escapes"])
for name, combine in (("nccl", api.COMBINE_NCCL), ("peer", api.COMBINE_PEER), ("fused", api.COMBINE_FUSED)):
    if n < 2 and combine == api.COMBINE_NCCL:
        continue
    with api.Group(list(range(n)), sharding=api.SHARD_SPP, combine=combine, profile=api.PROFILE_V2, num_bounces=BOUNCES) as g:
        g.resize(W, H, NTX, NTY)
        g.render_frames(SPP)
        b = g.download_target()
        c = g.counters()
    d = np.abs(b.astype(np.float64) - exact)
    out["spp_shard_" + name] = {"max_abs": float(d.max()), "max_rel": float((d / np.maximum(np.abs(exact), 1e-3)).max()),
                                "ms": c.get("last_render_ms"), "segments_equal": c["segments"] == out["segments"]}
print(json.dumps(out), flush=True)
