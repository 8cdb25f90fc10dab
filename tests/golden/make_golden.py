"""make_golden.py -- regenerates tests/golden/*.npz from the REFERENCE's own code.

Runs the binaries that oracle/ref_build/build_ref.sh compiles in place from /root/reference
("exact" mode: see oracle/ref_build/shim.h) and stores their f32 accumulation buffers.  The
fixtures pin oracle/pt_oracle.c (tests/test_oracle_golden.py) and the CUDA path
(tests/test_gpu_parity.py) on machines where /root/reference does not exist.

    python tests/golden/make_golden.py                   # needs /root/reference (this container)
    python tests/golden/make_golden.py --only NAME ...   # (re)generate the named fixtures only, keep the rest of the index

Env textures are the deterministic synthetic ones of oracle.pyoracle.synthetic_env, regenerated
by the tests from (width, height), so no texture data is stored.
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import pyoracle as po  # noqa: E402

# name, ref binary, oracle profile, env kind, env sampler, env (w,h) or None, W, H, ntx, nty, bounces, frames
CASES = [
    ("v2_b4", "ref_v2_exact", po.PROFILE_V2, po.ENV_NONE, po.SAMPLER_POINT, None, 96, 64, 2, 4, 4, 6),
    ("v2_b8", "ref_v2_exact", po.PROFILE_V2, po.ENV_NONE, po.SAMPLER_POINT, None, 96, 64, 2, 4, 8, 6),
    ("v2_b16", "ref_v2_exact", po.PROFILE_V2, po.ENV_NONE, po.SAMPLER_POINT, None, 64, 64, 4, 2, 16, 3),
    ("simt_textured_b4", "ref_simt_textured_exact", po.PROFILE_SIMT_TEXTURED, po.ENV_EQUIRECT, po.SAMPLER_POINT,
     (128, 64), 96, 64, 2, 4, 4, 6),
    ("v4_equirect_random", "ref_v4_equirect_random_exact", po.PROFILE_V4, po.ENV_EQUIRECT, po.SAMPLER_RANDOM,
     (128, 64), 128, 72, 4, 6, 8, 6),
    ("v4_equirect_bilinear", "ref_v4_equirect_bilinear_exact", po.PROFILE_V4, po.ENV_EQUIRECT, po.SAMPLER_BILINEAR,
     (128, 64), 128, 72, 4, 6, 8, 6),
    ("v4_cubemap_random", "ref_v4_cubemap_random_exact", po.PROFILE_V4, po.ENV_CUBEMAP, po.SAMPLER_RANDOM,
     (32, 192), 128, 72, 4, 6, 8, 6),
    ("v4_cubemap_bilinear", "ref_v4_cubemap_bilinear_exact", po.PROFILE_V4, po.ENV_CUBEMAP, po.SAMPLER_BILINEAR,
     (32, 192), 128, 72, 4, 6, 8, 6),
    ("v3redo", "ref_v3redo_exact", po.PROFILE_V3REDO, po.ENV_EQUIRECT, po.SAMPLER_BILINEAR, (128, 64), 128, 72, 2, 4, 8, 6),
    ("v3redo_scene0", "ref_v3redo_scene0_exact", po.PROFILE_V3REDO_SCENE0, po.ENV_EQUIRECT, po.SAMPLER_BILINEAR, (128, 64), 128, 72, 2, 4, 8, 6),
    ("v4_b16", "ref_v4_equirect_random_exact", po.PROFILE_V4, po.ENV_EQUIRECT, po.SAMPLER_RANDOM,
     (128, 64), 64, 40, 2, 5, 16, 4),
]
# the NON-default sides of global_preprocessor_flags.h:63-65 (build_ref.sh compiles one binary per combination):
# name, ref binary, oracle v4_flags (pyoracle.V4_EXACT_EXP | V4_SINCOS_UNIT_VECTORS)
FLAG_CASES = [
    ("v4_exact_exp", "ref_v4_equirect_random_expexact_exact", po.V4_EXACT_EXP),
    ("v4_sincos_unit_vectors", "ref_v4_equirect_random_sincos_exact", po.V4_SINCOS_UNIT_VECTORS),
    ("v4_all_exact", "ref_v4_equirect_random_allexact_exact", po.V4_EXACT_EXP | po.V4_SINCOS_UNIT_VECTORS),
]


def main():
    only = sys.argv[sys.argv.index("--only") + 1:] if "--only" in sys.argv else None
    index = []
    if only is not None:
        with open(os.path.join(HERE, "index.json")) as f:
            index = [e for e in json.load(f) if e["name"] not in only]
    cases = [c + (0,) for c in CASES]
    cases += [(name, binary, po.PROFILE_V4, po.ENV_EQUIRECT, po.SAMPLER_RANDOM, (128, 64), 128, 72, 4, 6, 8, 6, flags)
              for (name, binary, flags) in FLAG_CASES]
    for (name, binary, profile, ek, es, envshape, W, H, ntx, nty, bounces, frames, v4_flags) in cases:
        if only is not None and name not in only:
            continue
        env = po.synthetic_env(*envshape) if envshape else None
        res = po.run_ref(binary, W, H, ntx, nty, frames, bounces=bounces, env=env)
        buf = res["buffer"]
        assert buf.size == W * H * 3 and np.isfinite(buf).all()
        # a second buffer: the same render continued for 2 more frames from the stored state
        cont = po.run_ref(binary, W, H, ntx, nty, 2, bounces=bounces, env=env, start_frame=frames, target=buf)["buffer"]
        np.savez_compressed(os.path.join(HERE, name + ".npz"), buffer=buf, continued=cont)
        index.append(dict(name=name, binary=binary, profile=profile, env_kind=ek, env_sampler=es,
                          env_shape=list(envshape) if envshape else None, width=W, height=H, ntx=ntx, nty=nty,
                          bounces=bounces, frames=frames, continued_frames=2, v4_flags=v4_flags))
        print(name, "mean", float(buf.mean()))
    # LDR goldens from the reference's CopyOutputToFile (single queue participant: deterministic); the second one is the
    # build with USE_FAST_APPROXIMATE_ACES_TONEMAP 0 (and the two shading switches off as well), the last two are builds with
    # USE_FAST_APPROXIMATE_GAMMA 0 alone and together with the exact ACES curve (pow_ps = portable_math.h's pm_powf).
    # `ldr_mode` = the oracle_resolve_ldr mode bits (2 exact ACES, 4 exact gamma) that reproduce the file
    env = po.synthetic_env(128, 64)
    for name, binary, exact_aces in (("v4_ldr", "ref_v4_equirect_random_exact", 0), ("v4_ldr_exact_aces", "ref_v4_equirect_random_allexact_exact", 1),
                                     ("v4_ldr_exact_gamma", "ref_v4_equirect_random_gammaexact_exact", 0),
                                     ("v4_ldr_exact_aces_gamma", "ref_v4_equirect_random_ldrexact_exact", 1)):
        if only is not None and name not in only:
            continue
        res = po.run_ref(binary, 128, 72, 4, 6, 6, bounces=8, env=env, threads=1, ldr=True)
        np.savez_compressed(os.path.join(HERE, name + ".npz"), buffer=res["buffer"], ldr=res["ldr"])
        index.append(dict(name=name, binary=binary, kind="ldr", width=128, height=72, ntx=4, nty=6, bounces=8, frames=6,
                          env_shape=[128, 64], exact_aces=exact_aces, ldr_mode=2 * exact_aces + (4 if "gamma" in name else 0)))
    with open(os.path.join(HERE, "index.json"), "w") as f:
        json.dump(index, f, indent=1)


if __name__ == "__main__":
    main()
