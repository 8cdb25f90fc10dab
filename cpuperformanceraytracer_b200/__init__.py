"""cpuperformanceraytracer_b200 -- B200-native engine for the path-tracing hot path of
torgeiba/CPUPerformanceRayTracer (demofox per-pixel path loop + env lookup + accumulation).

The implementation is libb200pt.so: hand-written sm_100a CUDA behind the C ABI of
include/b200pt.h.  `api` is the ctypes handle; `build` compiles the library in-tree.
"""
from . import api  # noqa: F401
from .api import (ACCUM_RUNNING_AVERAGE, ACCUM_SUM, ENV_CUBEMAP, ENV_EQUIRECT, ENV_NONE, MATH_FAST, MATH_PARITY,  # noqa: F401
                  PROFILE_OPT_V4, PROFILE_SIMT_TEXTURED, PROFILE_V2, PROFILE_V3_REDO, SAMPLER_BILINEAR, SAMPLER_POINT, SAMPLER_RANDOM,
                  B200PTError, Renderer, detile, load_library)
