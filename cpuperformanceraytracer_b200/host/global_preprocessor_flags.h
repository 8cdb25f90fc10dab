// global_preprocessor_flags.h -- host-side defaults with the macro names the reference uses for
// its compile-time configuration (reference global_preprocessor_flags.h:3-109).  In the reference
// these can only be changed by recompiling; here they are the DEFAULTS of runtime options
// (B200RenderOptions in demofox_render.h, fields of b200pt_params in include/b200pt.h).
// Define any of them before including demofox_render.h (or with -D) to change a default.
#pragma once

#ifndef RENDER_OFFLINE
#define RENDER_OFFLINE 1  // this host layer only mirrors ApplicationState::RenderOffline
#endif
#ifndef NUM_SAMPLES_PER_FRAME
#define NUM_SAMPLES_PER_FRAME 1  // :30,33
#endif
#ifndef NUM_FRAMES_TO_RENDER
#define NUM_FRAMES_TO_RENDER 600  // :31
#endif
#ifndef RENDER_BUFFER_PIXEL_WIDTH
#define RENDER_BUFFER_PIXEL_WIDTH 1280  // :40
#endif
#ifndef RENDER_BUFFER_PIXEL_HEIGHT
#define RENDER_BUFFER_PIXEL_HEIGHT 720  // :39
#endif
#ifndef USE_ENV_MAP
#define USE_ENV_MAP 1  // :56
#endif
#ifndef USE_ENV_CUBEMAP
#define USE_ENV_CUBEMAP 0  // :57
#endif
#ifndef OUTPUT_TO_SCREEN
#define OUTPUT_TO_SCREEN (!RENDER_OFFLINE)  // :58
#endif
#ifndef ACCUMULATE_FRAMES
#define ACCUMULATE_FRAMES 1  // :60 (always on in the kernel)
#endif
#ifndef USE_FAST_APPROXIMATE_GAMMA
#define USE_FAST_APPROXIMATE_GAMMA 1  // :62
#endif
#ifndef USE_FAST_APPROXIMATE_ACES_TONEMAP
#define USE_FAST_APPROXIMATE_ACES_TONEMAP 1  // :63
#endif
#ifndef USE_FAST_APPROXIMATE_EXP
#define USE_FAST_APPROXIMATE_EXP 1  // :64
#endif
#ifndef USE_UNIT_VECTOR_REJECTION_SAMPLING
#define USE_UNIT_VECTOR_REJECTION_SAMPLING 1  // :65
#endif
#ifndef USE_RANDOM_JITTER_TEXTURE_SAMPLING
#define USE_RANDOM_JITTER_TEXTURE_SAMPLING 1  // :66
#endif
#ifndef NUM_TILES_X
#define NUM_TILES_X 10  // :85
#endif
#ifndef NUM_TILES_Y
#define NUM_TILES_Y 15  // :86
#endif
// NUM_THREADS (:69) has no counterpart: the persistent kernel sizes itself from the SM count.
