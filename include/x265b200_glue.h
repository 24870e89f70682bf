/*
 * x265b200_glue.h -- entry points of the per-depth table fillers
 * x265-mod-by-patman_b200/lib/libx265b200_glue_{8,10,12}.so (csrc/setup_b200_primitives.cpp).
 *
 * One glue library exists per X265_DEPTH because the reference bakes the pixel type into the build
 * (source/CMakeLists.txt:787-798) and EncoderPrimitives' typedefs depend on it.  A C++ caller inside the
 * encoder uses  X265_NS::setupB200Primitives(EncoderPrimitives&)  directly, exactly where
 * x265_setup_primitives (source/common/primitives.cpp:355-367) layers setupIntrinsicPrimitives /
 * setupAssemblyPrimitives over the C table; these C handles are for drivers that cannot name the C++
 * symbol (ctypes, dlopen, the TestBench driver oracle/harness_main.cpp).
 */
#ifndef X265B200_GLUE_H
#define X265B200_GLUE_H

#include "x265b200.h"

#ifdef __cplusplus
extern "C" {
#endif

/* Opens the process-wide context on `device` (once, thread-safe) and overwrites every hot-path and adjacent slot of
 * `table` (an EncoderPrimitives of this library's bit depth) with a thunk onto the host entries of x265b200.h.
 * Replaces the setupAssemblyPrimitives call of source/common/primitives.cpp:363.  Returns 0 or X265B200_ERR_*. */
int x265b200_setup_primitives(void* table, int device);

/* the context the slots run on (NULL before x265b200_setup_primitives): batched entries of x265b200.h take it */
x265b200_ctx* x265b200_glue_context(void);

/* X265_DEPTH this glue library was compiled for (8, 10 or 12) */
int x265b200_glue_depth(void);

/* Sticky status of the slots' context, with the first error's text in *message (may be NULL).  A slot has no way to
 * return an error, so the encoder polls this once per frame (e.g. at the top of FrameEncoder::compressFrame,
 * source/encoder/frameencoder.cpp) and aborts the encode when it is non-zero. */
int x265b200_glue_status(const char** message);

#ifdef __cplusplus
}
#endif
#endif /* X265B200_GLUE_H */
