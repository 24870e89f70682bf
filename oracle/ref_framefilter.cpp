/* ref_framefilter.cpp -- compiles the reference's SEA integral primitives (static functions of
 * encoder/framefilter.cpp:38-156, bound by setupSeaIntegralPrimitives_c) into oracle/_ref/libx265ref_<depth>.so.
 * The reference source is included from where it lies (-I$(REF)/encoder); nothing is copied.  The rest of that file
 * (FrameFilter, which needs the whole encoder) is compiled too but never linked: this TU is built with
 * -ffunction-sections -fdata-sections -fvisibility=hidden and the library is linked with --gc-sections, so only what
 * setupSeaIntegralPrimitives_c references survives.  Test infrastructure, never shipped. */
#include "framefilter.cpp"
