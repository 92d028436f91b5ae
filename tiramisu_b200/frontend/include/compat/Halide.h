/* Halide.h — generator-side stand-in: the reference's old-style generators include Halide.h although the
 * functions they call are tiramisu's; this front-end has no Halide dependency (the wrappers' Halide::Buffer
 * stand-in lives in tests/compat/include). */
#ifndef TIRAMISU_B200_FRONTEND_HALIDE_STANDIN_H
#define TIRAMISU_B200_FRONTEND_HALIDE_STANDIN_H
namespace Halide { }
#endif
