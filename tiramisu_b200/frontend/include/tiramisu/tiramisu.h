// tiramisu/tiramisu.h — umbrella header of the tiramisu_b200 front-end.
//
// Generators written for the reference include this one header (same role as the reference's
// include/tiramisu/tiramisu.h) and get the whole generator-side API: the enumerations, expression trees,
// function / buffer / computation / input / constant / view, `block`, and the no-op tracing macros.  Unlike the
// reference's, nothing here pulls in ISL, Halide or LLVM headers; the standard headers generators tend to rely
// on transitively (<cassert>, <cmath>, <cstring>, <algorithm>) come with core.h.
#ifndef TIRAMISU_B200_FRONTEND_TIRAMISU_H
#define TIRAMISU_B200_FRONTEND_TIRAMISU_H

#include <tiramisu/type.h>      // expr_t, primitive_t, op_t, argument_t
#include <tiramisu/expr.h>      // expr, var, sync, memcpy / allocate / cublas_gemm helpers
#include <tiramisu/core.h>      // function, buffer, computation, input, constant, view, init(), codegen()
#include <tiramisu/block.h>     // block: one scheduling command applied to several computations
#include <tiramisu/debug.h>     // DEBUG* / ERROR macros (tracing is a no-op here)
#include <tiramisu/utils.h>     // forward declarations wrapper headers need

#endif
