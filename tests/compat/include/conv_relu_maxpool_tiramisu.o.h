// Stand-in for the header Halide emits next to the generated object
// (benchmarks/DNN/blocks/Conv-ReLU-MaxPool/cpu/dense/wrapper_nn_block.cpp:8 includes it).  TEST INFRASTRUCTURE.
#pragma once
#include "HalideRuntime.h"
#ifdef __cplusplus
extern "C" {
#endif
int conv_relu_maxpool_block(struct halide_buffer_t *input, struct halide_buffer_t *filter, struct halide_buffer_t *bias,
                            struct halide_buffer_t *output);
#ifdef __cplusplus
}
#endif
