/* TEST INFRASTRUCTURE: a small instance of benchmarks/DNN/layers/convolution/direct/cpu/configure.h (same macros, same
 * derived quantities) so that the UNMODIFIED conv_layer_generator_tiramisu.cpp can be lowered at a size the host emulation
 * of generated sources finishes in seconds.  N = 10 keeps a non-empty conclude column (X_BOUND = 9), FOut = 64 keeps the
 * fout_b split (8 blocks, factor 4). */
#ifndef __CONV_CONF_HEADER_
#define __CONV_CONF_HEADER_
#define BATCH_SIZE 2
#define N 10
#define FIn 32
#define FOut 64
#define K 3
#define FOUT_BLOCKING 8
#define FOUT_NB_BLOCKS FOut/FOUT_BLOCKING
#define FOUT_B_SPLIT_FACTOR 4
#define FIN_BLOCKING 16
#define FIN_NB_BLOCKS FIn/FIN_BLOCKING
#define X_BLOCKING 3
#define X_NB_BLOCKS N/X_BLOCKING
#define X_BOUND X_NB_BLOCKS*X_BLOCKING
#endif
