/* TEST INFRASTRUCTURE: a small instance of benchmarks/DNN/blocks/vggBlock/cpu/dense/configure.h (same macros, same derived
 * quantities) so that the UNMODIFIED vgg_block_generator_tiramisu.cpp can be lowered at a size the host emulation of
 * generated sources finishes in seconds.  N = 16: two x blocks of 8 for the first convolution, X2_BOUND = 15 and a
 * non-empty conclude column for the second (N % X2_BLOCKING = 1, the generator's `(N-1)/2` branch). */
#ifndef __CONV_CONF_HEADER_
#define __CONV_CONF_HEADER_
#define BATCH_SIZE 2
#define N 16
#define FIn 3
#define FOut 16
#define K 3
#define FIN2_BLOCKING 8
#define FIN2_NB_BLOCKS FOut/FOUT_BLOCKING
#define FOUT_BLOCKING 8
#define FOUT_NB_BLOCKS FOut/FOUT_BLOCKING
#define X1_BLOCKING 8
#define X1_NB_BLOCKS N/X1_BLOCKING
#define X2_BLOCKING 3
#define X2_NB_BLOCKS N/X2_BLOCKING
#define X2_BOUND X2_NB_BLOCKS*X2_BLOCKING
#define VEC_LEN 8
#endif
