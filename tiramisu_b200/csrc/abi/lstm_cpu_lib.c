/* libtiramisu_b200_lstm_cpu_lib.so — the symbol `lstm` with the 6-argument ABI of the reference's
 * benchmarks/DNN/blocks/LSTM/cpu_lib (wrapper.cpp:65-72: lstm(weights, biases, x, h, c, out)); see lstm_gpu_lib.c. */
struct halide_buffer_t;
int tiramisu_b200_lstm_cpu_lib(struct halide_buffer_t *, struct halide_buffer_t *, struct halide_buffer_t *, struct halide_buffer_t *,
                               struct halide_buffer_t *, struct halide_buffer_t *);
__attribute__((visibility("default"))) int lstm(struct halide_buffer_t *weights, struct halide_buffer_t *biases, struct halide_buffer_t *x,
                                                struct halide_buffer_t *h, struct halide_buffer_t *c, struct halide_buffer_t *out)
{
    return tiramisu_b200_lstm_cpu_lib(weights, biases, x, h, c, out);
}
