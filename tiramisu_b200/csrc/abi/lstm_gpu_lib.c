/* libtiramisu_b200_lstm_gpu_lib.so — the symbol `lstm` with the 6-argument ABI of the reference's
 * benchmarks/DNN/blocks/LSTM/gpu_lib (wrapper.h: lstm(Weights, biases, x, y, time_start, time_end)).  The three LSTM
 * benchmarks of the reference all export `lstm` with different arities, so each ABI gets a shared library of its own;
 * the work is done by tiramisu_b200_lstm_gpu_lib in libtiramisu_b200.so. */
struct halide_buffer_t;
int tiramisu_b200_lstm_gpu_lib(struct halide_buffer_t *, struct halide_buffer_t *, struct halide_buffer_t *, struct halide_buffer_t *,
                               struct halide_buffer_t *, struct halide_buffer_t *);
__attribute__((visibility("default"))) int lstm(struct halide_buffer_t *Weights, struct halide_buffer_t *biases, struct halide_buffer_t *x,
                                                struct halide_buffer_t *y, struct halide_buffer_t *time_start, struct halide_buffer_t *time_end)
{
    return tiramisu_b200_lstm_gpu_lib(Weights, biases, x, y, time_start, time_end);
}
