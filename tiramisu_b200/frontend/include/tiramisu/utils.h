// tiramisu/utils.h — generator-side part only.  The wrapper-side helpers (init_buffer,
// compare_buffers, print_test_results; reference include/tiramisu/utils.h) need Halide::Buffer and
// stay with the wrappers' own include path.
#ifndef TIRAMISU_B200_FRONTEND_UTILS_H
#define TIRAMISU_B200_FRONTEND_UTILS_H
// Generators include their wrapper header (which declares the generated function over halide_buffer_t*);
// pointers to incomplete types are all those declarations need.
struct halide_buffer_t;
struct halide_filter_metadata_t;
#endif
