/* halide_buffer_t — the argument type of every Tiramisu-generated function.
 *
 * ABI mirror (56 bytes on x86-64) of the struct the reference's wrappers pass
 * (reference: 3rdParty/Halide/src/runtime/HalideRuntime.h:1434-1460; used by
 * every `extern "C" int <fn>(halide_buffer_t*...)` that function::codegen emits,
 * src/tiramisu_codegen_halide.cpp:4013-4017).  This is a re-declaration of the
 * layout, not the Halide header: only the fields the hot path reads are named
 * the same way so that wrapper code compiled against the real HalideRuntime.h
 * and code compiled against this file agree byte-for-byte.
 *
 * If the real Halide runtime header was included first we step aside.
 */
#ifndef TIRAMISU_B200_HALIDE_BUFFER_H
#define TIRAMISU_B200_HALIDE_BUFFER_H

#include <stdint.h>

#ifndef HALIDE_HALIDERUNTIME_H

#ifdef __cplusplus
extern "C" {
#endif

struct halide_device_interface_t;

typedef enum halide_type_code_t {
    halide_type_int = 0,
    halide_type_uint = 1,
    halide_type_float = 2,
    halide_type_handle = 3,
    halide_type_bfloat = 4
} halide_type_code_t;

typedef struct halide_type_t {
    uint8_t code;   /* halide_type_code_t */
    uint8_t bits;
    uint16_t lanes;
} halide_type_t;

typedef struct halide_dimension_t {
    int32_t min, extent, stride;
    uint32_t flags;
} halide_dimension_t;

typedef struct halide_buffer_t {
    uint64_t device;                                   /* device handle (we accept a raw CUDA pointer here) */
    const struct halide_device_interface_t *device_interface;
    uint8_t *host;                                     /* host pointer; the only field the reference reads */
    uint64_t flags;
    struct halide_type_t type;
    int32_t dimensions;
    halide_dimension_t *dim;                           /* innermost dimension first */
    void *padding;
} halide_buffer_t;

#ifdef __cplusplus
}
#endif

#endif /* HALIDE_HALIDERUNTIME_H */

#if defined(__cplusplus) && __cplusplus >= 201103L
static_assert(sizeof(halide_buffer_t) == 56, "halide_buffer_t ABI drifted");
static_assert(sizeof(halide_dimension_t) == 16, "halide_dimension_t ABI drifted");
#endif

#endif /* TIRAMISU_B200_HALIDE_BUFFER_H */
