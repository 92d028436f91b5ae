// tiramisu/type.h — enumerations of the tiramisu_b200 front-end.
//
// Source-compatible with the names the reference's generators use
// (/root/reference/include/tiramisu/type.h:14-139: expr_t, primitive_t, op_t, argument_t); the
// values and grouping are this front-end's own.
#ifndef TIRAMISU_B200_FRONTEND_TYPE_H
#define TIRAMISU_B200_FRONTEND_TYPE_H

namespace tiramisu {

enum expr_t { e_val, e_var, e_sync, e_op, e_none };

enum primitive_t {
    p_uint8, p_uint16, p_uint32, p_uint64,
    p_int8, p_int16, p_int32, p_int64,
    p_float32, p_float64,
    p_boolean,
    p_async, p_wait_ptr, p_void_ptr,
    p_none
};

enum op_t {
    // unary
    o_minus, o_logical_not, o_floor, o_ceil, o_round, o_trunc, o_abs, o_sqrt, o_expo, o_log,
    o_sin, o_cos, o_tan, o_asin, o_acos, o_atan, o_sinh, o_cosh, o_tanh, o_asinh, o_acosh, o_atanh,
    o_cast, o_address,
    // named (operand is a buffer name)
    o_allocate, o_free, o_buffer,
    // binary
    o_add, o_sub, o_mul, o_div, o_mod,
    o_logical_and, o_logical_or,
    o_eq, o_ne, o_le, o_lt, o_ge, o_gt,
    o_max, o_min, o_right_shift, o_left_shift,
    o_memcpy,
    // ternary
    o_select, o_cond, o_lerp,
    // name + argument list
    o_call, o_access, o_address_of, o_lin_index, o_type, o_dummy,
    o_none
};

enum argument_t { a_input, a_output, a_temporary };

enum rank_t { r_sender, r_receiver };

}  // namespace tiramisu

#endif
