// tiramisu/expr.h — expression trees of the tiramisu_b200 front-end.
//
// Mirrors the part of the reference's expression API its generators use
// (/root/reference/include/tiramisu/expr.h:156 `class expr`, :1729 `class var`, :1713 `class sync`,
// :1851-1973 value_cast / cast / memcpy / allocate helpers; cublas_gemm from
// /root/reference/src/tiramisu_expr.cpp:195-237).  The representation is this front-end's own: an
// immutable, shared node per expression, no ISL, no Halide.
#ifndef TIRAMISU_B200_FRONTEND_EXPR_H
#define TIRAMISU_B200_FRONTEND_EXPR_H

#include <cstdint>
#include <memory>
#include <string>
#include <type_traits>
#include <vector>

#include <tiramisu/type.h>

namespace tiramisu {

class buffer;
class computation;
class function;
class expr;
class var;

std::string str_from_tiramisu_type_primitive(primitive_t t);
bool is_float_type(primitive_t t);
bool is_int_type(primitive_t t);
int type_bits(primitive_t t);

class global {
public:
    static void set_default_tiramisu_options();
    static void set_loop_iterator_type(primitive_t t);
    static void set_loop_iterator_default_data_type(primitive_t t) { set_loop_iterator_type(t); }
    static primitive_t get_loop_iterator_data_type();
    static function *get_implicit_function();
    static void set_implicit_function(function *f);
    static bool is_auto_data_mapping_set() { return true; }
    static void set_auto_data_mapping(bool) {}
};

class expr {
public:
    struct node {
        expr_t kind = e_none;
        op_t op = o_none;
        primitive_t dtype = p_none;
        std::string name;            // e_var, o_access/o_call/o_allocate/o_free/o_buffer
        int64_t ival = 0;            // integer / boolean literals
        uint64_t uval = 0;           // p_uint64 literals
        double fval = 0.0;           // floating-point literals
        std::vector<expr> ops;       // operands / call arguments / access indices
    };

    expr() : p_(nullptr) {}
    expr(uint8_t v) { lit_int(p_uint8, v); }
    expr(int8_t v) { lit_int(p_int8, v); }
    expr(uint16_t v) { lit_int(p_uint16, v); }
    expr(int16_t v) { lit_int(p_int16, v); }
    expr(uint32_t v) { lit_int(p_uint32, v); }
    expr(int32_t v) { lit_int(p_int32, v); }
    expr(uint64_t v) { lit_int(p_uint64, (int64_t) v); }
    expr(int64_t v) { lit_int(p_int64, v); }
    expr(long long v) { lit_int(p_int64, (int64_t) v); }
    expr(unsigned long long v) { lit_int(p_uint64, (int64_t) v); }
    expr(float v) { lit_float(p_float32, v); }
    expr(double v) { lit_float(p_float64, v); }
    // exactly `bool` (a plain `expr(bool)` would also accept string literals through pointer-to-bool)
    template <typename T, typename std::enable_if<std::is_same<T, bool>::value, int>::type = 0> expr(T v) { lit_int(p_boolean, v ? 1 : 0); }

    expr(op_t o, expr a);
    expr(op_t o, expr a, expr b);
    expr(op_t o, expr a, expr b, expr c);
    expr(op_t o, std::string name);                                               // o_allocate / o_free / o_buffer
    expr(op_t o, std::string name, std::vector<expr> args, primitive_t t);       // o_call / o_access
    expr(op_t o, primitive_t t, expr a);                                          // o_cast

    static expr make_var(primitive_t t, const std::string &name);
    static expr make_sync();

    bool is_defined() const { return p_ != nullptr && p_->kind != e_none; }
    expr_t get_expr_type() const { return p_ ? p_->kind : e_none; }
    op_t get_op_type() const { return p_ ? p_->op : o_none; }
    primitive_t get_data_type() const { return p_ ? p_->dtype : p_none; }
    const std::string &get_name() const;
    int get_n_arg() const { return p_ ? (int) p_->ops.size() : 0; }
    const expr &get_operand(int i) const { return p_->ops.at(i); }
    const std::vector<expr> &get_operands() const { return p_->ops; }
    bool is_int_literal() const { return p_ && p_->kind == e_val && !is_float_type(p_->dtype); }
    bool is_literal() const { return p_ && p_->kind == e_val; }
    int64_t get_int_val() const { return p_ ? p_->ival : 0; }
    double get_double_val() const { return p_ ? p_->fval : 0.0; }
    /** Typed views of a literal (reference expr.h:537-616). */
    uint8_t get_uint8_value() const { return (uint8_t) get_int_val(); }
    int8_t get_int8_value() const { return (int8_t) get_int_val(); }
    uint16_t get_uint16_value() const { return (uint16_t) get_int_val(); }
    int16_t get_int16_value() const { return (int16_t) get_int_val(); }
    uint32_t get_uint32_value() const { return (uint32_t) get_int_val(); }
    int32_t get_int32_value() const { return (int32_t) get_int_val(); }
    uint64_t get_uint64_value() const { return (uint64_t) get_int_val(); }
    int64_t get_int64_value() const { return get_int_val(); }
    float get_float32_value() const { return (float) get_double_val(); }
    double get_float64_value() const { return get_double_val(); }
    const node *get() const { return p_.get(); }

    /** Replace every variable named names[i] by with[i] (simultaneously). */
    expr substitute(const std::vector<std::string> &names, const std::vector<expr> &with) const;
    /** Structural equality. */
    bool is_equal(const expr &o) const;
    /** Debug text (C-like). */
    std::string to_str() const;
    void dump(bool = true) const;

    expr operator-() const;
    expr operator!() const { return expr(o_logical_not, *this); }
    template <typename T> expr cast() const;

protected:
    std::shared_ptr<const node> p_;
    void lit_int(primitive_t t, int64_t v);
    void lit_float(primitive_t t, double v);
    explicit expr(std::shared_ptr<const node> n) : p_(std::move(n)) {}
    friend class var;
};

/** A literal of type \p t holding \p v (reference: expr.h:1851 value_cast). */
expr value_cast_i(primitive_t t, int64_t v);
expr value_cast_f(primitive_t t, double v);
template <typename T> expr value_cast(primitive_t t, T v)
{
    if (std::is_floating_point<T>::value || is_float_type(t)) return is_float_type(t) ? value_cast_f(t, (double) v) : value_cast_i(t, (int64_t) v);
    return value_cast_i(t, (int64_t) v);
}
expr cast(primitive_t t, const expr &e);

// A literal operand takes the type of the expression it is combined with (reference expr.h:1888-1970).
template <typename T> using tb_arith = typename std::enable_if<std::is_arithmetic<T>::value, int>::type;
expr lit_like(const expr &e, double fv, int64_t iv, bool is_fp);
template <typename T> expr lit_like(const expr &e, T v)
{
    return lit_like(e, (double) v, (int64_t) v, std::is_floating_point<T>::value);
}

#define TB_BINOP(sym, op)                                                                         \
    expr operator sym(const expr &a, const expr &b);                                              \
    template <typename T, tb_arith<T> = 0> expr operator sym(const expr &a, T b) { return a sym lit_like(a, b); } \
    template <typename T, tb_arith<T> = 0> expr operator sym(T a, const expr &b) { return lit_like(b, a) sym b; }
TB_BINOP(+, o_add)
TB_BINOP(-, o_sub)
TB_BINOP(*, o_mul)
TB_BINOP(/, o_div)
TB_BINOP(%, o_mod)
TB_BINOP(==, o_eq)
TB_BINOP(!=, o_ne)
TB_BINOP(<, o_lt)
TB_BINOP(<=, o_le)
TB_BINOP(>, o_gt)
TB_BINOP(>=, o_ge)
TB_BINOP(>>, o_right_shift)
TB_BINOP(<<, o_left_shift)
#undef TB_BINOP
expr operator&&(const expr &a, const expr &b);
expr operator||(const expr &a, const expr &b);

/** Loop iterator / named scalar.  `var("i", 0, N)` carries the half-open range [0, N). */
class var : public expr {
public:
    var();
    var(std::string name);
    var(std::string name, bool save);
    var(primitive_t t, std::string name);
    var(primitive_t t, std::string name, bool save);
    var(std::string name, expr lower_bound, expr upper_bound);
    const expr &get_lower() const { return lower_; }
    const expr &get_upper() const { return upper_; }
    bool has_range() const { return lower_.is_defined() && upper_.is_defined(); }

private:
    expr lower_, upper_;
};

/** Barrier between the threads of a block (reference expr.h:1713). */
class sync : public expr {
public:
    sync();
};

expr memcpy(const buffer &from, const buffer &to);
expr allocate(const buffer &b);
expr cuda_stream_synchronize();
expr cublas_gemm(const buffer &A, const buffer &B, buffer &C, expr M, expr N, expr K,
                 expr alpha = 1.f, expr beta = 0.f, expr ldA = 0, expr ldB = 0, expr ldC = 0,
                 expr offsetA = 0, expr offsetB = 0, expr offsetC = 0,
                 expr transposeA = false, expr transposeB = false);

}  // namespace tiramisu

#endif
