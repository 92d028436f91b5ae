// expr.cpp — expression nodes of the tiramisu_b200 front-end (see include/tiramisu/expr.h).
#include <tiramisu/core.h>

#include <cstdio>
#include <cstdlib>
#include <iostream>
#include <sstream>

namespace tiramisu {

static primitive_t g_iter_type = p_int32;
static function *g_implicit = nullptr;

void global::set_default_tiramisu_options() { g_iter_type = p_int32; }
void global::set_loop_iterator_type(primitive_t t) { g_iter_type = t; }
primitive_t global::get_loop_iterator_data_type() { return g_iter_type; }
function *global::get_implicit_function() { return g_implicit; }
void global::set_implicit_function(function *f) { g_implicit = f; }

bool is_float_type(primitive_t t) { return t == p_float32 || t == p_float64; }
bool is_int_type(primitive_t t) { return t <= p_int64 || t == p_boolean; }
int type_bits(primitive_t t)
{
    switch (t) {
    case p_uint8: case p_int8: case p_boolean: return 8;
    case p_uint16: case p_int16: return 16;
    case p_uint32: case p_int32: case p_float32: return 32;
    case p_uint64: case p_int64: case p_float64: return 64;
    default: return 0;
    }
}
std::string str_from_tiramisu_type_primitive(primitive_t t)
{
    switch (t) {
    case p_uint8: return "uint8"; case p_uint16: return "uint16"; case p_uint32: return "uint32"; case p_uint64: return "uint64";
    case p_int8: return "int8"; case p_int16: return "int16"; case p_int32: return "int32"; case p_int64: return "int64";
    case p_float32: return "float32"; case p_float64: return "float64"; case p_boolean: return "bool";
    case p_async: return "async"; case p_wait_ptr: return "wait_ptr"; case p_void_ptr: return "void*";
    default: return "none";
    }
}

static int64_t wrap_int(primitive_t t, int64_t v)
{
    switch (t) {
    case p_uint8: return (uint8_t) v; case p_int8: return (int8_t) v;
    case p_uint16: return (uint16_t) v; case p_int16: return (int16_t) v;
    case p_uint32: return (uint32_t) v; case p_int32: return (int32_t) v;
    case p_boolean: return v != 0;
    default: return v;
    }
}

void expr::lit_int(primitive_t t, int64_t v)
{
    auto n = std::make_shared<node>();
    n->kind = e_val; n->dtype = t; n->ival = wrap_int(t, v); n->uval = (uint64_t) v; n->fval = (double) n->ival;
    p_ = n;
}
void expr::lit_float(primitive_t t, double v)
{
    auto n = std::make_shared<node>();
    n->kind = e_val; n->dtype = t; n->fval = (t == p_float32) ? (double) (float) v : v; n->ival = (int64_t) v;
    p_ = n;
}
expr value_cast_i(primitive_t t, int64_t v)
{
    switch (t) {
    case p_uint8: return expr((uint8_t) v); case p_int8: return expr((int8_t) v);
    case p_uint16: return expr((uint16_t) v); case p_int16: return expr((int16_t) v);
    case p_uint32: return expr((uint32_t) v); case p_int32: return expr((int32_t) v);
    case p_uint64: return expr((uint64_t) v); case p_int64: return expr((int64_t) v);
    case p_boolean: return expr(v != 0);
    case p_float32: return expr((float) v); case p_float64: return expr((double) v);
    default: return expr((int32_t) v);
    }
}
expr value_cast_f(primitive_t t, double v)
{
    if (t == p_float32) return expr((float) v);
    if (t == p_float64) return expr(v);
    return value_cast_i(t, (int64_t) v);
}
expr lit_like(const expr &e, double fv, int64_t iv, bool is_fp)
{
    primitive_t t = e.get_data_type();
    if (t == p_none || t > p_boolean) return is_fp ? expr(fv) : expr((int32_t) iv);
    if (is_float_type(t)) return value_cast_f(t, is_fp ? fv : (double) iv);
    if (is_fp) return value_cast_f(p_float32, fv);      // int expression combined with a fp literal
    return value_cast_i(t == p_boolean ? p_int32 : t, iv);
}

const std::string &expr::get_name() const
{
    static const std::string empty;
    return p_ ? p_->name : empty;
}

static bool is_cmp(op_t o) { return o == o_eq || o == o_ne || o == o_le || o == o_lt || o == o_ge || o == o_gt; }

static primitive_t binary_type(op_t o, const expr &a, const expr &b)
{
    if (is_cmp(o) || o == o_logical_and || o == o_logical_or) return p_boolean;
    primitive_t ta = a.get_data_type(), tb = b.get_data_type();
    if (ta == tb) return ta;
    if (ta == p_none) return tb;
    if (tb == p_none) return ta;
    if (o == o_right_shift || o == o_left_shift) return ta;
    if (a.is_literal() && !b.is_literal()) return tb;
    if (b.is_literal() && !a.is_literal()) return ta;
    if (is_float_type(ta) != is_float_type(tb)) return is_float_type(ta) ? ta : tb;
    return type_bits(ta) >= type_bits(tb) ? ta : tb;
}

static int64_t floordiv(int64_t a, int64_t b) { int64_t q = a / b; return ((a % b != 0) && ((a < 0) != (b < 0))) ? q - 1 : q; }

expr::expr(op_t o, expr a)
{
    auto n = std::make_shared<node>();
    n->kind = e_op; n->op = o; n->dtype = (o == o_logical_not) ? p_boolean : a.get_data_type();
    n->ops = {a};
    p_ = n;
}

expr::expr(op_t o, expr a, expr b)
{
    primitive_t t = binary_type(o, a, b);
    if (o != o_memcpy) {
        // bring a literal operand to the type of the other one
        if (a.is_literal() && a.get_data_type() != t && !is_cmp(o)) a = is_float_type(t) ? value_cast_f(t, a.is_int_literal() ? (double) a.get_int_val() : a.get_double_val()) : (a.is_int_literal() ? value_cast_i(t, a.get_int_val()) : a);
        if (b.is_literal() && b.get_data_type() != t && !is_cmp(o) && o != o_right_shift && o != o_left_shift) b = is_float_type(t) ? value_cast_f(t, b.is_int_literal() ? (double) b.get_int_val() : b.get_double_val()) : (b.is_int_literal() ? value_cast_i(t, b.get_int_val()) : b);
        // fold integer literals so extents and divisibility stay decidable
        if (a.is_int_literal() && b.is_int_literal() && a.get_data_type() != p_boolean && t != p_uint64) {
            int64_t x = a.get_int_val(), y = b.get_int_val();
            bool ok = true; int64_t r = 0;
            switch (o) {
            case o_add: r = x + y; break;
            case o_sub: r = x - y; break;
            case o_mul: r = x * y; break;
            case o_div: if (y == 0) ok = false; else r = (x >= 0 && y > 0) ? x / y : x / y; break;
            case o_mod: if (y == 0) ok = false; else r = x % y; break;
            case o_max: r = x > y ? x : y; break;
            case o_min: r = x < y ? x : y; break;
            default: ok = false;
            }
            if (ok) { *this = value_cast_i(t, r); return; }
        }
        // (x + c1) +- c2  ->  x + (c1 +- c2): keeps `(lo + extent) + -1 + 1` style ISL bounds readable
        if (is_int_type(t) && (o == o_add || o == o_sub) && b.is_int_literal() && a.get_expr_type() == e_op &&
            a.get_op_type() == o_add && a.get_operand(1).is_int_literal() && a.get_data_type() == t) {
            int64_t c = a.get_operand(1).get_int_val() + (o == o_add ? b.get_int_val() : -b.get_int_val());
            *this = expr(o_add, a.get_operand(0), value_cast_i(t, c));
            return;
        }
        // x + 0, x * 1, x - 0, 0 + x, 1 * x, x * 0
        if (is_int_type(t)) {
            if (b.is_int_literal() && ((b.get_int_val() == 0 && (o == o_add || o == o_sub)) || (b.get_int_val() == 1 && (o == o_mul || o == o_div)))) { *this = a; return; }
            if (a.is_int_literal() && ((a.get_int_val() == 0 && o == o_add) || (a.get_int_val() == 1 && o == o_mul))) { *this = b; return; }
        }
        (void) floordiv;
    }
    auto n = std::make_shared<node>();
    n->kind = e_op; n->op = o; n->dtype = (o == o_memcpy) ? p_none : t;
    n->ops = {a, b};
    p_ = n;
}

expr::expr(op_t o, expr a, expr b, expr c)
{
    auto n = std::make_shared<node>();
    n->kind = e_op; n->op = o;
    n->dtype = (o == o_select || o == o_cond) ? binary_type(o_add, b, c) : a.get_data_type();
    n->ops = {a, b, c};
    p_ = n;
}

expr::expr(op_t o, std::string name)
{
    auto n = std::make_shared<node>();
    n->kind = e_op; n->op = o; n->name = std::move(name); n->dtype = p_none;
    p_ = n;
}

expr::expr(op_t o, std::string name, std::vector<expr> args, primitive_t t)
{
    auto n = std::make_shared<node>();
    n->kind = e_op; n->op = o; n->name = std::move(name); n->dtype = t; n->ops = std::move(args);
    p_ = n;
}

expr::expr(op_t o, primitive_t t, expr a)
{
    if (a.is_literal()) {   // fold literal casts
        *this = a.is_int_literal() ? (is_float_type(t) ? value_cast_f(t, (double) a.get_int_val()) : value_cast_i(t, a.get_int_val()))
                                   : value_cast_f(t, a.get_double_val());
        return;
    }
    auto n = std::make_shared<node>();
    n->kind = e_op; n->op = o; n->dtype = t; n->ops = {a};
    p_ = n;
}

expr expr::make_var(primitive_t t, const std::string &name)
{
    auto n = std::make_shared<node>();
    n->kind = e_var; n->dtype = t; n->name = name;
    return expr(std::shared_ptr<const node>(n));
}
expr expr::make_sync()
{
    auto n = std::make_shared<node>();
    n->kind = e_sync; n->dtype = p_none;
    return expr(std::shared_ptr<const node>(n));
}

expr cast(primitive_t t, const expr &e) { return e.get_data_type() == t ? e : expr(o_cast, t, e); }

expr expr::operator-() const
{
    if (is_int_literal()) return value_cast_i(get_data_type(), -get_int_val());
    if (is_literal()) return value_cast_f(get_data_type(), -get_double_val());
    return expr(o_minus, *this);
}

#define TB_DEF(sym, op) expr operator sym(const expr &a, const expr &b) { return expr(op, a, b); }
TB_DEF(+, o_add) TB_DEF(-, o_sub) TB_DEF(*, o_mul) TB_DEF(/, o_div) TB_DEF(%, o_mod)
TB_DEF(==, o_eq) TB_DEF(!=, o_ne) TB_DEF(<, o_lt) TB_DEF(<=, o_le) TB_DEF(>, o_gt) TB_DEF(>=, o_ge)
TB_DEF(>>, o_right_shift) TB_DEF(<<, o_left_shift) TB_DEF(&&, o_logical_and) TB_DEF(||, o_logical_or)
#undef TB_DEF

expr expr::substitute(const std::vector<std::string> &names, const std::vector<expr> &with) const
{
    if (!p_) return *this;
    if (p_->kind == e_var) {
        for (size_t i = 0; i < names.size(); i++)
            if (names[i] == p_->name) return with[i];
        return *this;
    }
    if (p_->kind != e_op || p_->ops.empty()) return *this;
    bool changed = false;
    std::vector<expr> ops;
    ops.reserve(p_->ops.size());
    for (const expr &o : p_->ops) {
        expr s = o.substitute(names, with);
        changed |= (s.get() != o.get());
        ops.push_back(s);
    }
    if (!changed) return *this;
    switch (p_->op) {
    case o_call: case o_access: case o_address_of: case o_lin_index:
        return expr(p_->op, p_->name, ops, p_->dtype);
    case o_cast:
        return expr(o_cast, p_->dtype, ops[0]);
    default:
        if (ops.size() == 1) return expr(p_->op, ops[0]);
        if (ops.size() == 2) return expr(p_->op, ops[0], ops[1]);
        return expr(p_->op, ops[0], ops[1], ops[2]);
    }
}

bool expr::is_equal(const expr &o) const
{
    if (p_.get() == o.p_.get()) return true;
    if (!p_ || !o.p_) return false;
    const node &a = *p_, &b = *o.p_;
    if (a.kind != b.kind || a.dtype != b.dtype) return false;
    switch (a.kind) {
    case e_val: return is_float_type(a.dtype) ? a.fval == b.fval : a.ival == b.ival;
    case e_var: return a.name == b.name;
    case e_sync: return true;
    case e_op:
        if (a.op != b.op || a.name != b.name || a.ops.size() != b.ops.size()) return false;
        for (size_t i = 0; i < a.ops.size(); i++)
            if (!a.ops[i].is_equal(b.ops[i])) return false;
        return true;
    default: return true;
    }
}

static const char *op_sym(op_t o)
{
    switch (o) {
    case o_add: return "+"; case o_sub: return "-"; case o_mul: return "*"; case o_div: return "/"; case o_mod: return "%";
    case o_logical_and: return "&&"; case o_logical_or: return "||"; case o_eq: return "=="; case o_ne: return "!=";
    case o_le: return "<="; case o_lt: return "<"; case o_ge: return ">="; case o_gt: return ">";
    case o_right_shift: return ">>"; case o_left_shift: return "<<";
    default: return "?";
    }
}

std::string expr::to_str() const
{
    if (!p_) return "<none>";
    std::ostringstream s;
    switch (p_->kind) {
    case e_val: if (is_float_type(p_->dtype)) s << p_->fval; else s << p_->ival; break;
    case e_var: s << p_->name; break;
    case e_sync: s << "sync()"; break;
    case e_op:
        if (!p_->name.empty()) {
            s << p_->name << "(";
            for (size_t i = 0; i < p_->ops.size(); i++) s << (i ? ", " : "") << p_->ops[i].to_str();
            s << ")";
        } else if (p_->ops.size() == 2 && op_sym(p_->op)[0] != '?') {
            s << "(" << p_->ops[0].to_str() << " " << op_sym(p_->op) << " " << p_->ops[1].to_str() << ")";
        } else {
            s << "op" << (int) p_->op << "(";
            for (size_t i = 0; i < p_->ops.size(); i++) s << (i ? ", " : "") << p_->ops[i].to_str();
            s << ")";
        }
        break;
    default: s << "<none>";
    }
    return s.str();
}
void expr::dump(bool) const { std::cerr << to_str() << "\n"; }
void str_dump(const std::string &s) { std::cerr << s; }

// ---- var / sync ---------------------------------------------------------------------------------
static int g_var_counter = 0;
var::var() : var("_v" + std::to_string(g_var_counter++)) {}
// A name declared once with a type keeps it for later untyped declarations (reference expr.h: var::declared_vars;
// tests/test_76.cpp).
static std::map<std::string, primitive_t> &declared_var_types()
{
    static std::map<std::string, primitive_t> m;
    return m;
}
static primitive_t type_of_var(const std::string &name)
{
    auto it = declared_var_types().find(name);
    return it == declared_var_types().end() ? g_iter_type : it->second;
}
static const std::string &declare_var(primitive_t t, const std::string &name)
{
    declared_var_types()[name] = t;
    return name;
}
var::var(std::string name) : expr(expr::make_var(type_of_var(name), name)) {}
var::var(std::string name, bool) : var(std::move(name)) {}
var::var(primitive_t t, std::string name) : expr(expr::make_var(t, declare_var(t, name))) {}
var::var(primitive_t t, std::string name, bool) : var(t, std::move(name)) {}
var::var(std::string name, expr lower_bound, expr upper_bound) : expr(expr::make_var(type_of_var(name), name)), lower_(lower_bound), upper_(upper_bound) {}

sync::sync() : expr(expr::make_sync()) {}

// ---- statement-like expressions -------------------------------------------------------------------
expr memcpy(const buffer &from, const buffer &to)
{
    return expr(o_memcpy, expr(o_buffer, from.get_name()), expr(o_buffer, to.get_name()));
}
expr allocate(const buffer &b) { return expr(o_allocate, b.get_name()); }
expr cuda_stream_synchronize()
{
    return expr(o_call, "tiramisu_cuda_stream_synchronize", {expr((int32_t) 0)}, p_int32);
}
expr cublas_gemm(const buffer &A, const buffer &B, buffer &C, expr M, expr N, expr K, expr alpha, expr beta,
                 expr ldA, expr ldB, expr ldC, expr offsetA, expr offsetB, expr offsetC, expr transposeA, expr transposeB)
{
    // same operand order as the runtime entry point it lowers to (reference src/tiramisu_expr.cpp:195-237)
    primitive_t t = C.get_elements_type();
    const char *fn = (t == p_float64) ? "tiramisu_cublas_dgemm" : "tiramisu_cublas_sgemm";
    auto u64 = [](const expr &e) { return cast(p_uint64, e); };
    return expr(o_call, fn,
                {expr(o_buffer, A.get_name()), expr(o_buffer, B.get_name()), expr(o_buffer, C.get_name()),
                 u64(M), u64(N), u64(K), cast(t, alpha), cast(t, beta), u64(ldA), u64(ldB), u64(ldC),
                 u64(offsetA), u64(offsetB), u64(offsetC), cast(p_boolean, transposeA), cast(p_boolean, transposeB)},
                p_uint8);
}

}  // namespace tiramisu
