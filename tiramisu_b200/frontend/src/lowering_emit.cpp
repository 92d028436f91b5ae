// lowering_emit.cpp — index simplification, expression printing, statement / kernel / host emission (see lowering.h).
#include "lowering.h"

namespace tiramisu {
namespace lowering_detail {

// ------------------------------------------------------------------------------------------------ index simplification
// After split/tile an original iterator is `outer * f + inner`; Layer-III accesses such as
// `acc[i % 4][j % 4]` must fold to `acc[i1][j1]` so that per-thread arrays stay in registers once the
// inner loops are unrolled.  Everything here assumes non-negative terms (checked through `bounds`).
void Lowering::enter_stmt(computation *c)
{
    cur_ = c;
    env_.clear();
    if (!c) return;
    for (auto &d : c->dims())
        if (d.lo.is_int_literal() && d.hi.is_int_literal() && d.hi.get_int_val() > d.lo.get_int_val())
            env_[d.name] = {d.lo.get_int_val(), d.hi.get_int_val() - 1};
        else if (d.lo.is_int_literal() && d.lo.get_int_val() >= 0)
            env_[d.name] = {d.lo.get_int_val(), INT64_MAX / 4};
}

expr Lowering::inline_defs(const expr &e) const
{
    if (!cur_) return e;
    expr r = e;
    for (auto &d : cur_->defs()) r = r.substitute({d.first}, {d.second});
    return r;
}

bool Lowering::bounds(const expr &e, const Env &env, int64_t &lo, int64_t &hi)
{
    if (!e.is_defined()) return false;
    if (e.is_int_literal()) { lo = hi = e.get_int_val(); return true; }
    if (e.get_expr_type() == e_var) {
        auto it = env.find(e.get_name());
        if (it == env.end()) return false;
        lo = it->second.first; hi = it->second.second;
        return true;
    }
    if (e.get_expr_type() != e_op || e.get_n_arg() != 2) return false;
    int64_t al, ah, bl, bh;
    if (!bounds(e.get_operand(0), env, al, ah) || !bounds(e.get_operand(1), env, bl, bh)) return false;
    const int64_t big = INT64_MAX / 4;
    auto sat = [&](__int128 v) { return (int64_t) (v > big ? big : (v < -big ? -big : v)); };
    switch (e.get_op_type()) {
    case o_add: lo = sat((__int128) al + bl); hi = sat((__int128) ah + bh); return true;
    case o_sub: lo = sat((__int128) al - bh); hi = sat((__int128) ah - bl); return true;
    case o_mul: {
        __int128 c[4] = {(__int128) al * bl, (__int128) al * bh, (__int128) ah * bl, (__int128) ah * bh};
        __int128 mn = c[0], mx = c[0];
        for (auto v : c) { mn = v < mn ? v : mn; mx = v > mx ? v : mx; }
        lo = sat(mn); hi = sat(mx);
        return true;
    }
    case o_div: if (bl != bh || bl <= 0 || al < 0) return false; lo = al / bl; hi = ah / bl; return true;
    case o_mod: if (bl != bh || bl <= 0 || al < 0) return false; if (ah < bl) { lo = al; hi = ah; } else { lo = 0; hi = bl - 1; } return true;
    default: return false;
    }
}

static void flatten_sum(const expr &e, std::vector<expr> &terms)
{
    if (e.get_expr_type() == e_op && e.get_op_type() == o_add) {
        flatten_sum(e.get_operand(0), terms);
        flatten_sum(e.get_operand(1), terms);
    } else terms.push_back(e);
}

// term = q * c exactly?  returns q
static bool multiple_of(const expr &t, int64_t c, expr &q)
{
    if (t.is_int_literal()) { if (t.get_int_val() % c) return false; q = value_cast_i(t.get_data_type(), t.get_int_val() / c); return true; }
    if (t.get_expr_type() == e_op && t.get_op_type() == o_mul) {
        for (int s = 0; s < 2; s++) {
            const expr &l = t.get_operand(s), &o = t.get_operand(1 - s);
            if (l.is_int_literal() && l.get_int_val() % c == 0) { q = o * value_cast_i(l.get_data_type(), l.get_int_val() / c); return true; }
            expr inner;
            if (multiple_of(o, c, inner)) { q = inner * l; return true; }
        }
    }
    return false;
}

expr Lowering::simp(const expr &e, const Env &env)
{
    if (!e.is_defined() || e.get_expr_type() != e_op) return e;
    op_t o = e.get_op_type();
    if (o == o_access || o == o_call || o == o_buffer || o == o_allocate || o == o_free || o == o_memcpy) return e;
    std::vector<expr> ops;
    for (const expr &x : e.get_operands()) ops.push_back(simp(x, env));
    if ((o == o_mod || o == o_div) && ops[1].is_int_literal() && ops[1].get_int_val() > 0 && is_int_type(e.get_data_type())) {
        const int64_t c = ops[1].get_int_val();
        int64_t lo, hi;
        if (bounds(ops[0], env, lo, hi) && lo >= 0 && hi < c) return o == o_mod ? ops[0] : value_cast_i(e.get_data_type(), 0);
        std::vector<expr> terms, rest, quot;
        flatten_sum(ops[0], terms);
        bool nonneg = true;
        for (const expr &t : terms) {
            int64_t tl, th;
            if (!bounds(t, env, tl, th) || tl < 0) { nonneg = false; break; }
            expr q;
            if (multiple_of(t, c, q)) quot.push_back(q); else rest.push_back(t);
        }
        if (nonneg && !quot.empty()) {
            expr r;
            for (const expr &t : rest) r = r.is_defined() ? r + t : t;
            if (o == o_mod) return r.is_defined() ? simp(expr(o_mod, r, ops[1]), env) : value_cast_i(e.get_data_type(), 0);
            expr qs;
            for (const expr &t : quot) qs = qs.is_defined() ? qs + t : t;
            if (r.is_defined()) {
                expr rq = simp(expr(o_div, r, ops[1]), env);
                if (!(rq.is_int_literal() && rq.get_int_val() == 0)) qs = qs + rq;
            }
            return qs;
        }
    }
    switch (o) {
    case o_cast: return expr(o_cast, e.get_data_type(), ops[0]);
    default:
        if (ops.size() == 1) return expr(o, ops[0]);
        if (ops.size() == 2) return expr(o, ops[0], ops[1]);
        if (ops.size() == 3) return expr(o, ops[0], ops[1], ops[2]);
        return e;
    }
}

// ------------------------------------------------------------------------------------------------ printing
std::string Lowering::literal(const expr &e)
{
    primitive_t t = e.get_data_type();
    char buf[96];
    if (t == p_float32) {
        float v = (float) e.get_double_val();
        if (v != v) return "__int_as_float(0x7fc00000)";
        snprintf(buf, sizeof buf, "%af", (double) v);
        return buf;
    }
    if (t == p_float64) { snprintf(buf, sizeof buf, "%a", e.get_double_val()); return buf; }
    if (t == p_boolean) return e.get_int_val() ? "true" : "false";
    if (t == p_int32) { snprintf(buf, sizeof buf, "%" PRId64, e.get_int_val()); return e.get_int_val() < 0 ? "(" + std::string(buf) + ")" : buf; }
    if (t == p_int64) { snprintf(buf, sizeof buf, "((int64_t) %" PRId64 "LL)", e.get_int_val()); return buf; }
    if (t == p_uint64) { snprintf(buf, sizeof buf, "((uint64_t) %" PRIu64 "ULL)", (uint64_t) e.get_int_val()); return buf; }
    snprintf(buf, sizeof buf, "((%s) %" PRId64 ")", ctype(t), e.get_int_val());
    return buf;
}

std::string Lowering::lin_index(buffer *b, const std::vector<expr> &idx, bool device)
{
    const auto &dims = b->get_dim_sizes();
    if (idx.empty() && !dims.empty()) {
        // `store_in(&buf, {})`: the computation owns the buffer's first element (reference tests/test_106.cpp:21)
        std::vector<expr> zeros(dims.size(), expr((int32_t) 0));
        return lin_index(b, zeros, device);
    }
    if (idx.size() != dims.size())
        die("buffer '" + b->get_name() + "' has " + std::to_string(dims.size()) + " dimensions but is accessed with " + std::to_string(idx.size()) + " indices");
    bool small = b->get_location() == mem_loc::gpu_shared || b->get_location() == mem_loc::gpu_local || b->get_location() == mem_loc::gpu_register;
    if (b->get_location() == mem_loc::gpu_register) {
        // a register buffer of one element is a scalar: the index only names which value it holds
        // (generator1.cpp stores c_A(i, k) in b_A_reg{1} at {i % R_BLOCK_I})
        int64_t total = 1;
        for (const expr &d : dims) total = d.is_int_literal() ? total * d.get_int_val() : 0;
        if (total == 1) return "0";
    }
    std::string s;
    for (size_t d = 0; d < idx.size(); d++) {
        std::string e = print(cur_ ? simp(inline_defs(idx[d]), env_) : idx[d], device);
        if (d == 0) s = small ? e : "(int64_t) " + e;
        else s = "(" + s + ") * " + print(dims[d], device) + " + " + e;
    }
    return s;
}

std::string Lowering::print_access(const expr &e, bool device)
{
    computation *c = nullptr;
    buffer *b = buffer_of_access(e, &c);
    if (!b) return "c_" + san(c->get_name());
    const auto &iters = c->get_iteration_variables();
    if ((int) iters.size() != e.get_n_arg())
        die("'" + c->get_name() + "' has " + std::to_string(iters.size()) + " iterators but is accessed with " + std::to_string(e.get_n_arg()) + " indices");
    std::vector<std::string> names;
    for (const var &v : iters) names.push_back(v.get_name());
    std::vector<expr> idx;
    for (const expr &a : c->access()) idx.push_back(a.substitute(names, e.get_operands()));
    auto loc = b->get_location();
    bool dev_buf = loc != mem_loc::host;
    if (device && !dev_buf) die("kernel code reads host buffer '" + b->get_name() + "' (through '" + c->get_name() + "'): tag it gpu_global and copy it");
    if (!device && mirrored_.count(b)) return "h_" + san(b->get_name()) + "[" + lin_index(b, idx, device) + "]";
    if (!device && dev_buf) die("host code accesses device buffer '" + b->get_name() + "' (through '" + c->get_name() + "')");
    return "b_" + san(b->get_name()) + "[" + lin_index(b, idx, device) + "]";
}

static const char *math_fn(op_t o, bool f32)
{
    switch (o) {
    case o_floor: return f32 ? "floorf" : "floor"; case o_ceil: return f32 ? "ceilf" : "ceil";
    case o_round: return f32 ? "roundf" : "round"; case o_trunc: return f32 ? "truncf" : "trunc";
    case o_sqrt: return f32 ? "sqrtf" : "sqrt"; case o_expo: return f32 ? "expf" : "exp"; case o_log: return f32 ? "logf" : "log";
    case o_sin: return f32 ? "sinf" : "sin"; case o_cos: return f32 ? "cosf" : "cos"; case o_tan: return f32 ? "tanf" : "tan";
    case o_asin: return f32 ? "asinf" : "asin"; case o_acos: return f32 ? "acosf" : "acos"; case o_atan: return f32 ? "atanf" : "atan";
    case o_sinh: return f32 ? "sinhf" : "sinh"; case o_cosh: return f32 ? "coshf" : "cosh"; case o_tanh: return f32 ? "tanhf" : "tanh";
    case o_asinh: return f32 ? "asinhf" : "asinh"; case o_acosh: return f32 ? "acoshf" : "acosh"; case o_atanh: return f32 ? "atanhf" : "atanh";
    default: return nullptr;
    }
}

void Lowering::need_proto(const expr &call)
{
    const std::string &n = call.get_name();
    if (n.rfind("tiramisu_", 0) == 0 || !proto_seen_.insert(n).second) return;
    protos_ << "extern \"C\" " << (call.get_data_type() == p_none ? "void" : ctype(call.get_data_type())) << " " << n << "(";
    for (int i = 0; i < call.get_n_arg(); i++) {
        const expr &a = call.get_operand(i);
        protos_ << (i ? ", " : "") << ((a.get_expr_type() == e_op && a.get_op_type() == o_buffer) ? "void *" : ctype(a.get_data_type()));
    }
    protos_ << ");\n";
}

std::string Lowering::print(const expr &e, bool device)
{
    if (!e.is_defined()) die("undefined expression reached the code generator");
    switch (e.get_expr_type()) {
    case e_val: return literal(e);
    case e_var: return (constants_.count(e.get_name()) ? "c_" : "i_") + san(e.get_name());
    case e_sync: return "__syncthreads()";
    default: break;
    }
    op_t o = e.get_op_type();
    primitive_t t = e.get_data_type();
    auto A = [&](int i) { return print(e.get_operand(i), device); };
    auto narrow = [&](const std::string &s) {   // C++ promotes narrow integers to int: re-wrap to the expression's type
        return (is_int_type(t) && type_bits(t) < 32 && t != p_boolean) ? "((" + std::string(ctype(t)) + ") " + s + ")" : s;
    };
    switch (o) {
    case o_add: return narrow("(" + A(0) + " + " + A(1) + ")");
    case o_sub: return narrow("(" + A(0) + " - " + A(1) + ")");
    case o_mul: return narrow("(" + A(0) + " * " + A(1) + ")");
    case o_div: case o_mod: {
        // Integer division rounds towards minus infinity and the remainder is never negative in the reference's
        // lowering (Halide Div / Mod, src/tiramisu_codegen_halide.cpp:3563-3569); C's operators truncate.  The plain
        // operators are emitted only when interval analysis proves both operands non-negative (or the type is
        // unsigned): A((i - 1) % 16) at i = 0 must read A[15], not A[-1].
        if (is_float_type(t)) {
            if (o == o_div) return narrow("(" + A(0) + " / " + A(1) + ")");
            return std::string(t == p_float32 ? "fmodf(" : "fmod(") + A(0) + ", " + A(1) + ")";
        }
        const bool is_unsigned = t == p_uint8 || t == p_uint16 || t == p_uint32 || t == p_uint64 || t == p_boolean;
        int64_t al = 0, ah = 0, bl = 0, bh = 0;
        const bool nonneg = is_unsigned || (bounds(inline_defs(e.get_operand(0)), env_, al, ah) && al >= 0 &&
                                            bounds(inline_defs(e.get_operand(1)), env_, bl, bh) && bl > 0);
        if (nonneg) return narrow("(" + A(0) + (o == o_div ? " / " : " % ") + A(1) + ")");
        return narrow(std::string(o == o_div ? "tb_floordiv<" : "tb_floormod<") + ctype(t) + ">(" + A(0) + ", " + A(1) + ")");
    }
    case o_logical_and: return "(" + A(0) + " && " + A(1) + ")";
    case o_logical_or: return "(" + A(0) + " || " + A(1) + ")";
    case o_logical_not: return "(!" + A(0) + ")";
    case o_eq: return "(" + A(0) + " == " + A(1) + ")";
    case o_ne: return "(" + A(0) + " != " + A(1) + ")";
    case o_le: return "(" + A(0) + " <= " + A(1) + ")";
    case o_lt: return "(" + A(0) + " < " + A(1) + ")";
    case o_ge: return "(" + A(0) + " >= " + A(1) + ")";
    case o_gt: return "(" + A(0) + " > " + A(1) + ")";
    case o_max: return "tb_max<" + std::string(ctype(t)) + ">(" + A(0) + ", " + A(1) + ")";
    case o_min: return "tb_min<" + std::string(ctype(t)) + ">(" + A(0) + ", " + A(1) + ")";
    case o_right_shift: return narrow("(" + A(0) + " >> " + A(1) + ")");
    case o_left_shift: return narrow("(" + A(0) + " << " + A(1) + ")");
    case o_minus: return narrow("(-" + A(0) + ")");
    case o_abs: return is_float_type(t) ? std::string(t == p_float32 ? "fabsf(" : "fabs(") + A(0) + ")" : "tb_abs<" + std::string(ctype(t)) + ">(" + A(0) + ")";
    case o_cast: return "((" + std::string(ctype(t)) + ") " + A(0) + ")";
    case o_select: case o_cond: return "(" + A(0) + " ? " + A(1) + " : " + A(2) + ")";
    case o_lerp: return "tb_lerp<" + std::string(ctype(t)) + ">(" + A(0) + ", " + A(1) + ", " + A(2) + ")";
    case o_access: return print_access(e, device);
    case o_buffer: {
        buffer *b = f_->find_buffer(e.get_name());
        if (!b) die("unknown buffer '" + e.get_name() + "'");
        return "b_" + san(b->get_name());
    }
    case o_call: {
        need_proto(e);
        std::string s = e.get_name() + "(";
        for (int i = 0; i < e.get_n_arg(); i++) s += (i ? ", " : "") + A(i);
        return s + ")";
    }
    default: break;
    }
    if (const char *fn = math_fn(o, t == p_float32)) return std::string(fn) + "(" + A(0) + ")";
    die("operator " + std::to_string((int) o) + " is not supported by the sm_100a lowering (in " + e.to_str() + ")");
}

std::string Lowering::bytes_of(buffer *b, bool device)
{
    std::string s = "(uint64_t) sizeof(" + std::string(ctype(b->get_elements_type())) + ")";
    for (const expr &d : b->get_dim_sizes()) s += " * (uint64_t) " + print(d, device);
    return s;
}

// ------------------------------------------------------------------------------------------------ statements
void Lowering::emit_bindings(Node *n, std::ostringstream &os, int ind, std::vector<std::string> &conds, bool device)
{
    computation *c = n->comp;
    auto &dims = c->dims();
    const std::string it = ctype(global::get_loop_iterator_data_type());
    for (size_t d = 0; d < dims.size(); d++) {
        Node *l = n->path[d];
        os << pad(ind) << "const " << it << " i_" << san(dims[d].name) << " = " << l->cname() << ";\n";
        if (!dims[d].lo.is_equal(l->lo)) conds.push_back("i_" + san(dims[d].name) + " >= " + print(dims[d].lo, device));
        if (!dims[d].hi.is_equal(l->hi) || l->launch_guard) conds.push_back("i_" + san(dims[d].name) + " < " + print(dims[d].hi, device));
    }
    const auto &defs = c->defs();
    for (auto it2 = defs.rbegin(); it2 != defs.rend(); ++it2)
        os << pad(ind) << "const " << it << " i_" << san(it2->first) << " = " << print(it2->second, device) << ";\n";
    for (const expr &g : c->guards()) conds.push_back(print(g, device));
    if (c->get_predicate().is_defined()) conds.push_back(print(c->get_predicate(), device));
}

void Lowering::emit_assignment(Node *n, std::ostringstream &os, int ind, bool device)
{
    computation *c = n->comp;
    os << pad(ind) << "{   // " << c->get_name() << "\n";
    std::vector<std::string> conds;
    emit_bindings(n, os, ind + 1, conds, device);
    enter_stmt(c);
    buffer *b = c->get_buffer();
    bool dev_buf = b->get_location() != mem_loc::host;
    if (device != dev_buf)
        die("'" + c->get_name() + "' runs on the " + (device ? "device" : "host") + " but stores into " + (dev_buf ? "device" : "host") + " buffer '" + b->get_name() + "'");
    std::vector<std::string> names;
    for (const var &v : c->get_iteration_variables()) names.push_back(v.get_name());
    std::string lhs = "b_" + san(b->get_name()) + "[" + lin_index(b, c->access(), device) + "]";
    std::string rhs = print(c->get_expr(), device);
    std::string cond;
    for (size_t i = 0; i < conds.size(); i++) cond += (i ? " && " : "") + conds[i];
    os << pad(ind + 1) << (cond.empty() ? "" : "if (" + cond + ") ") << lhs << " = (" << ctype(b->get_elements_type()) << ") " << rhs << ";\n";
    os << pad(ind) << "}\n";
    enter_stmt(nullptr);
}

// An accumulation whose in-thread loops are <reduction loops> around <a short data-parallel register tile> — the shape
// a CPU register-blocked schedule leaves behind (conv layer: fin_b, k_y, k_x, ffin around fout_b_low (4) x xx (3)) — is
// emitted with the tile's accumulators in registers: load the tile, run the reduction on the registers, store the tile.
// Written as one statement per iteration the compiler does not promote the elements (the tile loops sit inside the
// reduction) and every addition is a global read-modify-write: 54.7 ms for the conv layer against 17.3 without the tile
// and <this form> with it (r02 runs 29-31).  The additions and their order are those of the statement.
bool Lowering::try_emit_accumulation_tile(Node *n, std::ostringstream &os, int ind)
{
    if (!pattern_ || getenv("TIRAMISU_B200_NO_REGISTER_TILE")) return false;
    std::vector<Node *> chain;
    Node *x = n;
    while (x->is_loop) {
        if (x->kids.size() != 1 || x->tag == loop_dim::gpu_block || x->tag == loop_dim::gpu_thread || x->warp_reduce) return false;
        chain.push_back(x);
        x = x->kids.front().get();
    }
    Node *st = x;
    computation *c = st->comp;
    if (!c || !c->has_access() || !c->get_expr().is_defined() || c->get_predicate().is_defined() || c->is_communicator()) return false;
    if (c->get_expr().get_expr_type() != e_op || c->get_expr().get_op_type() != o_add) return false;
    // reduction loops first, then the tile
    size_t n_red = 0;
    while (n_red < chain.size() && !dim_is_parallel(c, chain[n_red]->dim_name)) n_red++;
    if (n_red == 0 || n_red == chain.size()) return false;
    int64_t tile = 1;
    for (size_t q = n_red; q < chain.size(); q++) {
        Node *l = chain[q];
        if (!dim_is_parallel(c, l->dim_name) || !l->lo.is_int_literal() || !l->hi.is_int_literal()) return false;
        const int64_t ext = l->hi.get_int_val() - l->lo.get_int_val();
        if (ext < 1 || ext > 8) return false;
        tile *= ext;
    }
    if (tile < 2 || tile > 32) return false;
    // the loops are the statement's own (no hull of several statements' ranges)
    for (Node *l : chain) {
        const int d = c->find_dim(l->dim_name);
        if (d < 0 || !c->dims()[(size_t) d].lo.is_equal(l->lo) || !c->dims()[(size_t) d].hi.is_equal(l->hi) || l->launch_guard) return false;
    }
    // exactly one term of the sum reads the element the statement stores; it is replaced in place (same association)
    int selfs = 0;
    bool bad = false;
    std::function<expr(const expr &)> swap_self = [&](const expr &e) -> expr {
        if (e.get_expr_type() == e_op && e.get_op_type() == o_add) return expr(o_add, swap_self(e.get_operand(0)), swap_self(e.get_operand(1)));
        if (e.get_expr_type() == e_op && e.get_op_type() == o_access) {
            computation *oc = nullptr;
            if (buffer_of_access(e, &oc) == c->get_buffer()) {
                bool same = oc->access().size() == c->access().size() && (int) oc->get_iteration_variables().size() == e.get_n_arg();
                if (same) {
                    std::vector<std::string> names;
                    for (const var &v : oc->get_iteration_variables()) names.push_back(v.get_name());
                    for (size_t q = 0; q < c->access().size(); q++) same &= oc->access()[q].substitute(names, e.get_operands()).is_equal(c->access()[q]);
                }
                if (!same) { bad = true; return e; }
                selfs++;
                return expr::make_var(c->get_data_type(), "tb_self");
            }
        }
        std::vector<expr> stack{e};
        while (!stack.empty()) {
            expr y = stack.back(); stack.pop_back();
            if (!y.is_defined() || y.get_expr_type() != e_op) continue;
            if (y.get_op_type() == o_access) { computation *oc = nullptr; if (buffer_of_access(y, &oc) == c->get_buffer()) bad = true; }
            for (const expr &o : y.get_operands()) stack.push_back(o);
        }
        return e;
    };
    const expr rhs = swap_self(c->get_expr());
    if (bad || selfs != 1) return false;

    buffer *b = c->get_buffer();
    const std::string it = ctype(global::get_loop_iterator_data_type()), ty = ctype(b->get_elements_type());
    std::string flat;
    {
        int64_t stride = 1;
        for (size_t q = chain.size(); q-- > n_red;) {
            Node *l = chain[q];
            flat = "(" + l->cname() + " - " + std::to_string(l->lo.get_int_val()) + ") * " + std::to_string(stride) + (flat.empty() ? "" : " + " + flat);
            stride *= l->hi.get_int_val() - l->lo.get_int_val();
        }
    }
    auto tile_open = [&](int i0) {
        int i = i0;
        for (size_t q = n_red; q < chain.size(); q++, i++) {
            Node *l = chain[q];
            os << pad(i) << "#pragma unroll\n" << pad(i) << "for (" << it << " " << l->cname() << " = " << l->lo.get_int_val() << "; " << l->cname() << " < "
               << l->hi.get_int_val() << "; " << l->cname() << "++) {\n";
        }
        return i;
    };
    auto tile_close = [&](int i) { for (size_t q = n_red; q < chain.size(); q++) os << pad(--i) << "}\n"; };
    auto element = [&](int i, int phase) {
        os << pad(i) << "{   // " << c->get_name() << (phase == 0 ? ": load the tile" : phase == 1 ? "" : ": store the tile") << "\n";
        std::vector<std::string> conds;
        emit_bindings(st, os, i + 1, conds, true);
        enter_stmt(c);
        std::string cond;
        for (size_t q = 0; q < conds.size(); q++) cond += (q ? " && " : "") + conds[q];
        const std::string lhs = "b_" + san(b->get_name()) + "[" + lin_index(b, c->access(), true) + "]";
        if (phase == 0) os << pad(i + 1) << "tb_acc[" << flat << "] = " << (cond.empty() ? lhs : "(" + cond + ") ? " + lhs + " : (" + ty + ") 0") << ";\n";
        else if (phase == 1) {
            os << pad(i + 1) << ty << " &i_tb_self = tb_acc[" << flat << "];\n";
            os << pad(i + 1) << (cond.empty() ? "" : "if (" + cond + ") ") << "i_tb_self = (" << ty << ") " << print(rhs, true) << ";\n";
        } else os << pad(i + 1) << (cond.empty() ? "" : "if (" + cond + ") ") << lhs << " = tb_acc[" << flat << "];\n";
        enter_stmt(nullptr);
        os << pad(i) << "}\n";
    };
    auto pinned_reduction = [&](int i) {          // the reduction iterators do not reach the stored element: any value will do
        for (size_t q = 0; q < n_red; q++) os << pad(i) << "const " << it << " " << chain[q]->cname() << " = " << print(chain[q]->lo, true) << ";\n";
    };
    os << pad(ind) << "{   // '" << c->get_name() << "': register tile of " << tile << " accumulators\n";
    os << pad(ind + 1) << ty << " tb_acc[" << tile << "];\n";
    os << pad(ind + 1) << "{\n";
    pinned_reduction(ind + 2);
    { int i = tile_open(ind + 2); element(i, 0); tile_close(i); }
    os << pad(ind + 1) << "}\n";
    int i = ind + 1;
    for (size_t q = 0; q < n_red; q++, i++) {
        Node *l = chain[q];
        expr ext = l->hi - l->lo;
        if (l->tag == loop_dim::unrolled || (ext.is_int_literal() && ext.get_int_val() <= 4)) os << pad(i) << "#pragma unroll\n";
        os << pad(i) << "for (" << it << " " << l->cname() << " = " << print(l->lo, true) << "; " << l->cname() << " < " << print(l->hi, true) << "; " << l->cname() << "++) {\n";
    }
    { int j = tile_open(i); element(j, 1); tile_close(j); }
    for (size_t q = 0; q < n_red; q++) os << pad(--i) << "}\n";
    os << pad(ind + 1) << "{\n";
    pinned_reduction(ind + 2);
    { int j = tile_open(ind + 2); element(j, 2); tile_close(j); }
    os << pad(ind + 1) << "}\n";
    os << pad(ind) << "}\n";
    return true;
}

void Lowering::emit_device(Node *n, Kernel &k, std::ostringstream &os, int ind)
{
    if (n->is_loop && n->tag != loop_dim::gpu_block && n->tag != loop_dim::gpu_thread && try_emit_accumulation_tile(n, os, ind)) return;
    if (n->is_loop) {
        const std::string it = ctype(global::get_loop_iterator_data_type());
        const char *ax[] = {"x", "y", "z"};
        if (n->tag == loop_dim::gpu_block || n->tag == loop_dim::gpu_thread) {
            const bool zero = n->lo.is_int_literal() && n->lo.get_int_val() == 0;
            // blockIdx.y / .z are limited to 65535: the launch walks taller grids in slices and passes the slice origin
            const std::string idx = n->tag == loop_dim::gpu_thread ? std::string("threadIdx.") + ax[n->axis]
                                    : n->axis == 0 ? std::string("blockIdx.x")
                                                   : std::string("(blockIdx.") + ax[n->axis] + " + tb_b" + ax[n->axis] + "0)";
            os << pad(ind) << "const " << it << " " << n->cname() << " = " << (zero ? "" : print(n->lo, true) + " + ") << "(" << it << ") " << idx << ";\n";
            for (auto &c : n->kids) emit_device(c.get(), k, os, ind);
        } else if (n->warp_reduce && n->kids.size() == 1 && !n->kids.front()->is_loop) {
            emit_warp_reduction(n, os, ind);
        } else {
            expr ext = n->hi - n->lo;
            if (n->tag == loop_dim::unrolled || (ext.is_int_literal() && ext.get_int_val() <= 32)) os << pad(ind) << "#pragma unroll\n";
            os << pad(ind) << "for (" << it << " " << n->cname() << " = " << print(n->lo, true) << "; " << n->cname() << " < "
               << print(n->hi, true) << "; " << n->cname() << "++) {\n";
            for (auto &c : n->kids) emit_device(c.get(), k, os, ind + 1);
            os << pad(ind) << "}\n";
        }
        return;
    }
    const expr &e = n->comp->get_expr();
    if (e.get_expr_type() == e_sync) { os << pad(ind) << "__syncthreads();   // " << n->comp->get_name() << "\n"; return; }
    if (e.get_expr_type() == e_op && (e.get_op_type() == o_allocate || e.get_op_type() == o_free)) {
        buffer *b = f_->find_buffer(e.get_name());
        if (!b) die("allocate of unknown buffer '" + e.get_name() + "'");
        if (b->get_location() == mem_loc::gpu_global || b->get_location() == mem_loc::host)
            die("buffer '" + b->get_name() + "' cannot be allocated inside a kernel");
        return;   // declared at the top of the kernel
    }
    if (e.get_expr_type() == e_op && (e.get_op_type() == o_memcpy || e.get_op_type() == o_call))
        die("'" + n->comp->get_name() + "': host operation inside a gpu loop nest");
    emit_assignment(n, os, ind, true);
}

// Which thread loop becomes threadIdx.x is the back end's choice, not the schedule's: every thread-tagged loop is
// parallel and all indices are computed from loop iterators, never from threadIdx itself.  The reference numbers the
// axes from the innermost tagged loop outwards (src/tiramisu_codegen_cuda.cpp: gpu_tile(i, j, ...) makes j threadIdx.x),
// which in the `*gpu` Halide benchmarks — `gpu_tile(i, j, 16, 16)` on images stored [c][j][i] — puts the ROW index on
// threadIdx.x: the 32 lanes of a warp touch 16 different rows per access (r02 run 4: cvtcolorgpu at 0.08 of the HBM
// roofline).  Here the thread loop whose iterator is the unit-stride index of most global accesses (stores count
// twice) takes axis x; the block loops are renumbered the same way.  Kernels with a warp-shuffle reduction keep their
// lane dimension on x.
void Lowering::coalesce_thread_axes(Node *root)
{
    using lowering_detail::affine_form;
    std::vector<Node *> loops, stmts, stack{root};
    while (!stack.empty()) {
        Node *x = stack.back(); stack.pop_back();
        if (x->is_loop) { if (x->axis >= 0) loops.push_back(x); }
        else if (x->comp) stmts.push_back(x);
        for (auto &c : x->kids) stack.push_back(c.get());
    }
    for (Node *l : loops) if (l->warp_reduce || l->dim_name.rfind("_lane_", 0) == 0) return;
    long votes[3] = {0, 0, 0};
    for (Node *st : stmts) {
        computation *c = st->comp;
        const expr &e = c->get_expr();
        if (!e.is_defined() || e.get_expr_type() == e_sync) continue;
        if (e.get_expr_type() == e_op && (e.get_op_type() == o_memcpy || e.get_op_type() == o_allocate || e.get_op_type() == o_free)) continue;
        for (const Access &a : accesses_of(c)) {
            if (!a.b || a.idx.empty() || a.b->get_location() != mem_loc::gpu_global) continue;
            expr last = a.idx.back();
            for (auto &d : c->defs()) last = last.substitute({d.first}, {d.second});
            std::map<std::string, int64_t> coef; int64_t cst = 0;
            if (!affine_form(last, coef, cst)) continue;
            for (Node *l : st->path)
                if (l->tag == loop_dim::gpu_thread && l->axis >= 0 && l->axis < 3) {
                    auto it = coef.find(l->dim_name);
                    if (it != coef.end() && (it->second == 1 || it->second == -1)) votes[l->axis] += a.write ? 2 : 1;
                }
        }
    }
    int best = 0;
    for (int a = 1; a < 3; a++) if (votes[a] > votes[best]) best = a;
    if (best == 0 || getenv("TIRAMISU_B200_KEEP_AXES")) return;
    for (Node *l : loops) {
        if (l->axis == best) l->axis = 0;
        else if (l->axis == 0) l->axis = best;
    }
}

void Lowering::emit_kernel(Node *n, std::ostringstream &host, int ind, std::vector<Node *> &host_loops)
{
    Kernel k;
    k.root = n;
    k.host_loops = host_loops;
    {   // a kernel inside host loops is launched once per iteration: refuse programs where that is millions of
        // launches (a CPU schedule whose fused loops had to stay sequential) instead of emitting code that runs for hours
        double launches = 1;
        for (Node *l : host_loops) {
            expr ext = l->hi - l->lo;
            if (ext.is_int_literal()) launches *= (double) std::max<int64_t>(1, ext.get_int_val());
        }
        if (launches > 2e6) {
            Node *first_stmt = n;
            while (first_stmt->is_loop && !first_stmt->kids.empty()) first_stmt = first_stmt->kids.front().get();
            die("'" + (first_stmt->comp ? first_stmt->comp->get_name() : n->dim_name) + "' would be launched " + std::to_string((long long) launches) +
                " times per call: the schedule fuses it inside loops that carry dependences between its statements, which this back end keeps as host loops");
        }
    }
    collect_kernel(n, k, false);
    // first statement's name labels the kernel (ncu-friendly)
    Node *first = n;
    while (first->is_loop && !first->kids.empty()) first = first->kids.front().get();
    k.name = "tiramisu_sm100_" + san(f_->get_name()) + "_k" + std::to_string(next_kernel_++) + "_" + san(first->comp ? first->comp->get_name() : n->dim_name);

    if (try_gemm_pattern(k, host, ind)) return;

    coalesce_thread_axes(n);

    // launch dimensions: per axis, the largest extent among the tagged loops on that axis
    std::vector<Node *> tagged;
    std::vector<Node *> stack{n};
    while (!stack.empty()) {
        Node *x = stack.back(); stack.pop_back();
        if (x->is_loop && x->axis >= 0) tagged.push_back(x);
        for (auto &c : x->kids) stack.push_back(c.get());
    }
    for (int t = 0; t < 2; t++) {
        loop_dim::tag_t tg = t == 0 ? loop_dim::gpu_block : loop_dim::gpu_thread;
        expr *dst = t == 0 ? k.grid : k.block;
        for (int a = 0; a < 3; a++) {
            expr best;
            bool mixed = false;
            for (Node *x : tagged) {
                if (x->tag != tg || x->axis != a) continue;
                expr ext = x->hi - x->lo;
                if (!best.is_defined()) best = ext;
                else if (!best.is_equal(ext)) { best = expr(o_max, best, ext); mixed = true; }
            }
            if (mixed) for (Node *x : tagged) if (x->tag == tg && x->axis == a) x->launch_guard = true;
            dst[a] = best.is_defined() ? best : expr((int32_t) 1);
            note_expr(dst[a], k, false);
        }
    }

    // ---- device code
    std::ostringstream &os = kernels_;
    std::ostringstream body;
    for (buffer *b : k.locals) {
        std::string n_elems;
        int64_t total = 1;
        for (const expr &d : b->get_dim_sizes()) {
            if (!d.is_int_literal()) die("on-chip buffer '" + b->get_name() + "' needs compile-time sizes");
            total *= d.get_int_val();
        }
        body << pad(1) << (b->get_location() == mem_loc::gpu_shared ? "__shared__ " : "") << ctype(b->get_elements_type()) << " b_" << san(b->get_name())
             << "[" << total << "];\n";
    }
    emit_device(n, k, body, 1);

    os << "static __global__ void ";          // internal linkage (as the reference emits): two generated objects in one process may reuse names
    int64_t nthreads = 1;
    bool known = true;
    for (int a = 0; a < 3; a++) { if (k.block[a].is_int_literal()) nthreads *= k.block[a].get_int_val(); else known = false; }
    if (known && nthreads >= 1 && nthreads <= 1024) os << "__launch_bounds__(" << nthreads << ") ";
    os << k.name << "(";
    std::vector<std::string> params, actuals;
    for (buffer *b : k.globals) {
        if (b->get_location() == mem_loc::host) die("kernel '" + k.name + "' uses host buffer '" + b->get_name() + "'");
        bool ro = !k.written.count(b);
        params.push_back(std::string(ro ? "const " : "") + ctype(b->get_elements_type()) + " *" + (ro ? " __restrict__ " : " ") + "b_" + san(b->get_name()));
        actuals.push_back("b_" + san(b->get_name()));
    }
    for (const std::string &cn : k.constants) {
        params.push_back(std::string("const ") + ctype(constants_[cn]->get_data_type()) + " c_" + san(cn));
        actuals.push_back("c_" + san(cn));
    }
    for (Node *l : k.host_loops) {
        params.push_back(std::string("const ") + ctype(global::get_loop_iterator_data_type()) + " " + l->cname());
        actuals.push_back(l->cname());
    }
    params.push_back("const unsigned tb_by0");
    params.push_back("const unsigned tb_bz0");
    actuals.push_back("tb_by0");
    actuals.push_back("tb_bz0");
    for (size_t i = 0; i < params.size(); i++) os << (i ? ", " : "") << params[i];
    os << ")\n{\n" << body.str() << "}\n\n";

    // ---- launch
    host << pad(ind) << "{\n";
    host << pad(ind + 1) << "const dim3 full((unsigned) (" << print(k.grid[0], false) << "), (unsigned) (" << print(k.grid[1], false) << "), (unsigned) (" << print(k.grid[2], false) << "));\n";
    host << pad(ind + 1) << "dim3 block((unsigned) (" << print(k.block[0], false) << "), (unsigned) (" << print(k.block[1], false) << "), (unsigned) (" << print(k.block[2], false) << "));\n";
    // gridDim.y / .z hold at most 65535 blocks: taller grids are launched in slices, the slice origin is a kernel argument
    host << pad(ind + 1) << "for (unsigned tb_bz0 = 0; full.x && tb_bz0 < full.z; tb_bz0 += 65535u)\n";
    host << pad(ind + 1) << "for (unsigned tb_by0 = 0; tb_by0 < full.y; tb_by0 += 65535u) {\n";
    host << pad(ind + 2) << "const dim3 grid(full.x, tb_min<unsigned>(65535u, full.y - tb_by0), tb_min<unsigned>(65535u, full.z - tb_bz0));\n";
    host << pad(ind + 2) << k.name << "<<<grid, block, 0, (cudaStream_t) tiramisu_b200_get_stream()>>>(";
    for (size_t i = 0; i < actuals.size(); i++) host << (i ? ", " : "") << actuals[i];
    host << ");\n";
    host << pad(ind + 2) << "if (tiramisu_b200_post_launch(\"" << k.name << "\")) return 1;\n";
    host << pad(ind + 1) << "}\n" << pad(ind) << "}\n";
}

void Lowering::emit_host_stmt(Node *n, std::ostringstream &os, int ind)
{
    computation *c = n->comp;
    if (c->is_communicator()) { emit_communicator(n, os, ind); return; }
    const expr &e = c->get_expr();
    if (e.get_expr_type() == e_sync) die("'" + c->get_name() + "': sync() outside a gpu loop nest");
    if (e.get_expr_type() == e_op && e.get_op_type() == o_memcpy) {
        buffer *src = f_->find_buffer(e.get_operand(0).get_name()), *dst = f_->find_buffer(e.get_operand(1).get_name());
        if (!src || !dst) die("memcpy between unknown buffers in '" + c->get_name() + "'");
        bool sd = src->get_location() != mem_loc::host, dd = dst->get_location() != mem_loc::host;
        os << pad(ind) << "{   // " << c->get_name() << ": " << src->get_name() << " -> " << dst->get_name() << "\n";
        std::vector<std::string> conds;
        emit_bindings(n, os, ind + 1, conds, false);
        std::string S = "b_" + san(src->get_name()), D = "b_" + san(dst->get_name());
        os << pad(ind + 1) << "const uint64_t nbytes = tb_min<uint64_t>(" << bytes_of(src, false) << ", " << bytes_of(dst, false) << ");\n";
        if (!sd && dd) os << pad(ind + 1) << "tiramisu_cuda_memcpy_to_device(" << D << ", (void *) " << S << ", nbytes);\n";
        else if (sd && !dd) os << pad(ind + 1) << "tiramisu_cuda_memcpy_to_host(" << D << ", (void *) " << S << ", nbytes);\n";
        else if (sd && dd) os << pad(ind + 1) << "cudaMemcpyAsync(" << D << ", " << S << ", nbytes, cudaMemcpyDeviceToDevice, (cudaStream_t) tiramisu_b200_get_stream());\n";
        else os << pad(ind + 1) << "memcpy(" << D << ", " << S << ", nbytes);\n";
        os << pad(ind) << "}\n";
        return;
    }
    if (e.get_expr_type() == e_op && (e.get_op_type() == o_allocate || e.get_op_type() == o_free)) {
        buffer *b = f_->find_buffer(e.get_name());
        if (!b) die("allocate/free of unknown buffer '" + e.get_name() + "'");
        std::string B = "b_" + san(b->get_name());
        os << pad(ind) << "{   // " << c->get_name() << "\n";
        std::vector<std::string> conds;
        emit_bindings(n, os, ind + 1, conds, false);
        bool dev = b->get_location() == mem_loc::gpu_global;
        if (b->get_location() != mem_loc::host && !dev) die("on-chip buffer '" + b->get_name() + "' allocated outside a kernel");
        if (e.get_op_type() == o_allocate)
            os << pad(ind + 1) << B << " = (" << ctype(b->get_elements_type()) << " *) " << (dev ? "tiramisu_cuda_malloc(" : "malloc(") << bytes_of(b, false) << ");\n";
        else
            os << pad(ind + 1) << (dev ? "tiramisu_cuda_free(" : "free(") << B << "); " << B << " = nullptr;\n";
        os << pad(ind) << "}\n";
        return;
    }
    if (e.get_expr_type() == e_op && e.get_op_type() == o_call && !c->has_access()) {
        os << pad(ind) << "{   // " << c->get_name() << "\n";
        std::vector<std::string> conds;
        emit_bindings(n, os, ind + 1, conds, false);
        os << pad(ind + 1) << print(e, false) << ";\n" << pad(ind) << "}\n";
        return;
    }
    if (c->get_buffer()->get_location() == mem_loc::host && !(e.get_expr_type() == e_op && e.get_op_type() == o_call))
        std::cerr << "tiramisu_b200 lowering: note: '" << c->get_name() << "' has no gpu level and runs in the host function\n";
    emit_assignment(n, os, ind, false);
}

void Lowering::emit_host(Node *n, std::ostringstream &os, int ind, std::vector<Node *> &host_loops)
{
    if (!n->is_loop) { emit_host_stmt(n, os, ind); return; }
    if (n->tag == loop_dim::gpu_block || n->tag == loop_dim::gpu_thread) { emit_kernel(n, os, ind, host_loops); return; }
    const std::string it = ctype(global::get_loop_iterator_data_type());
    if (n->tag == loop_dim::distributed) {
        // one iteration per rank: this process runs the one whose index is its rank (reference: `if (rank == i)`)
        os << pad(ind) << "for (" << it << " " << n->cname() << " = (" << it << ") c_rank, tb_once = 0; tb_once == 0 && " << n->cname() << " >= " << print(n->lo, false)
           << " && " << n->cname() << " < " << print(n->hi, false) << "; tb_once++) {\n";
    } else
    os << pad(ind) << "for (" << it << " " << n->cname() << " = " << print(n->lo, false) << "; " << n->cname() << " < " << print(n->hi, false) << "; "
       << n->cname() << "++) {\n";
    host_loops.push_back(n);
    for (auto &c : n->kids) emit_host(c.get(), os, ind + 1, host_loops);
    host_loops.pop_back();
    os << pad(ind) << "}\n";
}

}  // namespace lowering_detail
}  // namespace tiramisu
