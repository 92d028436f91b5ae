// lowering.cpp — `tiramisu::codegen(args, "x.o", true)` for sm_100a.
//
// Replaces the reference's generator-time back end for GPU functions: the cuda_ast emitter
// (/root/reference/src/tiramisu_codegen_cuda.cpp: kernel extraction :548-640, kernel wrappers
// :855-866, nvcc driver :1100-1200) and the Halide host-code generator
// (/root/reference/src/tiramisu_codegen_halide.cpp:2488-2603 temporaries, :1840-1859 runtime calls).
//
// Pipeline (all of it this file):
//   1. order the scheduled computations by their then()/after() chain;
//   2. build ONE loop tree: consecutive computations share the loops down to their fuse level, a
//      shared loop takes the union of the ranges and every statement re-checks its own range;
//   3. every outermost gpu-tagged loop becomes a kernel; block/thread loops become blockIdx/threadIdx
//      (x,y,z from the innermost tagged loop outwards, like the reference) so the innermost — unit
//      stride — iterator is threadIdx.x and global accesses coalesce; shared/local/register buffers
//      become __shared__ / per-thread arrays; read-only global buffers are `const __restrict__`;
//   4. pattern lowering: a kernel that is exactly one fp32 contraction C(i,j) += A(i,k)*B(k,j) over
//      row-major global buffers is not emitted as SIMT code at all — it becomes a call of the
//      runtime's tcgen05 GEMM entry point (tiramisu_cublas_sgemm in libtiramisu_b200.so);
//   5. the host function keeps the reference ABI `extern "C" int f(halide_buffer_t*...)`; copies,
//      allocation, synchronisation and launches go through the tiramisu_cuda_* runtime;
//   6. nvcc -gencode arch=compute_100a,code=sm_100a -fmad=false (x86 Tiramisu code has no FMA
//      contraction, src/tiramisu_codegen_halide.cpp:3988-3995) compiles the result to the object.
#include <tiramisu/core.h>

#include <algorithm>
#include <cinttypes>
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <iostream>
#include <map>
#include <memory>
#include <set>
#include <sstream>

namespace tiramisu {
namespace {

[[noreturn]] void die(const std::string &msg)
{
    std::cerr << "tiramisu_b200 lowering: " << msg << std::endl;
    std::exit(1);
}

std::string san(const std::string &s)
{
    std::string o;
    for (char c : s) o += (isalnum((unsigned char) c) || c == '_') ? c : '_';
    return o;
}

const char *ctype(primitive_t t)
{
    switch (t) {
    case p_uint8: return "uint8_t"; case p_int8: return "int8_t"; case p_uint16: return "uint16_t"; case p_int16: return "int16_t";
    case p_uint32: return "uint32_t"; case p_int32: return "int32_t"; case p_uint64: return "uint64_t"; case p_int64: return "int64_t";
    case p_float32: return "float"; case p_float64: return "double"; case p_boolean: return "bool";
    default: return "uint8_t";
    }
}

struct Node {
    bool is_loop = false;
    // loop
    int id = 0;
    std::string dim_name;
    expr lo, hi;
    loop_dim::tag_t tag = loop_dim::none;
    int axis = -1;               // gpu axis (0 = x)
    bool launch_guard = false;   // launch dimension may exceed this loop's extent
    bool warp_reduce = false;    // reduction loop strided over the lanes of a warp (see Lowering::warp_reduction_terms)
    std::vector<std::unique_ptr<Node>> kids;
    // statement
    computation *comp = nullptr;
    std::vector<Node *> path;    // enclosing loops, outermost first (one per scheduled dimension)
    std::string cname() const { return "v" + std::to_string(id) + "_" + san(dim_name); }
};

struct Kernel {
    Node *root = nullptr;
    std::string name;
    std::vector<Node *> host_loops;        // enclosing host loops (passed by value)
    std::set<std::string> constants;       // function constants referenced
    std::vector<buffer *> globals;         // global buffers referenced, first-use order
    std::set<buffer *> written;
    std::vector<buffer *> locals;          // shared / local / register buffers
    expr grid[3], block[3];
};

class Lowering {
public:
    Lowering(function *f, const std::vector<buffer *> &args, bool cuda) : f_(f), args_(args), cuda_(cuda) {}
    std::string run();

private:
    function *f_;
    std::vector<buffer *> args_;
    bool cuda_;
    Node root_;
    int next_id_ = 0;
    int next_kernel_ = 0;
    std::map<std::string, computation *> constants_;
    std::vector<computation *> constant_order_;
    std::ostringstream kernels_;            // device code
    std::ostringstream protos_;             // extern "C" prototypes of o_call targets
    std::set<std::string> proto_seen_;
    bool pattern_ = true;
    bool retargeted_ = false;                 // CPU schedule mapped onto the GPU; host arguments mirrored on the device
    std::set<buffer *> mirrored_;             // argument buffers with a device mirror (b_X device, h_X host)
    std::set<buffer *> no_copy_in_;           // outputs that are completely overwritten before being read
    std::set<buffer *> written_args_;         // mirrored arguments some statement stores into (copied back whatever their declared type)
    std::set<buffer *> direct_;               // mirrored arguments handed to a runtime entry point as halide buffers
    size_t n_statements_ = 0;
    void propagate_constants();
    void hoist_buffer_extents();
    struct ExtentRewrite { buffer *b; expr from, to; };
    std::vector<ExtentRewrite> extent_rewrites_;
    void retarget(const std::vector<computation *> &ord);
    void mirror_buffers(const std::vector<computation *> &ord, bool gpu_only);
    void auto_map(computation *c, int shared_depth);
    void fuse_pointwise_nests(const std::vector<computation *> &ord, const std::vector<int> &shared);
    void distribute_parallel_shared_loops(const std::vector<computation *> &ord);
    bool dim_is_parallel(computation *c, const std::string &dim);
    struct Access { buffer *b; std::vector<expr> idx; bool write; };
    std::vector<Access> accesses_of(computation *c);
    std::set<computation *> plainly_mapped_;      // statements auto_map gave the standard element-per-thread mapping
    bool warp_reduction_terms(computation *c, const std::string &r, std::vector<expr> *summands);
    void emit_warp_reduction(Node *loop, std::ostringstream &os, int ind);
    static int fuse_depth(computation *prev, computation *c);
    bool try_heat2d_pattern(Node *zero, Node *heat, std::ostringstream &host, int ind);
    bool try_conv_block_idiom(const std::vector<computation *> &ord, std::ostringstream &host);
    // index simplification context of the statement being emitted
    using Env = std::map<std::string, std::pair<int64_t, int64_t>>;   // loop dimension -> inclusive [lo, hi]
    const computation *cur_ = nullptr;
    Env env_;
    expr inline_defs(const expr &e) const;
    static bool bounds(const expr &e, const Env &env, int64_t &lo, int64_t &hi);
    static expr simp(const expr &e, const Env &env);
    void enter_stmt(computation *c);

    // 1-2
    std::vector<computation *> order();
    void build_tree(const std::vector<computation *> &ord);
    // 3
    void assign_axes(Node *n);
    int count_below(const Node *n, loop_dim::tag_t t) const;
    void collect_kernel(Node *n, Kernel &k, bool inside);
    void note_expr(const expr &e, Kernel &k, bool is_store_target);
    buffer *buffer_of_access(const expr &e, computation **who = nullptr);
    // printing
    std::string print(const expr &e, bool device);
    std::string print_access(const expr &e, bool device);
    std::string lin_index(buffer *b, const std::vector<expr> &idx, bool device);
    std::string literal(const expr &e);
    std::string bytes_of(buffer *b, bool device);
    // emission
    void emit_host(Node *n, std::ostringstream &os, int ind, std::vector<Node *> &host_loops);
    void emit_host_stmt(Node *n, std::ostringstream &os, int ind);
    void emit_kernel(Node *n, std::ostringstream &host, int ind, std::vector<Node *> &host_loops);
    void emit_device(Node *n, Kernel &k, std::ostringstream &os, int ind);
    void emit_assignment(Node *n, std::ostringstream &os, int ind, bool device);
    void emit_bindings(Node *n, std::ostringstream &os, int ind, std::vector<std::string> &conds, bool device);
    struct Contraction { buffer *C = nullptr, *A = nullptr, *B = nullptr; bool tA = false, tB = false; expr M, N, K, offA, offB, offC; };
    bool match_contraction(computation *c, Contraction &m);
    bool match_contractions(computation *c, std::vector<Contraction> &out, std::vector<std::string> *outer);
    bool match_zero_fill(computation *z, buffer *C, const expr &M, const expr &N);
    void emit_sgemm_call(std::ostringstream &host, int ind, const Contraction &m, buffer *C, const std::string &alpha, const std::string &beta);
    bool try_gemm_pattern(Kernel &k, std::ostringstream &host, int ind);
    bool try_sgemm_triple(Node *kz, Node *kr, Node *ks, const std::vector<computation *> &ord, std::ostringstream &host, int ind);
    void need_proto(const expr &call);
};

std::string pad(int n) { return std::string((size_t) n * 4, ' '); }
bool affine_form(const expr &e, std::map<std::string, int64_t> &coef, int64_t &cst, int64_t scale = 1);

// ------------------------------------------------------------------------------------------------ order
std::vector<computation *> Lowering::order()
{
    std::vector<computation *> ord;
    std::set<computation *> seen;
    for (computation *c : f_->get_computations()) {
        if (c->is_input() || c->is_constant() || !c->should_schedule_this_computation()) continue;
        if (c->pred() != nullptr) continue;
        for (computation *p = c; p; p = p->succ()) {
            if (!seen.insert(p).second) die("cycle in the then()/after() order at '" + p->get_name() + "'");
            ord.push_back(p);
        }
    }
    for (computation *c : f_->get_computations())
        if (!c->is_input() && !c->is_constant() && c->should_schedule_this_computation() && !seen.count(c))
            die("computation '" + c->get_name() + "' is not reachable in the then()/after() order");
    return ord;
}

// number of loops `c` shares with its predecessor in the order
int Lowering::fuse_depth(computation *prev, computation *c)
{
    if (!prev || c->pred() != prev) return 0;
    int depth = 0;
    if (c->pred_level_index() >= 0) depth = c->pred_level_index() + 1;
    else if (c->pred_level_index() == -2) {
        int d = c->find_dim(c->pred_level());
        if (d < 0) d = prev->find_dim(c->pred_level());
        if (d < 0) die("'" + c->get_name() + "' is ordered after '" + prev->get_name() + "' at unknown level '" + c->pred_level() + "'");
        depth = d + 1;
    }
    return std::min(depth, std::min((int) prev->dims().size(), (int) c->dims().size()));
}

static expr union_lo(const expr &a, const expr &b) { return a.is_equal(b) ? a : expr(o_min, a, b); }
static expr union_hi(const expr &a, const expr &b) { return a.is_equal(b) ? a : expr(o_max, a, b); }

void Lowering::build_tree(const std::vector<computation *> &ord)
{
    std::vector<Node *> open;
    computation *prev = nullptr;
    for (computation *c : ord) {
        if (!c->get_expr().is_defined()) die("computation '" + c->get_name() + "' has no expression");
        auto &dims = c->dims();
        int depth = fuse_depth(prev, c);
        // a level deeper than the predecessor's own nest only orders the two (reference generators write
        // `copy.then(kernel_stmt, i1)` after a loop-less copy): share what both have
        depth = std::min(depth, std::min((int) open.size(), (int) dims.size()));
        open.resize(depth);
        for (int d = 0; d < depth; d++) {
            Node *n = open[d];
            n->lo = union_lo(n->lo, dims[d].lo);
            n->hi = union_hi(n->hi, dims[d].hi);
            if (n->tag == loop_dim::none) n->tag = dims[d].tag;
        }
        for (int d = depth; d < (int) dims.size(); d++) {
            auto n = std::make_unique<Node>();
            n->is_loop = true; n->id = next_id_++; n->dim_name = dims[d].name;
            n->lo = dims[d].lo; n->hi = dims[d].hi; n->tag = dims[d].tag; n->warp_reduce = dims[d].warp_reduce;
            Node *raw = n.get();
            (open.empty() ? &root_ : open.back())->kids.push_back(std::move(n));
            open.push_back(raw);
        }
        auto s = std::make_unique<Node>();
        s->comp = c; s->path = open; s->id = next_id_++;
        (open.empty() ? &root_ : open.back())->kids.push_back(std::move(s));
        prev = c;
    }
}

// ------------------------------------------------------------------------------------------------ gpu axes
int Lowering::count_below(const Node *n, loop_dim::tag_t t) const
{
    int best = 0;
    for (auto &k : n->kids) {
        if (!k->is_loop) continue;
        int c = count_below(k.get(), t) + (k->tag == t ? 1 : 0);
        best = std::max(best, c);
    }
    return best;
}

void Lowering::assign_axes(Node *n)
{
    if (n->is_loop && (n->tag == loop_dim::gpu_block || n->tag == loop_dim::gpu_thread)) {
        n->axis = count_below(n, n->tag);
        if (n->axis > 2) die("more than three nested gpu " + std::string(n->tag == loop_dim::gpu_block ? "block" : "thread") + " levels at '" + n->dim_name + "'");
    }
    for (auto &k : n->kids) assign_axes(k.get());
}

buffer *Lowering::buffer_of_access(const expr &e, computation **who)
{
    computation *c = f_->find_computation(e.get_name());
    if (!c) die("access to unknown computation '" + e.get_name() + "'");
    if (who) *who = c;
    if (c->is_constant()) return nullptr;
    return c->get_buffer();
}

void Lowering::note_expr(const expr &e, Kernel &k, bool)
{
    if (!e.is_defined()) return;
    if (e.get_expr_type() == e_var) {
        if (constants_.count(e.get_name())) k.constants.insert(e.get_name());
        return;
    }
    if (e.get_expr_type() != e_op) return;
    if (e.get_op_type() == o_access) {
        computation *c = nullptr;
        buffer *b = buffer_of_access(e, &c);
        if (!b) k.constants.insert(c->get_name());
        else {
            auto loc = b->get_location();
            auto &list = (loc == mem_loc::gpu_global || loc == mem_loc::gpu_constant || loc == mem_loc::host) ? k.globals : k.locals;
            if (std::find(list.begin(), list.end(), b) == list.end()) list.push_back(b);
            for (const expr &d : b->get_dim_sizes()) note_expr(d, k, false);
            for (const expr &a : c->access()) note_expr(a, k, false);
        }
    }
    if (e.get_op_type() == o_allocate || e.get_op_type() == o_buffer) {
        buffer *b = f_->find_buffer(e.get_name());
        if (b && (b->get_location() == mem_loc::gpu_shared || b->get_location() == mem_loc::gpu_local || b->get_location() == mem_loc::gpu_register))
            if (std::find(k.locals.begin(), k.locals.end(), b) == k.locals.end()) k.locals.push_back(b);
    }
    for (const expr &o : e.get_operands()) note_expr(o, k, false);
}

void Lowering::collect_kernel(Node *n, Kernel &k, bool)
{
    if (n->is_loop) {
        note_expr(n->lo, k, false);
        note_expr(n->hi, k, false);
        for (auto &c : n->kids) collect_kernel(c.get(), k, true);
        return;
    }
    computation *c = n->comp;
    note_expr(c->get_expr(), k, false);
    note_expr(c->get_predicate(), k, false);
    for (auto &d : c->dims()) { note_expr(d.lo, k, false); note_expr(d.hi, k, false); }
    for (auto &d : c->defs()) note_expr(d.second, k, false);
    for (auto &g : c->guards()) note_expr(g, k, false);
    const expr &e = c->get_expr();
    bool is_stmt_only = e.get_expr_type() == e_sync || (e.get_expr_type() == e_op && (e.get_op_type() == o_allocate || e.get_op_type() == o_free));
    if (!is_stmt_only) {
        buffer *b = c->get_buffer();
        auto loc = b->get_location();
        auto &list = (loc == mem_loc::gpu_shared || loc == mem_loc::gpu_local || loc == mem_loc::gpu_register) ? k.locals : k.globals;
        if (std::find(list.begin(), list.end(), b) == list.end()) list.push_back(b);
        k.written.insert(b);
        for (const expr &d : b->get_dim_sizes()) note_expr(d, k, false);
        for (const expr &a : c->access()) note_expr(a, k, false);
    }
}

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
    case o_div: return narrow("(" + A(0) + " / " + A(1) + ")");
    case o_mod: return is_float_type(t) ? std::string(t == p_float32 ? "fmodf(" : "fmod(") + A(0) + ", " + A(1) + ")" : narrow("(" + A(0) + " % " + A(1) + ")");
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

void Lowering::emit_device(Node *n, Kernel &k, std::ostringstream &os, int ind)
{
    if (n->is_loop) {
        const std::string it = ctype(global::get_loop_iterator_data_type());
        const char *ax[] = {"x", "y", "z"};
        if (n->tag == loop_dim::gpu_block || n->tag == loop_dim::gpu_thread) {
            const bool zero = n->lo.is_int_literal() && n->lo.get_int_val() == 0;
            os << pad(ind) << "const " << it << " " << n->cname() << " = " << (zero ? "" : print(n->lo, true) + " + ") << "(" << it << ") "
               << (n->tag == loop_dim::gpu_block ? "blockIdx." : "threadIdx.") << ax[n->axis] << ";\n";
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

// Dense fp32 contraction over row-major global buffers:
//     [ X(I,J) = 0 ]   X(.., R ..) = X(..) + P(I,R) * Q(R,J)
// (one statement, or a zero-init fused in front of it as in chain_mm/gpu/generator.cpp:38-43; the
// reduction iterator R may sit anywhere in the iterator list, operands may be stored transposed).
// The kernel is not emitted: it becomes one call of the runtime's tcgen05 GEMM, beta = 1 (accumulate
// into what the buffer holds) or beta = 0 (the zero-init is part of the pattern).
static std::vector<Node *> statements_of(Node *root)
{
    std::vector<Node *> stmts, order{root};
    for (size_t i = 0; i < order.size(); i++) {               // pre-order, children left to right
        Node *n = order[i];
        if (!n->is_loop) { stmts.push_back(n); continue; }
        size_t at = i + 1;
        for (auto &c : n->kids) order.insert(order.begin() + (at++), c.get());
    }
    return stmts;
}

// Generalised matcher.  The statement may have more iterators than (I, J, R): the others ("outer" iterators:
// the time / layer loops around an LSTM cell) must not index rows or columns — they may only appear in the
// leading indices of higher-dimensional buffers, where they select a 2-D slice (an element offset).  The sum may
// hold several products: each becomes one GEMM call accumulating into the same element (beta = 1).
bool Lowering::match_contractions(computation *c, std::vector<Contraction> &out, std::vector<std::string> *outer)
{
    if (c->get_data_type() != p_float32 || c->get_predicate().is_defined() || !c->get_expr().is_defined() || !c->has_access()) return false;
    const auto &iters = c->get_iteration_variables();
    if (iters.size() < 3) return false;
    auto is_var = [](const expr &x, const std::string &n) { return x.get_expr_type() == e_var && x.get_name() == n; };
    auto mentions = [](const expr &x, const std::string &v) { return !x.substitute({v}, {expr((int32_t) 0)}).is_equal(x) || (x.get_expr_type() == e_var && x.get_name() == v); };
    buffer *C = c->get_buffer();
    const size_t nc = c->access().size();
    if (C->get_location() != mem_loc::gpu_global || nc < 2 || (size_t) C->get_n_dims() != nc) return false;
    if (c->access()[nc - 2].get_expr_type() != e_var || c->access()[nc - 1].get_expr_type() != e_var) return false;
    const std::string I = c->access()[nc - 2].get_name(), J = c->access()[nc - 1].get_name();
    if (I == J) return false;
    // flatten the sum
    std::vector<expr> terms, stack{c->get_expr()};
    while (!stack.empty()) {
        expr x = stack.back(); stack.pop_back();
        if (x.get_expr_type() == e_op && x.get_op_type() == o_add) { stack.push_back(x.get_operand(1)); stack.push_back(x.get_operand(0)); }
        else terms.push_back(x);
    }
    if (terms.size() < 2) return false;
    // resolve an access to (buffer, index tuple in this statement's iterators)
    auto resolve = [&](const expr &x, buffer *&b, std::vector<expr> &idx) {
        if (!(x.get_expr_type() == e_op && x.get_op_type() == o_access)) return false;
        computation *oc = nullptr;
        b = buffer_of_access(x, &oc);
        if (!b || (int) oc->get_iteration_variables().size() != x.get_n_arg() || oc->access().size() != (size_t) b->get_n_dims()) return false;
        std::vector<std::string> names;
        for (const var &v : oc->get_iteration_variables()) names.push_back(v.get_name());
        idx.clear();
        for (const expr &a : oc->access()) idx.push_back(a.substitute(names, x.get_operands()));
        return true;
    };
    // the reduction iterator: the one (other than I, J) that indexes the last two dimensions of an operand
    std::string R;
    int self = -1;
    for (size_t t = 0; t < terms.size(); t++) {
        buffer *b; std::vector<expr> idx;
        if (resolve(terms[t], b, idx) && b == C && idx.size() == nc) {
            bool same = true;
            for (size_t q = 0; q < nc; q++) same &= idx[q].is_equal(c->access()[q]);
            if (same) { if (self >= 0) return false; self = (int) t; continue; }
        }
    }
    if (self < 0) return false;
    struct Opnd { buffer *b = nullptr; bool trans = false; int role = -1; expr off; };
    auto lead_offset = [&](buffer *b, const std::vector<expr> &idx, expr &off) {
        // element offset of the 2-D slice selected by the leading indices; they must not move with I, J or R
        const auto &d = b->get_dim_sizes();
        off = expr((int32_t) 0);
        for (size_t q = 0; q + 2 < idx.size(); q++) {
            if (mentions(idx[q], I) || mentions(idx[q], J) || (!R.empty() && mentions(idx[q], R))) return false;
            off = off * cast(p_int64, d[q]) + cast(p_int64, idx[q]);
        }
        if (idx.size() > 2) off = off * cast(p_int64, d[idx.size() - 2]) * cast(p_int64, d[idx.size() - 1]);
        return true;
    };
    for (size_t t = 0; t < terms.size(); t++) {
        if ((int) t == self) continue;
        const expr &prod = terms[t];
        if (!(prod.get_expr_type() == e_op && prod.get_op_type() == o_mul)) return false;
        Opnd o[2];
        for (int side = 0; side < 2; side++) {
            buffer *b; std::vector<expr> idx;
            if (!resolve(prod.get_operand(side), b, idx) || b == C || idx.size() < 2) return false;
            if (b->get_location() != mem_loc::gpu_global || b->get_elements_type() != p_float32) return false;
            const expr &r = idx[idx.size() - 2], &s2 = idx[idx.size() - 1];
            if (r.get_expr_type() != e_var || s2.get_expr_type() != e_var) return false;
            // which of I / J it carries decides the role; the other variable is the reduction
            auto other_is_R = [&](const std::string &cand) {
                if (cand == I || cand == J) return false;
                if (R.empty()) R = cand;
                return R == cand;
            };
            o[side].b = b;
            if (is_var(r, I) && other_is_R(s2.get_name())) { o[side].role = 0; o[side].trans = false; }
            else if (is_var(s2, I) && other_is_R(r.get_name())) { o[side].role = 0; o[side].trans = true; }
            else if (is_var(s2, J) && other_is_R(r.get_name())) { o[side].role = 1; o[side].trans = false; }
            else if (is_var(r, J) && other_is_R(s2.get_name())) { o[side].role = 1; o[side].trans = true; }
            else return false;
            if (!lead_offset(b, idx, o[side].off)) return false;
        }
        if (o[0].role == o[1].role) return false;
        const Opnd &a = o[0].role == 0 ? o[0] : o[1], &b = o[0].role == 0 ? o[1] : o[0];
        Contraction m;
        m.C = C; m.A = a.b; m.B = b.b; m.tA = a.trans; m.tB = b.trans; m.offA = a.off; m.offB = b.off;
        out.push_back(m);
    }
    if (out.empty() || R.empty()) return false;
    expr offC;
    {
        std::vector<expr> idx(c->access().begin(), c->access().end());
        if (!lead_offset(C, idx, offC)) return false;
    }
    expr M, N, K;
    std::vector<std::string> outer_names;
    for (const var &v : iters) {
        if (!(v.get_lower().is_int_literal() && v.get_lower().get_int_val() == 0) && (v.get_name() == I || v.get_name() == J || v.get_name() == R)) return false;
        if (v.get_name() == I) M = v.get_upper();
        else if (v.get_name() == J) N = v.get_upper();
        else if (v.get_name() == R) K = v.get_upper();
        else outer_names.push_back(v.get_name());
    }
    if (!M.is_defined() || !N.is_defined() || !K.is_defined()) return false;
    if (!outer_names.empty() && !outer) return false;
    if (outer) *outer = outer_names;
    for (Contraction &m : out) { m.M = M; m.N = N; m.K = K; m.offC = offC; }
    return true;
}

bool Lowering::match_contraction(computation *c, Contraction &m)
{
    std::vector<Contraction> all;
    if (!match_contractions(c, all, nullptr) || all.size() != 1) return false;
    if (c->get_iteration_variables().size() != 3 || all[0].C->get_n_dims() != 2 || all[0].A->get_n_dims() != 2 || all[0].B->get_n_dims() != 2) return false;
    m = all[0];
    return true;
}

// X(i,j) = 0 over exactly [0,M) x [0,N) of buffer C
bool Lowering::match_zero_fill(computation *z, buffer *C, const expr &M, const expr &N)
{
    auto is_var = [](const expr &x, const std::string &n) { return x.get_expr_type() == e_var && x.get_name() == n; };
    const auto &zi = z->get_iteration_variables();
    if (!z->get_expr().is_defined() || z->get_predicate().is_defined() || zi.size() != 2) return false;
    if (z->get_buffer() != C || z->access().size() != 2) return false;
    if (!(z->get_expr().is_literal() && z->get_expr().get_double_val() == 0.0)) return false;
    if (!is_var(z->access()[0], zi[0].get_name()) || !is_var(z->access()[1], zi[1].get_name())) return false;
    if (!zi[0].get_upper().is_equal(M) || !zi[1].get_upper().is_equal(N)) return false;
    for (const var &v : zi) if (!(v.get_lower().is_int_literal() && v.get_lower().get_int_val() == 0)) return false;
    return true;
}

void Lowering::emit_sgemm_call(std::ostringstream &host, int ind, const Contraction &m, buffer *C, const std::string &alpha, const std::string &beta)
{
    auto ld = [&](buffer *b) { return print(b->get_dim_sizes().back(), false); };
    auto off = [&](const expr &o) { return o.is_defined() ? "(uint64_t) (" + print(o, false) + ")" : std::string("0"); };
    host << pad(ind) << "if (tiramisu_cublas_sgemm((float *) b_" << san(m.A->get_name()) << ", (float *) b_" << san(m.B->get_name()) << ", b_" << san(C->get_name())
         << ", (uint64_t) " << print(m.M, false) << ", (uint64_t) " << print(m.N, false) << ", (uint64_t) " << print(m.K, false)
         << ", " << alpha << ", " << beta << ", (uint64_t) " << ld(m.A) << ", (uint64_t) " << ld(m.B) << ", (uint64_t) " << ld(C)
         << ", " << off(m.offA) << ", " << off(m.offB) << ", " << off(C == m.C ? m.offC : expr()) << ", " << (m.tA ? "true" : "false") << ", " << (m.tB ? "true" : "false") << ")) return 1;\n";
}

bool Lowering::try_gemm_pattern(Kernel &k, std::ostringstream &host, int ind)
{
    if (!pattern_) return false;
    std::vector<Node *> stmts = statements_of(k.root);
    if (stmts.empty() || stmts.size() > 2) return false;
    Node *stmt = stmts.back();
    computation *c = stmt->comp;
    std::vector<Contraction> all;
    std::vector<std::string> outer;
    if (!match_contractions(c, all, &outer)) return false;
    // outer iterators must be exactly the host loops around the kernel (in order), untouched by splits
    size_t n_host = 0;
    while (n_host < stmt->path.size() && stmt->path[n_host] != k.root) n_host++;
    if (n_host == stmt->path.size() || n_host != outer.size()) return false;
    for (size_t i = 0; i < n_host; i++)
        if (c->dims()[i].name != outer[i]) return false;
    for (Node *s : stmts) if (s->path.size() <= n_host || s->path[n_host] != k.root) return false;
    const bool zero_init = stmts.size() == 2;
    if (zero_init && (n_host != 0 || all.size() != 1 || !match_zero_fill(stmts[0]->comp, all[0].C, all[0].M, all[0].N))) return false;
    host << pad(ind) << "{   // pattern lowering: '" << c->get_name() << "' is " << (all.size() > 1 ? std::to_string(all.size()) + " dense fp32 contractions" : "a dense fp32 contraction")
         << (zero_init ? " (zero-init fused: beta = 0)" : "") << " -> tcgen05 GEMM of the runtime\n";
    const std::string it = ctype(global::get_loop_iterator_data_type());
    for (size_t i = 0; i < n_host; i++)
        host << pad(ind + 1) << "const " << it << " i_" << san(c->dims()[i].name) << " = " << stmt->path[i]->cname() << ";\n";
    for (const Contraction &m : all) emit_sgemm_call(host, ind + 1, m, m.C, "1.0f", zero_init ? "0.0f" : "1.0f");
    host << pad(ind) << "}\n";
    return true;
}

// The sgemm_tiramisu statement triple (benchmarks/linear_algebra/blas/level3/sgemm/cpu/sgemm_generator.cpp:90-110)
// as three consecutive kernels:  T = 0;  T += A*B over k;  C = a*T + b*C   ->   one runtime GEMM call with
// alpha = a, beta = b.  T must be a temporary nobody reads afterwards.
bool Lowering::try_sgemm_triple(Node *kz, Node *kr, Node *ks, const std::vector<computation *> &ord, std::ostringstream &host, int ind)
{
    if (!pattern_) return false;
    auto single = [](Node *k) -> Node * {
        std::vector<Node *> st = statements_of(k);
        return (st.size() == 1 && !st[0]->path.empty() && st[0]->path[0] == k) ? st[0] : nullptr;
    };
    Node *z = single(kz), *r = single(kr), *s = single(ks);
    if (!z || !r || !s) return false;
    Contraction m;
    if (!match_contraction(r->comp, m) || !match_zero_fill(z->comp, m.C, m.M, m.N)) return false;
    buffer *T = m.C;
    if (T->get_argument_type() != a_temporary) return false;
    // S: C(i,j) = a * T(i,j) + b * C(i,j)
    computation *sc = s->comp;
    const auto &si = sc->get_iteration_variables();
    auto is_var = [](const expr &x, const std::string &n) { return x.get_expr_type() == e_var && x.get_name() == n; };
    if (sc->get_data_type() != p_float32 || sc->get_predicate().is_defined() || si.size() != 2 || sc->access().size() != 2) return false;
    buffer *C = sc->get_buffer();
    if (C == T || C == m.A || C == m.B || C->get_location() != mem_loc::gpu_global || C->get_n_dims() != 2) return false;
    if (!is_var(sc->access()[0], si[0].get_name()) || !is_var(sc->access()[1], si[1].get_name())) return false;
    if (!si[0].get_upper().is_equal(m.M) || !si[1].get_upper().is_equal(m.N)) return false;
    for (const var &v : si) if (!(v.get_lower().is_int_literal() && v.get_lower().get_int_val() == 0)) return false;
    const expr &e = sc->get_expr();
    if (!(e.get_expr_type() == e_op && e.get_op_type() == o_add)) return false;
    auto element_of = [&](const expr &x) -> buffer * {     // access resolving to element (i,j) of a 2-D buffer
        if (!(x.get_expr_type() == e_op && x.get_op_type() == o_access)) return nullptr;
        computation *oc = nullptr;
        buffer *b = buffer_of_access(x, &oc);
        if (!b || oc->access().size() != 2 || (int) oc->get_iteration_variables().size() != x.get_n_arg()) return nullptr;
        std::vector<std::string> names;
        for (const var &v : oc->get_iteration_variables()) names.push_back(v.get_name());
        if (!is_var(oc->access()[0].substitute(names, x.get_operands()), si[0].get_name()) ||
            !is_var(oc->access()[1].substitute(names, x.get_operands()), si[1].get_name())) return nullptr;
        return b;
    };
    auto scalar = [&](const expr &x) { return x.get_data_type() == p_float32 && (x.is_literal() || (x.get_expr_type() == e_var && constants_.count(x.get_name()))); };
    auto scaled = [&](const expr &x, expr &k, buffer *&b) {
        if (!(x.get_expr_type() == e_op && x.get_op_type() == o_mul)) return false;
        for (int side = 0; side < 2; side++)
            if (scalar(x.get_operand(side)) && (b = element_of(x.get_operand(1 - side)))) { k = x.get_operand(side); return true; }
        return false;
    };
    expr k0, k1;
    buffer *b0 = nullptr, *b1 = nullptr;
    if (!scaled(e.get_operand(0), k0, b0) || !scaled(e.get_operand(1), k1, b1)) return false;
    if (b0 == C && b1 == T) { std::swap(k0, k1); std::swap(b0, b1); }
    if (b0 != T || b1 != C) return false;
    // nobody reads T after the triple
    bool after = false;
    for (computation *c : ord) {
        if (c == sc) { after = true; continue; }
        if (!after) continue;
        std::vector<expr> stack{c->get_expr(), c->get_predicate()};
        while (!stack.empty()) {
            expr x = stack.back(); stack.pop_back();
            if (!x.is_defined() || x.get_expr_type() != e_op) continue;
            if (x.get_op_type() == o_access) {
                computation *oc = f_->find_computation(x.get_name());
                if (oc && !oc->is_constant() && oc->get_buffer() == T) return false;
            }
            if (x.get_op_type() == o_buffer && x.get_name() == T->get_name()) return false;
            for (const expr &o : x.get_operands()) stack.push_back(o);
        }
    }
    host << pad(ind) << "// pattern lowering: '" << z->comp->get_name() << "' = 0; '" << r->comp->get_name() << "' += A*B; '" << sc->get_name()
         << "' = a*acc + b*C  (the sgemm_tiramisu statement) -> one tcgen05 GEMM call of the runtime\n";
    emit_sgemm_call(host, ind, m, C, print(k0, false), print(k1, false));
    return true;
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
    for (size_t i = 0; i < params.size(); i++) os << (i ? ", " : "") << params[i];
    os << ")\n{\n" << body.str() << "}\n\n";

    // ---- launch
    host << pad(ind) << "{\n";
    host << pad(ind + 1) << "dim3 grid((unsigned) (" << print(k.grid[0], false) << "), (unsigned) (" << print(k.grid[1], false) << "), (unsigned) (" << print(k.grid[2], false) << "));\n";
    host << pad(ind + 1) << "dim3 block((unsigned) (" << print(k.block[0], false) << "), (unsigned) (" << print(k.block[1], false) << "), (unsigned) (" << print(k.block[2], false) << "));\n";
    host << pad(ind + 1) << "if (grid.x && grid.y && grid.z) {\n";
    host << pad(ind + 2) << k.name << "<<<grid, block, 0, (cudaStream_t) tiramisu_b200_get_stream()>>>(";
    for (size_t i = 0; i < actuals.size(); i++) host << (i ? ", " : "") << actuals[i];
    host << ");\n";
    host << pad(ind + 2) << "if (tiramisu_b200_post_launch(\"" << k.name << "\")) return 1;\n";
    host << pad(ind + 1) << "}\n" << pad(ind) << "}\n";
}

void Lowering::emit_host_stmt(Node *n, std::ostringstream &os, int ind)
{
    computation *c = n->comp;
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
    os << pad(ind) << "for (" << it << " " << n->cname() << " = " << print(n->lo, false) << "; " << n->cname() << " < " << print(n->hi, false) << "; "
       << n->cname() << "++) {\n";
    host_loops.push_back(n);
    for (auto &c : n->kids) emit_host(c.get(), os, ind + 1, host_loops);
    host_loops.pop_back();
    os << pad(ind) << "}\n";
}

// ------------------------------------------------------------------------------------------------ constants
// Function constants whose definition folds to a literal are substituted everywhere (the Halide-derived
// generators name every loop bound: `heat2d_s1_r__y_loop_min` = 1 ...), so that extents, divisibility
// and the pattern matchers see numbers.
// A buffer extent written as an access (`buffer S0_b("S0_b", {N_input(0), N_input(0)}, ...)`, reference
// tests/test_87.cpp:50) would make every kernel dereference a host parameter vector: such extents are
// evaluated once at function entry, like declared constants.
void Lowering::hoist_buffer_extents()
{
    int n = 0;
    for (buffer *b : f_->get_buffers()) {
        std::vector<std::string> names;
        std::vector<expr> values;
        for (const expr &d : b->get_dim_sizes()) {
            bool has_access = false;
            std::vector<expr> stack{d};
            while (!stack.empty()) {
                expr x = stack.back(); stack.pop_back();
                if (!x.is_defined() || x.get_expr_type() != e_op) continue;
                if (x.get_op_type() == o_access) {
                    computation *oc = f_->find_computation(x.get_name());
                    if (oc && !oc->is_constant()) has_access = true;
                }
                for (const expr &o : x.get_operands()) stack.push_back(o);
            }
            if (!has_access) continue;
            std::string cname = "_extent" + std::to_string(n++) + "_" + b->get_name();
            f_->own(new constant(cname, d, global::get_loop_iterator_data_type(), true, nullptr, 0, f_));
            extent_rewrites_.push_back({b, d, expr::make_var(global::get_loop_iterator_data_type(), cname)});
        }
    }
    for (auto &rw : extent_rewrites_) rw.b->replace_dim(rw.from, rw.to);
}

void Lowering::propagate_constants()
{
    std::vector<std::string> names;
    std::vector<expr> values;
    for (computation *c : f_->get_computations()) {
        if (!c->is_constant()) continue;
        expr v = c->get_expr().substitute(names, values);
        if (v.is_literal()) {
            names.push_back(c->get_name());
            values.push_back(cast(c->get_data_type(), v));
        }
    }
    if (names.empty()) return;
    for (computation *c : f_->get_computations()) {
        c->substitute_everywhere(names, values);
    }
    for (buffer *b : f_->get_buffers()) b->substitute_dims(names, values);
}

// ------------------------------------------------------------------------------------------------ CPU schedule -> GPU
// `codegen(args, obj)` / `gen_halide_obj()` without gpu levels is a CPU schedule (parallelize /
// vectorize / unroll).  There is no CPU code generator here, so each loop nest is mapped onto the GPU:
//   * the vectorized loop (unit stride by construction) becomes threadIdx.x and its outer part is cut
//     into threadIdx.y x blockIdx.x so that a block covers 128 consecutive elements (coalesced);
//     without a vector tag the innermost data-parallel loop is cut by 128;
//   * the outermost data-parallel loops become blockIdx.y / blockIdx.z; reduction loops (iterators
//     that do not index the stored element) stay sequential inside the thread;
//   * host buffers get device mirrors: inputs are copied in at entry, outputs copied back at exit —
//     an output that the first statement overwrites completely is not copied in.
// Warp-shuffle reduction (pattern mode only: it re-associates the sum).  A statement
//     X(..) = X(..) + t1(.., r) + t2(.., r) ...
// whose reduction iterator r is the unit-stride index of everything that reads it (a matrix-vector
// product against row-major operands: the gate pre-activations of benchmarks/DNN/blocks/LSTM/cpu/
// generator.cpp:40-41) would, mapped one output per thread, make the 32 lanes of a warp read 32
// different rows.  Instead one warp owns one output: lane L accumulates r = L, L+32, ... in a register
// (coalesced 128-byte reads), the partial sums are combined with __shfl_xor_sync and lane 0 adds the result to X.
bool Lowering::warp_reduction_terms(computation *c, const std::string &r, std::vector<expr> *summands)
{
    const expr &e = c->get_expr();
    if (!e.is_defined() || e.get_expr_type() != e_op || e.get_op_type() != o_add || !c->has_access()) return false;
    std::vector<expr> terms, stack{e};
    while (!stack.empty()) {              // flatten the left-leaning + chain, keeping the written order
        expr x = stack.back(); stack.pop_back();
        if (x.get_expr_type() == e_op && x.get_op_type() == o_add) { stack.push_back(x.get_operand(1)); stack.push_back(x.get_operand(0)); }
        else terms.push_back(x);
    }
    auto mentions = [&](const expr &x, const std::string &v) { return !x.substitute({v}, {expr((int32_t) 0)}).is_equal(x) || (x.get_expr_type() == e_var && x.get_name() == v); };
    int self = -1;
    for (size_t t = 0; t < terms.size(); t++) {
        const expr &x = terms[t];
        if (!(x.get_expr_type() == e_op && x.get_op_type() == o_access)) continue;
        computation *oc = nullptr;
        buffer *b = buffer_of_access(x, &oc);
        if (b != c->get_buffer() || oc->access().size() != c->access().size() || (int) oc->get_iteration_variables().size() != x.get_n_arg()) continue;
        std::vector<std::string> names;
        for (const var &v : oc->get_iteration_variables()) names.push_back(v.get_name());
        bool same = true;
        for (size_t q = 0; q < c->access().size(); q++) same &= oc->access()[q].substitute(names, x.get_operands()).is_equal(c->access()[q]);
        if (same) { if (self >= 0) return false; self = (int) t; }
    }
    if (self < 0 || terms.size() < 2) return false;
    for (const expr &a : c->access()) if (mentions(a, r)) return false;
    // every access that reads r reads it as its last (unit-stride) index, and nothing else may read the buffer written
    bool any = false;
    for (size_t t = 0; t < terms.size(); t++) {
        if ((int) t == self) continue;
        std::vector<expr> st{terms[t]};
        while (!st.empty()) {
            expr x = st.back(); st.pop_back();
            if (!x.is_defined()) continue;
            if (x.get_expr_type() == e_var && x.get_name() == r) return false;          // r used as a value
            if (x.get_expr_type() != e_op) continue;
            if (x.get_op_type() == o_access) {
                computation *oc = nullptr;
                buffer *b = buffer_of_access(x, &oc);
                if (!b) continue;
                if (b == c->get_buffer()) return false;
                if ((int) oc->get_iteration_variables().size() != x.get_n_arg() || oc->access().empty()) return false;
                std::vector<std::string> names;
                for (const var &v : oc->get_iteration_variables()) names.push_back(v.get_name());
                for (size_t q = 0; q < oc->access().size(); q++) {
                    expr idx = oc->access()[q].substitute(names, x.get_operands());
                    const bool last = q + 1 == oc->access().size();
                    if (!mentions(idx, r)) continue;
                    if (!last || !(idx.get_expr_type() == e_var && idx.get_name() == r)) return false;
                    any = true;
                }
                continue;
            }
            for (const expr &o : x.get_operands()) st.push_back(o);
        }
        if (summands) summands->push_back(terms[t]);
    }
    return any;
}

void Lowering::emit_warp_reduction(Node *loop, std::ostringstream &os, int ind)
{
    Node *stmt = loop->kids.front().get();
    computation *c = stmt->comp;
    std::vector<expr> summands;
    if (!warp_reduction_terms(c, loop->dim_name, &summands)) die("internal: warp reduction of '" + c->get_name() + "' no longer matches");
    const std::string it = ctype(global::get_loop_iterator_data_type()), T = ctype(c->get_data_type());
    Node *lane = nullptr;
    for (Node *p : stmt->path) if (p->dim_name == "_lane_" + c->get_name()) lane = p;
    if (!lane) die("internal: warp reduction of '" + c->get_name() + "' has no lane dimension");
    os << pad(ind) << "{   // " << c->get_name() << ": one warp per element, '" << loop->dim_name << "' strided over the lanes, shuffle reduction\n";
    os << pad(ind + 1) << T << " acc = (" << T << ") 0;\n";
    os << pad(ind + 1) << "for (" << it << " " << loop->cname() << " = " << print(loop->lo, true) << " + " << lane->cname() << "; " << loop->cname() << " < "
       << print(loop->hi, true) << "; " << loop->cname() << " += 32) {\n";
    std::vector<std::string> conds;
    emit_bindings(stmt, os, ind + 2, conds, true);
    enter_stmt(c);
    std::string cond;
    for (size_t i = 0; i < conds.size(); i++) cond += (i ? " && " : "") + conds[i];
    if (!cond.empty()) os << pad(ind + 2) << "if (" << cond << ") {\n";
    for (const expr &t : summands) os << pad(ind + 2 + !cond.empty()) << "acc = acc + " << print(t, true) << ";\n";
    if (!cond.empty()) os << pad(ind + 2) << "}\n";
    os << pad(ind + 1) << "}\n";
    os << pad(ind + 1) << "#pragma unroll\n" << pad(ind + 1) << "for (int off = 16; off > 0; off >>= 1) acc = acc + __shfl_xor_sync(0xffffffffu, acc, off);\n";
    // the store: bindings of the outer dimensions only (the reduction iterator is not part of the element's address)
    os << pad(ind + 1) << "if (" << lane->cname() << " == 0) {\n";
    os << pad(ind + 2) << "const " << it << " " << loop->cname() << " = " << print(loop->lo, true) << ";\n";
    std::vector<std::string> conds2;
    emit_bindings(stmt, os, ind + 2, conds2, true);
    std::string cond2;
    for (size_t i = 0; i < conds2.size(); i++) cond2 += (i ? " && " : "") + conds2[i];
    buffer *b = c->get_buffer();
    std::string lhs = "b_" + san(b->get_name()) + "[" + lin_index(b, c->access(), true) + "]";
    os << pad(ind + 2) << (cond2.empty() ? "" : "if (" + cond2 + ") ") << lhs << " = (" << T << ") (" << lhs << " + acc);\n";
    os << pad(ind + 1) << "}\n" << pad(ind) << "}\n";
    enter_stmt(nullptr);
}

// A scheduled dimension is data-parallel iff every original iterator rebuilt from it indexes the stored element
// (each iteration writes its own element); a dimension that feeds no original iterator is not.
bool Lowering::dim_is_parallel(computation *c, const std::string &dim)
{
    if (!c->has_access() && !c->get_expr().is_defined()) return false;
    c->get_buffer();
    bool feeds_any = false;
    for (const var &v : c->get_iteration_variables()) {
        expr full = expr::make_var(p_int32, v.get_name());
        for (auto &d : c->defs()) full = full.substitute({d.first}, {d.second});
        const bool feeds = !full.substitute({dim}, {expr((int32_t) 7)}).is_equal(full);
        if (!feeds) continue;
        feeds_any = true;
        // ... and indexes it injectively: an index such as y/2 or i%4 sends several iterations to one element
        // (the 2x2 max-pool of conv_relu_maxpool_generator_tiramisu.cpp:104 is a reduction over y and x)
        bool stored = false;
        for (const expr &a : c->access()) {
            const bool uses = !a.substitute({v.get_name()}, {expr((int32_t) 0)}).is_equal(a) || (a.get_expr_type() == e_var && a.get_name() == v.get_name());
            if (!uses) continue;
            std::map<std::string, int64_t> coef; int64_t cst = 0;
            if (!affine_form(a, coef, cst) || coef[v.get_name()] == 0) return false;
            stored = true;
        }
        if (!stored) return false;
    }
    return feeds_any;
}

// Loop distribution for re-targeting.  A CPU schedule that fuses statements inside loops (`S1.after(S0, j)`) keeps
// those loops as host loops around one kernel per statement — a launch per iteration.  When the innermost shared
// loops are data parallel for every statement of the fused run and the statements only exchange values
// element-wise (every two accesses to one buffer, one of them a write, use identical index expressions), running
// S0 over the whole box and then S1 is equivalent: the shared loops are un-fused down to the outermost one that
// must stay sequential, each statement is mapped over its full box, and nest fusion may join identical ones again
// inside ONE kernel (reference tests/test_94.cpp, test_155.cpp).
void Lowering::distribute_parallel_shared_loops(const std::vector<computation *> &ord)
{
    auto plain = [](computation *c) {
        const expr &e = c->get_expr();
        return e.is_defined() && e.get_expr_type() != e_sync &&
               !(e.get_expr_type() == e_op && (e.get_op_type() == o_memcpy || e.get_op_type() == o_allocate || e.get_op_type() == o_free || e.get_op_type() == o_call));
    };
    size_t i = 0;
    while (i < ord.size()) {
        // a maximal run of consecutive statements that share at least one loop with their predecessor
        size_t j = i + 1;
        while (j < ord.size() && fuse_depth(ord[j - 1], ord[j]) > 0) j++;
        if (j - i >= 2) {
            bool ok = true;
            for (size_t k = i; k < j; k++) ok &= plain(ord[k]) && !ord[k]->get_predicate().is_defined();
            // how many leading loops the whole run shares
            int common = INT32_MAX;
            for (size_t k = i + 1; k < j; k++) common = std::min(common, fuse_depth(ord[k - 1], ord[k]));
            // keep = the shared loops that must stay sequential: everything down to the innermost non-parallel one
            int keep = 0;
            if (ok) {
                for (int d = 0; d < common; d++)
                    for (size_t k = i; k < j; k++)
                        if (!dim_is_parallel(ord[k], ord[k]->dims()[d].name)) keep = d + 1;
                // element-wise exchange only
                std::vector<Access> seen;
                for (size_t k = i; k < j && ok; k++) {
                    std::vector<Access> mine = accesses_of(ord[k]);
                    for (const Access &p : mine)
                        for (const Access &q : seen) {
                            if (p.b != q.b || (!p.write && !q.write)) continue;
                            bool identical = p.idx.size() == q.idx.size();
                            for (size_t t = 0; identical && t < p.idx.size(); t++) identical = p.idx[t].is_equal(q.idx[t]);
                            if (!identical) ok = false;
                        }
                    seen.insert(seen.end(), mine.begin(), mine.end());
                }
            }
            if (ok && keep < common)
                for (size_t k = i + 1; k < j; k++)
                    if (fuse_depth(ord[k - 1], ord[k]) > keep) ord[k]->set_pred_level_index(keep - 1);
        }
        i = j;
    }
}

void Lowering::auto_map(computation *c, int shared_depth)
{
    const expr &e = c->get_expr();
    if (e.get_expr_type() == e_sync || (e.get_expr_type() == e_op && (e.get_op_type() == o_memcpy || e.get_op_type() == o_allocate ||
                                                                       e.get_op_type() == o_free || (e.get_op_type() == o_call && !c->has_access()))))
        return;
    auto &dims = c->dims();
    auto one_thread = [&]() {
        // nothing to spread over threads (a scalar statement, or every loop carries a dependence): the
        // statement's own loops run sequentially inside a one-thread kernel
        loop_dim one; one.name = "_one_" + c->get_name(); one.lo = expr((int32_t) 0); one.hi = expr((int32_t) 1); one.tag = loop_dim::gpu_block;
        for (auto &d : dims) d.tag = loop_dim::none;
        dims.insert(dims.begin() + shared_depth, one);
    };
    if (dims.empty()) { one_thread(); return; }
    // A statement that reads the buffer it writes at another element than its own (in-place updates such as
    // reference tests/test_20.cpp: S1(i,j) = S0(i,j) + S0(9,9), all in buf0) carries a dependence between its
    // iterations; without a dependence test the only safe mapping keeps its order: one thread.
    if (c->has_access()) {
        std::vector<expr> stack{c->get_expr(), c->get_predicate()};
        while (!stack.empty()) {
            expr x = stack.back(); stack.pop_back();
            if (!x.is_defined() || x.get_expr_type() != e_op) continue;
            if (x.get_op_type() == o_access) {
                computation *oc = f_->find_computation(x.get_name());
                if (oc && !oc->is_constant() && oc->get_buffer() == c->get_buffer()) {
                    bool same = oc->access().size() == c->access().size() && (int) oc->get_iteration_variables().size() == x.get_n_arg();
                    if (same) {
                        std::vector<std::string> names;
                        for (const var &v : oc->get_iteration_variables()) names.push_back(v.get_name());
                        for (size_t q = 0; q < c->access().size(); q++) same &= oc->access()[q].substitute(names, x.get_operands()).is_equal(c->access()[q]);
                    }
                    if (!same) { one_thread(); return; }
                }
            }
            for (const expr &o : x.get_operands()) stack.push_back(o);
        }
    }
    c->get_buffer();
    auto parallel_safe = [&](const std::string &dim) { return dim_is_parallel(c, dim); };
    // loops shared with a neighbour in the order (fused nests: the time / layer loops of the LSTM) stay
    // sequential host loops around the kernels; only the statement's own loops are mapped
    int vec = -1;
    for (size_t i = shared_depth; i < dims.size(); i++) if (dims[i].tag == loop_dim::vector && parallel_safe(dims[i].name)) vec = (int) i;
    for (auto &d : dims) d.tag = loop_dim::none;
    if ((int) dims.size() == shared_depth) { one_thread(); return; }
    // warp-shuffle reduction: the innermost loop is the only reduction and is unit stride in what it reads
    if (pattern_ && is_float_type(c->get_data_type()) && (int) dims.size() >= shared_depth + 2) {
        const loop_dim &rd = dims.back();
        int n_red = 0, q = -1;
        for (int i = shared_depth; i < (int) dims.size(); i++) if (!parallel_safe(dims[i].name)) n_red++;
        for (int i = (int) dims.size() - 2; i >= shared_depth; i--) if (parallel_safe(dims[i].name)) { q = i; break; }
        expr ext = rd.hi - rd.lo;
        const bool long_enough = !ext.is_int_literal() || ext.get_int_val() >= 64;
        const bool is_iter = std::any_of(c->get_iteration_variables().begin(), c->get_iteration_variables().end(), [&](const var &v) { return v.get_name() == rd.name; });
        if (n_red == 1 && q >= 0 && !parallel_safe(rd.name) && is_iter && long_enough && c->guards().empty() && warp_reduction_terms(c, rd.name, nullptr)) {
            const std::string inner = dims[q].name, red = rd.name;
            c->split(var(inner), 8, var(inner + "_b"), var(inner + "_t"));          // 8 warps (outputs) per block
            c->tag_gpu_block_level(var(inner + "_b"));
            c->tag_gpu_thread_level(var(inner + "_t"));
            loop_dim lane; lane.name = "_lane_" + c->get_name(); lane.lo = expr((int32_t) 0); lane.hi = expr((int32_t) 32); lane.tag = loop_dim::gpu_thread;
            auto &dd = c->dims();
            dd.insert(dd.end() - 1, lane);                                          // just outside the reduction loop: the innermost thread level (x)
            dd.back().warp_reduce = true;
            int extra = 0;
            for (size_t i = shared_depth; i < dd.size() && extra < 2; i++) {
                if (dd[i].tag != loop_dim::none || !parallel_safe(dd[i].name)) break;
                dd[i].tag = loop_dim::gpu_block;
                extra++;
            }
            (void) red;
            return;
        }
    }
    std::string tx, ty, bx;
    if (vec >= shared_depth + 1 && parallel_safe(dims[vec - 1].name) && dims[vec].hi.is_int_literal() && dims[vec].hi.get_int_val() <= 64) {
        const int lanes = (int) dims[vec].hi.get_int_val();
        const int rows = std::max(1, 256 / lanes);
        tx = dims[vec].name;
        std::string outer = dims[vec - 1].name;
        bx = outer + "_b"; ty = outer + "_t";
        c->split(var(outer), rows, var(bx), var(ty));
    } else {
        int q = -1;
        for (int i = (int) dims.size() - 1; i >= shared_depth; i--) if (parallel_safe(dims[i].name)) { q = i; break; }
        if (q < 0) { one_thread(); return; }
        std::string inner = dims[q].name;
        bx = inner + "_b"; tx = inner + "_t";
        c->split(var(inner), 256, var(bx), var(tx));
    }
    c->tag_gpu_block_level(var(bx));
    c->tag_gpu_thread_level(var(tx));
    if (!ty.empty()) c->tag_gpu_thread_level(var(ty));
    // Coarsening: one element per thread makes a 10000 x 10000 sweep 800k tiny CTAs and the launch rate,
    // not HBM, the limit (measured: 412 us for a 400 MB fill).  The data-parallel loop just outside the
    // x mapping is cut by 8: the outer part becomes a block index, the inner part an unrolled loop in the
    // thread (8 independent row accesses in flight per thread, neighbouring rows of a stencil shared).
    {
        int y = -1;
        for (size_t i = shared_depth; i < c->dims().size(); i++) {
            const loop_dim &d = c->dims()[i];
            if (d.tag != loop_dim::none) break;
            if (parallel_safe(d.name)) y = (int) i; else break;
        }
        if (y >= 0) {
            const loop_dim d = c->dims()[y];
            expr ext = d.hi - d.lo;
            if (!ext.is_int_literal() || ext.get_int_val() >= 32) {
                c->split(var(d.name), 8, var(d.name + "_o"), var(d.name + "_r"));
                c->tag_unroll_level(var(d.name + "_r"));
            }
        }
    }
    int extra = 0;
    for (size_t i = shared_depth; i < c->dims().size() && extra < 2; i++) {
        loop_dim &d = c->dims()[i];
        if (d.tag != loop_dim::none) break;                        // only loops outside the x mapping
        if (!parallel_safe(d.name)) break;                         // a sequential loop outside would become a host loop anyway
        d.tag = loop_dim::gpu_block;
        extra++;
    }
    // pointwise (no sequential loop left inside): a candidate for nest fusion
    bool pointwise = true;
    for (size_t i = shared_depth; i < c->dims().size(); i++)
        pointwise &= c->dims()[i].tag == loop_dim::gpu_block || c->dims()[i].tag == loop_dim::gpu_thread || c->dims()[i].tag == loop_dim::unrolled;
    if (pointwise) plainly_mapped_.insert(c);
}

// Host buffers touched by device statements move to the device: arguments get a mirror (b_X device,
// h_X host, copies at entry / exit), temporaries are simply allocated there.  Buffers only read by function
// constants (SIZES / alpha / beta style parameter vectors) stay where the caller put them.
// `gpu_only`: consider only statements that carry gpu levels (old-style gpu tests that tag loops but keep
// plain host buffers, reference tests/test_17.cpp) — the rest of such a function stays host code.
void Lowering::mirror_buffers(const std::vector<computation *> &ord, bool gpu_only)
{
    auto on_gpu = [](computation *c) {
        for (auto &d : c->dims()) if (d.tag == loop_dim::gpu_block || d.tag == loop_dim::gpu_thread) return true;
        return false;
    };
    std::set<buffer *> used;
    for (computation *c : ord) {
        if (gpu_only && !on_gpu(c)) continue;
        const expr &e = c->get_expr();
        const bool stmt_only = e.get_expr_type() == e_sync || (e.get_expr_type() == e_op && (e.get_op_type() == o_memcpy || e.get_op_type() == o_allocate || e.get_op_type() == o_free));
        if (!stmt_only && (c->has_access() || e.get_expr_type() != e_op || e.get_op_type() != o_call)) {
            used.insert(c->get_buffer());
            written_args_.insert(c->get_buffer());
        }
        std::vector<expr> stack{c->get_expr(), c->get_predicate()};
        for (auto &d : c->dims()) { stack.push_back(d.lo); stack.push_back(d.hi); }
        while (!stack.empty()) {
            expr x = stack.back(); stack.pop_back();
            if (!x.is_defined() || x.get_expr_type() != e_op) continue;
            if (x.get_op_type() == o_access) {
                computation *oc = f_->find_computation(x.get_name());
                if (oc && !oc->is_constant()) used.insert(oc->get_buffer());
            }
            for (const expr &o : x.get_operands()) stack.push_back(o);
        }
    }
    std::set<buffer *> is_arg(args_.begin(), args_.end());
    for (buffer *b : f_->get_buffers()) {
        if (b->get_location() != mem_loc::host || !used.count(b)) continue;
        b->tag_gpu_global();
        if (is_arg.count(b)) mirrored_.insert(b);
    }
}

// ------------------------------------------------------------------------------------------------ nest fusion
// Consecutive statements of a re-targeted CPU schedule that run over the same box, got the same mapping and only
// exchange values element-wise (the nine pointwise statements of an LSTM cell, LSTM/cpu/generator.cpp:43-51) are
// fused into ONE kernel: a launch per statement costs more than the statement.  Legality without a dependence
// analysis: every two accesses to the same buffer inside the group (at least one of them a write) must either
// use identical index expressions — same thread, program order — or differ by a non-zero constant in a
// dimension that does not depend on the mapped iterators (c(m-1) read while c(m) is written: never the same
// element inside one kernel instance).
std::vector<Lowering::Access> Lowering::accesses_of(computation *c)
{
    std::vector<Access> out;
    if (c->has_access()) out.push_back({c->get_buffer(), std::vector<expr>(c->access().begin(), c->access().end()), true});
    std::vector<expr> stack{c->get_expr(), c->get_predicate()};
    while (!stack.empty()) {
        expr x = stack.back(); stack.pop_back();
        if (!x.is_defined() || x.get_expr_type() != e_op) continue;
        if (x.get_op_type() == o_access) {
            computation *oc = f_->find_computation(x.get_name());
            if (oc && !oc->is_constant() && (int) oc->get_iteration_variables().size() == x.get_n_arg()) {
                std::vector<std::string> names;
                for (const var &v : oc->get_iteration_variables()) names.push_back(v.get_name());
                Access a{oc->get_buffer(), {}, false};
                for (const expr &i : oc->access()) a.idx.push_back(i.substitute(names, x.get_operands()));
                out.push_back(a);
            }
        }
        for (const expr &o : x.get_operands()) stack.push_back(o);
    }
    return out;
}

bool affine_form(const expr &e, std::map<std::string, int64_t> &coef, int64_t &cst, int64_t scale)
{
    if (!e.is_defined()) return false;
    if (e.is_int_literal()) { cst += scale * e.get_int_val(); return true; }
    if (e.get_expr_type() == e_var) { coef[e.get_name()] += scale; return true; }
    if (e.get_expr_type() != e_op) return false;
    switch (e.get_op_type()) {
    case o_add: return affine_form(e.get_operand(0), coef, cst, scale) && affine_form(e.get_operand(1), coef, cst, scale);
    case o_sub: return affine_form(e.get_operand(0), coef, cst, scale) && affine_form(e.get_operand(1), coef, cst, -scale);
    case o_minus: return affine_form(e.get_operand(0), coef, cst, -scale);
    case o_mul:
        if (e.get_operand(0).is_int_literal()) return affine_form(e.get_operand(1), coef, cst, scale * e.get_operand(0).get_int_val());
        if (e.get_operand(1).is_int_literal()) return affine_form(e.get_operand(0), coef, cst, scale * e.get_operand(1).get_int_val());
        return false;
    default: return false;
    }
}

void Lowering::fuse_pointwise_nests(const std::vector<computation *> &ord, const std::vector<int> &shared)
{
    auto same_dims = [](computation *a, computation *b) {
        if (a->dims().size() != b->dims().size()) return false;
        for (size_t i = 0; i < a->dims().size(); i++) {
            const loop_dim &x = a->dims()[i], &y = b->dims()[i];
            if (x.name != y.name || x.tag != y.tag || !x.lo.is_equal(y.lo) || !x.hi.is_equal(y.hi) || x.warp_reduce || y.warp_reduce) return false;
        }
        if (a->guards().size() != b->guards().size() || a->defs().size() != b->defs().size()) return false;
        for (size_t i = 0; i < a->guards().size(); i++) if (!a->guards()[i].is_equal(b->guards()[i])) return false;
        for (size_t i = 0; i < a->defs().size(); i++) if (a->defs()[i].first != b->defs()[i].first || !a->defs()[i].second.is_equal(b->defs()[i].second)) return false;
        return true;
    };
    std::vector<computation *> group;
    std::vector<Access> group_acc;
    for (size_t i = 0; i < ord.size(); i++) {
        computation *c = ord[i];
        const bool candidate = plainly_mapped_.count(c) && !c->get_predicate().is_defined();
        bool joined = false;
        if (candidate && !group.empty() && shared[i] == shared[i - 1] && same_dims(group.back(), c) &&
            (shared[i] == 0 || fuse_depth(ord[i - 1], c) >= shared[i])) {
            // mapped iterators = the original iterators rebuilt from the dimensions below the shared host loops
            std::set<std::string> mapped;
            for (size_t d = shared[i]; d < c->dims().size(); d++) mapped.insert(c->dims()[d].name);
            for (auto &df : c->defs()) mapped.insert(df.first);
            auto uses_mapped = [&](const expr &x) {
                for (const std::string &m : mapped)
                    if (!x.substitute({m}, {expr((int32_t) 0)}).is_equal(x) || (x.get_expr_type() == e_var && x.get_name() == m)) return true;
                return false;
            };
            std::vector<Access> mine = accesses_of(c);
            bool legal = true;
            for (const Access &p : mine) {
                for (const Access &q : group_acc) {
                    if (p.b != q.b || (!p.write && !q.write)) continue;
                    bool identical = p.idx.size() == q.idx.size();
                    for (size_t k = 0; identical && k < p.idx.size(); k++) identical = p.idx[k].is_equal(q.idx[k]);
                    if (identical) continue;
                    bool disjoint = false;
                    for (size_t k = 0; !disjoint && k < std::min(p.idx.size(), q.idx.size()); k++) {
                        if (uses_mapped(p.idx[k]) || uses_mapped(q.idx[k])) continue;
                        std::map<std::string, int64_t> coef; int64_t cst = 0;
                        if (!affine_form(p.idx[k] - q.idx[k], coef, cst)) continue;
                        bool vars = false;
                        for (auto &kv : coef) vars |= kv.second != 0;
                        disjoint = !vars && cst != 0;
                    }
                    if (!disjoint) legal = false;
                }
            }
            if (legal) {
                c->set_pred_level_index((int) c->dims().size() - 1);
                group.push_back(c);
                group_acc.insert(group_acc.end(), mine.begin(), mine.end());
                joined = true;
            }
        }
        if (!joined) {
            group.clear(); group_acc.clear();
            if (candidate) { group.push_back(c); group_acc = accesses_of(c); }
        }
    }
}

void Lowering::retarget(const std::vector<computation *> &ord)
{
    retargeted_ = true;
    if (pattern_) distribute_parallel_shared_loops(ord);
    std::vector<int> shared(ord.size(), 0);
    for (size_t i = 0; i < ord.size(); i++) {
        int with_prev = i ? fuse_depth(ord[i - 1], ord[i]) : 0;
        int with_next = i + 1 < ord.size() ? fuse_depth(ord[i], ord[i + 1]) : 0;
        shared[i] = std::max(with_prev, with_next);
    }
    for (size_t i = 0; i < ord.size(); i++) auto_map(ord[i], shared[i]);
    if (pattern_) fuse_pointwise_nests(ord, shared);
    mirror_buffers(ord, false);
    // outputs whose first use is a statement that overwrites every element need no copy-in
    for (buffer *b : args_) {
        if (b->get_argument_type() != a_output) continue;
        for (computation *c : ord) {
            bool reads = false;
            std::vector<expr> stack{c->get_expr(), c->get_predicate()};
            while (!stack.empty()) {
                expr x = stack.back(); stack.pop_back();
                if (!x.is_defined() || x.get_expr_type() != e_op) continue;
                if (x.get_op_type() == o_access) {
                    computation *oc = f_->find_computation(x.get_name());
                    if (oc && !oc->is_constant() && oc->get_buffer() == b) reads = true;
                }
                for (const expr &o : x.get_operands()) stack.push_back(o);
            }
            bool writes = c->has_access() && c->get_buffer() == b;
            if (!reads && !writes) continue;
            if (writes && !reads && c->guards_only_trim() && !c->get_predicate().is_defined()) {
                const auto &it = c->get_iteration_variables();
                bool full = it.size() == b->get_dim_sizes().size() && c->access().size() == it.size();
                for (size_t i = 0; full && i < it.size(); i++)
                    full = c->access()[i].get_expr_type() == e_var && c->access()[i].get_name() == it[i].get_name() &&
                           it[i].get_lower().is_int_literal() && it[i].get_lower().get_int_val() == 0 &&
                           it[i].get_upper().is_equal(b->get_dim_sizes()[i]);
                if (full) no_copy_in_.insert(b);
            }
            break;
        }
    }
}

// ------------------------------------------------------------------------------------------------ stencil pattern
// zero-fill of OUT followed by the 5-point heat2d update of its interior from IN (the two statements of
// benchmarks/halide/heat2d_tiramisu.cpp:52-69 / heat2dgpu_tiramisu.cpp:50-56): both kernels are replaced
// by ONE call of the runtime's register-tiled, 128-bit heat2d kernel (stencil.cu), which writes the zero
// border itself — the grid is read once and written once.
bool Lowering::try_heat2d_pattern(Node *zero, Node *heat, std::ostringstream &host, int ind)
{
    if (!pattern_) return false;
    auto single = [](Node *k) -> Node * {
        Node *stmt = nullptr;
        std::vector<Node *> st{k};
        while (!st.empty()) {
            Node *n = st.back(); st.pop_back();
            if (!n->is_loop) { if (stmt) return nullptr; stmt = n; continue; }
            for (auto &c : n->kids) st.push_back(c.get());
        }
        return (stmt && !stmt->path.empty() && stmt->path[0] == k) ? stmt : nullptr;
    };
    Node *zs = single(zero), *hs = single(heat);
    if (!zs || !hs) return false;
    computation *z = zs->comp, *h = hs->comp;
    if (z->get_data_type() != p_float32 || h->get_data_type() != p_float32 || h->get_predicate().is_defined() || z->get_predicate().is_defined()) return false;
    buffer *out = h->get_buffer();
    if (z->get_buffer() != out || out->get_n_dims() != 2 || out->get_location() != mem_loc::gpu_global) return false;
    const expr &H = out->get_dim_sizes()[0], &W = out->get_dim_sizes()[1];
    // zero: out[y][x] = 0 over the whole buffer
    const auto &zi = z->get_iteration_variables();
    if (!(z->get_expr().is_literal() && z->get_expr().get_double_val() == 0.0) || zi.size() != 2) return false;
    auto is_var = [](const expr &x, const std::string &n) { return x.get_expr_type() == e_var && x.get_name() == n; };
    auto is_lit = [](const expr &x, int64_t v) { return x.is_int_literal() && x.get_int_val() == v; };
    for (int i = 0; i < 2; i++)
        if (!is_var(z->access()[i], zi[i].get_name()) || !is_lit(zi[i].get_lower(), 0) || !zi[i].get_upper().is_equal(out->get_dim_sizes()[i])) return false;
    // heat: out[y][x] over [1, H-1) x [1, W-1)
    const auto &hi = h->get_iteration_variables();
    if (hi.size() != 2) return false;
    const std::string Y = hi[0].get_name(), X = hi[1].get_name();
    if (!is_var(h->access()[0], Y) || !is_var(h->access()[1], X)) return false;
    for (int i = 0; i < 2; i++)
        if (!is_lit(hi[i].get_lower(), 1) || !hi[i].get_upper().is_equal(out->get_dim_sizes()[i] - 1)) return false;
    // expression: a*C + b*(((p+q)+r)+s), products and sums in either operand order
    auto offset = [&](const expr &x, const std::string &v, int &off) {
        if (is_var(x, v)) { off = 0; return true; }
        if (x.get_expr_type() != e_op || x.get_n_arg() != 2) return false;
        const expr &a = x.get_operand(0), &b = x.get_operand(1);
        if (x.get_op_type() == o_add && is_var(a, v) && b.is_int_literal()) { off = (int) b.get_int_val(); return true; }
        if (x.get_op_type() == o_add && is_var(b, v) && a.is_int_literal()) { off = (int) a.get_int_val(); return true; }
        if (x.get_op_type() == o_sub && is_var(a, v) && b.is_int_literal()) { off = -(int) b.get_int_val(); return true; }
        return false;
    };
    buffer *in = nullptr;
    auto tap = [&](const expr &x, int &dy, int &dx) {
        if (!(x.get_expr_type() == e_op && x.get_op_type() == o_access)) return false;
        computation *oc = nullptr;
        buffer *b = buffer_of_access(x, &oc);
        if (!b || b == out || b->get_n_dims() != 2 || b->get_elements_type() != p_float32 || oc->access().size() != 2) return false;
        if ((int) oc->get_iteration_variables().size() != x.get_n_arg()) return false;
        if (in && in != b) return false;
        std::vector<std::string> names;
        for (const var &v : oc->get_iteration_variables()) names.push_back(v.get_name());
        expr r = oc->access()[0].substitute(names, x.get_operands()), c = oc->access()[1].substitute(names, x.get_operands());
        if (!offset(r, Y, dy) || !offset(c, X, dx)) return false;
        in = b;
        return true;
    };
    auto is_f = [](const expr &x, float v) { return x.is_literal() && x.get_data_type() == p_float32 && (float) x.get_double_val() == v; };
    auto scaled = [&](const expr &x, float k, expr &rest) {
        if (!(x.get_expr_type() == e_op && x.get_op_type() == o_mul)) return false;
        if (is_f(x.get_operand(0), k)) { rest = x.get_operand(1); return true; }
        if (is_f(x.get_operand(1), k)) { rest = x.get_operand(0); return true; }
        return false;
    };
    const expr &e = h->get_expr();
    if (!(e.get_expr_type() == e_op && e.get_op_type() == o_add)) return false;
    expr centre, sum;
    if (!((scaled(e.get_operand(0), 0.3f, centre) && scaled(e.get_operand(1), 0.4f, sum)) ||
          (scaled(e.get_operand(1), 0.3f, centre) && scaled(e.get_operand(0), 0.4f, sum)))) return false;
    int dy, dx;
    if (!tap(centre, dy, dx) || dy || dx) return false;
    // ((p+q)+r)+s
    auto add2 = [](const expr &x, expr &a, expr &b) {
        if (!(x.get_expr_type() == e_op && x.get_op_type() == o_add)) return false;
        a = x.get_operand(0); b = x.get_operand(1);
        return true;
    };
    auto dir = [&](const expr &x) -> char {
        int ty, tx;
        if (!tap(x, ty, tx)) return 0;
        if (ty == 0 && tx == 1) return 'E';
        if (ty == 0 && tx == -1) return 'W';
        if (ty == 1 && tx == 0) return 'S';
        if (ty == -1 && tx == 0) return 'N';
        return 0;
    };
    expr l3, s4, l2, s3, p, q;
    if (!add2(sum, l3, s4)) return false;
    if (!dir(s4) && dir(l3)) std::swap(l3, s4);          // the single tap may be written on either side
    if (!add2(l3, l2, s3)) return false;
    if (!dir(s3) && dir(l2)) std::swap(l2, s3);
    if (!add2(l2, p, q)) return false;
    char a = dir(p), b = dir(q), c3 = dir(s3), c4 = dir(s4);
    if (!a || !b || !c3 || !c4) return false;
    auto pair = [&](char u, char v) { return (a == u && b == v) || (a == v && b == u); };
    int form = -1;
    if (pair('E', 'W') && c3 == 'S' && c4 == 'N') form = 0;          // heat2d_tiramisu.cpp:68
    else if (pair('S', 'E') && c3 == 'N' && c4 == 'W') form = 1;     // heat2dgpu_tiramisu.cpp:54
    if (form < 0 || !in) return false;
    if (!in->get_dim_sizes()[0].is_equal(H) || !in->get_dim_sizes()[1].is_equal(W)) return false;
    host << pad(ind) << "// pattern lowering: zero-fill '" << z->get_name() << "' + 5-point update '" << h->get_name()
         << "' -> the runtime's heat2d kernel (one pass over the grid)\n";
    if (retargeted_ && n_statements_ == 2 && mirrored_.count(in) && mirrored_.count(out)) {
        // the whole function is this stencil on the caller's buffers: let the runtime entry point do the
        // host<->device traffic too (row-chunk pipeline for host grids, in place for device-resident ones)
        auto arg_of = [&](buffer *b) { return (int) (std::find(args_.begin(), args_.end(), b) - args_.begin()); };
        host << pad(ind) << "if (tiramisu_b200_heat2d_buffers(arg" << arg_of(in) << ", arg" << arg_of(out) << ", (int32_t) (" << print(H, false)
             << "), (int32_t) (" << print(W, false) << "), " << form << ")) return 1;\n";
        direct_.insert(in);
        direct_.insert(out);
        return true;
    }
    host << pad(ind) << "if (tiramisu_b200_heat2d(b_" << san(in->get_name()) << ", b_" << san(out->get_name()) << ", (int32_t) (" << print(H, false)
         << "), (int32_t) (" << print(W, false) << "), " << form << ", 0, (int32_t) (" << print(H, false) << "))) return 1;\n";
    return true;
}
// ------------------------------------------------------------------------------------------------ conv block idiom
// The whole conv + ReLU + max-pool block of benchmarks/DNN/blocks/Conv-ReLU-MaxPool/cpu/dense/
// conv_relu_maxpool_generator_tiramisu.cpp:28-49, recognised from Layers I and III only (what is computed and where
// it is stored; the CPU schedule of Layer II only orders it):
//     Z: OUT[n][fb][yo][xo][ff] = 0                                    over the whole output
//     I: R[..] = BIAS[fb][ff]
//     C: R[..] = R[..] + IN[n][y+ky][x+kx][fin] * W[fb][ky][kx][fin][ff]      (ky, kx in [0,3), fin in [0,FIn))
//     P: OUT[n][fb][y/2][x/2][ff] = max(OUT[same element], R[..])
// plus dead statements (the weight "prefetch" copy).  That is the runtime's fused implicit-GEMM conv kernel
// (conv.cu: tcgen05 when the shape allows, the exact SIMT kernel otherwise), so the function body becomes one call.
bool Lowering::try_conv_block_idiom(const std::vector<computation *> &ord, std::ostringstream &host)
{
    if (!pattern_) return false;
    auto is_var = [](const expr &x) { return x.get_expr_type() == e_var; };
    auto resolve = [&](const expr &x, buffer *&b, std::vector<expr> &idx) {
        if (!(x.get_expr_type() == e_op && x.get_op_type() == o_access)) return false;
        computation *oc = nullptr;
        b = buffer_of_access(x, &oc);
        if (!b || (int) oc->get_iteration_variables().size() != x.get_n_arg() || oc->access().size() != (size_t) b->get_n_dims()) return false;
        std::vector<std::string> names;
        for (const var &v : oc->get_iteration_variables()) names.push_back(v.get_name());
        idx.clear();
        for (const expr &a : oc->access()) idx.push_back(a.substitute(names, x.get_operands()));
        return true;
    };
    auto same = [](const std::vector<expr> &a, const std::vector<expr> &b) {
        if (a.size() != b.size()) return false;
        for (size_t i = 0; i < a.size(); i++) if (!a[i].is_equal(b[i])) return false;
        return true;
    };
    auto lit = [](const expr &e, int64_t &v) { if (!e.is_int_literal()) return false; v = e.get_int_val(); return true; };
    auto range_is = [&](computation *c, const std::string &name, int64_t hi) {
        for (const var &v : c->get_iteration_variables())
            if (v.get_name() == name) { int64_t l, h; return lit(v.get_lower(), l) && lit(v.get_upper(), h) && l == 0 && h == hi; }
        return false;
    };
    // P: the only max statement
    computation *P = nullptr, *C = nullptr, *I = nullptr, *Z = nullptr;
    for (computation *c : ord)
        if (c->get_expr().get_expr_type() == e_op && c->get_expr().get_op_type() == o_max && c->get_data_type() == p_float32) { if (P) return false; P = c; }
    if (!P || !P->has_access()) return false;
    buffer *OUT = P->get_buffer(), *R = nullptr, *IN = nullptr, *W = nullptr, *BIAS = nullptr;
    if (OUT->get_n_dims() != 5 || P->access().size() != 5) return false;
    std::string n, fb, y, x, ff;
    {
        const auto &a = P->access();
        auto half = [&](const expr &e, std::string &v) {
            if (!(e.get_expr_type() == e_op && e.get_op_type() == o_div && is_var(e.get_operand(0)) && e.get_operand(1).is_int_literal() && e.get_operand(1).get_int_val() == 2)) return false;
            v = e.get_operand(0).get_name();
            return true;
        };
        if (!is_var(a[0]) || !is_var(a[1]) || !half(a[2], y) || !half(a[3], x) || !is_var(a[4])) return false;
        n = a[0].get_name(); fb = a[1].get_name(); ff = a[4].get_name();
        buffer *b0, *b1; std::vector<expr> i0, i1;
        if (!resolve(P->get_expr().get_operand(0), b0, i0) || !resolve(P->get_expr().get_operand(1), b1, i1)) return false;
        if (b0 != OUT) { std::swap(b0, b1); std::swap(i0, i1); }
        if (b0 != OUT || b1 == OUT || !same(i0, std::vector<expr>(a.begin(), a.end()))) return false;
        R = b1;
    }
    std::vector<expr> r_idx_P;
    { buffer *b; std::vector<expr> i0; const expr &e = P->get_expr();
      if (resolve(e.get_operand(0), b, i0) && b == R) r_idx_P = i0; else if (resolve(e.get_operand(1), b, i0) && b == R) r_idx_P = i0; else return false; }
    // C, I, Z and the dead rest
    std::string ky, kx, fin;
    for (computation *c : ord) {
        if (c == P || !c->has_access()) { if (c != P) return false; continue; }
        buffer *b = c->get_buffer();
        const expr &e = c->get_expr();
        if (b == R && e.get_expr_type() == e_op && e.get_op_type() == o_add) {
            if (C) return false;
            // self + IN * W
            expr self = e.get_operand(0), prod = e.get_operand(1);
            if (!(prod.get_expr_type() == e_op && prod.get_op_type() == o_mul)) std::swap(self, prod);
            if (!(prod.get_expr_type() == e_op && prod.get_op_type() == o_mul)) return false;
            buffer *sb; std::vector<expr> si;
            if (!resolve(self, sb, si) || sb != R || !same(si, std::vector<expr>(c->access().begin(), c->access().end()))) return false;
            buffer *b0, *b1; std::vector<expr> i0, i1;
            if (!resolve(prod.get_operand(0), b0, i0) || !resolve(prod.get_operand(1), b1, i1)) return false;
            if (i0.size() == 5) { std::swap(b0, b1); std::swap(i0, i1); }
            if (i0.size() != 4 || i1.size() != 5) return false;
            for (const expr &q : i1) if (!is_var(q)) return false;
            if (i1[0].get_name() != fb || i1[4].get_name() != ff) return false;
            ky = i1[1].get_name(); kx = i1[2].get_name(); fin = i1[3].get_name();
            auto sum_of = [&](const expr &q, const std::string &a, const std::string &bb) {
                return q.get_expr_type() == e_op && q.get_op_type() == o_add && is_var(q.get_operand(0)) && is_var(q.get_operand(1)) &&
                       ((q.get_operand(0).get_name() == a && q.get_operand(1).get_name() == bb) || (q.get_operand(0).get_name() == bb && q.get_operand(1).get_name() == a));
            };
            if (!is_var(i0[0]) || i0[0].get_name() != n || !sum_of(i0[1], y, ky) || !sum_of(i0[2], x, kx) || !is_var(i0[3]) || i0[3].get_name() != fin) return false;
            if (!same(std::vector<expr>(c->access().begin(), c->access().end()), r_idx_P)) return false;
            IN = b0; W = b1; C = c;
        } else if (b == R) {
            if (I) return false;
            buffer *bb; std::vector<expr> bi;
            if (!resolve(e, bb, bi) || bi.size() != 2 || !is_var(bi[0]) || !is_var(bi[1]) || bi[0].get_name() != fb || bi[1].get_name() != ff) return false;
            if (!same(std::vector<expr>(c->access().begin(), c->access().end()), r_idx_P)) return false;
            BIAS = bb; I = c;
        } else if (b == OUT) {
            if (Z || !(e.is_literal() && e.get_double_val() == 0.0) || c->get_iteration_variables().size() != 5) return false;
            for (int q = 0; q < 5; q++) {
                const var &v = c->get_iteration_variables()[q];
                if (!(c->access()[q].get_expr_type() == e_var && c->access()[q].get_name() == v.get_name())) return false;
                if (!(v.get_lower().is_int_literal() && v.get_lower().get_int_val() == 0) || !v.get_upper().is_equal(OUT->get_dim_sizes()[q])) return false;
            }
            Z = c;
        }
    }
    if (!C || !I || !Z || !IN || !W || !BIAS) return false;
    if (IN == OUT || W == OUT || BIAS == OUT || IN->get_n_dims() != 4 || W->get_n_dims() != 5 || BIAS->get_n_dims() != 2) return false;
    // program order: Z and I before C before P
    auto pos = [&](computation *c) { return std::find(ord.begin(), ord.end(), c) - ord.begin(); };
    if (!(pos(Z) < pos(C) && pos(I) < pos(C) && pos(C) < pos(P))) return false;
    // everything else is dead: writes a buffer nobody reads and none of the five above
    std::set<buffer *> core{OUT, R, IN, W, BIAS};
    for (computation *c : ord) {
        if (c == P || c == C || c == I || c == Z) continue;
        buffer *b = c->get_buffer();
        if (core.count(b)) return false;
        for (computation *d : ord)
            for (const Access &a : accesses_of(d)) if (!a.write && a.b == b) return false;
    }
    // shapes: all literal, consistent with the runtime entry point
    int64_t B, FOB, HO, WO, F8, Bi, Hp, Wp, FIn, w0, w1, w2, w3, w4, bb0, bb1;
    const auto &od = OUT->get_dim_sizes(); const auto &id = IN->get_dim_sizes(); const auto &wd = W->get_dim_sizes(); const auto &bd = BIAS->get_dim_sizes();
    if (!lit(od[0], B) || !lit(od[1], FOB) || !lit(od[2], HO) || !lit(od[3], WO) || !lit(od[4], F8) || !lit(id[0], Bi) || !lit(id[1], Hp) || !lit(id[2], Wp) || !lit(id[3], FIn) ||
        !lit(wd[0], w0) || !lit(wd[1], w1) || !lit(wd[2], w2) || !lit(wd[3], w3) || !lit(wd[4], w4) || !lit(bd[0], bb0) || !lit(bd[1], bb1)) return false;
    const int64_t N = 2 * HO;
    if (F8 != 8 || HO != WO || Bi != B || Hp != N + 2 || Wp != N + 2 || w0 != FOB || w1 != 3 || w2 != 3 || w3 != FIn || w4 != 8 || bb0 != FOB || bb1 != 8) return false;
    if (!range_is(C, n, B) || !range_is(C, fb, FOB) || !range_is(C, y, N) || !range_is(C, x, N) || !range_is(C, ky, 3) || !range_is(C, kx, 3) || !range_is(C, fin, FIn) || !range_is(C, ff, 8)) return false;
    if (!range_is(P, n, B) || !range_is(P, fb, FOB) || !range_is(P, y, N) || !range_is(P, x, N) || !range_is(P, ff, 8)) return false;
    if (!range_is(I, n, B) || !range_is(I, fb, FOB) || !range_is(I, y, N) || !range_is(I, x, N) || !range_is(I, ff, 8)) return false;
    for (computation *c : {C, P, I, Z}) if (c->get_predicate().is_defined()) return false;
    for (buffer *b : {IN, W, BIAS, OUT}) if (b->get_elements_type() != p_float32 || b->get_location() != mem_loc::host) return false;
    std::set<buffer *> is_arg(args_.begin(), args_.end());
    for (buffer *b : {IN, W, BIAS, OUT}) if (!is_arg.count(b)) return false;
    // the four tensors live on the device for the call; OUT is written completely
    retargeted_ = true;
    for (buffer *b : {IN, W, BIAS, OUT}) { b->tag_gpu_global(); mirrored_.insert(b); }
    no_copy_in_.insert(OUT);
    written_args_.insert(OUT);
    host << pad(1) << "// idiom lowering: zero-init + bias + 3x3 convolution + max(out, .) over 2x2 windows = the conv + ReLU + max-pool block\n"
         << pad(1) << "// ('" << Z->get_name() << "', '" << I->get_name() << "', '" << C->get_name() << "', '" << P->get_name() << "') -> the runtime's fused implicit-GEMM kernel\n"
         << pad(1) << "if (tiramisu_b200_conv_relu_maxpool(b_" << san(IN->get_name()) << ", b_" << san(W->get_name()) << ", b_" << san(BIAS->get_name()) << ", b_" << san(OUT->get_name())
         << ", " << B << ", " << N << ", " << FIn << ", " << FOB * 8 << ")) return 1;\n";
    return true;
}

// ------------------------------------------------------------------------------------------------ driver
std::string Lowering::run()
{
    if (const char *m = getenv("TIRAMISU_B200_LOWERING")) pattern_ = std::string(m) != "generic";
    hoist_buffer_extents();
    propagate_constants();
    for (computation *c : f_->get_computations())
        if (c->is_constant()) { constants_[c->get_name()] = c; constant_order_.push_back(c); }
    for (buffer *b : args_)
        if (b->get_location() != mem_loc::host && b->get_location() != mem_loc::gpu_global)
            die("argument buffer '" + b->get_name() + "' is tagged on-chip");

    std::vector<computation *> ord = order();
    n_statements_ = ord.size();
    bool any_gpu = false;
    for (computation *c : ord)
        for (auto &d : c->dims()) any_gpu |= d.tag == loop_dim::gpu_block || d.tag == loop_dim::gpu_thread;
    std::ostringstream idiom_body;
    const bool whole_function_idiom = !cuda_ && !any_gpu && try_conv_block_idiom(ord, idiom_body);
    if (whole_function_idiom) {
        std::cerr << "tiramisu_b200 lowering: '" << f_->get_name() << "' is the conv + ReLU + max-pool block: lowered to the runtime's fused kernel\n";
    } else if (!cuda_ && !any_gpu) {
        std::cerr << "tiramisu_b200 lowering: '" << f_->get_name() << "' is a CPU schedule; there is no CPU code generator: "
                     "re-targeting its loop nests to sm_100a (parallel -> blocks, vector -> threads, device mirrors for host buffers)\n";
        retarget(ord);
    } else if (any_gpu) {
        mirror_buffers(ord, true);          // gpu-tagged statements over plain host buffers (no memcpy computations)
    }
    if (!whole_function_idiom) build_tree(ord);
    assign_axes(&root_);

    std::ostringstream body;
    body << idiom_body.str();
    std::vector<Node *> host_loops;
    for (size_t i = 0; i < root_.kids.size(); i++) {
        Node *k = root_.kids[i].get();
        auto is_kernel = [](Node *n) { return n->is_loop && (n->tag == loop_dim::gpu_block || n->tag == loop_dim::gpu_thread); };
        if (i + 1 < root_.kids.size() && is_kernel(k) && is_kernel(root_.kids[i + 1].get()) &&
            try_heat2d_pattern(k, root_.kids[i + 1].get(), body, 1)) {
            i++;
            continue;
        }
        if (i + 2 < root_.kids.size() && is_kernel(k) && is_kernel(root_.kids[i + 1].get()) && is_kernel(root_.kids[i + 2].get()) &&
            try_sgemm_triple(k, root_.kids[i + 1].get(), root_.kids[i + 2].get(), ord, body, 1)) {
            i += 2;
            continue;
        }
        emit_host(k, body, 1, host_loops);
    }

    std::ostringstream out;
    out << "// Generated by the tiramisu_b200 sm_100a lowering for Tiramisu function '" << f_->get_name() << "'. Do not edit.\n"
        << "#include <cstdint>\n#include <cstdlib>\n#include <cstring>\n#include <cmath>\n#include <type_traits>\n#include <cuda_runtime.h>\n\n"
        << "#ifndef HALIDE_HALIDERUNTIME_H\n"
        << "struct halide_dimension_t { int32_t min, extent, stride; uint32_t flags; };\n"
        << "struct halide_type_t { uint8_t code, bits; uint16_t lanes; };\n"
        << "struct halide_buffer_t { uint64_t device; const void *device_interface; uint8_t *host; uint64_t flags; halide_type_t type;\n"
        << "                         int32_t dimensions; halide_dimension_t *dim; void *padding; };\n"
        << "#endif\n"
        << "extern \"C\" {\n"
        << "int tiramisu_cuda_memcpy_to_device(void *, void *, uint64_t);\nint tiramisu_cuda_memcpy_to_host(void *, void *, uint64_t);\n"
        << "void *tiramisu_cuda_malloc(uint64_t);\nint tiramisu_cuda_free(void *);\nint32_t tiramisu_cuda_stream_synchronize(int32_t);\n"
        << "void *tiramisu_b200_get_stream(void);\nint tiramisu_b200_post_launch(const char *);\n"
        << "int tiramisu_cuda_memcpy_to_device_async(void *, const void *, uint64_t);\n"
        << "int tiramisu_b200_heat2d(const float *, float *, int32_t, int32_t, int32_t, int32_t, int32_t);\n"
        << "int tiramisu_b200_heat2d_buffers(halide_buffer_t *, halide_buffer_t *, int32_t, int32_t, int32_t);\n"
        << "int tiramisu_b200_conv_relu_maxpool(const float *, const float *, const float *, float *, int32_t, int32_t, int32_t, int32_t);\n"
        << "int tiramisu_cublas_sgemm(float *, float *, float *, uint64_t, uint64_t, uint64_t, float, float, uint64_t, uint64_t, uint64_t, uint64_t, uint64_t, uint64_t, bool, bool);\n"
        << "int tiramisu_cublas_dgemm(double *, double *, double *, uint64_t, uint64_t, uint64_t, double, double, uint64_t, uint64_t, uint64_t, uint64_t, uint64_t, uint64_t, bool, bool);\n"
        << "}\n"
        << "template <typename T> __host__ __device__ __forceinline__ T tb_max(T a, T b) { return a > b ? a : b; }\n"
        << "template <typename T> __host__ __device__ __forceinline__ T tb_min(T a, T b) { return a < b ? a : b; }\n"
        << "template <typename T> __host__ __device__ __forceinline__ T tb_abs(T a) { return a < 0 ? -a : a; }\n"
        << "// Halide's lerp: an unsigned integer weight is a fixed-point fraction of its type's maximum\n"
        << "template <typename T, typename W> __host__ __device__ __forceinline__ T tb_lerp(T a, T b, W w) {\n"
        << "    const double wf = std::is_floating_point<W>::value ? (double) w : (double) w / (double) (W) ~(W) 0;\n"
        << "    const double r = (double) a * (1.0 - wf) + (double) b * wf;\n"
        << "    return std::is_floating_point<T>::value ? (T) r : (T) (r + 0.5);\n}\n"
        << "static inline void *tb_arg(halide_buffer_t *b) { return b->host ? (void *) b->host : (void *) b->device; }\n\n";
    out << protos_.str() << "\n" << kernels_.str();

    out << "extern \"C\" int " << f_->get_name() << "(";
    for (size_t i = 0; i < args_.size(); i++) out << (i ? ", " : "") << "halide_buffer_t *arg" << i;
    out << ")\n{\n";
    std::set<buffer *> is_arg(args_.begin(), args_.end());
    for (size_t i = 0; i < args_.size(); i++) {
        const std::string T = ctype(args_[i]->get_elements_type()), N = san(args_[i]->get_name());
        if (mirrored_.count(args_[i])) {
            // host argument of a re-targeted CPU schedule: h_X is the caller's memory, b_X its device mirror
            // (a device-resident argument, ->host == NULL, is used in place)
            out << pad(1) << T << " *h_" << N << " = (" << T << " *) arg" << i << "->host;\n";
            out << pad(1) << T << " *b_" << N << " = h_" << N << " ? nullptr : (" << T << " *) arg" << i << "->device;\n";
        } else {
            out << pad(1) << T << " *b_" << N << " = (" << T << " *) tb_arg(arg" << i << ");\n";
        }
    }
    for (buffer *b : f_->get_buffers())
        if (!is_arg.count(b) && (b->get_location() == mem_loc::host || b->get_location() == mem_loc::gpu_global))
            out << pad(1) << ctype(b->get_elements_type()) << " *b_" << san(b->get_name()) << " = nullptr;\n";
    for (computation *c : constant_order_)
        out << pad(1) << "const " << ctype(c->get_data_type()) << " c_" << san(c->get_name()) << " = (" << ctype(c->get_data_type()) << ") "
            << print(c->get_expr(), false) << ";\n";
    // temporaries: allocated at entry, freed at exit (reference src/tiramisu_codegen_halide.cpp:2488-2530)
    std::vector<buffer *> autos;
    for (buffer *b : f_->get_buffers())
        if (!is_arg.count(b) && b->get_auto_allocate() && (b->get_location() == mem_loc::host || b->get_location() == mem_loc::gpu_global))
            autos.push_back(b);
    for (buffer *b : autos)
        out << pad(1) << "b_" << san(b->get_name()) << " = (" << ctype(b->get_elements_type()) << " *) "
            << (b->get_location() == mem_loc::gpu_global ? "tiramisu_cuda_malloc(" : "malloc(") << bytes_of(b, false) << ");\n";
    for (buffer *b : args_) {
        if (!mirrored_.count(b) || direct_.count(b)) continue;
        const std::string N = san(b->get_name());
        out << pad(1) << "if (h_" << N << ") {\n" << pad(2) << "b_" << N << " = (" << ctype(b->get_elements_type()) << " *) tiramisu_cuda_malloc(" << bytes_of(b, false) << ");\n";
        if (!no_copy_in_.count(b))
            out << pad(2) << "tiramisu_cuda_memcpy_to_device_async(b_" << N << ", h_" << N << ", " << bytes_of(b, false) << ");\n";
        out << pad(1) << "}\n";
    }
    out << body.str();
    for (buffer *b : args_) {
        if (!mirrored_.count(b) || direct_.count(b)) continue;
        const std::string N = san(b->get_name());
        out << pad(1) << "if (h_" << N << ") {\n";
        if (b->get_argument_type() == a_output || written_args_.count(b))
            out << pad(2) << "tiramisu_cuda_memcpy_to_host(h_" << N << ", b_" << N << ", " << bytes_of(b, false) << ");\n";
        out << pad(2) << "tiramisu_cuda_free(b_" << N << ");\n" << pad(1) << "}\n";
    }
    for (auto it = autos.rbegin(); it != autos.rend(); ++it)
        out << pad(1) << ((*it)->get_location() == mem_loc::gpu_global ? "tiramisu_cuda_free(" : "free(") << "b_" << san((*it)->get_name()) << ");\n";
    out << pad(1) << "return 0;\n}\n\n";
    out << "extern \"C\" int " << f_->get_name() << "_argv(void **args)\n{\n" << pad(1) << "return " << f_->get_name() << "(";
    for (size_t i = 0; i < args_.size(); i++) out << (i ? ", " : "") << "(halide_buffer_t *) args[" << i << "]";
    out << ");\n}\n";
    return out.str();
}

}  // namespace

void function::codegen(const std::vector<buffer *> &arguments, const std::string obj_filename, bool gen_cuda_stmt)
{
    set_arguments(arguments);
    Lowering l(this, arguments, gen_cuda_stmt);
    std::string src = l.run();
    std::string cu = obj_filename + ".b200.cu";
    {
        std::ofstream o(cu);
        if (!o) die("cannot write " + cu);
        o << src;
    }
    const char *nvcc = getenv("NVCC");
    const char *extra = getenv("TIRAMISU_B200_NVCC_FLAGS");
    std::string cmd = std::string(nvcc ? nvcc : "/usr/local/cuda/bin/nvcc") +
                      " -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -fmad=false -Xcompiler -fPIC " +
                      (extra ? extra : "") + " -c " + cu + " -o " + obj_filename;
    // the image's $CC/$CXX wrappers are not usable nvcc host compilers
    cmd = "env -u CC -u CXX " + cmd;
    int rc = std::system(cmd.c_str());
    if (rc != 0) die("nvcc failed (" + std::to_string(rc) + ") on " + cu);
}

}  // namespace tiramisu
