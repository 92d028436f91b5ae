// lowering_retarget.cpp — CPU schedules re-targeted to sm_100a (see lowering.h).
#include "lowering.h"

namespace tiramisu {
namespace lowering_detail {

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

// Uses of inline computations (`Si.set_inline()`, reference tests/test_77.cpp) are replaced by their
// expressions with the iterators substituted; the inline computations themselves are neither scheduled nor stored.
void Lowering::expand_inline_computations()
{
    std::function<expr(const expr &, int)> expand = [&](const expr &e, int depth) -> expr {
        if (!e.is_defined() || e.get_expr_type() != e_op) return e;
        if (depth > 64) die("inline computations expand without end (a cycle?)");
        if (e.get_op_type() == o_access) {
            computation *oc = f_->find_computation(e.get_name());
            if (oc && oc->is_inline_computation()) {
                if ((int) oc->get_iteration_variables().size() != e.get_n_arg()) die("inline computation '" + oc->get_name() + "' used with the wrong number of indices");
                std::vector<std::string> names;
                for (const var &v : oc->get_iteration_variables()) names.push_back(v.get_name());
                std::vector<expr> args;
                for (const expr &a : e.get_operands()) args.push_back(expand(a, depth + 1));
                return expand(oc->get_expr().substitute(names, args), depth + 1);
            }
        }
        bool changed = false;
        std::vector<expr> ops;
        for (const expr &o : e.get_operands()) { expr x = expand(o, depth + 1); changed |= x.get() != o.get(); ops.push_back(x); }
        if (!changed) return e;
        switch (e.get_op_type()) {
        case o_call: case o_access: case o_address_of: case o_lin_index: return expr(e.get_op_type(), e.get_name(), ops, e.get_data_type());
        case o_cast: return expr(o_cast, e.get_data_type(), ops[0]);
        default:
            if (ops.size() == 1) return expr(e.get_op_type(), ops[0]);
            if (ops.size() == 2) return expr(e.get_op_type(), ops[0], ops[1]);
            return expr(e.get_op_type(), ops[0], ops[1], ops[2]);
        }
    };
    for (computation *c : f_->get_computations())
        if (!c->is_inline_computation() && c->get_expr().is_defined()) c->replace_expression(expand(c->get_expr(), 0));
}

// Loop-level constants (`constant(name, e, t, false, &S, level)`, reference tests/test_04.cpp): lets.  Each is
// replaced by its (cast) value in every statement that names it; later lets may use earlier ones.
void Lowering::expand_lets()
{
    std::vector<std::string> names;
    std::vector<expr> values;
    for (computation *c : f_->get_computations()) {
        if (!c->is_let()) continue;
        names.push_back(c->get_name());
        values.push_back(cast(c->get_data_type(), c->get_expr().substitute(names, values)));
    }
    if (names.empty()) return;
    for (computation *c : f_->get_computations())
        if (!c->is_let()) c->substitute_everywhere(names, values);
}

void Lowering::propagate_constants()
{
    std::vector<std::string> names;
    std::vector<expr> values;
    for (computation *c : f_->get_computations()) {
        if (!c->is_constant() || c->is_let()) continue;
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

// Which original iterators of a statement may run in parallel: the map (iterators) -> (stored element) must be
// JOINTLY injective over them, otherwise two threads would read-modify-write one element (S(i,j) stored at
// b_O[i+j], a polynomial multiply, has every iterator in the store index and is still a reduction along the
// anti-diagonals).  Triangular elimination over the store index expressions: an iterator that is the only
// not-yet-determined one of an affine index is determined by it; iterators that share an index are all determined
// only when strides and extents separate them (mixed radix: i*16 + j with 0 <= j < 16), else one of them (largest
// stride, outermost on ties) is kept and the others are sequential.  An iterator that occurs in a non-affine index
// (y/2, i%4: several iterations per element, the 2x2 max-pool of conv_relu_maxpool_generator_tiramisu.cpp:104) or
// in no index at all (a reduction) is sequential.
std::set<std::string> Lowering::parallel_iterators(computation *c)
{
    std::set<std::string> det, seq;
    std::vector<std::string> iters;
    for (const var &v : c->get_iteration_variables()) iters.push_back(v.get_name());
    auto is_iter = [&](const std::string &n) { return std::find(iters.begin(), iters.end(), n) != iters.end(); };
    struct Idx { bool affine; std::map<std::string, int64_t> coef; };
    std::vector<Idx> idx;
    for (const expr &a : c->access()) {
        Idx x; int64_t cst = 0;
        x.affine = affine_form(a, x.coef, cst);
        if (!x.affine) {
            for (const std::string &n : iters)
                if (!a.substitute({n}, {expr((int32_t) 0)}).is_equal(a) || (a.get_expr_type() == e_var && a.get_name() == n)) seq.insert(n);
            continue;
        }
        for (auto it = x.coef.begin(); it != x.coef.end();) it = (it->second == 0 || !is_iter(it->first)) ? x.coef.erase(it) : std::next(it);
        idx.push_back(x);
    }
    auto extent_of = [&](const std::string &n, int64_t &ext) {
        for (const var &v : c->get_iteration_variables())
            if (v.get_name() == n && v.has_range()) {
                expr e = v.get_upper() - v.get_lower();
                if (e.is_int_literal()) { ext = e.get_int_val(); return true; }
            }
        return false;
    };
    for (;;) {
        bool progress = false, open = false;
        for (const Idx &x : idx) {
            std::vector<std::string> u;
            for (auto &kv : x.coef) if (!det.count(kv.first) && !seq.count(kv.first)) u.push_back(kv.first);
            if (u.size() == 1) { det.insert(u[0]); progress = true; }
            else if (u.size() > 1) open = true;
        }
        if (progress) continue;
        if (!open) break;
        for (const Idx &x : idx) {
            std::vector<std::pair<int64_t, std::string>> u;                 // (|stride|, name) of the undetermined iterators
            for (auto &kv : x.coef) if (!det.count(kv.first) && !seq.count(kv.first)) u.push_back({kv.second < 0 ? -kv.second : kv.second, kv.first});
            if (u.size() < 2) continue;
            std::stable_sort(u.begin(), u.end(), [](const std::pair<int64_t, std::string> &a, const std::pair<int64_t, std::string> &b) { return a.first < b.first; });
            bool separable = true;
            for (size_t k = 0; k + 1 < u.size() && separable; k++) {
                int64_t ext = 0;
                separable = extent_of(u[k].second, ext) && u[k + 1].first >= u[k].first * ext;
            }
            if (separable) {
                for (auto &e : u) det.insert(e.second);
            } else {
                // keep the iterator with the largest stride (the outermost in iteration order among equals)
                size_t keep = u.size() - 1;
                for (size_t k = 0; k < u.size(); k++)
                    if (u[k].first == u.back().first) { keep = k; break; }
                for (size_t k = 0; k < u.size(); k++) (k == keep ? det : seq).insert(u[k].second);
            }
            break;
        }
    }
    for (const std::string &n : seq) det.erase(n);
    // A sequential iterator runs INSIDE the thread, at a different value in every thread at any moment: a parallel
    // iterator that shares a store index with it would let two threads meet on one element after all (thread i at
    // step j and thread i+1 at step j-1 both update b_O[i+j]).  Whatever shares an index with a sequential iterator
    // is therefore sequential too — a statement like the polynomial multiply ends up in one thread, in its own order.
    std::vector<std::set<std::string>> groups;
    for (const Idx &x : idx) {
        std::set<std::string> g;
        for (auto &kv : x.coef) g.insert(kv.first);
        groups.push_back(g);
    }
    for (const expr &a : c->access()) {
        std::map<std::string, int64_t> coef; int64_t cst = 0;
        if (affine_form(a, coef, cst)) continue;
        std::set<std::string> g;
        for (const std::string &n : iters)
            if (!a.substitute({n}, {expr((int32_t) 0)}).is_equal(a)) g.insert(n);
        groups.push_back(g);
    }
    for (bool changed = true; changed;) {
        changed = false;
        for (const auto &g : groups) {
            bool has_seq = false;
            for (const std::string &n : g) has_seq |= !det.count(n);
            if (!has_seq) continue;
            for (const std::string &n : g) changed |= det.erase(n) > 0;
        }
    }
    return det;
}

// A scheduled dimension is data-parallel iff every original iterator rebuilt from it is one of the jointly
// injective ones above (each iteration writes its own element); a dimension that feeds no original iterator is not.
bool Lowering::dim_is_parallel(computation *c, const std::string &dim)
{
    if (!c->has_access() && !c->get_expr().is_defined()) return false;
    if (c->is_communicator()) return false;
    { const int l = c->find_dim(dim); if (l >= 0 && c->dims()[(size_t) l].tag == loop_dim::distributed) return false; }   // the rank scope is not a loop to map
    c->get_buffer();
    auto hit = parallel_iters_.find(c);
    if (hit == parallel_iters_.end()) hit = parallel_iters_.emplace(c, parallel_iterators(c)).first;
    bool feeds_any = false;
    for (const var &v : c->get_iteration_variables()) {
        expr full = expr::make_var(p_int32, v.get_name());
        for (auto &d : c->defs()) full = full.substitute({d.first}, {d.second});
        const bool feeds = !full.substitute({dim}, {expr((int32_t) 7)}).is_equal(full);
        if (!feeds) continue;
        feeds_any = true;
        if (!hit->second.count(v.get_name())) return false;
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
            // keep = the shared loops that must stay fused.  A loop that is data parallel for every statement of the run can be
            // un-fused when the statements only exchange values element-wise (below).  A loop that is NOT (a max-pool
            // accumulating over y and x into out[y/2][x/2], vgg_block_generator_tiramisu.cpp:75-84) can still be split BETWEEN
            // statements — each keeps its own iterations in order — when every value that crosses from one statement to another
            // does so inside one iteration of it: some dimension of the buffer is indexed, in both accesses, by the same plain
            // iterator built the same way from that loop (same element => same iteration).
            int keep = 0;
            if (ok) {
                std::vector<bool> par_all((size_t) common, true);
                for (int d = 0; d < common; d++)
                    for (size_t k = i; k < j; k++)
                        if (!dim_is_parallel(ord[k], ord[k]->dims()[d].name)) par_all[(size_t) d] = false;
                auto in_dims = [](computation *c, expr e) { for (auto &df : c->defs()) e = e.substitute({df.first}, {df.second}); return e; };
                auto mentions = [](const expr &e, const std::string &n) { return !e.substitute({n}, {expr((int32_t) 0)}).is_equal(e) || (e.get_expr_type() == e_var && e.get_name() == n); };
                auto same_dims = [&](computation *a, computation *b, const expr &e) {      // the loops e names have the same range in a and b
                    for (const loop_dim &x : a->dims()) {
                        if (!mentions(e, x.name)) continue;
                        const int o = b->find_dim(x.name);
                        if (o < 0 || !b->dims()[(size_t) o].lo.is_equal(x.lo) || !b->dims()[(size_t) o].hi.is_equal(x.hi)) return false;
                    }
                    return true;
                };
                // loops [from, common) un-fused: are p (statement a) and q (statement b) the same element only in the same iteration?
                auto compatible = [&](computation *a, const Access &p, computation *b, const Access &q, int from) {
                    if (p.idx.size() != q.idx.size()) return false;
                    std::set<int> pinned, pinned_plain;
                    bool all_identical = true;
                    for (size_t t = 0; t < p.idx.size(); t++) {
                        const expr pa = in_dims(a, p.idx[t]), qb = in_dims(b, q.idx[t]);
                        const bool identical = p.idx[t].is_equal(q.idx[t]);
                        all_identical &= identical;
                        bool p_sh = false, q_sh = false;
                        for (int d = from; d < common; d++) { p_sh |= mentions(pa, a->dims()[d].name); q_sh |= mentions(qb, b->dims()[d].name); }
                        if (identical && pa.is_equal(qb) && same_dims(a, b, pa)) {
                            for (int d = from; d < common; d++)
                                if (mentions(pa, a->dims()[d].name)) { pinned.insert(d); if (p.idx[t].get_expr_type() == e_var) pinned_plain.insert(d); }
                        } else if (!identical && (p_sh || q_sh)) return false;
                    }
                    for (int d = from; d < common; d++) {
                        if (!par_all[(size_t) d]) { if (!pinned_plain.count(d)) return false; }
                        else if (!all_identical && !pinned.count(d)) return false;
                    }
                    return true;
                };
                // out[f] = max(out[f], v) with v not reading out (min likewise): updates of this kind on one buffer commute
                // exactly, whichever statement they come from (a max-pool and its conclude column meet in the same output
                // element, vgg_block_generator_tiramisu.cpp:75-96); 0 = not such a statement
                auto commutative_update = [&](computation *c) -> int {
                    const expr &e = c->get_expr();
                    if (!c->has_access() || !e.is_defined() || e.get_expr_type() != e_op || (e.get_op_type() != o_max && e.get_op_type() != o_min)) return 0;
                    int self = 0, others = 0;
                    for (int side = 0; side < 2; side++) {
                        const expr &x = e.get_operand(side);
                        bool is_self = false;
                        if (x.get_expr_type() == e_op && x.get_op_type() == o_access) {
                            computation *oc = nullptr;
                            if (buffer_of_access(x, &oc) == c->get_buffer() && (int) oc->get_iteration_variables().size() == x.get_n_arg() &&
                                oc->access().size() == c->access().size()) {
                                std::vector<std::string> names;
                                for (const var &v : oc->get_iteration_variables()) names.push_back(v.get_name());
                                is_self = true;
                                for (size_t t = 0; t < c->access().size(); t++) is_self &= oc->access()[t].substitute(names, x.get_operands()).is_equal(c->access()[t]);
                            }
                        }
                        if (is_self) { self++; continue; }
                        std::vector<expr> stack{x};
                        while (!stack.empty()) {
                            expr y = stack.back(); stack.pop_back();
                            if (!y.is_defined() || y.get_expr_type() != e_op) continue;
                            if (y.get_op_type() == o_access) { computation *oc = nullptr; if (buffer_of_access(y, &oc) == c->get_buffer()) others++; }
                            for (const expr &o : y.get_operands()) stack.push_back(o);
                        }
                    }
                    return self == 1 && others == 0 ? (e.get_op_type() == o_max ? 1 : 2) : 0;
                };
                std::map<computation *, int> comm;
                for (size_t k = i; k < j; k++) comm[ord[k]] = commutative_update(ord[k]);
                std::vector<std::pair<computation *, Access>> all;
                for (size_t k = i; k < j; k++)
                    for (const Access &p : accesses_of(ord[k])) all.push_back({ord[k], p});
                auto exchange_ok = [&](int from) {
                    for (size_t x = 0; x < all.size(); x++)
                        for (size_t y = 0; y < x; y++) {
                            if (all[x].first == all[y].first) continue;
                            const Access &p = all[x].second, &q = all[y].second;
                            if (p.b != q.b || (!p.write && !q.write)) continue;
                            if (comm[all[x].first] && comm[all[x].first] == comm[all[y].first] && all[x].first->get_buffer() == p.b &&
                                all[y].first->get_buffer() == p.b) continue;
                            if (!compatible(all[x].first, p, all[y].first, q, from)) return false;
                        }
                    return true;
                };
                // the outermost level from which everything below can be un-fused
                keep = common;
                for (int from = 0; from < common; from++)
                    if (exchange_ok(from)) { keep = from; break; }
                if (keep == common) ok = false;
            }
            // Only levels below `common` were checked, so only the boundaries fused exactly `common` deep are lowered: a
            // boundary fused deeper than that (S1.then(S2, k) inside a run fused at i) keeps every level it shares — S2 goes
            // on reading the running t[i] inside (i, k) — and the loops [keep, common) are split BETWEEN the sub-runs on
            // either side of the lowered boundaries, which is what the pairwise check above covers.  The sub-runs are
            // looked at again on the next pass, with their own (deeper) common level.
            if (ok && keep < common)
                for (size_t k = i + 1; k < j; k++)
                    if (fuse_depth(ord[k - 1], ord[k]) == common) { ord[k]->set_pred_level_index(keep - 1); changed_ = true; }
        }
        i = j;
    }
}

// Fused runs whose OUTER shared loops are data parallel while an inner one is not (reference tests/test_85.cpp: a
// reduction over k with its zero-init fused at k, inside i and j).  Distribution cannot un-fuse them (k carries the
// dependence) and as host loops they cost a launch per (i, j, k).  Instead the shared data-parallel loops become
// the grid and everything inside — the sequential shared loops and the statements' own loops — runs in the
// thread, in program order.  Legal under the same two facts distribution relies on: every statement writes an
// element of its own per iteration of the mapped loops (dim_is_parallel), and the statements only exchange values
// element-wise (identical index expressions for every pair of accesses to one buffer that includes a write), so two
// threads never touch the same element of a written buffer.
// Checked in host emulation (tests/emulation: reference test_85 passes its wrapper, also under AddressSanitizer) and
// it leaves every generated source that has run on a B200 byte-identical; it has not itself run on a B200 yet.
// TIRAMISU_B200_MAP_FUSED=0 at generator time keeps such runs as host loops.
void Lowering::map_fused_parallel_loops(const std::vector<computation *> &ord)
{
    auto plain = [](computation *c) {
        const expr &e = c->get_expr();
        return e.is_defined() && e.get_expr_type() != e_sync && c->has_access() && !c->get_predicate().is_defined() &&
               !(e.get_expr_type() == e_op && (e.get_op_type() == o_memcpy || e.get_op_type() == o_allocate || e.get_op_type() == o_free || e.get_op_type() == o_call));
    };
    size_t i = 0;
    while (i < ord.size()) {
        size_t j = i + 1;
        while (j < ord.size() && fuse_depth(ord[j - 1], ord[j]) > 0) j++;
        const size_t first = i, last = j;
        i = j;
        if (last - first < 2) continue;
        bool ok = true;
        for (size_t k = first; k < last; k++) ok &= plain(ord[k]);
        if (!ok) continue;
        std::vector<int> depth(last - first, 0);
        int common = INT32_MAX;
        for (size_t k = first + 1; k < last; k++) { depth[k - first] = fuse_depth(ord[k - 1], ord[k]); common = std::min(common, depth[k - first]); }
        // P leading shared loops: same name (index expressions are compared by iterator name) and bounds in every statement, data parallel for every statement, no tag of the user's that maps
        int P = 0;
        for (int d = 0; d < common && ok; d++) {
            const loop_dim &ref = ord[first]->dims()[d];
            bool par = true;
            for (size_t k = first; k < last; k++) {
                const loop_dim &x = ord[k]->dims()[d];
                par &= x.name == ref.name && x.lo.is_equal(ref.lo) && x.hi.is_equal(ref.hi) && x.tag != loop_dim::gpu_block && x.tag != loop_dim::gpu_thread &&
                       dim_is_parallel(ord[k], x.name);
            }
            if (!par) break;
            P = d + 1;
        }
        if (P == 0 || P == common) continue;           // P == common: nothing sequential is shared — distribution's case
        // element-wise exchange only
        std::vector<Access> seen;
        for (size_t k = first; k < last && ok; k++) {
            std::vector<Access> mine = accesses_of(ord[k]);
            for (const Access &p : mine)
                for (const Access &q : seen) {
                    if (p.b != q.b || (!p.write && !q.write)) continue;
                    bool identical = p.idx.size() == q.idx.size();
                    for (size_t t = 0; identical && t < p.idx.size(); t++) identical = p.idx[t].is_equal(q.idx[t]);
                    if (!identical) ok = false;
                }
            // ... and inside one statement (a read of the written buffer at another element)
            for (const Access &p : mine)
                for (const Access &q : mine) {
                    if (&p == &q || p.b != q.b || (!p.write && !q.write)) continue;
                    bool identical = p.idx.size() == q.idx.size();
                    for (size_t t = 0; identical && t < p.idx.size(); t++) identical = p.idx[t].is_equal(q.idx[t]);
                    if (!identical) ok = false;
                }
            seen.insert(seen.end(), mine.begin(), mine.end());
        }
        if (!ok) continue;
        for (size_t k = first; k < last; k++) {
            computation *c = ord[k];
            for (auto &d : c->dims()) if (d.tag != loop_dim::distributed) d.tag = loop_dim::none;
            const std::string inner = c->dims()[P - 1].name;
            if (P >= 2) c->dims()[P - 2].tag = loop_dim::gpu_block;
            if (P >= 3) c->dims()[P - 3].tag = loop_dim::gpu_block;
            c->split(var(inner), 256, var(inner + "_b"), var(inner + "_t"));
            c->tag_gpu_block_level(var(inner + "_b"));
            c->tag_gpu_thread_level(var(inner + "_t"));
            if (k > first) c->set_pred_level_index(depth[k - first]);      // one more shared loop above the fuse level
            grid_fused_.insert(c);
        }
    }
}

void Lowering::auto_map(computation *c, int shared_depth)
{
    if (c->is_communicator()) return;                                // host loops around tiramisu_MPI_* calls
    if (c->get_distributed_dimension() >= 0) shared_depth = std::max(shared_depth, c->get_distributed_dimension() + 1);   // the rank scope stays on the host
    const expr &e = c->get_expr();
    if (e.get_expr_type() == e_sync || (e.get_expr_type() == e_op && (e.get_op_type() == o_memcpy || e.get_op_type() == o_allocate ||
                                                                       e.get_op_type() == o_free || (e.get_op_type() == o_call && !c->has_access()))))
        return;
    auto &dims = c->dims();
    if (grid_fused_.count(c)) return;             // mapped as part of its fused run: its own loops stay in the thread
    auto one_thread = [&]() {
        // nothing to spread over threads (a scalar statement, or every loop carries a dependence): the
        // statement's own loops run sequentially inside a one-thread kernel
        loop_dim one; one.name = "_one_" + c->get_name(); one.lo = expr((int32_t) 0); one.hi = expr((int32_t) 1); one.tag = loop_dim::gpu_block;
        for (auto &d : dims) if (d.tag != loop_dim::distributed) d.tag = loop_dim::none;
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
    for (auto &d : dims) if (d.tag != loop_dim::distributed) d.tag = loop_dim::none;
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
    // no vector loop, but the innermost data-parallel loop is short (a blocked channel dimension of 8): it plays that role —
    // cut by 256 it would leave most of the block idle
    if (vec < 0) {
        int q = -1;
        for (int i = (int) dims.size() - 1; i >= shared_depth; i--) if (parallel_safe(dims[(size_t) i].name)) { q = i; break; }
        if (q > shared_depth) {
            const expr ext = dims[(size_t) q].hi - dims[(size_t) q].lo;
            bool outer_par = false;
            for (int i = shared_depth; i < q; i++) outer_par |= parallel_safe(dims[(size_t) i].name);
            if (outer_par && ext.is_int_literal() && ext.get_int_val() >= 2 && ext.get_int_val() <= 32 && dims[(size_t) q].lo.is_int_literal() &&
                dims[(size_t) q].lo.get_int_val() == 0 && parallel_safe(dims[(size_t) q - 1].name))
                vec = q;
        }
    }
    std::string tx, ty, bx;
    // the row loop of the thread block is the nearest data-parallel loop outside the vector loop that runs more than once
    // (vectorize(ffout, 8) of an 8-wide iterator leaves an outer part of extent 1: mapping THAT onto threadIdx.y left 31 of 32
    // rows idle, r02 run 28: 85 ms for the conv layer); it is moved next to the vector loop — both are data parallel
    if (vec >= shared_depth + 1) {
        auto once = [&](const loop_dim &d) { return d.lo.is_int_literal() && d.hi.is_int_literal() && d.hi.get_int_val() - d.lo.get_int_val() <= 1; };
        int o = vec - 1;
        while (o >= shared_depth && once(dims[(size_t) o]) && parallel_safe(dims[(size_t) o].name)) o--;
        if (o >= shared_depth && o != vec - 1 && parallel_safe(dims[(size_t) o].name)) {
            bool all_par = true;
            for (int q = o; q < vec; q++) all_par &= parallel_safe(dims[(size_t) q].name);
            if (all_par) std::rotate(dims.begin() + o, dims.begin() + o + 1, dims.begin() + vec);
        }
    }
    if (vec >= shared_depth + 1 && parallel_safe(dims[vec - 1].name) && dims[vec].hi.is_int_literal() && dims[vec].hi.get_int_val() <= 64) {
        const int lanes = (int) dims[vec].hi.get_int_val();
        int rows = std::max(1, 256 / lanes);
        {
            const expr ext = dims[vec - 1].hi - dims[vec - 1].lo;
            if (ext.is_int_literal() && ext.get_int_val() >= 1 && ext.get_int_val() < rows) rows = (int) ext.get_int_val();
            else if (ext.is_int_literal() && ext.get_int_val() > rows) {
                // equal blocks: 37 rows are two blocks of 19, not 32 + 5 (the second block would be 5 / 32 full)
                const int64_t nblk = (ext.get_int_val() + rows - 1) / rows;
                rows = (int) ((ext.get_int_val() + nblk - 1) / nblk);
            }
        }
        tx = dims[vec].name;
        std::string outer = dims[vec - 1].name;
        const expr oext = dims[vec - 1].hi - dims[vec - 1].lo;
        if (oext.is_int_literal() && oext.get_int_val() == rows) ty = outer;          // the whole row loop fits the block: no block part
        else { bx = outer + "_b"; ty = outer + "_t"; c->split(var(outer), rows, var(bx), var(ty)); }
    } else {
        int q = -1;
        for (int i = (int) dims.size() - 1; i >= shared_depth; i--) if (parallel_safe(dims[i].name)) { q = i; break; }
        if (q < 0) { one_thread(); return; }
        std::string inner = dims[q].name;
        bx = inner + "_b"; tx = inner + "_t";
        c->split(var(inner), 256, var(bx), var(tx));
    }
    if (!bx.empty()) c->tag_gpu_block_level(var(bx));
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
    const int max_extra = bx.empty() ? 3 : 2;
    for (size_t i = shared_depth; i < c->dims().size() && extra < max_extra; i++) {
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
        if (c->is_communicator()) {
            // messages travel between the DEVICE mirrors: a receive writes its buffer, a send reads its payload's
            if (c->is_recv() && c->has_access()) { used.insert(c->get_buffer()); written_args_.insert(c->get_buffer()); }
            if (c->is_send()) {
                computation *oc = f_->find_computation(c->get_expr().get_name());
                if (oc && !oc->is_constant()) used.insert(oc->get_buffer());
            }
            continue;
        }
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

// A CPU schedule puts the loops of an accumulation in cache order: reduction loops outside, register-blocked output
// loops inside (conv_layer_generator_tiramisu.cpp: n, y, fout_b_up, fin_b, x_b, k_y, k_x, ffin, fout_b_low, ffout, xx).
// For the device the output loops are the grid and the reduction runs inside the thread, so below the loops the statement
// shares with its neighbours the data-parallel loops are moved outside the others, each class keeping its order.  Legal
// for a statement whose only dependence on itself is the accumulation into the element it stores: every element still
// sees its own additions in the original order.
void Lowering::sink_reduction_loops(const std::vector<computation *> &ord)
{
    for (size_t i = 0; i < ord.size(); i++) {
        computation *c = ord[i];
        if (!c->has_access() || !c->get_expr().is_defined() || c->is_communicator()) continue;
        const int keep = std::max(i ? fuse_depth(ord[i - 1], c) : 0, i + 1 < ord.size() ? fuse_depth(c, ord[i + 1]) : 0);
        auto &dims = c->dims();
        if ((int) dims.size() - keep < 2) continue;
        // reads of the statement's own buffer: only the element it stores
        bool ok = true;
        std::vector<expr> stack{c->get_expr(), c->get_predicate()};
        while (!stack.empty() && ok) {
            expr x = stack.back(); stack.pop_back();
            if (!x.is_defined() || x.get_expr_type() != e_op) continue;
            for (const expr &o : x.get_operands()) stack.push_back(o);
            if (x.get_op_type() != o_access) continue;
            computation *oc = nullptr;
            if (buffer_of_access(x, &oc) != c->get_buffer()) continue;
            if ((int) oc->get_iteration_variables().size() != x.get_n_arg() || oc->access().size() != c->access().size()) { ok = false; break; }
            std::vector<std::string> names;
            for (const var &v : oc->get_iteration_variables()) names.push_back(v.get_name());
            for (size_t q = 0; q < c->access().size() && ok; q++) ok = oc->access()[q].substitute(names, x.get_operands()).is_equal(c->access()[q]);
        }
        if (!ok) continue;
        // ... except the CPU author's register tile: short data-parallel loops tagged for unrolling INSIDE the reduction
        // (xx and fout_b_low of the conv layer: 3 x 4 outputs per filter / input value loaded) stay there; the emitter
        // keeps the tile's accumulators in registers (lowering_emit.cpp: try_emit_accumulation_tile — without that the
        // tile inside is a loss, 54.7 vs 17.3 ms, r02 runs 29 / 30)
        std::vector<loop_dim> par, seq, tile;
        bool seq_seen = false, moved = false;
        for (size_t q = (size_t) keep; q < dims.size(); q++) {
            if (dims[q].tag == loop_dim::distributed || dims[q].tag == loop_dim::gpu_block || dims[q].tag == loop_dim::gpu_thread) { par.clear(); seq.clear(); tile.clear(); moved = false; break; }
            const expr ext = dims[q].hi - dims[q].lo;
            const bool short_unrolled = dims[q].tag == loop_dim::unrolled && ext.is_int_literal() && ext.get_int_val() <= 8 && !getenv("TIRAMISU_B200_NO_REGISTER_TILE");
            if (!dim_is_parallel(c, dims[q].name)) { seq.push_back(dims[q]); seq_seen = true; }
            else if (seq_seen && short_unrolled) tile.push_back(dims[q]);
            else { par.push_back(dims[q]); moved |= seq_seen; }
        }
        if (!moved) continue;
        std::copy(par.begin(), par.end(), dims.begin() + keep);
        std::copy(seq.begin(), seq.end(), dims.begin() + keep + (long) par.size());
        std::copy(tile.begin(), tile.end(), dims.begin() + keep + (long) (par.size() + seq.size()));
    }
}

// A pooling statement stores out[..][y / 2][x / 2][..] = max(out[..], v(y, x)): as iterators of the stored element y and x
// are not injective, so nothing about them could be mapped and the statement ran its whole (y, x) plane inside each
// thread (VGG block: 9.2 ms of an 11.8 ms call, r02 run 29).  An iterator that reaches the store only as `it / k` is cut
// into it = k q + r (Layer I, not the schedule): q indexes the store directly and is data parallel, r (k iterations)
// stays the sequential accumulation.  Only for a statement that stands alone in the order (its schedule described loops
// that no longer exist and is replaced by the plain nest) and that nothing names in an access.
void Lowering::strip_divided_iterators(const std::vector<computation *> &ord)
{
    for (size_t i = 0; i < ord.size(); i++) {
        computation *c = ord[i];
        if (!c->has_access() || !c->get_expr().is_defined() || c->is_communicator()) continue;
        if ((i && fuse_depth(ord[i - 1], c) > 0) || (i + 1 < ord.size() && fuse_depth(c, ord[i + 1]) > 0)) continue;
        bool tagged = false;
        for (const loop_dim &d : c->dims()) tagged |= d.tag == loop_dim::gpu_block || d.tag == loop_dim::gpu_thread || d.tag == loop_dim::distributed;
        if (tagged) continue;
        // named in somebody's access?
        bool named = false;
        for (computation *o : f_->get_computations()) {
            std::vector<expr> stack{o->get_expr(), o->get_predicate()};
            while (!stack.empty() && !named) {
                expr x = stack.back(); stack.pop_back();
                if (!x.is_defined() || x.get_expr_type() != e_op) continue;
                if (x.get_op_type() == o_access && x.get_name() == c->get_name()) named = true;
                for (const expr &a : x.get_operands()) stack.push_back(a);
            }
        }
        if (named) continue;
        auto mentions = [](const expr &e, const std::string &n) { return !e.substitute({n}, {expr((int32_t) 0)}).is_equal(e) || (e.get_expr_type() == e_var && e.get_name() == n); };
        std::vector<var> iters = c->get_iteration_variables();
        std::vector<std::string> names; std::vector<expr> values; std::vector<expr> guards;
        std::vector<expr> access = c->access();
        std::vector<var> out;
        bool any = false;
        for (const var &v : iters) {
            int64_t k = 0; int where = -1, uses = 0;
            for (size_t t = 0; t < access.size(); t++) {
                if (!mentions(access[t], v.get_name())) continue;
                uses++;
                const expr &a = access[t];
                if (a.get_expr_type() == e_op && a.get_op_type() == o_div && a.get_operand(0).get_expr_type() == e_var && a.get_operand(0).get_name() == v.get_name() &&
                    a.get_operand(1).is_int_literal() && a.get_operand(1).get_int_val() >= 2) { k = a.get_operand(1).get_int_val(); where = (int) t; }
            }
            const bool zero_based = v.has_range() && v.get_lower().is_int_literal() && v.get_lower().get_int_val() == 0;
            if (uses != 1 || where < 0 || !zero_based) { out.push_back(v); continue; }
            const expr hi = v.get_upper();
            const bool divides = hi.is_int_literal() && hi.get_int_val() % k == 0;
            const expr q_hi = hi.is_int_literal() ? expr((int32_t) ((hi.get_int_val() + k - 1) / k)) : (hi + (int32_t) (k - 1)) / expr((int32_t) k);
            var q(v.get_name() + "_q", expr((int32_t) 0), q_hi), r(v.get_name() + "_r", expr((int32_t) 0), expr((int32_t) k));
            out.push_back(q); out.push_back(r);
            const expr val = expr::make_var(global::get_loop_iterator_data_type(), q.get_name()) * expr((int32_t) k) + expr::make_var(global::get_loop_iterator_data_type(), r.get_name());
            names.push_back(v.get_name()); values.push_back(val);
            access[(size_t) where] = expr::make_var(global::get_loop_iterator_data_type(), q.get_name());
            if (!divides) guards.push_back(val < hi);
            any = true;
        }
        if (!any || c->get_predicate().is_defined()) continue;
        for (size_t t = 0; t < access.size(); t++) access[t] = access[t].substitute(names, values);     // (the rewritten dimensions no longer name v)
        // the expression in the new iterators; a read of the element the statement stores (out[y/2][x/2] through a view)
        // becomes an access to the statement itself at its own iteration, so that it still reads exactly `access`
        std::vector<expr> self_args;
        for (const var &v : out) self_args.push_back(expr::make_var(global::get_loop_iterator_data_type(), v.get_name()));
        const std::vector<expr> old_access = c->access();
        bool fail = false;
        std::function<expr(const expr &)> rebuild = [&](const expr &e) -> expr {
            if (!e.is_defined()) return e;
            if (e.get_expr_type() != e_op) return e.substitute(names, values);
            const op_t o = e.get_op_type();
            if (o == o_access) {
                computation *oc = nullptr;
                if (buffer_of_access(e, &oc) == c->get_buffer()) {
                    bool self = oc->access().size() == old_access.size() && (int) oc->get_iteration_variables().size() == e.get_n_arg();
                    if (self) {
                        std::vector<std::string> nn;
                        for (const var &v : oc->get_iteration_variables()) nn.push_back(v.get_name());
                        for (size_t q = 0; q < old_access.size(); q++) self &= oc->access()[q].substitute(nn, e.get_operands()).is_equal(old_access[q]);
                    }
                    if (!self) { fail = true; return e; }
                    return expr(o_access, c->get_name(), self_args, e.get_data_type());
                }
            }
            std::vector<expr> ops;
            for (const expr &x : e.get_operands()) ops.push_back(rebuild(x));
            if (o == o_access || o == o_call) return expr(o, e.get_name(), ops, e.get_data_type());
            if (o == o_cast) return expr(o_cast, e.get_data_type(), ops[0]);
            if (ops.size() == 1) return expr(o, ops[0]);
            if (ops.size() == 2) return expr(o, ops[0], ops[1]);
            if (ops.size() == 3) return expr(o, ops[0], ops[1], ops[2]);
            fail = true;
            return e;
        };
        const expr new_expr = rebuild(c->get_expr());
        if (fail) continue;
        c->replace_expression(new_expr);
        names.clear(); values.clear();                                   // the expression is already in the new iterators
        std::cerr << "tiramisu_b200 lowering: '" << c->get_name() << "' stores through `iterator / constant`: the iterator is cut into quotient (data parallel) and remainder\n";
        c->rewrite_iterators(out, names, values, access, guards);
        parallel_iters_.erase(c);
    }
}

void Lowering::retarget(const std::vector<computation *> &ord)
{
    retargeted_ = true;
    if (pattern_) sink_reduction_loops(ord);
    if (pattern_) {
        // until nothing moves: un-fusing an outer level turns its sub-runs into runs of their own
        for (int pass = 0; pass < 16; pass++) {
            changed_ = false;
            distribute_parallel_shared_loops(ord);
            if (!changed_) break;
        }
    }
    if (pattern_) { strip_divided_iterators(ord); sink_reduction_loops(ord); }      // the remainders are reduction loops now
    const char *map_fused = getenv("TIRAMISU_B200_MAP_FUSED");
    if (pattern_ && !(map_fused && map_fused[0] == '0')) map_fused_parallel_loops(ord);
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

}  // namespace lowering_detail

// Bound inference for box domains (reference core.h compute_bounds; tests/test_22-24.cpp).  A computation declared
// without bounds (`{S0[i,j]}`) gets, per iterator, the hull of what its consumers read: consumers are visited last
// declared first, so that a producer's box is complete when it is itself looked at as a consumer.  Each argument of
// an access must be affine in the consumer's iterators; its interval follows from their ranges.
void function::compute_bounds()
{
    using lowering_detail::affine_form;
    using lowering_detail::die;
    auto lit_or = [](const expr &a, const expr &b, bool want_min) {
        if (a.is_equal(b)) return a;
        if (a.is_int_literal() && b.is_int_literal())
            return expr((int32_t) (want_min ? std::min(a.get_int_val(), b.get_int_val()) : std::max(a.get_int_val(), b.get_int_val())));
        return expr(want_min ? o_min : o_max, a, b);
    };
    std::map<computation *, std::vector<std::pair<expr, expr>>> hull;
    for (auto it = comps_.rbegin(); it != comps_.rend(); ++it) {
        computation *c = *it;
        if (c->pending_bounds()) {
            auto h = hull.find(c);
            bool complete = h != hull.end();
            if (complete) for (auto &r : h->second) complete &= r.first.is_defined() && r.second.is_defined();
            if (!complete) die("compute_bounds: nothing that has bounds reads '" + c->get_name() + "'");
            c->set_iterator_ranges(h->second);
        }
        if (c->is_input() || !c->get_expr().is_defined()) continue;
        std::map<std::string, std::pair<expr, expr>> range;
        for (const var &v : c->get_iteration_variables()) if (v.has_range()) range[v.get_name()] = {v.get_lower(), v.get_upper()};
        std::vector<expr> stack{c->get_expr(), c->get_predicate()};
        while (!stack.empty()) {
            expr x = stack.back(); stack.pop_back();
            if (!x.is_defined() || x.get_expr_type() != e_op) continue;
            for (const expr &o : x.get_operands()) stack.push_back(o);
            if (x.get_op_type() != o_access) continue;
            computation *p = find_computation(x.get_name());
            if (!p || !p->pending_bounds() || (int) p->get_iteration_variables().size() != x.get_n_arg()) continue;
            auto &h = hull[p];
            h.resize(x.get_n_arg());
            for (int k = 0; k < x.get_n_arg(); k++) {
                std::map<std::string, int64_t> coef; int64_t cst = 0;
                if (!affine_form(x.get_operand(k), coef, cst)) die("compute_bounds: '" + c->get_name() + "' reads '" + p->get_name() + "' at a non-affine index");
                expr lo((int32_t) cst), hi((int32_t) cst);
                for (auto &kv : coef) {
                    if (kv.second == 0) continue;
                    const expr f((int32_t) kv.second);
                    auto r = range.find(kv.first);
                    if (r == range.end()) {                       // a parameter: the same term at both ends
                        expr v = expr::make_var(p_int32, kv.first);
                        lo = lo + f * v; hi = hi + f * v;
                    } else if (kv.second > 0) {
                        lo = lo + f * r->second.first; hi = hi + f * (r->second.second - 1);
                    } else {
                        lo = lo + f * (r->second.second - 1); hi = hi + f * r->second.first;
                    }
                }
                hi = hi + 1;
                h[k].first = h[k].first.is_defined() ? lit_or(h[k].first, lo, true) : lo;
                h[k].second = h[k].second.is_defined() ? lit_or(h[k].second, hi, false) : hi;
            }
        }
    }
}

}  // namespace tiramisu
