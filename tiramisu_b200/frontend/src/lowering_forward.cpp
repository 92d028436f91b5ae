// lowering_forward.cpp — register blocking of CPU schedules, undone (see lowering.h).
//
// A CPU generator that wants its accumulators in registers writes the update of an output element as THREE statements
// (benchmarks/DNN/layers/convolution/direct/cpu/conv_layer_generator_tiramisu.cpp:51-64, :173-183):
//     L:  T[g] = P[f]                 load the element into a small temporary T ("reg_buf")
//     U:  T[g] = T[g] + <terms>       accumulate over further (reduction) iterators
//     S:  P[f] = T[g]                 store it back
// fused inside the loops that walk the output, with g made of `iterator % block` terms so that T is a few vectors long.  On
// the CPU that is a performance idiom; for a data-parallel device it is a chain of dependences through one tiny buffer that
// every iteration of the outer loops shares, and the nest re-targets to millions of launches (round 1 refused the
// generator with that count).  Where T is PRIVATE to one iteration of the loops the three statements share — inside it, g
// is one distinct inner loop per dimension of T, so every element of T stands for exactly one element of P and nothing
// else touches it between its load and its store — the triple is the single statement
//     U': P[f] = P[f] + <terms>
// and T, L and S disappear.  What the sequential program computes is unchanged: U' performs the same additions on the same
// values in the same order, only in place.
#include "lowering.h"

namespace tiramisu {
namespace lowering_detail {

#define FWD_FAIL do { if (getenv("TIRAMISU_B200_DEBUG_FORWARD")) std::cerr << "forward: no triple at line " << __LINE__ << "\n"; return false; } while (0)
namespace {
bool is_access(const expr &x) { return x.is_defined() && x.get_expr_type() == e_op && x.get_op_type() == o_access; }

void flatten(const expr &e, std::vector<expr> &terms)
{
    if (e.is_defined() && e.get_expr_type() == e_op && e.get_op_type() == o_add) { flatten(e.get_operand(0), terms); flatten(e.get_operand(1), terms); }
    else terms.push_back(e);
}

// integer literal folding of + - * % / (what is left after single-iteration loops were replaced by their value)
expr fold(const expr &e)
{
    if (!e.is_defined() || e.get_expr_type() != e_op || e.get_n_arg() != 2) return e;
    const op_t o = e.get_op_type();
    if (o != o_add && o != o_sub && o != o_mul && o != o_mod && o != o_div) return e;
    const expr a = fold(e.get_operand(0)), b = fold(e.get_operand(1));
    if (a.is_int_literal() && b.is_int_literal()) {
        const int64_t x = a.get_int_val(), y = b.get_int_val();
        if ((o == o_mod || o == o_div) && (y <= 0 || x < 0)) return expr(o, a, b);
        const int64_t r = o == o_add ? x + y : o == o_sub ? x - y : o == o_mul ? x * y : o == o_mod ? x % y : x / y;
        return expr((int32_t) r);
    }
    if (o == o_mul && ((a.is_int_literal() && a.get_int_val() == 0) || (b.is_int_literal() && b.get_int_val() == 0))) return expr((int32_t) 0);
    if (o == o_mul && a.is_int_literal() && a.get_int_val() == 1) return b;
    if (o == o_mul && b.is_int_literal() && b.get_int_val() == 1) return a;
    if (o == o_add && a.is_int_literal() && a.get_int_val() == 0) return b;
    if ((o == o_add || o == o_sub) && b.is_int_literal() && b.get_int_val() == 0) return a;
    return expr(o, a, b);
}
}  // namespace

// ord[i], ord[i+1], ord[i+2] = L, U, S as above?  With apply: rewrite U, take L and S out of the program.
bool Lowering::staging_triple(const std::vector<computation *> &ord, size_t i, bool apply)
{
    if (i + 2 >= ord.size()) FWD_FAIL;
    computation *L = ord[i], *U = ord[i + 1], *S = ord[i + 2];
    for (computation *c : {L, U, S})
        if (!c->has_access() || !c->get_expr().is_defined() || c->get_predicate().is_defined() || c->is_communicator()) FWD_FAIL;
    buffer *T = L->get_buffer(), *P = S->get_buffer();
    if (T != U->get_buffer() || T == P || T->get_argument_type() != a_temporary) FWD_FAIL;
    // resolve an access to (buffer, index tuple in the reader's iterator names)
    auto resolve = [&](const expr &x, buffer *&b, std::vector<expr> &idx) {
        if (!is_access(x)) FWD_FAIL;
        computation *oc = nullptr;
        b = buffer_of_access(x, &oc);
        if (!b || (int) oc->get_iteration_variables().size() != x.get_n_arg()) FWD_FAIL;
        std::vector<std::string> names;
        for (const var &v : oc->get_iteration_variables()) names.push_back(v.get_name());
        idx.clear();
        for (const expr &a : oc->access()) idx.push_back(a.substitute(names, x.get_operands()));
        return true;
    };
    auto same = [](const std::vector<expr> &a, const std::vector<expr> &b) {
        if (a.size() != b.size()) FWD_FAIL;
        for (size_t q = 0; q < a.size(); q++) if (!a[q].is_equal(b[q])) FWD_FAIL;
        return true;
    };
    // L: a pure copy P[f] -> T[g]
    buffer *bl = nullptr; std::vector<expr> fL;
    if (!resolve(L->get_expr(), bl, fL) || bl != P) FWD_FAIL;
    const std::vector<expr> &gL = L->access();
    // S: a pure copy T[g] -> P[f], the same g and f by name
    buffer *bs = nullptr; std::vector<expr> gS;
    if (!resolve(S->get_expr(), bs, gS) || bs != T || !same(gS, gL) || !same(S->access(), fL)) FWD_FAIL;
    // U: T[g] = T[g] + terms that do not touch T
    if (!same(U->access(), gL)) FWD_FAIL;
    std::vector<expr> terms;
    flatten(U->get_expr(), terms);
    int self = -1;
    for (size_t t = 0; t < terms.size(); t++) {
        buffer *b = nullptr; std::vector<expr> idx;
        if (resolve(terms[t], b, idx) && b == T) {
            if (self >= 0 || !same(idx, gL)) FWD_FAIL;
            self = (int) t;
            continue;
        }
        std::vector<expr> stack{terms[t]};
        while (!stack.empty()) {
            expr x = stack.back(); stack.pop_back();
            if (!x.is_defined() || x.get_expr_type() != e_op) continue;
            if (is_access(x)) { computation *oc = nullptr; if (buffer_of_access(x, &oc) == T) FWD_FAIL; }
            for (const expr &o : x.get_operands()) stack.push_back(o);
        }
    }
    if (self < 0 || terms.size() < 2) FWD_FAIL;
    // the iterators L and S name are iterators of U, over the same ranges
    auto range_of = [](computation *c, const std::string &n, expr &lo, expr &hi) {
        for (const var &v : c->get_iteration_variables()) if (v.get_name() == n) { lo = v.get_lower(); hi = v.get_upper(); return v.has_range(); }
        FWD_FAIL;
    };
    if (L->get_iteration_variables().size() != S->get_iteration_variables().size()) FWD_FAIL;
    for (const var &v : L->get_iteration_variables()) {
        expr lo, hi;
        for (computation *c : {U, S})
            if (!range_of(c, v.get_name(), lo, hi) || !lo.is_equal(v.get_lower()) || !hi.is_equal(v.get_upper())) FWD_FAIL;
    }
    // privacy: inside one iteration of the loops all three share, every dimension of T is one distinct inner loop of L
    // (resp. S), and L (S) has no other inner loop that runs more than once
    const int d = std::min(fuse_depth(L, U), fuse_depth(U, S));
    if (d < 1) FWD_FAIL;
    auto private_in = [&](computation *c) {
        Env env;
        for (const loop_dim &ld : c->dims()) {
            if (!ld.lo.is_int_literal() || !ld.hi.is_int_literal()) continue;
            env[ld.name] = {ld.lo.get_int_val(), ld.hi.get_int_val() - 1};
        }
        std::set<std::string> used;
        for (expr g : gL) {
            for (auto &df : c->defs()) g = g.substitute({df.first}, {df.second});
            for (const loop_dim &ld : c->dims())                       // a loop that runs once is its lower bound
                if (ld.lo.is_int_literal() && ld.hi.is_int_literal() && ld.hi.get_int_val() == ld.lo.get_int_val() + 1)
                    g = g.substitute({ld.name}, {ld.lo});
            g = fold(simp(fold(g), env));
            if (g.is_int_literal()) continue;
            if (g.get_expr_type() != e_var || c->find_dim(g.get_name()) < d || !used.insert(g.get_name()).second) {
                if (getenv("TIRAMISU_B200_DEBUG_FORWARD")) { std::cerr << "forward: '" << c->get_name() << "' d=" << d << " index "; g.dump(false); std::cerr << "\n"; }
                FWD_FAIL;
            }
        }
        for (int q = d; q < (int) c->dims().size(); q++) {
            const loop_dim &ld = c->dims()[q];
            const bool once = ld.lo.is_int_literal() && ld.hi.is_int_literal() && ld.hi.get_int_val() - ld.lo.get_int_val() <= 1;
            if (!once && !used.count(ld.name)) FWD_FAIL;
        }
        return true;
    };
    if (!private_in(L) || !private_in(S)) FWD_FAIL;
    if (!apply) return true;

    std::cerr << "tiramisu_b200 lowering: '" << L->get_name() << "' / '" << U->get_name() << "' / '" << S->get_name() << "' load, accumulate and store '"
              << P->get_name() << "' through the private staging buffer '" << T->get_name() << "' (CPU register blocking): lowered as one in-place accumulation\n";
    expr sum;
    for (size_t t = 0; t < terms.size(); t++) {
        const expr &x = (int) t == self ? L->get_expr() : terms[t];
        sum = sum.is_defined() ? sum + x : x;
    }
    U->replace_expression(sum);
    U->set_access_exprs(P, fL);
    const int l_in = fuse_depth(L->pred(), L), l_out = fuse_depth(L, U), s_in = fuse_depth(U, S);
    const int s_out = S->succ() ? fuse_depth(S, S->succ()) : 0;
    L->remove_from_order(std::min(l_in, l_out) - 1);
    S->remove_from_order(std::min(s_in, s_out) - 1);
    return true;
}

// ------------------------------------------------------------------------------------------------ dead temporaries
// A statement that fills a temporary nothing reads (vgg_block_generator_tiramisu.cpp:109-113 `prefetch_weights`: a loop over
// the weights whose only purpose is to pull them into the CPU's cache) is not computed.
bool Lowering::drop_dead_temporaries()
{
    bool any = false;
    for (bool changed = true; changed;) {
        changed = false;
        std::vector<computation *> ord = order();
        for (computation *c : ord) {
            if (!c->has_access() || c->is_communicator() || c->get_buffer()->get_argument_type() != a_temporary) continue;
            buffer *T = c->get_buffer();
            bool read = false;
            for (computation *o : f_->get_computations()) {
                std::vector<expr> stack{o->get_expr(), o->get_predicate()};
                while (!stack.empty() && !read) {
                    expr x = stack.back(); stack.pop_back();
                    if (!x.is_defined() || x.get_expr_type() != e_op) continue;
                    if (x.get_op_type() == o_access) {
                        computation *oc = f_->find_computation(x.get_name());
                        if (oc && !oc->is_constant() && oc->has_access() && oc->get_buffer() == T) read = true;
                    }
                    if (x.get_op_type() == o_buffer && x.get_name() == T->get_name()) read = true;
                    for (const expr &a : x.get_operands()) stack.push_back(a);
                }
                if (read) break;
            }
            if (read) continue;
            std::cerr << "tiramisu_b200 lowering: '" << c->get_name() << "' fills the temporary '" << T->get_name() << "' that nothing reads; not computed\n";
            const int in = fuse_depth(c->pred(), c), out = c->succ() ? fuse_depth(c, c->succ()) : 0;
            c->remove_from_order(std::min(in, out) - 1);
            changed = any = true;
            break;
        }
    }
    return any;
}

// ------------------------------------------------------------------------------------------------ array expansion
// The other register-blocking idiom (vgg_block_generator_tiramisu.cpp:38-44, :230-246):
//     I:  T[g] = <init>             U:  T[g] = T[g] + <terms>            R:  out[f] = h(T[g])     (ReLU, max-pool, ...)
// all fused inside the loops that walk the output, T a few vectors long and private to one iteration of those loops.  The
// value R reads is not stored back where it came from, so there is nothing to accumulate in place; instead T gets its
// full extent back — one element per point of I's iteration box, T'[n][fin2_b][y][x][ffin2] — which removes the only
// dependence between iterations of the shared loops (the re-use of T) and lets them be distributed and mapped.  Readers
// reach T through accesses to I / U, which resolve through the statements' new mapping; nothing else changes.
bool Lowering::expand_private_temporaries()
{
    bool any = false;
    std::set<computation *> done;
    for (;;) {
        std::vector<computation *> ord = order();
        bool changed = false;
        for (size_t i = 0; i < ord.size() && !changed; i++) {
            computation *I = ord[i];
            if (done.count(I) || !I->has_access() || !I->get_expr().is_defined() || I->get_predicate().is_defined() || I->is_communicator()) continue;
            buffer *T = I->get_buffer();
            if (T->get_argument_type() != a_temporary) continue;
            auto reads_T = [&](const expr &e, computation *via_only, bool &other) {
                // does e read T?  `other`: through something that is not `via_only` / I (a view, another statement)
                bool r = false;
                std::vector<expr> stack{e};
                while (!stack.empty()) {
                    expr x = stack.back(); stack.pop_back();
                    if (!x.is_defined() || x.get_expr_type() != e_op) continue;
                    if (x.get_op_type() == o_access) {
                        computation *oc = f_->find_computation(x.get_name());
                        if (oc && !oc->is_constant() && oc->has_access() && oc->get_buffer() == T) { r = true; if (oc != I && oc != via_only) other = true; }
                    }
                    if (x.get_op_type() == o_buffer && x.get_name() == T->get_name()) { r = true; other = true; }
                    for (const expr &a : x.get_operands()) stack.push_back(a);
                }
                return r;
            };
            bool other = false;
            if (reads_T(I->get_expr(), nullptr, other)) continue;                 // I must be a pure initialisation
            // the group: I, an optional accumulation U that follows it, and the readers, all consecutive-fused
            std::vector<std::string> inames;
            for (const var &v : I->get_iteration_variables()) { if (!v.has_range()) { inames.clear(); break; } inames.push_back(v.get_name()); }
            if (inames.empty()) continue;
            computation *U = nullptr;
            size_t last = i;
            bool ok = true;
            for (size_t q = i + 1; q < ord.size() && ok; q++) {
                computation *c = ord[q];
                const bool writes = c->has_access() && c->get_buffer() == T;
                bool oth = false;
                const bool reads = reads_T(c->get_expr(), U, oth) || reads_T(c->get_predicate(), U, oth);
                if (writes) {
                    // a second (I, U) group on the same staging buffer starts here (the conclude column): this group ends
                    bool o2 = false;
                    if (!reads_T(c->get_expr(), nullptr, o2)) break;
                    if (U || oth) { ok = false; break; }
                    // U: same store index as I by name, reads T only at that element (through I)
                    if (c->access().size() != I->access().size()) { ok = false; break; }
                    for (size_t t = 0; t < c->access().size(); t++) ok &= c->access()[t].is_equal(I->access()[t]);
                    for (const std::string &n : inames) {
                        bool has = false;
                        for (const var &v : c->get_iteration_variables()) has |= v.get_name() == n;
                        ok &= has;
                    }
                    U = c;
                    last = q;
                } else if (reads) {
                    if (oth) { ok = false; break; }
                    last = q;
                }
            }
            if (!ok || last == i) continue;
            // every access to I / U (outside I and U themselves) names I's iterators by their own name (or is a literal at a
            // reduction position): the reader works on the element of its own iteration
            auto identity_args = [&](computation *reader) {
                std::vector<expr> stack{reader->get_expr(), reader->get_predicate()};
                while (!stack.empty()) {
                    expr x = stack.back(); stack.pop_back();
                    if (!x.is_defined() || x.get_expr_type() != e_op) continue;
                    for (const expr &a : x.get_operands()) stack.push_back(a);
                    if (x.get_op_type() != o_access) continue;
                    computation *oc = f_->find_computation(x.get_name());
                    if (oc != I && oc != U) continue;
                    const auto &its = oc->get_iteration_variables();
                    if ((int) its.size() != x.get_n_arg()) return false;
                    for (int a = 0; a < x.get_n_arg(); a++) {
                        const bool named = std::find(inames.begin(), inames.end(), its[(size_t) a].get_name()) != inames.end();
                        const expr &arg = x.get_operand(a);
                        if (named && !(arg.get_expr_type() == e_var && arg.get_name() == its[(size_t) a].get_name())) return false;
                    }
                }
                return true;
            };
            int d = INT32_MAX;
            for (size_t q = i + 1; q <= last; q++) d = std::min(d, fuse_depth(ord[q - 1], ord[q]));
            for (size_t q = i + 1; q <= last && ok; q++) ok = identity_args(ord[q]);
            if (!ok || d < 1 || d == INT32_MAX) continue;
            // privacy of T inside one iteration of the d shared loops (as in staging_triple)
            {
                Env env;
                for (const loop_dim &ld : I->dims())
                    if (ld.lo.is_int_literal() && ld.hi.is_int_literal()) env[ld.name] = {ld.lo.get_int_val(), ld.hi.get_int_val() - 1};
                std::set<std::string> used;
                for (expr g : I->access()) {
                    for (auto &df : I->defs()) g = g.substitute({df.first}, {df.second});
                    for (const loop_dim &ld : I->dims())
                        if (ld.lo.is_int_literal() && ld.hi.is_int_literal() && ld.hi.get_int_val() == ld.lo.get_int_val() + 1) g = g.substitute({ld.name}, {ld.lo});
                    g = fold(simp(fold(g), env));
                    if (g.is_int_literal()) continue;
                    if (g.get_expr_type() != e_var || I->find_dim(g.get_name()) < d || !used.insert(g.get_name()).second) { ok = false; break; }
                }
                for (int q = d; q < (int) I->dims().size() && ok; q++) {
                    const loop_dim &ld = I->dims()[(size_t) q];
                    const bool once = ld.lo.is_int_literal() && ld.hi.is_int_literal() && ld.hi.get_int_val() - ld.lo.get_int_val() <= 1;
                    if (!once && !used.count(ld.name)) ok = false;
                }
            }
            if (!ok) continue;
            std::vector<expr> sizes, idx;
            for (const var &v : I->get_iteration_variables()) { sizes.push_back(v.get_upper()); idx.push_back(expr::make_var(global::get_loop_iterator_data_type(), v.get_name())); }
            buffer *X = f_->own(new buffer(T->get_name() + "_of_" + I->get_name(), sizes, T->get_elements_type(), a_temporary, f_));
            std::cerr << "tiramisu_b200 lowering: '" << T->get_name() << "' is a staging buffer private to one iteration of the loops '" << I->get_name()
                      << "' shares with its " << (U ? "accumulation and its " : "") << "readers (CPU register blocking): given its full extent back as '" << X->get_name() << "'\n";
            I->set_access_exprs(X, idx);
            if (U) U->set_access_exprs(X, idx);
            done.insert(I);
            changed = any = true;
        }
        if (!changed) break;
    }
    return any;
}

bool Lowering::forward_private_temporaries()
{
    bool any = false;
    for (;;) {
        std::vector<computation *> ord = order();
        // every statement that touches a staging buffer must belong to such a triple (the buffer is not live across them)
        std::map<buffer *, std::vector<size_t>> triples;
        for (size_t i = 0; i + 2 < ord.size(); i++)
            if (staging_triple(ord, i, false)) triples[ord[i]->get_buffer()].push_back(i);
        bool changed = false;
        for (auto &kv : triples) {
            buffer *T = kv.first;
            std::set<size_t> covered;
            for (size_t i : kv.second) { covered.insert(i); covered.insert(i + 1); covered.insert(i + 2); }
            bool ok = true;
            for (size_t q = 0; q < ord.size() && ok; q++) {
                if (covered.count(q)) continue;
                if (ord[q]->has_access() && ord[q]->get_buffer() == T) ok = false;
                std::vector<expr> stack{ord[q]->get_expr(), ord[q]->get_predicate()};
                while (!stack.empty() && ok) {
                    expr x = stack.back(); stack.pop_back();
                    if (!x.is_defined() || x.get_expr_type() != e_op) continue;
                    if (x.get_op_type() == o_access) {
                        computation *oc = f_->find_computation(x.get_name());
                        if (oc && !oc->is_constant() && oc->has_access() && oc->get_buffer() == T) ok = false;
                    }
                    for (const expr &o : x.get_operands()) stack.push_back(o);
                }
            }
            if (!ok) continue;
            // one at a time: the order changes under our feet
            staging_triple(ord, kv.second[0], true);
            changed = any = true;
            break;
        }
        if (!changed) break;
    }
    return any;
}

}  // namespace lowering_detail
}  // namespace tiramisu
