// lowering_patterns.cpp — pattern lowerings onto the runtime's hand-written kernels (see lowering.h).
#include "lowering.h"

namespace tiramisu {
namespace lowering_detail {

// Dense fp32 contraction over row-major global buffers:
//     [ X(I,J) = 0 ]   X(.., R ..) = X(..) + P(I,R) * Q(R,J)
// (one statement, or a zero-init fused in front of it as in chain_mm/gpu/generator.cpp:38-43; the
// reduction iterator R may sit anywhere in the iterator list, operands may be stored transposed).
// The kernel is not emitted: it becomes one call of the runtime's tcgen05 GEMM, beta = 1 (accumulate
// into what the buffer holds) or beta = 0 (the zero-init is part of the pattern).
std::vector<Node *> statements_of(Node *root)
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
    if (!z->get_expr().is_defined() || z->get_predicate().is_defined() || zi.size() < 2) return false;
    // further iterators must be single points the store does not name (`{acc[i,j,0]: ...}`, sgemm_generator.cpp:90)
    for (size_t q = 2; q < zi.size(); q++) {
        if (!(zi[q].get_lower().is_int_literal() && zi[q].get_lower().get_int_val() == 0 && zi[q].get_upper().is_int_literal() && zi[q].get_upper().get_int_val() == 1)) return false;
        for (const expr &a : z->access())
            if (!a.substitute({zi[q].get_name()}, {expr((int32_t) 0)}).is_equal(a)) return false;
    }
    if (z->get_buffer() != C || z->access().size() != 2) return false;
    if (!(z->get_expr().is_literal() && z->get_expr().get_double_val() == 0.0)) return false;
    if (!is_var(z->access()[0], zi[0].get_name()) || !is_var(z->access()[1], zi[1].get_name())) return false;
    if (!zi[0].get_upper().is_equal(M) || !zi[1].get_upper().is_equal(N)) return false;
    for (size_t q = 0; q < 2; q++) if (!(zi[q].get_lower().is_int_literal() && zi[q].get_lower().get_int_val() == 0)) return false;
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

// ------------------------------------------------------------------------------------------------ uint8 3x3 blur pair
// benchmarks/halide/blurxy_tiramisu.cpp:61-77 (row a7) as two consecutive kernels:
//     bx(c, y, x) = ((in(c, y, x) + in(c, y, x+1)) + in(c, y, x+2)) / 3      over [C] x [H-6] x [W-8], a temporary
//     by(c, y, x) = ((bx(c, y, x) + bx(c, y+1, x)) + bx(c, y+2, x)) / 3      over [C] x [H-8] x [W-8]
// in uint8 (sums wrap mod 256, as the reference's generated code does) -> the runtime's blur kernel, one pass, no temporary.
bool Lowering::try_blurxy_pattern(Node *kx, Node *ky, const std::vector<computation *> &ord, std::ostringstream &host, int ind)
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
    Node *sx = single(kx), *sy = single(ky);
    if (!sx || !sy) return false;
    computation *bx = sx->comp, *by = sy->comp;
    for (computation *c : {bx, by})
        if (c->get_data_type() != p_uint8 || c->get_predicate().is_defined() || !c->has_access() || c->get_iteration_variables().size() != 3 || c->access().size() != 3) return false;
    buffer *T = bx->get_buffer(), *out = by->get_buffer();
    if (T == out || T->get_argument_type() != a_temporary || T->get_n_dims() != 3 || out->get_n_dims() != 3 || out->get_location() != mem_loc::gpu_global) return false;
    auto is_var = [](const expr &x, const std::string &n) { return x.get_expr_type() == e_var && x.get_name() == n; };
    auto is_lit = [](const expr &x, int64_t v) { return x.is_int_literal() && x.get_int_val() == v; };
    auto offset = [&](const expr &x, const std::string &v, int &off) {
        if (is_var(x, v)) { off = 0; return true; }
        if (x.get_expr_type() != e_op || x.get_n_arg() != 2 || x.get_op_type() != o_add) return false;
        const expr &a = x.get_operand(0), &b = x.get_operand(1);
        if (is_var(a, v) && b.is_int_literal()) { off = (int) b.get_int_val(); return true; }
        if (is_var(b, v) && a.is_int_literal()) { off = (int) a.get_int_val(); return true; }
        return false;
    };
    // ((t0 + t1) + t2) / 3 with the three taps of `src` along dimension `along` (1 = rows, 2 = columns), in that order
    auto three_taps = [&](computation *c, int along, buffer *&src) {
        const auto &it = c->get_iteration_variables();
        for (int d = 0; d < 3; d++) if (!is_var(c->access()[d], it[d].get_name()) || !is_lit(it[d].get_lower(), 0)) return false;
        const expr &e = c->get_expr();
        if (!(e.get_expr_type() == e_op && e.get_op_type() == o_div && is_lit(e.get_operand(1), 3) && e.get_operand(1).get_data_type() == p_uint8)) return false;
        const expr &s2 = e.get_operand(0);
        if (!(s2.get_expr_type() == e_op && s2.get_op_type() == o_add)) return false;
        const expr &s1 = s2.get_operand(0);
        if (!(s1.get_expr_type() == e_op && s1.get_op_type() == o_add)) return false;
        const expr taps[3] = {s1.get_operand(0), s1.get_operand(1), s2.get_operand(1)};
        src = nullptr;
        for (int k = 0; k < 3; k++) {
            const expr &x = taps[k];
            if (!(x.get_expr_type() == e_op && x.get_op_type() == o_access) || x.get_data_type() != p_uint8) return false;
            computation *oc = nullptr;
            buffer *b = buffer_of_access(x, &oc);
            if (!b || b->get_n_dims() != 3 || oc->access().size() != 3 || (int) oc->get_iteration_variables().size() != x.get_n_arg() || (src && src != b)) return false;
            src = b;
            std::vector<std::string> names;
            for (const var &v : oc->get_iteration_variables()) names.push_back(v.get_name());
            for (int d = 0; d < 3; d++) {
                int off = -1;
                if (!offset(oc->access()[d].substitute(names, x.get_operands()), it[d].get_name(), off) || off != (d == along ? k : 0)) return false;
            }
        }
        return true;
    };
    buffer *in = nullptr, *mid = nullptr;
    if (!three_taps(bx, 2, in) || !three_taps(by, 1, mid) || mid != T || in == T || in == out || in->get_location() != mem_loc::gpu_global) return false;
    // shapes: in [C][H][W], bx over [C][H-6][W-8] (the temporary at least that large), by = out = [C][H-8][W-8]
    const expr &Cn = in->get_dim_sizes()[0], &H = in->get_dim_sizes()[1], &W = in->get_dim_sizes()[2];
    auto same_sz = [&](const expr &a, const expr &b) {
        if (a.is_equal(b)) return true;
        Env none;
        return simp(a - b, none).is_int_literal() && simp(a - b, none).get_int_val() == 0;
    };
    auto lit_or_equal = [&](const expr &have, const expr &base, int64_t minus) {
        if (have.is_int_literal() && base.is_int_literal()) return have.get_int_val() == base.get_int_val() - minus;
        return same_sz(have, base - (int32_t) minus);
    };
    const auto &ix = bx->get_iteration_variables(), &iy = by->get_iteration_variables();
    if (!lit_or_equal(ix[0].get_upper(), Cn, 0) || !lit_or_equal(ix[1].get_upper(), H, 6) || !lit_or_equal(ix[2].get_upper(), W, 8)) return false;
    if (!lit_or_equal(iy[0].get_upper(), Cn, 0) || !lit_or_equal(iy[1].get_upper(), H, 8) || !lit_or_equal(iy[2].get_upper(), W, 8)) return false;
    if (!lit_or_equal(out->get_dim_sizes()[0], Cn, 0) || !lit_or_equal(out->get_dim_sizes()[1], H, 8) || !lit_or_equal(out->get_dim_sizes()[2], W, 8)) return false;
    // nobody else reads the temporary
    for (computation *c : ord) {
        if (c == bx || c == by) continue;
        if (c->has_access() && c->get_buffer() == T) return false;
        std::vector<expr> stack{c->get_expr(), c->get_predicate()};
        while (!stack.empty()) {
            expr x = stack.back(); stack.pop_back();
            if (!x.is_defined() || x.get_expr_type() != e_op) continue;
            if (x.get_op_type() == o_access) { computation *oc = f_->find_computation(x.get_name()); if (oc && !oc->is_constant() && oc->has_access() && oc->get_buffer() == T) return false; }
            for (const expr &o : x.get_operands()) stack.push_back(o);
        }
    }
    host << pad(ind) << "// pattern lowering: '" << bx->get_name() << "' (3 taps along x) + '" << by->get_name()
         << "' (3 taps along y), uint8 -> the runtime's 3x3 blur kernel (one pass, the temporary is not written)\n";
    host << pad(ind) << "if (tiramisu_b200_blurxy_u8(b_" << san(in->get_name()) << ", b_" << san(out->get_name()) << ", (int32_t) (" << print(Cn, false)
         << "), (int32_t) (" << print(H, false) << "), (int32_t) (" << print(W, false) << "))) return 1;\n";
    return true;
}

// ------------------------------------------------------------------------------------------------ staged contraction
// A hand-scheduled GPU GEMM (benchmarks/linear_algebra/blas/level3/sgemm/gpu/generator{1,2,3}.cpp): ONE kernel whose
// Layer-I arithmetic is still
//     acc(i,j) = 0;   acc(i,j,k) = acc(i,j,0) + P(i,k) * Q(k,j);   C(i,j) = acc(i,j,0) * a + C(i,j) * b
// and whose every other statement is scheduling plumbing — allocate, sync, and pure copies (the right-hand side is one
// access, no arithmetic) that stage P and Q from global memory through shared memory into registers.  Layer I says what
// P(i,k) IS — element (i,k) of the input P — whatever Layer III does with its storage; the staging is followed at the
// BUFFER level (register <- shared <- global: each on-chip buffer is written only by copies that read one other buffer)
// back to the global matrices of shape {M,K} and {K,N}, and the whole kernel becomes one tcgen05 GEMM call.  Lowered as
// written, the same kernel runs at ~20 TFLOP/s on the SIMT pipes (r01: 2.5-3.2 ms at 3072^3 against 0.1 ms).
bool Lowering::try_staged_contraction(Node *kernel, std::ostringstream &host, int ind)
{
    if (!pattern_) return false;
    std::vector<Node *> stmts = statements_of(kernel);
    if (stmts.size() < 4) return false;
    auto is_access = [](const expr &x) { return x.is_defined() && x.get_expr_type() == e_op && x.get_op_type() == o_access; };
    auto is_var = [](const expr &x, const std::string &n) { return x.get_expr_type() == e_var && x.get_name() == n; };
    computation *init = nullptr, *acc = nullptr, *fin = nullptr;
    std::map<buffer *, std::set<buffer *>> copied_from;                  // on-chip / global buffer <- buffers its copies read
    for (Node *n : stmts) {
        computation *c = n->comp;
        const expr &e = c->get_expr();
        if (e.get_expr_type() == e_sync) continue;
        if (e.get_expr_type() == e_op && (e.get_op_type() == o_allocate || e.get_op_type() == o_free)) continue;
        if (c->get_predicate().is_defined()) return false;
        if (is_access(e)) {                                              // a pure copy
            computation *src = nullptr;
            buffer *sb = buffer_of_access(e, &src);
            if (!sb) return false;
            copied_from[c->get_buffer()].insert(sb);
            continue;
        }
        if (e.is_literal() && e.get_double_val() == 0.0) { if (init) return false; init = c; continue; }
        if (e.get_expr_type() == e_op && e.get_op_type() == o_add) {
            bool self = false;
            for (int side = 0; side < 2; side++)
                self |= is_access(e.get_operand(side)) && e.get_operand(side).get_name() == c->get_name();
            if (self && c->get_iteration_variables().size() == 3) { if (acc) return false; acc = c; continue; }
            if (c->get_iteration_variables().size() == 2) { if (fin) return false; fin = c; continue; }
        }
        return false;                                                    // arithmetic this pattern does not know
    }
    if (!init || !acc || !fin || acc->get_data_type() != p_float32) return false;
    // acc(i,j,k) = acc(i,j,*) + P(i,k) * Q(k,j)
    const auto &it = acc->get_iteration_variables();
    for (const var &v : it) if (!(v.has_range() && v.get_lower().is_int_literal() && v.get_lower().get_int_val() == 0)) return false;
    const std::string I = it[0].get_name(), J = it[1].get_name(), Kn = it[2].get_name();
    const expr &ae = acc->get_expr();
    const int self_side = (is_access(ae.get_operand(0)) && ae.get_operand(0).get_name() == acc->get_name()) ? 0 : 1;
    const expr &selfx = ae.get_operand(self_side), &prod = ae.get_operand(1 - self_side);
    if (selfx.get_n_arg() != 3 || !is_var(selfx.get_operand(0), I) || !is_var(selfx.get_operand(1), J)) return false;
    if (!(prod.get_expr_type() == e_op && prod.get_op_type() == o_mul && is_access(prod.get_operand(0)) && is_access(prod.get_operand(1)))) return false;
    expr P, Q;
    for (int side = 0; side < 2; side++) {
        const expr &x = prod.get_operand(side);
        if (x.get_n_arg() != 2) return false;
        if (is_var(x.get_operand(0), I) && is_var(x.get_operand(1), Kn)) P = x;
        else if (is_var(x.get_operand(0), Kn) && is_var(x.get_operand(1), J)) Q = x;
    }
    if (!P.is_defined() || !Q.is_defined()) return false;
    if (init->get_buffer() != acc->get_buffer() || acc->get_buffer()->get_location() == mem_loc::gpu_global || acc->get_buffer()->get_location() == mem_loc::host)
        return false;
    const expr M = it[0].get_upper(), N = it[1].get_upper(), K = it[2].get_upper();
    // C(i,j) = acc(i,j,*) * a + C(i,j) * b, identity mapping into a global {M,N} buffer
    const auto &fi = fin->get_iteration_variables();
    buffer *C = fin->get_buffer();
    if (C->get_location() != mem_loc::gpu_global || C->get_n_dims() != 2 || fin->access().size() != 2 || !is_var(fin->access()[0], fi[0].get_name()) ||
        !is_var(fin->access()[1], fi[1].get_name()) || !fi[0].get_upper().is_equal(M) || !fi[1].get_upper().is_equal(N) ||
        !C->get_dim_sizes()[0].is_equal(M) || !C->get_dim_sizes()[1].is_equal(N))
        return false;
    auto scalar = [&](const expr &x) { return x.get_data_type() == p_float32 && (x.is_literal() || (x.get_expr_type() == e_var && constants_.count(x.get_name()))); };
    expr alpha, beta;
    for (int side = 0; side < 2; side++) {
        const expr &t = fin->get_expr().get_operand(side);
        if (!(t.get_expr_type() == e_op && t.get_op_type() == o_mul)) return false;
        for (int f = 0; f < 2; f++) {
            const expr &sc = t.get_operand(f), &el = t.get_operand(1 - f);
            if (!scalar(sc) || !is_access(el)) continue;
            if (el.get_name() == acc->get_name() && el.get_n_arg() == 3 && is_var(el.get_operand(0), fi[0].get_name()) && is_var(el.get_operand(1), fi[1].get_name())) alpha = sc;
            else if (el.get_name() == fin->get_name() && el.get_n_arg() == 2 && is_var(el.get_operand(0), fi[0].get_name()) && is_var(el.get_operand(1), fi[1].get_name())) beta = sc;
        }
    }
    if (!alpha.is_defined() || !beta.is_defined()) return false;
    // follow the staging of P and Q back to global memory
    auto global_source = [&](const expr &x, const expr &d0, const expr &d1) -> buffer * {
        // every copy path from the operand's storage must end in the same global buffer (a double-buffered staging writes
        // one shared tile from two places: generator2.cpp stages global -> temporary registers -> shared as well)
        std::set<buffer *> roots, seen;
        std::vector<buffer *> stack{buffer_of_access(x, nullptr)};
        while (!stack.empty()) {
            buffer *b = stack.back(); stack.pop_back();
            if (!b || !seen.insert(b).second) continue;
            if (b->get_location() == mem_loc::gpu_global) { roots.insert(b); continue; }
            auto w = copied_from.find(b);
            if (w == copied_from.end()) return nullptr;                  // an on-chip buffer nobody fills with a copy
            for (buffer *s2 : w->second) stack.push_back(s2);
        }
        if (roots.size() != 1) return nullptr;
        buffer *b = *roots.begin();
        return (b->get_n_dims() == 2 && b->get_dim_sizes()[0].is_equal(d0) && b->get_dim_sizes()[1].is_equal(d1) && !copied_from.count(b)) ? b : nullptr;
    };
    buffer *A = global_source(P, M, K), *B = global_source(Q, K, N);
    if (!A || !B || A == B || A == C || B == C || A->get_elements_type() != p_float32 || B->get_elements_type() != p_float32) return false;
    Contraction m;
    m.C = C; m.A = A; m.B = B; m.M = M; m.N = N; m.K = K;
    host << pad(ind) << "// pattern lowering: kernel of '" << acc->get_name() << "' = hand-staged contraction (Layer I: acc += " << P.get_name() << "(i,k) * " << Q.get_name()
         << "(k,j); staging followed back to '" << A->get_name() << "', '" << B->get_name() << "') -> one tcgen05 GEMM call of the runtime\n";
    emit_sgemm_call(host, ind, m, C, print(alpha, false), print(beta, false));
    return true;
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

}  // namespace lowering_detail
}  // namespace tiramisu
