// isl_box.cpp — the old-style, string-based declarations of the reference API, without ISL.
//
// The reference's Halide-derived generators (benchmarks/halide/heat2d_tiramisu.cpp:44-69 and friends)
// declare iteration domains as ISL set strings and data mappings as ISL map strings
// (reference include/tiramisu/core.h:3236 and :4180).  Every domain the hot-path generators write is
// a box — a conjunction of `lo <= it <= hi` chains over affine bounds in function constants — so a
// small recursive-descent reader is enough: it produces the same (name, [lo, hi)) list the modern
// `var("i", lo, hi)` API produces.  Anything that is not a box (disjunctions, constraints coupling two
// iterators) stops the generator with a message instead of being mis-compiled.
#include <tiramisu/core.h>

#include <algorithm>
#include <cctype>
#include <cstdint>
#include <cstdlib>
#include <iostream>

namespace tiramisu {
namespace {

[[noreturn]] void die(const std::string &msg)
{
    std::cerr << "tiramisu_b200 front-end: " << msg << std::endl;
    std::exit(1);
}

struct Tok { enum K { id, num, sym, end } k = end; std::string s; int64_t v = 0; };

class Reader {
public:
    explicit Reader(const std::string &text) : text_(text) { lex(); }

    const Tok &peek(int ahead = 0) const { return toks_[std::min(pos_ + (size_t) ahead, toks_.size() - 1)]; }
    Tok next() { Tok t = peek(); if (pos_ < toks_.size() - 1) pos_++; return t; }
    bool accept(const std::string &s) { if (peek().k != Tok::end && peek().s == s) { next(); return true; } return false; }
    void expect(const std::string &s) { if (!accept(s)) fail("expected '" + s + "'"); }
    size_t mark() const { return pos_; }
    void reset(size_t m) { pos_ = m; }
    [[noreturn]] void fail(const std::string &why) const { die(why + " near '" + peek().s + "' in \"" + text_ + "\""); }

    // affine-ish integer expression
    expr parse_expr()
    {
        expr e = parse_term();
        for (;;) {
            if (accept("+")) e = e + parse_term();
            else if (accept("-")) e = e - parse_term();
            else return e;
        }
    }

    static bool is_cmp(const std::string &s) { return s == "<=" || s == "<" || s == ">=" || s == ">" || s == "="; }

private:
    std::string text_;
    std::vector<Tok> toks_;
    size_t pos_ = 0;

    void lex()
    {
        size_t i = 0, n = text_.size();
        while (i < n) {
            char c = text_[i];
            if (isspace((unsigned char) c)) { i++; continue; }
            Tok t;
            if (isalpha((unsigned char) c) || c == '_') {
                size_t j = i;
                while (j < n && (isalnum((unsigned char) text_[j]) || text_[j] == '_' || text_[j] == '\'')) j++;
                t.k = Tok::id; t.s = text_.substr(i, j - i); i = j;
            } else if (isdigit((unsigned char) c)) {
                size_t j = i;
                while (j < n && isdigit((unsigned char) text_[j])) j++;
                t.k = Tok::num; t.s = text_.substr(i, j - i); t.v = std::strtoll(t.s.c_str(), nullptr, 10); i = j;
            } else {
                t.k = Tok::sym;
                t.s = std::string(1, c);
                size_t len = 1;
                for (const char *w : {"->", "<=", ">=", "==", "!="})
                    if (text_.compare(i, 2, w) == 0) { t.s = w; len = 2; }
                if (t.s == "==") t.s = "=";
                i += len;
            }
            toks_.push_back(t);
        }
        toks_.push_back(Tok());
    }

    expr parse_term()
    {
        expr e = parse_unary();
        for (;;) {
            if (accept("*")) e = e * parse_unary();
            else if (accept("/")) e = e / parse_unary();
            else if (accept("%") || (peek().k == Tok::id && peek().s == "mod" && (next(), true))) e = e % parse_unary();
            else return e;
        }
    }
    expr parse_unary()
    {
        if (accept("-")) return -parse_unary();
        if (accept("+")) return parse_unary();
        return parse_primary();
    }
    expr parse_primary()
    {
        if (peek().k == Tok::num) return expr((int32_t) next().v);
        if (accept("(")) { expr e = parse_expr(); expect(")"); return e; }
        if (peek().k == Tok::id) {
            std::string name = next().s;
            if (name == "floor" || name == "floord") { expect("("); expr e = parse_expr(); expect(")"); return e; }   // floor(a / b): operands are non-negative here
            if ((name == "min" || name == "max") && peek().s == "(") {                                               // clamped accesses
                expect("("); expr a = parse_expr(); expect(","); expr b = parse_expr(); expect(")");
                return expr(name == "min" ? o_min : o_max, a, b);
            }
            return expr::make_var(global::get_loop_iterator_data_type(), name);
        }
        fail("expected an expression");
    }
};

struct Cmp { expr l, r; std::string op; expr r2; };       // op "in2": l is one of the two points r, r2

void parse_conjunction(Reader &rd, std::vector<Cmp> &out)
{
    for (;;) {
        // a chain `e1 op e2 [op e3]`, or a parenthesised conjunction
        size_t m = rd.mark();
        bool chain = false;
        if (!(rd.peek().k == Tok::sym && rd.peek().s == "(")) chain = true;
        else {
            // '(' may open an expression operand: look for a comparison right after the matching ')'
            int depth = 0, k = 0;
            for (;; k++) {
                const Tok &t = rd.peek(k);
                if (t.k == Tok::end) break;
                if (t.s == "(") depth++;
                if (t.s == ")" && --depth == 0) break;
            }
            chain = Reader::is_cmp(rd.peek(k + 1).s);
        }
        if (chain) {
            expr prev = rd.parse_expr();
            if (!Reader::is_cmp(rd.peek().s)) rd.fail("expected a comparison");
            while (Reader::is_cmp(rd.peek().s)) {
                std::string op = rd.next().s;
                expr nxt = rd.parse_expr();
                out.push_back({prev, nxt, op});
                prev = nxt;
            }
        } else {
            rd.reset(m);
            rd.expect("(");
            std::vector<Cmp> inner;
            parse_conjunction(rd, inner);
            if (rd.peek().k == Tok::id && rd.peek().s == "or") {
                // the one disjunction that is still a product of intervals and points: `(i = a or i = b)`, the two
                // border lines of a grid (benchmarks/halide/heat2dgpu_tiramisu.cpp:51-52) -> Cmp with op "in2"
                rd.next();
                std::vector<Cmp> other;
                parse_conjunction(rd, other);
                const bool two_points = inner.size() == 1 && other.size() == 1 && inner[0].op == "=" && other[0].op == "=" &&
                                        inner[0].l.get_expr_type() == e_var && other[0].l.get_expr_type() == e_var &&
                                        inner[0].l.get_name() == other[0].l.get_name();
                if (!two_points || (rd.peek().k == Tok::id && rd.peek().s == "or"))
                    rd.fail("disjunctions are not boxes: only `(i = a or i = b)` is supported besides rectangular iteration domains");
                out.push_back({inner[0].l, inner[0].r, "in2", other[0].r});
            } else {
                out.insert(out.end(), inner.begin(), inner.end());
            }
            rd.expect(")");
        }
        if (rd.peek().k == Tok::id && rd.peek().s == "and") { rd.next(); continue; }
        if (rd.peek().k == Tok::id && rd.peek().s == "or") return;          // the caller decides (two-point form or error)
        return;
    }
}

// `fixed` (optional) receives, per position, the literal written in place of an iterator name
// (`{S[i,j,0]}`: a dimension pinned to one value), or INT64_MIN for a named iterator.
std::vector<std::string> parse_name_list(Reader &rd, std::vector<int64_t> *fixed = nullptr)
{
    std::vector<std::string> names;
    rd.expect("[");
    if (!rd.accept("]")) {
        do {
            if (fixed && (rd.peek().k == Tok::num || (rd.peek().s == "-" && rd.peek(1).k == Tok::num))) {
                const bool neg = rd.accept("-");
                fixed->push_back(neg ? -rd.next().v : rd.next().v);
                names.push_back("_fixed" + std::to_string(names.size()));
                continue;
            }
            if (rd.peek().k != Tok::id) rd.fail("expected an identifier");
            names.push_back(rd.next().s);
            if (fixed) fixed->push_back(INT64_MIN);
        } while (rd.accept(","));
        rd.expect("]");
    }
    return names;
}

}  // namespace

computation::computation(std::string iteration_domain, expr e, bool schedule_this_computation, primitive_t t, function *fct)
{
    Reader rd(iteration_domain);
    if (rd.peek().s == "[") { parse_name_list(rd); rd.expect("->"); }       // parameters: resolved by name later
    rd.expect("{");
    if (rd.peek().k != Tok::id) rd.fail("expected the computation name");
    std::string name = rd.next().s;
    std::vector<int64_t> fixed;
    std::vector<std::string> iters = parse_name_list(rd, &fixed);
    std::vector<Cmp> cmps;
    if (rd.accept(":")) parse_conjunction(rd, cmps);
    if (rd.peek().k == Tok::id && rd.peek().s == "or") rd.fail("disjunctions are not boxes: only rectangular iteration domains are supported");
    rd.expect("}");

    std::vector<expr> lo(iters.size()), hi(iters.size());
    for (size_t i = 0; i < iters.size(); i++)
        if (fixed[i] != INT64_MIN) { lo[i] = expr((int32_t) fixed[i]); hi[i] = expr((int32_t) (fixed[i] + 1)); }
    auto iter_index = [&](const expr &x) {
        if (x.get_expr_type() != e_var) return -1;
        for (size_t i = 0; i < iters.size(); i++) if (iters[i] == x.get_name()) return (int) i;
        return -1;
    };
    auto mentions_iter = [&](const expr &x) {
        for (const std::string &it : iters)
            if (!x.substitute({it}, {expr((int32_t) 0)}).is_equal(x)) return true;
        return false;
    };
    // `k * i` with a positive literal k (benchmarks/halide/rgbyuv420gpu_tiramisu.cpp: `0 <= 2 * i < N`): the iterator and k
    auto scaled_iter = [&](const expr &x, int64_t &k) {
        if (x.get_expr_type() != e_op || x.get_op_type() != o_mul) return -1;
        for (int a = 0; a < 2; a++)
            if (x.get_operand(a).is_int_literal() && x.get_operand(a).get_int_val() > 0 && iter_index(x.get_operand(1 - a)) >= 0) {
                k = x.get_operand(a).get_int_val();
                return iter_index(x.get_operand(1 - a));
            }
        return -1;
    };
    std::vector<std::pair<expr, expr>> two_points(iters.size());
    struct TileGuard { int outer, inner; int64_t k; expr R; };
    std::vector<TileGuard> tile_guards;
    int tp_outer = -1, tp_inner = -1;
    int64_t tp_k = 0;
    // x = k * o + i (either operand order) with o, i iterators and k a positive literal
    auto tiled_pair = [&](const expr &x, int &o, int &i, int64_t &k) {
        if (x.get_expr_type() != e_op || x.get_op_type() != o_add) return false;
        for (int a = 0; a < 2; a++) {
            const int so = scaled_iter(x.get_operand(a), k), si = iter_index(x.get_operand(1 - a));
            if (so >= 0 && si >= 0 && so != si && k >= 2) { o = so; i = si; return true; }
        }
        return false;
    };
    for (const Cmp &c : cmps) {
        int li = iter_index(c.l), ri = iter_index(c.r);
        int64_t k = 1;
        if (c.op == "in2") {
            if (li < 0 || mentions_iter(c.r) || mentions_iter(c.r2)) die("iteration domain of '" + name + "': unsupported disjunction in \"" + iteration_domain + "\"");
            two_points[li] = {c.r, c.r2};
            lo[li] = expr((int32_t) 0);
            hi[li] = expr((int32_t) 2);
        } else if (li < 0 && ri < 0 && scaled_iter(c.l, k) >= 0 && !mentions_iter(c.r)) {
            // k*i op R (operands non-negative): i < ceil(R / k), i <= floor(R / k), ...
            const int d = scaled_iter(c.l, k);
            const expr K((int32_t) k);
            if (c.op == "<") hi[d] = (c.r + (int32_t) (k - 1)) / K; else if (c.op == "<=") hi[d] = c.r / K + 1;
            else if (c.op == ">=") lo[d] = (c.r + (int32_t) (k - 1)) / K; else if (c.op == ">") lo[d] = c.r / K + 1;
            else die("iteration domain of '" + name + "' pins a scaled iterator: \"" + iteration_domain + "\"");
        } else if (li < 0 && ri < 0 && scaled_iter(c.r, k) >= 0 && !mentions_iter(c.l)) {
            const int d = scaled_iter(c.r, k);
            const expr K((int32_t) k);
            if (c.op == "<=") lo[d] = (c.l + (int32_t) (k - 1)) / K; else if (c.op == "<") lo[d] = c.l / K + 1;
            else if (c.op == ">") hi[d] = (c.l + (int32_t) (k - 1)) / K; else if (c.op == ">=") hi[d] = c.l / K + 1;
            else die("iteration domain of '" + name + "' pins a scaled iterator: \"" + iteration_domain + "\"");
        } else if (li >= 0 && ri < 0 && !mentions_iter(c.r)) {
            if (c.op == "<=") hi[li] = c.r + 1; else if (c.op == "<") hi[li] = c.r;
            else if (c.op == ">=") lo[li] = c.r; else if (c.op == ">") lo[li] = c.r + 1;
            else { lo[li] = c.r; hi[li] = c.r + 1; }
        } else if (ri >= 0 && li < 0 && !mentions_iter(c.l)) {
            if (c.op == "<=") lo[ri] = c.l; else if (c.op == "<") lo[ri] = c.l + 1;
            else if (c.op == ">=") hi[ri] = c.l + 1; else if (c.op == ">") hi[ri] = c.l;
            else { lo[ri] = c.l; hi[ri] = c.l + 1; }
        } else if (li < 0 && ri < 0 && !mentions_iter(c.l) && !mentions_iter(c.r)) {
            continue;                                    // a constraint on parameters only: context, not a bound
        } else if (tiled_pair(c.l, tp_outer, tp_inner, tp_k) && !mentions_iter(c.r) && (c.op == "<" || c.op == "<=")) {
            // a hand-tiled iterator pair, `k * o + i < R` with 0 <= i < k declared on its own (benchmarks/halide/
            // warp_affinegpu_tiramisu.cpp:56: `0 <= i0 * 16 + i1 < N and 0 <= i1 < 16`): the box o in [0, ceil(R / k)) x i in
            // [0, k) with the excess of the last tile cut by a guard — what split() produces for a non-dividing factor
            const expr R = c.op == "<" ? c.r : c.r + 1;
            hi[tp_outer] = (R + (int32_t) (tp_k - 1)) / expr((int32_t) tp_k);
            tile_guards.push_back({tp_outer, tp_inner, tp_k, R});
        } else if (tiled_pair(c.r, tp_outer, tp_inner, tp_k) && c.l.is_int_literal() && c.l.get_int_val() <= 0 && (c.op == "<=" || (c.op == "<" && c.l.get_int_val() < 0))) {
            continue;                                    // `0 <= k * o + i`: implied by o >= 0 and i >= 0 (checked below)
        } else {
            die("iteration domain of '" + name + "' is not a box (constraint couples iterators): \"" + iteration_domain + "\"");
        }
    }
    const bool asked_to_schedule = schedule_this_computation;
    bool unbounded = false;
    if (!e.is_defined()) schedule_this_computation = false;      // a declaration only (input / wrapper computation)
    std::vector<var> vars;
    for (size_t i = 0; i < iters.size(); i++) {
        if (lo[i].is_defined() && hi[i].is_defined()) vars.push_back(var(iters[i], lo[i], hi[i]));
        else if (!schedule_this_computation) vars.push_back(var(iters[i]));
        else if (!lo[i].is_defined() && !hi[i].is_defined()) { vars.push_back(var(iters[i])); unbounded = true; }   // compute_bounds()
        else die("iterator '" + iters[i] + "' of '" + name + "' has no lower or no upper bound in \"" + iteration_domain + "\"");
    }
    if (unbounded) {
        // every bound must come from the consumers: no partial boxes
        vars.clear();
        for (const std::string &it : iters) vars.push_back(var(it));
    }
    for (const TileGuard &g : tile_guards) {
        const bool inner_ok = lo[g.inner].is_defined() && hi[g.inner].is_defined() && lo[g.inner].is_int_literal() && lo[g.inner].get_int_val() == 0 &&
                              hi[g.inner].is_int_literal() && hi[g.inner].get_int_val() == g.k;
        const bool outer_ok = lo[g.outer].is_defined() && lo[g.outer].is_int_literal() && lo[g.outer].get_int_val() == 0;
        if (!inner_ok || !outer_ok)
            die("iteration domain of '" + name + "' is not a box (constraint couples iterators): \"" + iteration_domain + "\"");
    }
    construct(name, vars, e, schedule_this_computation && !unbounded, t, fct);
    for (const TileGuard &g : tile_guards)
        guards_.push_back(var(iters[g.outer]) * expr((int32_t) g.k) + var(iters[g.inner]) < g.R);
    if (unbounded) { scheduled_ = true; pending_bounds_ = true; }
    // a two-point iterator runs as a loop over {0, 1}; the iterator the expression and the access name is a + t * (b - a)
    for (size_t i = 0; i < iters.size(); i++)
        if (two_points[i].first.is_defined() && !unbounded) {
            const int d = find_dim(iters[i]);
            if (d < 0) continue;
            const std::string nn = iters[i] + "_pt";
            defs_.push_back({iters[i], two_points[i].first + var(nn) * (two_points[i].second - two_points[i].first)});
            dims_[d].name = nn;
        }
    is_input_ = !schedule_this_computation && !e.is_defined();
    // declared with an empty expression but to be scheduled: a statement whose expression comes later through
    // set_expression (reference tests/test_04.cpp); until then it reads like an input
    schedule_when_defined_ = asked_to_schedule && !e.is_defined();
}

void computation::add_definitions(std::string iteration_domain, expr e, bool schedule_this_computation, primitive_t t, function *fct)
{
    function *f = fct ? fct : fct_;
    computation *u = f->own(new computation(iteration_domain, e, schedule_this_computation, t, f));
    if (u->get_name() != name_) die("add_definitions: the new definition is named '" + u->get_name() + "', not '" + name_ + "'");
    updates_.push_back(u);
    if (buf_) set_access_exprs(buf_, access_);           // same mapping until told otherwise (re-propagates to every unmapped update)
}

void computation::set_access(std::string access_str)
{
    Reader rd(access_str);
    if (rd.peek().s == "[") { parse_name_list(rd); rd.expect("->"); }
    rd.expect("{");
    if (rd.peek().k != Tok::id) rd.fail("expected the computation name");
    rd.next();
    std::vector<int64_t> fixed;
    std::vector<std::string> names = parse_name_list(rd, &fixed);
    rd.expect("->");
    if (rd.peek().k != Tok::id) rd.fail("expected a buffer name");
    std::string bname = rd.next().s;
    std::vector<expr> idx;
    rd.expect("[");
    if (!rd.accept("]")) {
        do idx.push_back(rd.parse_expr()); while (rd.accept(","));
        rd.expect("]");
    }
    rd.accept(":");      // (conditions on accesses are not used by the hot-path generators)
    if (names.size() != iters_.size())
        die("set_access of '" + name_ + "': " + std::to_string(names.size()) + " iterators named, the computation has " + std::to_string(iters_.size()));
    buffer *b = fct_->find_buffer(bname);
    if (!b) die("set_access of '" + name_ + "': unknown buffer '" + bname + "'");
    // the names in the map are positional aliases of the computation's iterators
    std::vector<expr> own;
    for (const var &v : iters_) own.push_back(expr::make_var(global::get_loop_iterator_data_type(), v.get_name()));
    for (expr &e : idx) e = e.substitute(names, own);
    set_access_exprs(b, idx);
    own_access_ = true;
}

}  // namespace tiramisu
