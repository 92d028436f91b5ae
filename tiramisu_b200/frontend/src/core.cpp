// core.cpp — buffers, computations and scheduling commands of the tiramisu_b200 front-end
// (see include/tiramisu/core.h).  A schedule is a list of named loop dimensions; every command
// rewrites the list and records how older names are rebuilt from newer ones.
#include <algorithm>
#include <set>
#include <tiramisu/block.h>
#include <tiramisu/core.h>

#include <cstdlib>
#include <iostream>
#include <stdexcept>

namespace tiramisu {

[[noreturn]] static void die(const std::string &msg)
{
    std::cerr << "tiramisu_b200 front-end: " << msg << std::endl;
    std::exit(1);
}

// ---- function -----------------------------------------------------------------------------------
function::function(std::string name) : name_(std::move(name))
{
    // an explicitly declared function becomes the implicit one for objects declared without it
    if (!global::get_implicit_function()) global::set_implicit_function(this);
}

void function::gen_halide_obj(const std::string &obj_file_name, bool)
{
    codegen(args_, obj_file_name, cuda_requested_);
}

void function::resolve_static_order()
{
    if (!static_order_) return;
    std::vector<computation *> sched;
    for (computation *c : comps_) {
        if (c->is_input() || c->is_constant() || !c->should_schedule_this_computation()) continue;
        if (c->pred_ || c->succ_)
            die("'" + c->get_name() + "' is ordered with then()/after() in a function that also uses after_low_level: one style per function");
        sched.push_back(c);
    }
    auto num = [](const computation *c, size_t k) { return k < c->static_.size() ? c->static_[k] : 0; };
    size_t depth = 0;
    for (computation *c : sched) depth = std::max(depth, c->static_.size());
    // lexicographic order of the number vectors; statements with equal vectors keep their declaration order
    std::stable_sort(sched.begin(), sched.end(), [&](const computation *a, const computation *b) {
        for (size_t k = 0; k < depth; k++)
            if (num(a, k) != num(b, k)) return num(a, k) < num(b, k);
        return false;
    });
    // ---- statements cut into pieces: the older generators write one statement per full / partial tile region
    // (sgemm_generator.cpp:90-110: `reduced_AB_1`, `_p0`, `_p1`, `_p2`) because their CPU code generator wants the full
    // tiles free of guards.  Consecutive statements with the same expression, the same store and iterator boxes that
    // partition one box ARE one statement over that box; the device lowering (and its pattern recognition) sees that one.
    std::set<computation *> merged;
    auto same_statement = [](computation *a, computation *b) {
        if (a->dtype_ != b->dtype_ || a->buf_ != b->buf_ || !a->buf_ || a->predicate_.is_defined() || b->predicate_.is_defined()) return false;
        if (!a->expr_.is_defined() || !b->expr_.is_defined() || !a->expr_.is_equal(b->expr_)) return false;
        if (a->iters_.size() != b->iters_.size() || a->access_.size() != b->access_.size()) return false;
        for (size_t d = 0; d < a->iters_.size(); d++)
            if (a->iters_[d].get_name() != b->iters_[d].get_name() || !a->iters_[d].has_range() || !b->iters_[d].has_range()) return false;
        for (size_t d = 0; d < a->access_.size(); d++)
            if (!a->access_[d].is_equal(b->access_[d])) return false;
        return true;
    };
    auto referenced_elsewhere = [&](computation *c, size_t lo_i, size_t hi_i) {
        for (size_t q = 0; q < sched.size(); q++) {
            if (q >= lo_i && q < hi_i) continue;
            std::vector<expr> stack{sched[q]->expr_, sched[q]->predicate_};
            while (!stack.empty()) {
                expr x = stack.back(); stack.pop_back();
                if (!x.is_defined() || x.get_expr_type() != e_op) continue;
                if (x.get_op_type() == o_access && x.get_name() == c->get_name()) return true;
                for (const expr &o : x.get_operands()) stack.push_back(o);
            }
        }
        return false;
    };
    for (size_t i = 0; i < sched.size();) {
        size_t e = i + 1;
        while (e < sched.size() && same_statement(sched[i], sched[e])) e++;
        const size_t n = e - i, nd = sched[i]->iters_.size();
        bool ok = n > 1;
        std::vector<std::pair<expr, expr>> box(nd);
        size_t combos = 1;
        for (size_t d = 0; d < nd && ok; d++) {
            std::vector<std::pair<expr, expr>> uniq;
            for (size_t q = i; q < e; q++) {
                const var &v = sched[q]->iters_[d];
                bool seen = false;
                for (auto &u : uniq) seen |= u.first.is_equal(v.get_lower()) && u.second.is_equal(v.get_upper());
                if (!seen) uniq.push_back({v.get_lower(), v.get_upper()});
            }
            if (uniq.size() == 1) box[d] = uniq[0];
            else if (uniq.size() == 2 && uniq[0].second.is_equal(uniq[1].first)) box[d] = {uniq[0].first, uniq[1].second};
            else if (uniq.size() == 2 && uniq[1].second.is_equal(uniq[0].first)) box[d] = {uniq[1].first, uniq[0].second};
            else ok = false;
            combos *= uniq.size();
        }
        ok = ok && combos == n;
        for (size_t q = i; q < e && ok; q++)            // every piece a different region (with combos == n: every region once)
            for (size_t r = q + 1; r < e && ok; r++) {
                bool same = true;
                for (size_t d = 0; d < nd; d++)
                    same &= sched[q]->iters_[d].get_lower().is_equal(sched[r]->iters_[d].get_lower()) && sched[q]->iters_[d].get_upper().is_equal(sched[r]->iters_[d].get_upper());
                ok = !same;
            }
        computation *base = nullptr;
        for (size_t q = i; q < e && ok; q++)
            if (referenced_elsewhere(sched[q], i, e)) { if (base) ok = false; base = sched[q]; }
        if (!ok) { i++; continue; }
        if (!base) base = sched[i];
        std::string names;
        for (size_t q = i; q < e; q++) names += (q > i ? ", " : "") + sched[q]->get_name();
        std::cerr << "tiramisu_b200 front-end: note: statements {" << names << "} are the pieces of one statement over one box; lowered as '" << base->get_name() << "'\n";
        for (size_t d = 0; d < nd; d++) base->iters_[d] = var(base->iters_[d].get_name(), box[d].first, box[d].second);
        base->dims_.clear(); base->defs_.clear(); base->guards_.clear();
        for (const var &v : base->iters_) { loop_dim ld; ld.name = v.get_name(); ld.lo = v.get_lower(); ld.hi = v.get_upper(); base->dims_.push_back(ld); }
        // the schedule of the pieces (tile sizes, unroll / vector levels) described the CPU loop nest of each piece;
        // the merged statement gets the plain data-parallel CPU schedule the re-targeting starts from
        base->parallelize(base->iters_[0]);
        if (nd >= 2 && !(box[1].second - box[1].first).is_int_literal()) base->vectorize(base->iters_[1], 8);
        merged.insert(base);
        for (size_t q = i; q < e; q++)
            if (sched[q] != base) comps_.erase(std::remove(comps_.begin(), comps_.end(), sched[q]), comps_.end());
        sched.erase(sched.begin() + i, sched.begin() + e);
        sched.insert(sched.begin() + i, base);
        i++;
    }
    // ---- a temporary nobody reads (sgemm_generator.cpp packs B into buf_Bp and then multiplies with B itself): not computed
    for (size_t i = 0; i < sched.size();) {
        computation *c = sched[i];
        bool read = !c->buf_ || c->buf_->get_argument_type() != a_temporary;
        for (computation *o : comps_) {
            if (read) break;
            if (o->buf_ != c->buf_) continue;
            read = referenced_elsewhere(o, sched.size(), sched.size()) && true;
        }
        if (read) { i++; continue; }
        std::cerr << "tiramisu_b200 front-end: note: '" << c->get_name() << "' fills the temporary '" << c->buf_->get_name() << "' that nothing reads; not computed\n";
        comps_.erase(std::remove(comps_.begin(), comps_.end(), c), comps_.end());
        sched.erase(sched.begin() + i);
    }
    for (size_t i = 1; i < sched.size(); i++) {
        computation *a = sched[i - 1], *b = sched[i];
        if (merged.count(a) || merged.count(b)) { b->link_after(*a, "", -1); continue; }      // fresh loop nests: the old depths mean nothing
        size_t k = 0;
        while (k < depth && num(a, k) == num(b, k)) k++;
        // equal numbers at depths 0..k-1: the two share loops 0..k-2 and are ordered below loop k-2 ... the number at depth d
        // orders statements INSIDE loop d-1, so sharing loops 0..k-1 needs equal numbers at depths 0..k and a difference at
        // depth k+1.  Equal vectors altogether (the reference then interleaves by iterator value; the statements of such a
        // pair are independent): one after the other at root.
        int level = k >= depth ? -1 : (int) k - 1;
        level = std::min(level, std::min((int) a->dims().size(), (int) b->dims().size()) - 1);
        b->link_after(*a, "", level);
    }
}

computation *function::find_computation(const std::string &name) const
{
    for (auto it = comps_.rbegin(); it != comps_.rend(); ++it)
        if ((*it)->get_name() == name) return *it;
    return nullptr;
}
constant *function::get_invariant_by_name(std::string name) const
{
    computation *c = find_computation(name);
    return c && c->is_constant() ? static_cast<constant *>(c) : nullptr;
}
buffer *function::find_buffer(const std::string &name) const
{
    for (auto it = bufs_.rbegin(); it != bufs_.rend(); ++it)
        if ((*it)->get_name() == name) return *it;
    return nullptr;
}
std::string function::fresh_name(const std::string &prefix) { return prefix + std::to_string(counter_++); }
void function::dump(bool) const
{
    for (auto *c : comps_) c->dump();
}

void init(std::string name)
{
    global::set_default_tiramisu_options();
    global::set_implicit_function(new function(std::move(name)));   // lives for the generator's lifetime
}
void init() { init("foo"); }

void codegen(const std::vector<buffer *> &arguments, const std::string obj_filename, bool gen_cuda_stmt)
{
    function *f = global::get_implicit_function();
    if (!f) die("codegen() called before tiramisu::init()");
    f->codegen(arguments, obj_filename, gen_cuda_stmt);
}

// ---- buffer -------------------------------------------------------------------------------------
buffer::buffer() {}
buffer::buffer(std::string name, std::vector<expr> dim_sizes, primitive_t type, argument_t argt, function *fct)
    : name_(std::move(name)), dims_(std::move(dim_sizes)), dtype_(type), argt_(argt), fct_(fct)
{
    if (!fct_) die("buffer '" + name_ + "' declared before tiramisu::init()");
    fct_->add_buffer(this);
}

static std::vector<var> prefix_iters(computation &C, const std::string &level)
{
    // the iteration variables of C's scheduled nest down to (and including) `level`
    std::vector<var> out;
    if (level.empty()) return out;
    int d = C.find_dim(level);
    if (d < 0) die("allocate_at/deallocate_at: '" + C.get_name() + "' has no loop level '" + level + "'");
    for (int i = 0; i <= d; i++) out.push_back(var(C.dims()[i].name, C.dims()[i].lo, C.dims()[i].hi));
    return out;
}

computation *buffer::allocate_at(computation &C, var level)
{
    auto_allocate_ = false;
    auto *c = new computation(fct_->fresh_name("_allocate_" + name_ + "_"), prefix_iters(C, level.get_name()),
                              expr(o_allocate, name_), true, p_none);
    return fct_->own(c);
}
computation *buffer::allocate_at(computation &C, int level)
{
    if (level < 0) return allocate_at(C, computation::root);
    return allocate_at(C, var(C.dims().at(level).name));
}
computation *buffer::deallocate_at(computation &C, var level)
{
    auto *c = new computation(fct_->fresh_name("_free_" + name_ + "_"), prefix_iters(C, level.get_name()),
                              expr(o_free, name_), true, p_none);
    return fct_->own(c);
}
void buffer::dump(bool) const { std::cerr << "buffer " << name_ << "\n"; }
void buffer::substitute_dims(const std::vector<std::string> &names, const std::vector<expr> &values)
{
    for (expr &d : dims_) d = d.substitute(names, values);
}

// ---- computation ----------------------------------------------------------------------------------
const var computation::root("");

computation::computation() {}

void computation::construct(std::string name, std::vector<var> iters, expr e, bool schedule, primitive_t t, function *fct)
{
    fct_ = fct ? fct : global::get_implicit_function();
    if (!fct_) die("computation declared before tiramisu::init()");
    name_ = name.empty() ? fct_->fresh_name("_C") : std::move(name);
    iters_ = std::move(iters);
    expr_ = e;
    scheduled_ = schedule;
    dtype_ = t;
    for (const var &v : iters_) {
        loop_dim d;
        d.name = v.get_name();
        if (v.has_range()) { d.lo = v.get_lower(); d.hi = v.get_upper(); }
        else if (schedule) die("iterator '" + v.get_name() + "' of computation '" + name_ + "' has no range");
        dims_.push_back(d);
    }
    fct_->add_computation(this);
}

static primitive_t type_of(const expr &e)
{
    primitive_t t = e.get_data_type();
    return t;
}

computation::computation(std::string name, std::vector<var> it, expr e, bool s, primitive_t t) { construct(name, it, e, s, t); }
computation::computation(std::string name, std::vector<var> it, expr e, bool s) { construct(name, it, e, s, type_of(e)); }
computation::computation(std::string name, std::vector<var> it, expr e) { construct(name, it, e, true, type_of(e)); }
computation::computation(std::vector<var> it, expr e) { construct("", it, e, true, type_of(e)); }
computation::computation(std::vector<var> it, expr e, bool s) { construct("", it, e, s, type_of(e)); }
computation::computation(std::string name, std::vector<var> it, primitive_t t) { construct(name, it, expr(), true, t); }
computation::computation(std::vector<var> it, primitive_t t) { construct("", it, expr(), true, t); }
computation::computation(std::string name, std::vector<var> it, expr pred, expr e) { construct(name, it, e, true, type_of(e)); predicate_ = pred; }
computation::computation(std::vector<var> it, expr pred, expr e) { construct("", it, e, true, type_of(e)); predicate_ = pred; }

void computation::set_expression(const expr &e)
{
    expr_ = e;
    if (dtype_ == p_none) dtype_ = e.get_data_type();
    if (schedule_when_defined_ && e.is_defined()) {
        for (const var &v : iters_)
            if (!v.has_range()) die("set_expression: iterator '" + v.get_name() + "' of '" + name_ + "' has no bounds");
        schedule_when_defined_ = false;
        is_input_ = false;
        scheduled_ = true;
    }
}

expr computation::access_expr(const std::vector<expr> &idx)
{
    return expr(o_access, name_, idx, dtype_);
}

int computation::find_dim(const std::string &name) const
{
    for (size_t i = 0; i < dims_.size(); i++)
        if (dims_[i].name == name) return (int) i;
    return -1;
}

void computation::dump() const
{
    std::cerr << name_ << " [";
    for (auto &d : dims_) std::cerr << " " << d.name << ":" << d.lo.to_str() << ".." << d.hi.to_str() << (d.tag ? "#" + std::to_string((int) d.tag) : "");
    std::cerr << " ] = " << expr_.to_str() << "\n";
}

void computation::substitute_everywhere(const std::vector<std::string> &names, const std::vector<expr> &values)
{
    expr_ = expr_.substitute(names, values);
    predicate_ = predicate_.substitute(names, values);
    for (auto &d : dims_) { d.lo = d.lo.substitute(names, values); d.hi = d.hi.substitute(names, values); }
    for (auto &d : defs_) d.second = d.second.substitute(names, values);
    for (auto &g : guards_) g = g.substitute(names, values);
    for (auto &a : access_) a = a.substitute(names, values);
    for (var &v : iters_)
        if (v.has_range()) v = var(v.get_name(), v.get_lower().substitute(names, values), v.get_upper().substitute(names, values));
}

// -- Layer II
static bool divisible(const expr &extent, int f)
{
    return extent.is_int_literal() && extent.get_int_val() % f == 0;
}

void computation::split(var L0, int f, var outer, var inner)
{
    int d = find_dim(L0.get_name());
    if (d < 0) die("split: '" + name_ + "' has no loop level '" + L0.get_name() + "'");
    if (f <= 0) die("split: non-positive factor");
    loop_dim old = dims_[d];
    expr extent = old.hi - old.lo;
    loop_dim o, i;
    o.name = outer.get_name(); o.lo = expr((int32_t) 0);
    o.hi = divisible(extent, f) ? expr((int32_t) (extent.get_int_val() / f)) : (extent + (f - 1)) / f;
    i.name = inner.get_name(); i.lo = expr((int32_t) 0); i.hi = expr((int32_t) f);
    var vo(o.name), vi(i.name);
    defs_.push_back({old.name, old.lo + (vo * f + vi)});
    if (!divisible(extent, f)) guards_.push_back(var(old.name) < old.hi);
    dims_[d] = o;
    dims_.insert(dims_.begin() + d + 1, i);
}
void computation::split(var L0, int f) { split(L0, f, var(L0.get_name() + "_o"), var(L0.get_name() + "_i")); }

void computation::interchange(var L0, var L1)
{
    int a = find_dim(L0.get_name()), b = find_dim(L1.get_name());
    if (a < 0 || b < 0) die("interchange: '" + name_ + "' has no loop level '" + (a < 0 ? L0 : L1).get_name() + "'");
    std::swap(dims_[a], dims_[b]);
}

void computation::tile(var L0, var L1, int sx, int sy, var o0, var o1, var i0, var i1)
{
    int a = find_dim(L0.get_name()), b = find_dim(L1.get_name());
    if (a < 0 || b != a + 1) die("tile: levels '" + L0.get_name() + "','" + L1.get_name() + "' of '" + name_ + "' must be consecutive");
    split(L0, sx, o0, i0);          // o0 i0 L1
    split(L1, sy, o1, i1);          // o0 i0 o1 i1
    interchange(i0, o1);            // o0 o1 i0 i1
}
void computation::tile(var L0, var L1, int sx, int sy)
{
    tile(L0, L1, sx, sy, var(L0.get_name() + "_o"), var(L1.get_name() + "_o"), var(L0.get_name() + "_i"), var(L1.get_name() + "_i"));
}
void computation::tile(var L0, var L1, var L2, int sx, int sy, int sz, var o0, var o1, var o2, var i0, var i1, var i2)
{
    int a = find_dim(L0.get_name()), b = find_dim(L1.get_name()), c = find_dim(L2.get_name());
    if (a < 0 || b != a + 1 || c != a + 2) die("tile: three consecutive levels expected in '" + name_ + "'");
    split(L0, sx, o0, i0);          // o0 i0 L1 L2
    split(L1, sy, o1, i1);          // o0 i0 o1 i1 L2
    split(L2, sz, o2, i2);          // o0 i0 o1 i1 o2 i2
    interchange(i0, o1);            // o0 o1 i0 i1 o2 i2
    interchange(i1, o2);            // o0 o1 i0 o2 i1 i2
    interchange(i0, o2);            // o0 o1 o2 i0 i1 i2
}
void computation::tile(var L0, var L1, var L2, int sx, int sy, int sz)
{
    auto o = [](const var &v) { return var(v.get_name() + "_o"); };
    auto i = [](const var &v) { return var(v.get_name() + "_i"); };
    tile(L0, L1, L2, sx, sy, sz, o(L0), o(L1), o(L2), i(L0), i(L1), i(L2));
}

void computation::after_low_level(computation &b, int level)
{
    if (&b == this) die("a computation cannot be ordered after itself: " + name_);
    if (level < -1) die("after_low_level: level " + std::to_string(level) + " of '" + name_ + "'");
    const size_t need = (size_t) level + 2;
    if (static_.size() < need) static_.resize(need, 0);
    if (b.static_.size() < need) b.static_.resize(need, 0);
    for (int k = 0; k <= level; k++) static_[k] = b.static_[k];
    static_[level + 1] = b.static_[level + 1] + 10;
    if (fct_) fct_->note_static_order();
}

// ---- by level
static var level_var(computation &c, int L, const char *what)
{
    if (L < 0 || L >= (int) c.dims().size())
        die(std::string(what) + ": '" + c.get_name() + "' has " + std::to_string(c.dims().size()) + " loop levels, level " + std::to_string(L) + " asked for");
    return var(c.dims()[L].name);
}
void computation::split(int L0, int f) { split(level_var(*this, L0, "split"), f); }
void computation::tile(int L0, int L1, int sx, int sy) { tile(level_var(*this, L0, "tile"), level_var(*this, L1, "tile"), sx, sy); }
void computation::tile(int L0, int L1, int L2, int sx, int sy, int sz)
{
    tile(level_var(*this, L0, "tile"), level_var(*this, L1, "tile"), level_var(*this, L2, "tile"), sx, sy, sz);
}
void computation::interchange(int L0, int L1) { interchange(level_var(*this, L0, "interchange"), level_var(*this, L1, "interchange")); }
// A tag on a level the statement does not have matches no loop: the reference keeps (statement, level) pairs and looks them
// up while it walks the generated loops (src/tiramisu_core.cpp: tag_unroll_level / should_unroll), so such a tag is inert
// there too — sgemm_generator.cpp:262 relies on it.
static bool has_level(computation &c, int L) { return L >= 0 && L < (int) c.dims().size(); }
void computation::tag_parallel_level(int L) { if (has_level(*this, L)) tag_parallel_level(level_var(*this, L, "tag_parallel_level")); }
void computation::tag_vector_level(int L, int v) { if (has_level(*this, L)) tag_vector_level(level_var(*this, L, "tag_vector_level"), v); }
void computation::tag_unroll_level(int L) { if (has_level(*this, L)) tag_unroll_level(level_var(*this, L, "tag_unroll_level")); }
void computation::tag_gpu_level(int L0, int L1) { tag_gpu_level(level_var(*this, L0, "tag_gpu_level"), level_var(*this, L1, "tag_gpu_level")); }
void computation::tag_gpu_level(int L0, int L1, int L2, int L3)
{
    tag_gpu_level(level_var(*this, L0, "tag_gpu_level"), level_var(*this, L1, "tag_gpu_level"), level_var(*this, L2, "tag_gpu_level"),
                  level_var(*this, L3, "tag_gpu_level"));
}

void computation::tag(const var &L, loop_dim::tag_t t)
{
    int d = find_dim(L.get_name());
    // a constant declared at a loop level is a let: it lives inside the statements that name it and is mapped with them
    // (benchmarks/halide/warp_affinegpu_tiramisu.cpp:88 tags `o_r` like the statement it belongs to)
    if (d < 0 && is_let_) return;
    if (d < 0) die("'" + name_ + "' has no loop level '" + L.get_name() + "' to tag");
    dims_[d].tag = t;
}
void computation::tag_gpu_block_level(var a) { tag(a, loop_dim::gpu_block); }
void computation::tag_gpu_block_level(var a, var b) { tag(a, loop_dim::gpu_block); tag(b, loop_dim::gpu_block); }
void computation::tag_gpu_block_level(var a, var b, var c) { tag_gpu_block_level(a, b); tag(c, loop_dim::gpu_block); }
void computation::tag_gpu_thread_level(var a) { tag(a, loop_dim::gpu_thread); }
void computation::tag_gpu_thread_level(var a, var b) { tag(a, loop_dim::gpu_thread); tag(b, loop_dim::gpu_thread); }
void computation::tag_gpu_thread_level(var a, var b, var c) { tag_gpu_thread_level(a, b); tag(c, loop_dim::gpu_thread); }
void computation::tag_gpu_level(var a, var b) { tag_gpu_block_level(a); tag_gpu_thread_level(b); }
void computation::tag_gpu_level(var a, var b, var c, var d) { tag_gpu_block_level(a, b); tag_gpu_thread_level(c, d); }
void computation::tag_gpu_level(var a, var b, var c, var d, var e, var f) { tag_gpu_block_level(a, b, c); tag_gpu_thread_level(d, e, f); }

void computation::gpu_tile(var L0, int sx, var o, var i) { split(L0, sx, o, i); tag_gpu_level(o, i); }
void computation::gpu_tile(var L0, int sx) { gpu_tile(L0, sx, var(L0.get_name() + "_o"), var(L0.get_name() + "_i")); }
void computation::gpu_tile(var L0, var L1, int sx, int sy, var o0, var o1, var i0, var i1)
{
    tile(L0, L1, sx, sy, o0, o1, i0, i1);
    tag_gpu_level(o0, o1, i0, i1);
}
void computation::gpu_tile(var L0, var L1, int sx, int sy)
{
    gpu_tile(L0, L1, sx, sy, var(L0.get_name() + "_o"), var(L1.get_name() + "_o"), var(L0.get_name() + "_i"), var(L1.get_name() + "_i"));
}
void computation::gpu_tile(var L0, var L1, var L2, int sx, int sy, int sz, var o0, var o1, var o2, var i0, var i1, var i2)
{
    tile(L0, L1, L2, sx, sy, sz, o0, o1, o2, i0, i1, i2);
    tag_gpu_level(o0, o1, o2, i0, i1, i2);
}
void computation::gpu_tile(var L0, var L1, var L2, int sx, int sy, int sz)
{
    auto o = [](const var &v) { return var(v.get_name() + "_o"); };
    auto i = [](const var &v) { return var(v.get_name() + "_i"); };
    gpu_tile(L0, L1, L2, sx, sy, sz, o(L0), o(L1), o(L2), i(L0), i(L1), i(L2));
}

void computation::parallelize(var L) { tag(L, loop_dim::parallel); }
void computation::tag_parallel_level(var L) { tag(L, loop_dim::parallel); }
void computation::tag_vector_level(var L, int v) { tag(L, loop_dim::vector); dims_[find_dim(L.get_name())].vec_len = v; }
void computation::vectorize(var L, int v, var o, var i) { split(L, v, o, i); tag_vector_level(i, v); }
void computation::vectorize(var L, int v) { vectorize(L, v, var(L.get_name() + "_o"), var(L.get_name() + "_i")); }
void computation::tag_unroll_level(var L) { tag(L, loop_dim::unrolled); }
void computation::unroll(var L, int f, var o, var i) { split(L, f, o, i); tag_unroll_level(i); }
void computation::unroll(var L, int f) { unroll(L, f, var(L.get_name() + "_o"), var(L.get_name() + "_i")); }
void computation::shift(var L0, int n)
{
    // iterate L0' = L0 + n over [lo+n, hi+n); the original iterator is L0' - n
    int d = find_dim(L0.get_name());
    if (d < 0) die("shift: '" + name_ + "' has no loop level '" + L0.get_name() + "'");
    std::string nn = L0.get_name() + "_s";
    defs_.push_back({dims_[d].name, var(nn) - n});
    dims_[d].name = nn;
    dims_[d].lo = dims_[d].lo + n;
    dims_[d].hi = dims_[d].hi + n;
}

void computation::set_iterator_ranges(const std::vector<std::pair<expr, expr>> &ranges)
{
    if (ranges.size() != iters_.size() || dims_.size() != iters_.size()) die("bounds of '" + name_ + "' must be set before it is scheduled");
    for (size_t i = 0; i < iters_.size(); i++) {
        if (dims_[i].name != iters_[i].get_name()) die("bounds of '" + name_ + "' must be set before it is scheduled");
        iters_[i] = var(iters_[i].get_name(), ranges[i].first, ranges[i].second);
        dims_[i].lo = ranges[i].first;
        dims_[i].hi = ranges[i].second;
    }
    pending_bounds_ = false;
}

void computation::loop_reversal(var old_var, var new_var)
{
    // same range, opposite direction: the original iterator is lo + hi - 1 - new
    int d = find_dim(old_var.get_name());
    if (d < 0) die("loop_reversal: '" + name_ + "' has no loop level '" + old_var.get_name() + "'");
    const loop_dim was = dims_[d];
    defs_.push_back({was.name, (was.lo + was.hi - 1) - var(new_var.get_name())});
    dims_[d].name = new_var.get_name();
}

// -- ordering
void computation::link_after(computation &before_this, const std::string &level_name, int level_index)
{
    // a let lives inside the statement it was declared with: ordering relative to the let is ordering relative to that
    // statement (warp_affinegpu_tiramisu.cpp:107 `copy_to_device.before(o_r, root)`)
    if (is_let_ && let_host_ && let_host_ != &before_this) { let_host_->link_after(before_this, level_name, level_index); return; }
    if (before_this.is_let_ && before_this.let_host_ && before_this.let_host_ != this) { link_after(*before_this.let_host_, level_name, level_index); return; }
    if (&before_this == this) die("a computation cannot be ordered after itself: " + name_);
    // splice: `before_this` keeps at most one successor; a later successor takes over the chain position
    if (before_this.succ_ && before_this.succ_ != this) {
        computation *old = before_this.succ_;
        // this -> ... -> tail of this chain, then old follows it at root
        computation *tail = this;
        while (tail->succ_) tail = tail->succ_;
        if (tail != old) {
            tail->succ_ = old;
            old->pred_ = tail;
        }
    }
    if (pred_ && pred_ != &before_this) pred_->succ_ = nullptr;
    before_this.succ_ = this;
    pred_ = &before_this;
    pred_level_ = level_name;
    pred_level_index_ = level_index;
}

computation &computation::then(computation &next, var level)
{
    next.link_after(*this, level.get_name(), level.get_name().empty() ? -1 : -2);
    return next;
}
computation &computation::then(computation &next, int level)
{
    next.link_after(*this, "", level);
    return next;
}
void computation::after(computation &b, var level) { link_after(b, level.get_name(), level.get_name().empty() ? -1 : -2); }
void computation::after(computation &b, int level) { link_after(b, "", level); }
void computation::before(computation &a, var level) { a.after(*this, level); }
void computation::compute_at(computation &consumer, var)
{
    static bool told = false;
    if (!told) {
        std::cerr << "tiramisu_b200 front-end: note: compute_at('" << name_ << "' in '" << consumer.get_name()
                  << "') is a CPU locality schedule; the producer is computed once over its whole domain before the consumer\n";
        told = true;
    }
    if (consumer.pred_ != this) consumer.after(*this, computation::root);
}
void computation::compute_at(computation &consumer, int) { compute_at(consumer, var()); }
void computation::between(computation &before_c, var before_l, computation &after_c, var after_l)
{
    after(before_c, before_l);
    after_c.after(*this, after_l);
}

computation *computation::cache_shared(computation &inp, const var &level, std::vector<int>, std::vector<expr>, bool)
{
    // Not a transformation here, and said so at generator time: the reference rewrites the consumer to read a shared-memory
    // copy of `inp` filled below `level` (src/tiramisu_core.cpp:9297-9534).  This back end keeps the consumer reading `inp`
    // where it lives; a contraction written this way is recognised from Layer I and runs on the tensor cores
    // (lowering_patterns.cpp: try_staged_contraction / try_gemm_pattern), anything else reads through L1/L2.
    static bool told = false;
    if (!told) {
        std::cerr << "tiramisu_b200 front-end: note: cache_shared('" << inp.get_name() << "', level '" << level.get_name() << "') of '" << name_
                  << "' is recorded as a hint only: no shared-memory staging is generated for it (further notes of this kind are not repeated)\n";
        told = true;
    }
    return &inp;
}

// -- Layer III
void computation::set_access_exprs(buffer *b, std::vector<expr> idx)
{
    buf_ = b;
    access_ = std::move(idx);
    // the mapping of a computation also holds for its later definitions until they get their own
    // (reference tests/test_54.cpp maps S0 after add_definitions and never maps the update)
    std::vector<std::string> mine;
    for (const var &v : iters_) mine.push_back(v.get_name());
    for (computation *u : updates_) {
        if (u->buf_ && u->own_access_) continue;
        std::vector<expr> theirs;
        for (const var &v : u->iters_) theirs.push_back(expr::make_var(global::get_loop_iterator_data_type(), v.get_name()));
        if (theirs.size() != mine.size()) continue;
        u->buf_ = b;
        u->access_.clear();
        for (const expr &a : access_) u->access_.push_back(a.substitute(mine, theirs));
    }
}

void computation::store_in(buffer *buff)
{
    std::vector<expr> idx;
    int n = buff->get_n_dims();
    if ((int) iters_.size() < n) die("store_in: '" + name_ + "' has fewer iterators than buffer '" + buff->get_name() + "' has dimensions");
    for (int i = 0; i < n; i++) idx.push_back(expr::make_var(global::get_loop_iterator_data_type(), iters_[i].get_name()));
    set_access_exprs(buff, idx);
}
void computation::store_in(buffer *buff, std::vector<expr> iterators) { set_access_exprs(buff, std::move(iterators)); }
void computation::store_in(std::vector<expr> mapping, std::vector<expr> sizes)
{
    auto *b = fct_->own(new buffer("_" + name_ + "_buffer", sizes, dtype_, is_input_ ? a_input : a_temporary, fct_));
    set_access_exprs(b, std::move(mapping));
}

buffer *computation::get_buffer()
{
    if (!buf_) {
        // automatic data mapping: one buffer element per point of the original iteration box
        std::vector<expr> sizes, idx;
        for (const var &v : iters_) {
            if (!v.has_range()) die("get_buffer: iterator '" + v.get_name() + "' of '" + name_ + "' has no range");
            sizes.push_back(v.get_upper());
            idx.push_back(expr::make_var(global::get_loop_iterator_data_type(), v.get_name()));
        }
        if (sizes.empty()) { sizes.push_back(expr((int32_t) 1)); idx.push_back(expr((int32_t) 0)); }
        auto *b = fct_->own(new buffer("_" + name_ + "_buffer", sizes, dtype_, is_input_ ? a_input : a_output, fct_));
        set_access_exprs(b, idx);
    }
    return buf_;
}

// ---- input / constant -----------------------------------------------------------------------------
input::input(std::string name, std::vector<var> it, primitive_t t) { construct(name, it, expr(), false, t); is_input_ = true; }
input::input(std::vector<var> it, primitive_t t) { construct("", it, expr(), false, t); is_input_ = true; }
static std::vector<var> named_iters(const std::vector<std::string> &names, const std::vector<expr> &sizes)
{
    if (names.size() != sizes.size()) die("input: dimension names and sizes differ in number");
    std::vector<var> v;
    for (size_t i = 0; i < names.size(); i++) v.push_back(var(names[i], expr((int32_t) 0), sizes[i]));
    return v;
}
input::input(std::string name, std::vector<std::string> names, std::vector<expr> sizes, primitive_t t)
{
    construct(name, named_iters(names, sizes), expr(), false, t);
    is_input_ = true;
}
static std::vector<var> sized_iters(const std::vector<expr> &sizes)
{
    std::vector<var> v;
    for (size_t i = 0; i < sizes.size(); i++) v.push_back(var("_d" + std::to_string(i), expr((int32_t) 0), sizes[i]));
    return v;
}
Input::Input(std::string name, std::vector<expr> sizes, primitive_t t) : input(name, sized_iters(sizes), t) {}

constant::constant(std::string name, const expr &e, primitive_t t, bool function_wide, computation *with_computation, int, function *func)
{
    // not function wide (reference core.h: a let statement at a loop level of `with_computation`): its value is
    // substituted into the statements that name it (lowering: expand_lets), so it may use their iterators
    construct(name, {}, e, false, t, func);
    is_constant_ = true;
    is_let_ = !function_wide;
    if (is_let_) let_host_ = with_computation;
}
constant::constant(std::string name, const expr &e, primitive_t t, function *func)
{
    construct(name, {}, e, false, t, func);
    is_constant_ = true;
}
constant::constant(std::string name, const expr &e)
{
    construct(name, {}, e, false, e.get_data_type());
    is_constant_ = true;
}
constant::operator expr() { return expr::make_var(dtype_, name_); }

// ---- block ----------------------------------------------------------------------------------------
block::block(const std::vector<computation *> ch) : children(ch) {}
#define TB_FWD(call) do { for (computation *child_ : children) child_->call; } while (0)
void block::gpu_tile(var a, var b, int x, int y) { TB_FWD(gpu_tile(a, b, x, y)); }
void block::gpu_tile(var a, var b, int x, int y, var o0, var o1, var i0, var i1) { TB_FWD(gpu_tile(a, b, x, y, o0, o1, i0, i1)); }
void block::gpu_tile(var a, var b, var c, int x, int y, int z) { TB_FWD(gpu_tile(a, b, c, x, y, z)); }
void block::gpu_tile(var a, var b, var c, int x, int y, int z, var o0, var o1, var o2, var i0, var i1, var i2) { TB_FWD(gpu_tile(a, b, c, x, y, z, o0, o1, o2, i0, i1, i2)); }
void block::interchange(var a, var b) { TB_FWD(interchange(a, b)); }
void block::parallelize(var a) { TB_FWD(parallelize(a)); }
void block::shift(var a, int n) { TB_FWD(shift(a, n)); }
void block::split(var a, int x) { TB_FWD(split(a, x)); }
void block::split(var a, int x, var o, var i) { TB_FWD(split(a, x, o, i)); }
void block::tile(var a, var b, int x, int y) { TB_FWD(tile(a, b, x, y)); }
void block::tile(var a, var b, int x, int y, var o0, var o1, var i0, var i1) { TB_FWD(tile(a, b, x, y, o0, o1, i0, i1)); }
void block::tile(var a, var b, var c, int x, int y, int z) { TB_FWD(tile(a, b, c, x, y, z)); }
void block::tile(var a, var b, var c, int x, int y, int z, var o0, var o1, var o2, var i0, var i1, var i2) { TB_FWD(tile(a, b, c, x, y, z, o0, o1, o2, i0, i1, i2)); }
void block::tag_gpu_level(var a, var b) { TB_FWD(tag_gpu_level(a, b)); }
void block::tag_gpu_level(var a, var b, var c, var d) { TB_FWD(tag_gpu_level(a, b, c, d)); }
void block::unroll(var a, int f) { TB_FWD(unroll(a, f)); }
void block::unroll(var a, int f, var o, var i) { TB_FWD(unroll(a, f, o, i)); }
void block::vectorize(var a, int v) { TB_FWD(vectorize(a, v)); }
void block::vectorize(var a, int v, var o, var i) { TB_FWD(vectorize(a, v, o, i)); }
#undef TB_FWD

}  // namespace tiramisu
