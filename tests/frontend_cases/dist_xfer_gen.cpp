// Explicit transfers (create_xfer) between <ranks> ranks, float32 — the shapes of the reference's tests/test_99.cpp
// (ring shift with blocking sends), test_101.cpp (fan-out from rank 0, distributed computation, fan-in) and test_102.cpp
// (collapse_many), with the number of ranks taken from the generator's argument (`dist_xfer_gen_<ranks>`):
//   1. fan-out  : rank 0 holds data[ranks * R][C]; rank q receives rows [q R, (q+1) R) into its local rows [0, R)
//   2. compute  : every rank scales its R local rows: out = data * 3 + 1          (distributed, rank-local indices)
//   3. ring     : rank q sends its first result row to rank q+1 (the last one to rank 0), received into row R of `out`
//   4. fan-in   : rank 0 gathers every rank's R result rows into gathered[q R + x]
// Messages are element-wise in the program (as in the reference's tests); the lowering coalesces them.
#include <tiramisu/tiramisu.h>

#include <cstdlib>
#include <string>

using namespace tiramisu;

int main(int argc, char **argv)
{
    const std::string name = argc > 1 ? argv[1] : "dist_xfer_gen_2";
    const int ranks = std::atoi(name.substr(name.rfind('_') + 1).c_str());
    const int R = 40, C = 96;
    const std::string sR = std::to_string(R), sC = std::to_string(C), sK = std::to_string(ranks), sK1 = std::to_string(ranks - 1);
    global::set_default_tiramisu_options();
    function f(name);
    var x("x"), y("y"), q("q"), p("p"), x1("x1"), x2("x2");
    constant ONE("ONE", 1, p_int32, true, nullptr, 0, &f);
    constant LAST("LAST", ranks - 1, p_int32, true, nullptr, 0, &f);
    computation data("{data[x,y]: 0<=x<" + std::to_string(ranks * R) + " and 0<=y<" + sC + "}", expr(), false, p_float32, &f);
    computation scaled("{scaled[x,y]: 0<=x<" + std::to_string(ranks * R) + " and 0<=y<" + sC + "}", data(x, y) * 3.0f + 1.0f, true, p_float32, &f);
    xfer fan_out = computation::create_xfer("[ONE]->{fan_out_s[p,q,x,y]: 0<=p<ONE and 1<=q<" + sK + " and 0<=x<" + sR + " and 0<=y<" + sC + "}",
                                            "{fan_out_r[q,x,y]: 1<=q<" + sK + " and 0<=x<" + sR + " and 0<=y<" + sC + "}", q, 0,
                                            xfer_prop(p_float32, {MPI, BLOCK, ASYNC}), xfer_prop(p_float32, {MPI, BLOCK, ASYNC}),
                                            data(q * R + x, y), &f);
    xfer ring = computation::create_xfer("{ring_s[q,y]: 0<=q<" + sK1 + " and 0<=y<" + sC + "}", "{ring_r[q,y]: 1<=q<" + sK + " and 0<=y<" + sC + "}",
                                         q + 1, q - 1, xfer_prop(p_float32, {MPI, BLOCK, SYNC}), xfer_prop(p_float32, {MPI, BLOCK, SYNC}),
                                         scaled(0, y), &f);
    xfer wrap = computation::create_xfer("[LAST]->{wrap_s[q,y]: LAST<=q<" + sK + " and 0<=y<" + sC + "}", "[ONE]->{wrap_r[q,y]: 0<=q<ONE and 0<=y<" + sC + "}",
                                         0, ranks - 1, xfer_prop(p_float32, {MPI, BLOCK, SYNC}), xfer_prop(p_float32, {MPI, BLOCK, SYNC}),
                                         scaled(0, y), &f);
    xfer fan_in = computation::create_xfer("{fan_in_s[q,x,y]: 1<=q<" + sK + " and 0<=x<" + sR + " and 0<=y<" + sC + "}",
                                           "[ONE]->{fan_in_r[p,q,x,y]: 0<=p<ONE and 1<=q<" + sK + " and 0<=x<" + sR + " and 0<=y<" + sC + "}", 0, q,
                                           xfer_prop(p_float32, {MPI, BLOCK, ASYNC}), xfer_prop(p_float32, {MPI, BLOCK, ASYNC}),
                                           scaled(x, y), &f);
    // the fan-in is collapsed by hand (test_102), the rest is left to the lowering
    fan_in.s->collapse_many({collapse_group(2, 0, -1, C), collapse_group(1, 0, -1, R)});
    fan_in.r->collapse_many({collapse_group(3, 0, -1, C), collapse_group(2, 0, -1, R)});
    data.split(x, R, x1, x2);
    scaled.split(x, R, x1, x2);
    data.tag_distribute_level(x1);
    scaled.tag_distribute_level(x1);
    scaled.drop_rank_iter(x1);
    fan_out.s->tag_distribute_level(p);
    fan_out.r->tag_distribute_level(q);
    ring.s->tag_distribute_level(q);
    ring.r->tag_distribute_level(q);
    wrap.s->tag_distribute_level(q);
    wrap.r->tag_distribute_level(q);
    fan_in.s->tag_distribute_level(q);
    fan_in.r->tag_distribute_level(p);
    fan_out.s->before(*fan_out.r, computation::root);
    fan_out.r->before(scaled, computation::root);
    scaled.before(*ring.s, computation::root);
    ring.s->before(*ring.r, computation::root);
    ring.r->before(*wrap.s, computation::root);
    wrap.s->before(*wrap.r, computation::root);
    wrap.r->before(*fan_in.s, computation::root);
    fan_in.s->before(*fan_in.r, computation::root);
    // rank 0 holds everything, the others R rows (the receive buffer of the ring is row R of `out`)
    buffer b_data("b_data", {expr(o_select, var("rank") == 0, ranks * R, R), C}, p_float32, a_input, &f);
    buffer b_out("b_out", {R + 1, C}, p_float32, a_output, &f);
    buffer b_all("b_all", {expr(o_select, var("rank") == 0, ranks * R, 1), C}, p_float32, a_output, &f);
    data.set_access("{data[x,y]->b_data[x,y]}");
    scaled.set_access("{scaled[x,y]->b_out[x,y]}");
    fan_out.r->set_access("{fan_out_r[q,x,y]->b_data[x,y]}");
    ring.r->set_access("{ring_r[q,y]->b_out[" + sR + ",y]}");
    wrap.r->set_access("{wrap_r[q,y]->b_out[" + sR + ",y]}");
    fan_in.r->set_access("{fan_in_r[p,q,x,y]->b_all[q*" + sR + "+x,y]}");
    f.codegen({&b_data, &b_out, &b_all}, name + ".o");
    return 0;
}
