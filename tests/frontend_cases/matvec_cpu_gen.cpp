// y = A x + y0 with a CPU schedule (TEST INFRASTRUCTURE, written for this repo): the reduction iterator j is the
// unit-stride index of both operands, so the pattern lowering maps one warp per output row and combines the
// lanes' partial sums with shuffles (tiramisu_b200/frontend/src/lowering_retarget.cpp: warp_reduction_terms);
// TIRAMISU_B200_LOWERING=generic keeps one row per thread and the statement's own summation order.
#include <tiramisu/tiramisu.h>

using namespace tiramisu;

int main(int argc, char **argv)
{
    const std::string name = argc > 1 ? argv[1] : "matvec_cpu_gen";
    tiramisu::init(name);

    input sizes("sizes", {"s_i"}, {2}, p_int32);
    constant M("M", sizes(0)), N("N", sizes(1));
    var i("i", 0, M), j("j", 0, N);

    input A("A", {i, j}, p_float32);
    input x("x", {j}, p_float32);
    input y0("y0", {i}, p_float32);
    computation y_init("y_init", {i}, y0(i));
    computation y("y", {i, j}, p_float32);
    y.set_expression(y(i, j - 1) + A(i, j) * x(j));

    buffer b_A("b_A", {M, N}, p_float32, a_input);
    buffer b_x("b_x", {N}, p_float32, a_input);
    buffer b_y0("b_y0", {M}, p_float32, a_input);
    buffer b_y("b_y", {M}, p_float32, a_output);
    A.store_in(&b_A);
    x.store_in(&b_x);
    y0.store_in(&b_y0);
    y_init.store_in(&b_y);
    y.store_in(&b_y, {i});

    y_init.parallelize(i);
    y.parallelize(i);
    y_init.then(y, computation::root);

    tiramisu::codegen({sizes.get_buffer(), &b_A, &b_x, &b_y0, &b_y}, name + ".o");
    return 0;
}
