// heat2d through the front-end + sm_100a lowering (TEST INFRASTRUCTURE, written for this repo).
// Same statement as the reference's benchmarks/halide/heat2d_tiramisu.cpp:26-27,52-69
// (out = 0 everywhere; interior: alpha*in + beta*(((E + W) + S) + N)), expressed in the modern API
// with run-time sizes, a non-zero lower bound and a 2-D gpu_tile whose tiles do not divide the grid.
#include <tiramisu/tiramisu.h>

using namespace tiramisu;

int main(int argc, char **argv)
{
    tiramisu::init(argc > 1 ? argv[1] : "heat2d_gen");

    input sizes("sizes", {"s_i"}, {2}, p_int32);
    constant H("H", sizes(0));
    constant W("W", sizes(1));

    var y("y", 0, H), x("x", 0, W);
    var yi("yi", 1, H - 1), xi("xi", 1, W - 1);
    var y0("y0"), x0("x0"), y1("y1"), x1("x1");

    input in("in", {y, x}, p_float32);
    computation zero("zero", {y, x}, expr(0.f));
    computation heat("heat", {yi, xi},
                     in(yi, xi) * 0.3f + (((in(yi, xi + 1) + in(yi, xi - 1)) + in(yi + 1, xi)) + in(yi - 1, xi)) * 0.4f);

    buffer b_in("b_in", {H, W}, p_float32, a_input);
    buffer b_out("b_out", {H, W}, p_float32, a_output);
    buffer b_in_gpu("b_in_gpu", {H, W}, p_float32, a_temporary);
    buffer b_out_gpu("b_out_gpu", {H, W}, p_float32, a_temporary);
    b_in_gpu.tag_gpu_global();
    b_out_gpu.tag_gpu_global();

    in.store_in(&b_in_gpu);
    zero.store_in(&b_out_gpu);
    heat.store_in(&b_out_gpu, {yi, xi});

    computation copy_in({}, memcpy(b_in, b_in_gpu));
    computation copy_out({}, memcpy(b_out_gpu, b_out));

    zero.gpu_tile(y, x, 8, 32, y0, x0, y1, x1);
    heat.gpu_tile(yi, xi, 8, 32, y0, x0, y1, x1);

    copy_in.then(zero, computation::root)
           .then(heat, computation::root)
           .then(copy_out, computation::root);

    tiramisu::codegen({sizes.get_buffer(), &b_in, &b_out}, std::string(argc > 1 ? argv[1] : "heat2d_gen") + ".o", true);
    return 0;
}
