// fp32 3x3 separable blur with a CPU schedule (TEST INFRASTRUCTURE, written for this repo): host buffers,
// parallelize(y) + vectorize(x, 8), codegen WITHOUT the gpu flag.  The back end has no CPU code
// generator, so the lowering re-targets the nest to sm_100a (blocks over y, 128 consecutive x per block,
// device mirrors for the host arguments).  Arithmetic = SURVEY §8 a8' (blurxy_tiramisu.cpp:61,77 in fp32).
#include <tiramisu/tiramisu.h>

using namespace tiramisu;

int main(int argc, char **argv)
{
    tiramisu::init("blur_f32_cpu_gen");

    input sizes("sizes", {"s_i"}, {2}, p_int32);
    constant H("H", sizes(0)), W("W", sizes(1));
    var y("y", 0, H), x("x", 0, W), yo("yo", 0, H - 2), xo("xo", 0, W - 2);

    input in("in", {y, x}, p_float32);
    auto bx = [&](expr yy) { return ((in(yy, xo) + in(yy, xo + 1)) + in(yy, xo + 2)) / 3.0f; };
    computation by("by", {yo, xo}, ((bx(yo) + bx(yo + 1)) + bx(yo + 2)) / 3.0f);

    buffer b_in("b_in", {H, W}, p_float32, a_input);
    buffer b_out("b_out", {H - 2, W - 2}, p_float32, a_output);
    in.store_in(&b_in);
    by.store_in(&b_out);

    by.parallelize(yo);
    by.vectorize(xo, 8);

    tiramisu::codegen({sizes.get_buffer(), &b_in, &b_out}, "blur_f32_cpu_gen.o");
    return 0;
}
