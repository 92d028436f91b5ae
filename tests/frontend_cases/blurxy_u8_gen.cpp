// uint8 blur through the front-end + sm_100a lowering (TEST INFRASTRUCTURE, written for this repo).
// Same arithmetic as the reference's benchmarks/halide/blurxy_tiramisu.cpp:61,77: every sum is a
// uint8 sum (wraps mod 256) before the division by 3; output is (H-8) x (W-8) per channel.
#include <tiramisu/tiramisu.h>

using namespace tiramisu;

int main(int argc, char **argv)
{
    tiramisu::init("blurxy_u8_gen");

    input sizes("sizes", {"s_i"}, {2}, p_int32);
    constant H("H", sizes(0));
    constant W("W", sizes(1));

    var c("c", 0, 3), y("y", 0, H), x("x", 0, W);
    var yo("yo", 0, H - 8), xo("xo", 0, W - 8);
    var y0("y0"), x0("x0"), y1("y1"), x1("x1");

    input in("in", {c, y, x}, p_uint8);
    auto bx = [&](expr yy) { return ((in(c, yy, xo) + in(c, yy, xo + 1)) + in(c, yy, xo + 2)) / (uint8_t) 3; };
    computation by("by", {c, yo, xo}, ((bx(yo) + bx(yo + 1)) + bx(yo + 2)) / (uint8_t) 3);

    buffer b_in("b_in", {3, H, W}, p_uint8, a_input);
    buffer b_out("b_out", {3, H - 8, W - 8}, p_uint8, a_output);
    buffer b_in_gpu("b_in_gpu", {3, H, W}, p_uint8, a_temporary);
    buffer b_out_gpu("b_out_gpu", {3, H - 8, W - 8}, p_uint8, a_temporary);
    b_in_gpu.tag_gpu_global();
    b_out_gpu.tag_gpu_global();
    in.store_in(&b_in_gpu);
    by.store_in(&b_out_gpu);

    computation copy_in({}, memcpy(b_in, b_in_gpu));
    computation copy_out({}, memcpy(b_out_gpu, b_out));

    by.gpu_tile(yo, xo, 4, 64, y0, x0, y1, x1);

    copy_in.then(by, computation::root).then(copy_out, computation::root);

    tiramisu::codegen({sizes.get_buffer(), &b_in, &b_out}, "blurxy_u8_gen.o", true);
    return 0;
}
