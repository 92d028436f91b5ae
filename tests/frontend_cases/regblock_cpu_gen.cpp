// regblock_cpu_gen — TEST INFRASTRUCTURE, written for this repo.
// A CPU schedule with the register-blocking idioms of the reference's DNN generators (conv_layer_generator_tiramisu.cpp,
// vgg_block_generator_tiramisu.cpp) at a size a test can afford, without needing the reference checkout:
//   * conv(n, y, x, f) = bias(f) + sum_k w(k, f) * in(n, y, x + k), accumulated through a 3 x 8 staging buffer that is loaded
//     from `out`, accumulated over k and stored back (load / accumulate / store triple), x register-blocked by 3 and tagged
//     for unrolling, fused inside the loops that walk the output;
//   * pooled(n, y/2, x/2, f) = max(0, conv over the 2 x 2 window), the conv values consumed out of a second staging buffer
//     (init / accumulate / consume), a max-accumulation over y and x.
// Expected lowering: every statement one kernel with one launch, no host loop; fp32 in the statement's own order (bit-exact).
//   args: in[N][Y][X + K - 1], w[K][8], bias[8], out[N][Y][X][8] (conv, output), pooled[N][Y/2][X/2][8] (output)
#include <tiramisu/tiramisu.h>

using namespace tiramisu;

#define N_ 3
#define Y_ 6
#define X_ 12
#define K_ 5
#define XB 3

int main(int argc, char **argv)
{
    const std::string name = argc > 1 ? argv[1] : "regblock_cpu_gen";
    tiramisu::init(name);

    var n("n", 0, N_), y("y", 0, Y_), x("x", 0, X_), f("f", 0, 8), k("k", 0, K_), xp("xp", 0, X_ + K_ - 1);
    var oy("oy", 0, Y_ / 2), ox("ox", 0, X_ / 2);
    input in("in", {n, y, xp}, p_float32);
    input w("w", {k, f}, p_float32);
    input bias("bias", {f}, p_float32);

    // ---- first half: load / accumulate / store
    computation conv_init("conv_init", {n, y, x, f}, bias(f));
    view acc_v("acc_v", {n, y, x, f}, p_float32);
    computation reg_load("reg_load", {n, y, x, f}, conv_init(n, y, x, f));
    computation conv("conv", {n, y, x, k, f}, acc_v(n, y, x, f) + w(k, f) * in(n, y, x + k));
    computation reg_store("reg_store", {n, y, x, f}, conv(n, y, x, 0, f));
    // ---- second half: init / accumulate / consume into a max-pool
    computation pool_init("pool_init", {n, oy, ox, f}, expr(0.f));
    view pooled_v("pooled_v", {n, y, x, f}, p_float32);
    computation c2_init("c2_init", {n, y, x, f}, bias(f));
    computation c2("c2", {n, y, x, k, f}, c2_init(n, y, x, f) + w(k, f) * in(n, y, x + k));
    computation pool("pool", {n, y, x, f}, expr(o_max, pooled_v(n, y, x, f), c2(n, y, x, 0, f)));

    var xb("xb"), xx("xx");
    for (computation *c : std::vector<computation *>{&reg_load, &conv, &reg_store, &c2_init, &c2, &pool}) c->split(x, XB, xb, xx);
    conv.interchange(xx, k);  conv.interchange(xx, f);      // n y xb k f xx
    c2.interchange(xx, k);    c2.interchange(xx, f);
    for (computation *c : std::vector<computation *>{&reg_load, &reg_store, &c2_init, &pool}) c->interchange(xx, f);
    for (computation *c : std::vector<computation *>{&reg_load, &conv, &reg_store, &c2_init, &c2, &pool}) { c->vectorize(f, 8); c->tag_unroll_level(xx); }
    conv.tag_parallel_level(n);
    conv_init.then(reg_load, y).then(conv, xb).then(reg_store, xb)
             .then(pool_init, computation::root).then(c2_init, n).then(c2, xb).then(pool, xb);

    buffer b_out("b_out", {N_, Y_, X_, 8}, p_float32, a_output);
    buffer b_pool("b_pool", {N_, Y_ / 2, X_ / 2, 8}, p_float32, a_output);
    buffer reg1("reg1", {XB, 8}, p_float32, a_temporary);
    buffer reg2("reg2", {XB, 8}, p_float32, a_temporary);
    conv_init.store_in(&b_out);
    acc_v.store_in(&reg1, {x % XB, f});
    reg_load.store_in(&reg1, {x % XB, f});
    conv.store_in(&reg1, {x % XB, f});
    reg_store.store_in(&b_out, {n, y, x, f});
    pool_init.store_in(&b_pool);
    pooled_v.store_in(&b_pool, {n, y / 2, x / 2, f});
    c2_init.store_in(&reg2, {x % XB, f});
    c2.store_in(&reg2, {x % XB, f});
    pool.store_in(&b_pool, {n, y / 2, x / 2, f});

    tiramisu::codegen({in.get_buffer(), w.get_buffer(), bias.get_buffer(), &b_out, &b_pool}, name + ".o");
    return 0;
}
