// Distributed 3x3 box blur (uint32) — the shape of the reference's tests/test_160.cpp and
// benchmarks/automatic_comm/blurautodist_tiramisu.cpp, with the number of ranks taken from the generator's argument
// (`dist_blur_gen_<ranks>`) so that it runs one rank per GPU on the boxes the tests get (2, 4 or 8 GPUs; the reference
// hard-codes 10 and 16 ranks).  Rows are distributed; gen_communication() derives the two halo rows every rank needs
// from the next one, enlarges the rank-local input buffer and schedules the send / receive in front of blurx.
#include <tiramisu/tiramisu.h>

#include <cstdlib>
#include <string>

using namespace tiramisu;

int main(int argc, char **argv)
{
    const std::string name = argc > 1 ? argv[1] : "dist_blur_gen_2";
    const int nodes = std::atoi(name.substr(name.rfind('_') + 1).c_str());
    const int rows_per = 48, cols = 1022, rows = rows_per * nodes;
    global::set_default_tiramisu_options();
    function blur(name);
    var i("i"), j("j"), i0("i0"), i1("i1");
    constant ROWS("ROWS", expr((int32_t) rows), p_int32, true, nullptr, 0, &blur);
    constant COLS("COLS", expr((int32_t) cols), p_int32, true, nullptr, 0, &blur);
    computation img("[ROWS,COLS]->{img[i,j]: 0<=i<ROWS and 0<=j<COLS+2}", expr(), false, p_uint32, &blur);
    expr e1 = (img(i, j) + img(i + 1, j) + img(i + 2, j)) / ((uint32_t) 3);
    computation blurx("[ROWS,COLS]->{blurx[i,j]: 0<=i<ROWS and 0<=j<COLS+2}", e1, true, p_uint32, &blur);
    expr e2 = (blurx(i, j) + blurx(i, j + 1) + blurx(i, j + 2)) / ((uint32_t) 3);
    computation blury("[ROWS,COLS]->{blury[i,j]: 0<=i<ROWS and 0<=j<COLS}", e2, true, p_uint32, &blur);
    for (computation *c : {&img, &blurx, &blury}) {
        c->split(i, rows_per, i0, i1);
        c->tag_distribute_level(i0);
        c->drop_rank_iter(i0);
    }
    blurx.before(blury, i0);
    buffer b_img("b_img", {expr(rows_per), expr(cols) + 2}, p_uint32, a_input, &blur);
    buffer b_blurx("b_blurx", {expr(rows_per), expr(cols) + 2}, p_uint32, a_temporary, &blur);
    buffer b_blury("b_blury", {expr(rows_per), expr(cols)}, p_uint32, a_output, &blur);
    img.set_access("{img[i,j]->b_img[i,j]}");
    blurx.set_access("{blurx[i,j]->b_blurx[i,j]}");
    blury.set_access("{blury[i,j]->b_blury[i,j]}");
    blurx.gen_communication();
    blur.codegen({&b_img, &b_blury}, name + ".o");
    return 0;
}
