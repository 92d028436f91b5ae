// Unit checks of the front-end's schedule bookkeeping (TEST INFRASTRUCTURE, written for this repo):
// what split / tile / interchange / gpu_tile / then leave in a computation's loop-dimension list.
// Prints "unit_schedule ok" and exits 0 when every check holds.
#include <tiramisu/tiramisu.h>
#include <tiramisu/block.h>

#include <cstdio>
#include <cstdlib>
#include <string>

using namespace tiramisu;

static int failures = 0;
#define CHECK(cond) do { if (!(cond)) { std::fprintf(stderr, "FAILED %s:%d: %s\n", __FILE__, __LINE__, #cond); failures++; } } while (0)

static std::string names(computation &c)
{
    std::string s;
    for (auto &d : c.dims()) s += (s.empty() ? "" : ",") + d.name;
    return s;
}

int main()
{
    tiramisu::init("unit");
    var i("i", 0, 128), j("j", 0, 100), k("k", 0, 64);
    var i0("i0"), i1("i1"), j0("j0"), j1("j1"), k0("k0"), k1("k1"), i00("i00"), i01("i01"), j00("j00"), j01("j01");

    // expression folding keeps extents decidable
    CHECK((expr(6) * 4 + 2).is_int_literal() && (expr(6) * 4 + 2).get_int_val() == 26);
    CHECK((i + 0).is_equal(i));
    CHECK((i * 1).is_equal(i));
    CHECK((i + 1).get_data_type() == p_int32);
    CHECK((expr(2.5f) * 2).get_data_type() == p_float32);
    CHECK((i < 4).get_data_type() == p_boolean);

    input A({i, k}, p_float32), B({k, j}, p_float32);
    computation C({i, j, k}, p_float32);
    C.set_expression(C(i, j, 0) + A(i, k) * B(k, j));
    CHECK(names(C) == "i,j,k");

    C.tile(i, j, 4, 4, i0, j0, i1, j1);
    CHECK(names(C) == "i0,j0,i1,j1,k");
    CHECK(C.dims()[0].hi.is_int_literal() && C.dims()[0].hi.get_int_val() == 32);
    CHECK(C.dims()[1].hi.is_int_literal() && C.dims()[1].hi.get_int_val() == 25);
    CHECK(C.guards().empty());                       // 128 and 100 are multiples of 4

    C.interchange(j1, k);
    C.interchange(i1, k);
    CHECK(names(C) == "i0,j0,k,i1,j1");
    C.split(k, 32, k0, k1);
    CHECK(names(C) == "i0,j0,k0,k1,i1,j1");

    C.gpu_tile(i0, j0, 16, 16, i00, j00, i01, j01);
    CHECK(names(C) == "i00,j00,i01,j01,k0,k1,i1,j1");
    CHECK(C.dims()[0].tag == loop_dim::gpu_block && C.dims()[1].tag == loop_dim::gpu_block);
    CHECK(C.dims()[2].tag == loop_dim::gpu_thread && C.dims()[3].tag == loop_dim::gpu_thread);
    CHECK(C.dims()[1].hi.get_int_val() == 2);        // ceil(25 / 16)
    CHECK(C.guards().size() == 1);                   // j0 < 25 (25 is not a multiple of 16); 32 / 16 is exact

    // data mapping
    C.store_in({i, j}, {128, 100});
    CHECK(C.get_buffer()->get_n_dims() == 2 && C.access().size() == 2);
    CHECK(A.get_buffer()->get_argument_type() == a_input);

    // ordering
    computation init({i, j}, expr(0.f));
    computation fin({i, j}, C(i, j, 0) * 2.f);
    init.then(C, computation::root).then(fin, computation::root);
    CHECK(C.pred() == &init && fin.pred() == &C && init.succ() == &C);

    // a block forwards its commands to every member
    computation P({i, j}, expr(1.f)), Q({i, j}, expr(2.f));
    block blk({&P, &Q});
    blk.gpu_tile(i, j, 8, 8, i0, j0, i1, j1);
    CHECK(names(P) == "i0,j0,i1,j1" && names(Q) == "i0,j0,i1,j1");
    CHECK(Q.dims()[3].tag == loop_dim::gpu_thread);

    // allocate_at takes the loop prefix of its anchor
    var t("t", 0, 3);
    buffer tmp("tmp", {16}, p_int32, a_temporary);
    computation body({t, var("u", 0, 16)}, expr(1));
    computation *al = tmp.allocate_at(body, t);
    CHECK(al->dims().size() == 1 && al->dims()[0].name == "t" && !tmp.get_auto_allocate());

    // old-style declarations: ISL box strings
    {
        tiramisu::function f("isl_unit");
        constant N("N", expr((int32_t) 10), p_int32, true, nullptr, 0, &f);
        computation S("[N]->{S[i,j]: 0<=i<N and 2<=j<=(N + -1)}", expr((uint8_t) 1), true, p_uint8, &f);
        CHECK(names(S) == "i,j");
        CHECK(S.dims()[0].lo.is_int_literal() && S.dims()[0].lo.get_int_val() == 0);
        CHECK(S.dims()[0].hi.get_expr_type() == e_var && S.dims()[0].hi.get_name() == "N");
        CHECK(S.dims()[1].lo.get_int_val() == 2);
        CHECK(S.dims()[1].hi.get_expr_type() == e_var && S.dims()[1].hi.get_name() == "N");     // (N + -1) + 1 folds back to N
        computation T("{T[i,0,k]: (0 <= i <= 7) and (1 <= k < 4)}", expr((uint8_t) 2), true, p_uint8, &f);
        CHECK(T.dims().size() == 3 && T.dims()[1].lo.get_int_val() == 0 && T.dims()[1].hi.get_int_val() == 1);   // pinned dimension
        CHECK(T.dims()[0].hi.get_int_val() == 8 && T.dims()[2].lo.get_int_val() == 1 && T.dims()[2].hi.get_int_val() == 4);
        computation In("[N]->{In[a,b]}", expr(), false, p_uint8, &f);                                // a declaration only
        CHECK(In.is_input());
        buffer buf("buf", {expr(10), expr(12)}, p_uint8, a_output, &f);
        S.set_access("[N]->{S[a,b]->buf[a + 1, min(b, 9)]}");                                         // positional aliases, clamp
        CHECK(S.access().size() == 2 && S.get_buffer() == &buf);
        CHECK(S.access()[0].get_op_type() == o_add && S.access()[0].get_operand(0).get_name() == "i");
        CHECK(S.access()[1].get_op_type() == o_min && S.access()[1].get_operand(0).get_name() == "j");
    }

    if (failures) return 1;
    std::printf("unit_schedule ok\n");
    return 0;
}
