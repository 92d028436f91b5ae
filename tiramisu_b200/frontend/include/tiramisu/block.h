// tiramisu/block.h — a group of computations that receive the same scheduling commands
// (reference: include/tiramisu/block.h:15-100).
#ifndef TIRAMISU_B200_FRONTEND_BLOCK_H
#define TIRAMISU_B200_FRONTEND_BLOCK_H

#include <tiramisu/core.h>

namespace tiramisu {

class block : public computation {
public:
    block(const std::vector<computation *> children);

    void gpu_tile(var a, var b, int sx, int sy) override;
    void gpu_tile(var a, var b, int sx, int sy, var o0, var o1, var i0, var i1) override;
    void gpu_tile(var a, var b, var c, int sx, int sy, int sz) override;
    void gpu_tile(var a, var b, var c, int sx, int sy, int sz,
                  var o0, var o1, var o2, var i0, var i1, var i2) override;
    void interchange(var a, var b) override;
    void parallelize(var lvl) override;
    void shift(var a, int shift_by) override;
    void split(var a, int sx) override;
    void split(var a, int sx, var o0, var i0) override;
    void tile(var a, var b, int sx, int sy) override;
    void tile(var a, var b, int sx, int sy, var o0, var o1, var i0, var i1) override;
    void tile(var a, var b, var c, int sx, int sy, int sz) override;
    void tile(var a, var b, var c, int sx, int sy, int sz,
              var o0, var o1, var o2, var i0, var i1, var i2) override;
    void tag_gpu_level(var a, var b) override;
    void tag_gpu_level(var a, var b, var c, var L3) override;
    void unroll(var lvl, int factor) override;
    void unroll(var lvl, int factor, var outer, var inner) override;
    void vectorize(var lvl, int lanes) override;
    void vectorize(var lvl, int lanes, var outer, var inner) override;

private:
    const std::vector<computation *> children;
};

}  // namespace tiramisu

#endif
