// tiramisu/block.h — a group of computations that receive the same scheduling commands
// (reference: include/tiramisu/block.h:15-100).
#ifndef TIRAMISU_B200_FRONTEND_BLOCK_H
#define TIRAMISU_B200_FRONTEND_BLOCK_H

#include <tiramisu/core.h>

namespace tiramisu {

class block : public computation {
public:
    block(const std::vector<computation *> children);

    void gpu_tile(var L0, var L1, int sizeX, int sizeY) override;
    void gpu_tile(var L0, var L1, int sizeX, int sizeY, var L0_outer, var L1_outer, var L0_inner, var L1_inner) override;
    void gpu_tile(var L0, var L1, var L2, int sizeX, int sizeY, int sizeZ) override;
    void gpu_tile(var L0, var L1, var L2, int sizeX, int sizeY, int sizeZ,
                  var L0_outer, var L1_outer, var L2_outer, var L0_inner, var L1_inner, var L2_inner) override;
    void interchange(var L0, var L1) override;
    void parallelize(var L) override;
    void shift(var L0, int n) override;
    void split(var L0, int sizeX) override;
    void split(var L0, int sizeX, var L0_outer, var L0_inner) override;
    void tile(var L0, var L1, int sizeX, int sizeY) override;
    void tile(var L0, var L1, int sizeX, int sizeY, var L0_outer, var L1_outer, var L0_inner, var L1_inner) override;
    void tile(var L0, var L1, var L2, int sizeX, int sizeY, int sizeZ) override;
    void tile(var L0, var L1, var L2, int sizeX, int sizeY, int sizeZ,
              var L0_outer, var L1_outer, var L2_outer, var L0_inner, var L1_inner, var L2_inner) override;
    void tag_gpu_level(var L0, var L1) override;
    void tag_gpu_level(var L0, var L1, var L2, var L3) override;
    void unroll(var L, int fac) override;
    void unroll(var L, int fac, var L_outer, var L_inner) override;
    void vectorize(var L, int v) override;
    void vectorize(var L, int v, var L_outer, var L_inner) override;

private:
    const std::vector<computation *> children;
};

}  // namespace tiramisu

#endif
