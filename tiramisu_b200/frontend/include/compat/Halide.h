/* Halide.h — generator-side stand-in: the reference's old-style generators include Halide.h although the
 * functions they call are tiramisu's; this front-end has no Halide dependency (the wrappers' Halide::Buffer
 * stand-in lives in tests/compat/include).  The one Halide type some generators do touch is Halide::Buffer<T>, and
 * only to ask an input image for its extents (benchmarks/halide/blurxy_tiramisu.cpp:25-28): a shape-only buffer. */
#ifndef TIRAMISU_B200_FRONTEND_HALIDE_STANDIN_H
#define TIRAMISU_B200_FRONTEND_HALIDE_STANDIN_H
namespace Halide {
template <typename T> class Buffer {
public:
    Buffer() = default;
    Buffer(int e0, int e1 = 0, int e2 = 0, int e3 = 0) : ext_{e0, e1, e2, e3}, dims_(e3 ? 4 : (e2 ? 3 : (e1 ? 2 : 1))) {}
    int dimensions() const { return dims_; }
    int extent(int d) const { return d >= 0 && d < dims_ ? ext_[d] : 1; }
    int width() const { return extent(0); }
    int height() const { return extent(1); }
    int channels() const { return extent(2); }
private:
    int ext_[4] = {0, 0, 0, 0};
    int dims_ = 0;
};
}  // namespace Halide
#endif
