// Halide.h — TEST INFRASTRUCTURE.  Minimal stand-in for the Halide compiler header that the
// reference's wrapper_*.cpp harnesses include, so that those harnesses compile UNMODIFIED (from
// where they lie under /root/reference) and link against libtiramisu_b200.so.  Only the host-side
// buffer class is needed: Halide::Buffer<T> is the reference checkout's own
// Halide::Runtime::Buffer<T> (3rdParty/Halide/src/runtime/HalideBuffer.h, header-only) plus the
// `(extents..., "name")` constructors and the 3-D accessors the wrappers use.
#pragma once

#include "HalideBuffer.h"

#include <cmath>
#include <string>
#include <vector>

namespace Halide {

inline halide_type_t Float(int bits) { return halide_type_t(halide_type_float, (uint8_t)bits); }
inline halide_type_t Int(int bits) { return halide_type_t(halide_type_int, (uint8_t)bits); }
inline halide_type_t UInt(int bits) { return halide_type_t(halide_type_uint, (uint8_t)bits); }

template <typename T = void>
class Buffer : public Runtime::Buffer<T> {
    using Base = Runtime::Buffer<T>;
    std::string name_;

public:
    Buffer() = default;
    Buffer(const Base &b) : Base(b) {}
    Buffer(Base &&b) : Base(std::move(b)) {}
    explicit Buffer(int a, const std::string &n = "") : Base(a), name_(n) {}
    Buffer(int a, int b, const std::string &n = "") : Base(a, b), name_(n) {}
    Buffer(int a, int b, int c, const std::string &n = "") : Base(a, b, c), name_(n) {}
    Buffer(int a, int b, int c, int d, const std::string &n = "") : Base(a, b, c, d), name_(n) {}
    Buffer(int a, int b, int c, int d, int e, const std::string &n = "") : Base(std::vector<int>{a, b, c, d, e}), name_(n) {}
    Buffer(const std::vector<int> &sizes, const std::string &n = "") : Base(sizes), name_(n) {}
    Buffer(halide_type_t t, int a, const std::string &n = "") : Base(t, a), name_(n) {}
    Buffer(halide_type_t t, int a, int b, const std::string &n = "") : Base(t, a, b), name_(n) {}
    Buffer(halide_type_t t, int a, int b, int c, const std::string &n = "") : Base(t, a, b, c), name_(n) {}
    Buffer(halide_type_t t, int a, int b, int c, int d, const std::string &n = "") : Base(t, a, b, c, d), name_(n) {}
    Buffer(T *data, const std::vector<int> &sizes, const std::string &n = "") : Base(data, sizes), name_(n) {}
    // (type, untyped data pointer, sizes): wrapper_test_93.cpp wraps int arrays in Buffer<int64_t> this way
    template <typename U> Buffer(halide_type_t, U *data, const std::vector<int> &sizes, const std::string &n = "") : Base((T *) data, sizes), name_(n) {}
    Buffer(halide_type_t, const std::vector<int> &sizes, const std::string &n = "") : Base(sizes), name_(n) {}
    Buffer(T *data, int a, const std::string &n = "") : Base(data, a), name_(n) {}
    Buffer(T *data, int a, int b, const std::string &n = "") : Base(data, a, b), name_(n) {}
    Buffer(T *data, int a, int b, int c, const std::string &n = "") : Base(data, a, b, c), name_(n) {}
    Buffer(T *data, int a, int b, int c, int d, const std::string &n = "") : Base(data, a, b, c, d), name_(n) {}

    const std::string &name() const { return name_; }
    Base *get() { return this; }
    const Base *get() const { return this; }
};

}  // namespace Halide
