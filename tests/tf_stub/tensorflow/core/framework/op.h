// Minimal stand-in for the TensorFlow op-registration API: ONLY what kge-keras_b200/tf_shim/kgeb_tf_ops.cc touches, so that the shim's
// calls into include/kgeb200.h can be type-checked in an image without TensorFlow (tests/test_cpu_oracle_and_host.py compiles the
// shim against this directory with -fsyntax-only).  Nothing here executes; it is test infrastructure, not TensorFlow.
#pragma once
#include <cstdint>
#include <functional>
#include <initializer_list>
#include <string>
#include <vector>

namespace tensorflow {
typedef int32_t int32;
typedef uint8_t uint8;
typedef int64_t int64;
enum DataType { DT_FLOAT = 1, DT_INT32 = 3, DT_UINT8 = 4, DT_INT64 = 9 };
class Status {
 public:
  Status() {}
  bool ok() const { return true; }
};
namespace errors {
template <typename... A>
Status Internal(A...) { return Status(); }
template <typename... A>
Status InvalidArgument(A...) { return Status(); }
}  // namespace errors
namespace shape_inference {
struct ShapeHandle {};
struct DimensionHandle {};
class InferenceContext {
 public:
  ShapeHandle input(int) { return {}; }
  DimensionHandle Dim(ShapeHandle, int) { return {}; }
  ShapeHandle Vector(DimensionHandle) { return {}; }
  ShapeHandle Scalar() { return {}; }
  ShapeHandle Matrix(DimensionHandle, DimensionHandle) { return {}; }
  DimensionHandle UnknownDim() { return {}; }
  void set_output(int, ShapeHandle) {}
};
}  // namespace shape_inference
struct OpDefBuilderStub {
  OpDefBuilderStub& Input(const char*) { return *this; }
  OpDefBuilderStub& Output(const char*) { return *this; }
  OpDefBuilderStub& Attr(const char*) { return *this; }
  OpDefBuilderStub& SetShapeFn(std::function<Status(shape_inference::InferenceContext*)>) { return *this; }
};
}  // namespace tensorflow
#define KGEB_STUB_CAT2(a, b) a##b
#define KGEB_STUB_CAT(a, b) KGEB_STUB_CAT2(a, b)
#define REGISTER_OP(name) static ::tensorflow::OpDefBuilderStub KGEB_STUB_CAT(kgeb_stub_op_, __COUNTER__) = ::tensorflow::OpDefBuilderStub()
