// see op.h in this directory: a type-check stand-in, not TensorFlow
#pragma once
#include "tensorflow/core/framework/op.h"

namespace tensorflow {
class TensorShape {
 public:
  TensorShape() {}
  TensorShape(std::initializer_list<int64_t>) {}
};
template <typename T>
struct FlatStub {
  T* data() const { return nullptr; }
  T& operator()(int64_t) const { static T t; return t; }
};
class Tensor {
 public:
  int64_t dim_size(int) const { return 0; }
  int dims() const { return 0; }
  const TensorShape& shape() const { static TensorShape s; return s; }
  size_t TotalBytes() const { return 0; }
  template <typename T>
  FlatStub<T> flat() const { return {}; }
  template <typename T>
  FlatStub<T> scalar() const { return {}; }
};
struct StreamHandleStub { void* stream; };
struct StreamStub { StreamHandleStub platform_specific_handle() const { return {nullptr}; } };
struct DeviceContextStub { StreamStub* stream() const { return nullptr; } };
class OpKernelConstruction {
 public:
  template <typename T>
  Status GetAttr(const char*, T*) { return Status(); }
};
class OpKernelContext {
 public:
  const Tensor& input(int) { static Tensor t; return t; }
  Status allocate_output(int, const TensorShape&, Tensor**) { return Status(); }
  Status allocate_temp(DataType, const TensorShape&, Tensor*) { return Status(); }
  DeviceContextStub* op_device_context() { return nullptr; }
  void CtxFailure(const Status&) {}
};
class OpKernel {
 public:
  explicit OpKernel(OpKernelConstruction*) {}
  virtual ~OpKernel() {}
  virtual void Compute(OpKernelContext*) = 0;
};
struct KernelDefBuilderStub {
  KernelDefBuilderStub& Device(const char*) { return *this; }
  KernelDefBuilderStub& HostMemory(const char*) { return *this; }
};
inline KernelDefBuilderStub Name(const char*) { return {}; }
static const char* const DEVICE_GPU = "GPU";
}  // namespace tensorflow
#define OP_REQUIRES_OK(ctx, expr) do { ::tensorflow::Status s_ = (expr); if (!s_.ok()) return; } while (0)
#define OP_REQUIRES(ctx, cond, status) do { if (!(cond)) { (void)(status); return; } } while (0)
#define REGISTER_KERNEL_BUILDER(builder, cls) static auto KGEB_STUB_CAT(kgeb_stub_kernel_, __COUNTER__) = (builder); static_assert(sizeof(cls) > 0, "")
