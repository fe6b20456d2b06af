// kgeb_tf_ops.cc -- TensorFlow custom-op shim over libkgeb200 (SOURCE ONLY: TensorFlow is not installable in the build
// image, so this file has never been compiled or run here; it documents the op surface of INTEGRATION.md section 2).
//
// KgebScore(E, R, idx)            -> scores           replaces EmbeddingModel.call + func (KGEkeras/models.py:134-237)
// KgebScoreGrad(E, R, idx, ds)    -> (gE, gR)         registered gradient of KgebScore (dense buffers == IndexedSlices summed)
#include <cuda_runtime.h>

#include "kgeb200.h"
#include "tensorflow/core/framework/op.h"
#include "tensorflow/core/framework/op_kernel.h"
#include "tensorflow/core/framework/shape_inference.h"
#include "tensorflow/core/platform/stream_executor.h"

using namespace tensorflow;

REGISTER_OP("KgebScore")
    .Input("entities: float").Input("relations: float").Input("idx: int32")
    .Attr("score_kind: int").Attr("gamma: float = 12.0")
    .Output("scores: float")
    .SetShapeFn([](shape_inference::InferenceContext* c) { c->set_output(0, c->Vector(c->Dim(c->input(2), 0))); return Status(); });

REGISTER_OP("KgebScoreGrad")
    .Input("entities: float").Input("relations: float").Input("idx: int32").Input("dscore: float")
    .Attr("score_kind: int").Attr("gamma: float = 12.0")
    .Output("g_entities: float").Output("g_relations: float")
    .SetShapeFn([](shape_inference::InferenceContext* c) { c->set_output(0, c->input(0)); c->set_output(1, c->input(1)); return Status(); });

namespace {
cudaStream_t StreamOf(OpKernelContext* ctx) {
  return reinterpret_cast<cudaStream_t>(ctx->op_device_context()->stream()->platform_specific_handle().stream);
}
struct Ctx {
  kgeb_ctx* h = nullptr;
  Status Ensure(const Tensor& E, const Tensor& R, int kind, float gamma) {
    if (h) return Status();
    int dev = 0;
    cudaGetDevice(&dev);
    if (kgeb_create(&h, dev, E.dim_size(0), R.dim_size(0), E.dim_size(1), R.dim_size(1), E.dim_size(1), R.dim_size(1), kind, gamma))
      return errors::Internal("kgeb_create: ", kgeb_last_error(nullptr));
    return Status();
  }
  ~Ctx() { if (h) kgeb_destroy(h); }
};
}  // namespace

class KgebScoreOp : public OpKernel {
 public:
  explicit KgebScoreOp(OpKernelConstruction* c) : OpKernel(c) {
    OP_REQUIRES_OK(c, c->GetAttr("score_kind", &kind_));
    OP_REQUIRES_OK(c, c->GetAttr("gamma", &gamma_));
  }
  void Compute(OpKernelContext* c) override {
    const Tensor &E = c->input(0), &R = c->input(1), &idx = c->input(2);
    OP_REQUIRES_OK(c, ctx_.Ensure(E, R, kind_, gamma_));
    Tensor* out = nullptr;
    OP_REQUIRES_OK(c, c->allocate_output(0, TensorShape({idx.dim_size(0)}), &out));
    const int rc = kgeb_score_fwd(ctx_.h, E.flat<float>().data(), R.flat<float>().data(), idx.flat<int32>().data(),
                                  idx.dim_size(0), out->flat<float>().data(), StreamOf(c));
    OP_REQUIRES(c, rc == 0, errors::Internal("kgeb_score_fwd: ", kgeb_last_error(ctx_.h)));
  }
 private:
  int kind_; float gamma_; Ctx ctx_;
};

class KgebScoreGradOp : public OpKernel {
 public:
  explicit KgebScoreGradOp(OpKernelConstruction* c) : OpKernel(c) {
    OP_REQUIRES_OK(c, c->GetAttr("score_kind", &kind_));
    OP_REQUIRES_OK(c, c->GetAttr("gamma", &gamma_));
  }
  void Compute(OpKernelContext* c) override {
    const Tensor &E = c->input(0), &R = c->input(1), &idx = c->input(2), &ds = c->input(3);
    OP_REQUIRES_OK(c, ctx_.Ensure(E, R, kind_, gamma_));
    Tensor *gE = nullptr, *gR = nullptr, tE, tR;
    OP_REQUIRES_OK(c, c->allocate_output(0, E.shape(), &gE));
    OP_REQUIRES_OK(c, c->allocate_output(1, R.shape(), &gR));
    OP_REQUIRES_OK(c, c->allocate_temp(DT_UINT8, TensorShape({E.dim_size(0)}), &tE));
    OP_REQUIRES_OK(c, c->allocate_temp(DT_UINT8, TensorShape({R.dim_size(0)}), &tR));
    cudaStream_t st = StreamOf(c);
    cudaMemsetAsync(gE->flat<float>().data(), 0, gE->TotalBytes(), st);
    cudaMemsetAsync(gR->flat<float>().data(), 0, gR->TotalBytes(), st);
    const int rc = kgeb_score_bwd(ctx_.h, E.flat<float>().data(), R.flat<float>().data(), idx.flat<int32>().data(), idx.dim_size(0),
                                  ds.flat<float>().data(), gE->flat<float>().data(), gR->flat<float>().data(),
                                  tE.flat<uint8>().data(), tR.flat<uint8>().data(), st);
    OP_REQUIRES(c, rc == 0, errors::Internal("kgeb_score_bwd: ", kgeb_last_error(ctx_.h)));
  }
 private:
  int kind_; float gamma_; Ctx ctx_;
};

REGISTER_KERNEL_BUILDER(Name("KgebScore").Device(DEVICE_GPU), KgebScoreOp);
REGISTER_KERNEL_BUILDER(Name("KgebScoreGrad").Device(DEVICE_GPU), KgebScoreGradOp);
