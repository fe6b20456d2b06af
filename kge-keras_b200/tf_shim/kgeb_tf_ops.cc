// kgeb_tf_ops.cc -- TensorFlow custom-op shim over libkgeb200.  SOURCE ONLY: TensorFlow is not installable in the build image, so
// this file has never been linked or run here.  What IS checked (tests/test_cpu_oracle_and_host.py): it compiles with
// -fsyntax-only against tests/tf_stub/ -- a stand-in for the handful of TensorFlow classes used below -- so every call into
// include/kgeb200.h type-checks against the real header.  INTEGRATION.md section 2 describes the op surface.
//
// KgebScore(E, R, idx)                 -> scores                 EmbeddingModel.call + func (KGEkeras/models.py:134-237)
// KgebScoreGrad(E, R, idx, ds)         -> (gE, gR)               registered gradient of KgebScore, dense buffers
// KgebScoreGradSlices(E, R, idx, ds)   -> (ids, rows, n, gR)     the same as tf.IndexedSlices parts for the entity table (what a
//                                                                 tf.keras Embedding gives its optimizer, models.py:89-112)
// KgebTrainStep(E, R, mE, vE, mR, vR, pos, step) -> loss         one fused step, in place: Philox negatives, score, loss, gradient
//                                                                 scatter, touched-row Adam (tests.py:63-87,130-134 as one op)
// KgebRank1N(E, R, triples, offs, ids) -> counts                 filtered 1-N ranking of one side (utils.py:165-245); exact mode on
//                                                                 the tensor cores for the dot-product kinds, CUDA cores otherwise
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

REGISTER_OP("KgebScoreGradSlices")
    .Input("entities: float").Input("relations: float").Input("idx: int32").Input("dscore: float")
    .Attr("score_kind: int").Attr("gamma: float = 12.0")
    .Output("ids: int32").Output("rows: float").Output("count: int64").Output("g_relations: float")
    .SetShapeFn([](shape_inference::InferenceContext* c) {
      c->set_output(0, c->Vector(c->Dim(c->input(0), 0)));
      c->set_output(1, c->input(0));
      c->set_output(2, c->Scalar());
      c->set_output(3, c->input(1));
      return Status();
    });

REGISTER_OP("KgebTrainStep")
    .Input("entities: float").Input("relations: float").Input("m_entities: float").Input("v_entities: float")
    .Input("m_relations: float").Input("v_relations: float").Input("pos: int32").Input("step: int64")
    .Attr("score_kind: int").Attr("gamma: float = 12.0").Attr("negatives: int = 2").Attr("loss_kind: int = 3").Attr("margin: float = 1.0")
    .Attr("lr: float = 0.001").Attr("beta1: float = 0.9").Attr("beta2: float = 0.999").Attr("epsilon: float = 1e-7").Attr("seed: int = 0")
    .Output("loss: float")
    .SetShapeFn([](shape_inference::InferenceContext* c) { c->set_output(0, c->Scalar()); return Status(); });

REGISTER_OP("KgebRank1N")
    .Input("entities: float").Input("relations: float").Input("triples: int32").Input("filter_offsets: int64").Input("filter_ids: int32")
    .Attr("score_kind: int").Attr("gamma: float = 12.0").Attr("side: int = 0")
    .Output("counts: int32")
    .SetShapeFn([](shape_inference::InferenceContext* c) { c->set_output(0, c->Vector(c->Dim(c->input(2), 0))); return Status(); });

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

// gradient of KgebScore as the parts of a tf.IndexedSlices for the entity table: ids ascending, one per touched row
class KgebScoreGradSlicesOp : public OpKernel {
 public:
  explicit KgebScoreGradSlicesOp(OpKernelConstruction* c) : OpKernel(c) {
    OP_REQUIRES_OK(c, c->GetAttr("score_kind", &kind_));
    OP_REQUIRES_OK(c, c->GetAttr("gamma", &gamma_));
  }
  void Compute(OpKernelContext* c) override {
    const Tensor &E = c->input(0), &R = c->input(1), &idx = c->input(2), &ds = c->input(3);
    OP_REQUIRES_OK(c, ctx_.Ensure(E, R, kind_, gamma_));
    const int64_t N = E.dim_size(0);
    const int W = (int)E.dim_size(1);
    Tensor *ids = nullptr, *rows = nullptr, *count = nullptr, *gR = nullptr, gE, tE, tR;
    OP_REQUIRES_OK(c, c->allocate_output(0, TensorShape({N}), &ids));
    OP_REQUIRES_OK(c, c->allocate_output(1, E.shape(), &rows));
    OP_REQUIRES_OK(c, c->allocate_output(2, TensorShape({}), &count));
    OP_REQUIRES_OK(c, c->allocate_output(3, R.shape(), &gR));
    OP_REQUIRES_OK(c, c->allocate_temp(DT_FLOAT, E.shape(), &gE));
    OP_REQUIRES_OK(c, c->allocate_temp(DT_UINT8, TensorShape({N}), &tE));
    OP_REQUIRES_OK(c, c->allocate_temp(DT_UINT8, TensorShape({R.dim_size(0)}), &tR));
    cudaStream_t st = StreamOf(c);
    cudaMemsetAsync(gE.flat<float>().data(), 0, gE.TotalBytes(), st);
    cudaMemsetAsync(gR->flat<float>().data(), 0, gR->TotalBytes(), st);
    cudaMemsetAsync(tE.flat<uint8>().data(), 0, tE.TotalBytes(), st);
    cudaMemsetAsync(tR.flat<uint8>().data(), 0, tR.TotalBytes(), st);
    int rc = kgeb_score_bwd(ctx_.h, E.flat<float>().data(), R.flat<float>().data(), idx.flat<int32>().data(), idx.dim_size(0),
                            ds.flat<float>().data(), gE.flat<float>().data(), gR->flat<float>().data(), tE.flat<uint8>().data(),
                            tR.flat<uint8>().data(), st);
    OP_REQUIRES(c, rc == 0, errors::Internal("kgeb_score_bwd: ", kgeb_last_error(ctx_.h)));
    rc = kgeb_grad_rows_export(ctx_.h, gE.flat<float>().data(), tE.flat<uint8>().data(), N, W, W, ids->flat<int32>().data(),
                               rows->flat<float>().data(), N, reinterpret_cast<int64_t*>(count->flat<int64>().data()), 0, st);
    OP_REQUIRES(c, rc == 0, errors::Internal("kgeb_grad_rows_export: ", kgeb_last_error(ctx_.h)));
  }
 private:
  int kind_; float gamma_; Ctx ctx_;
};

// one fused training step, in place on the variables' buffers (the op owns the gradient buffers and flags between steps)
class KgebTrainStepOp : public OpKernel {
 public:
  explicit KgebTrainStepOp(OpKernelConstruction* c) : OpKernel(c) {
    OP_REQUIRES_OK(c, c->GetAttr("score_kind", &kind_));
    OP_REQUIRES_OK(c, c->GetAttr("gamma", &gamma_));
    OP_REQUIRES_OK(c, c->GetAttr("negatives", &K_));
    OP_REQUIRES_OK(c, c->GetAttr("loss_kind", &loss_kind_));
    OP_REQUIRES_OK(c, c->GetAttr("margin", &margin_));
    OP_REQUIRES_OK(c, c->GetAttr("lr", &lr_));
    OP_REQUIRES_OK(c, c->GetAttr("beta1", &b1_));
    OP_REQUIRES_OK(c, c->GetAttr("beta2", &b2_));
    OP_REQUIRES_OK(c, c->GetAttr("epsilon", &eps_));
    OP_REQUIRES_OK(c, c->GetAttr("seed", &seed_));
  }
  ~KgebTrainStepOp() override { cudaFree(gE_); cudaFree(gR_); cudaFree(tE_); cudaFree(tR_); }
  void Compute(OpKernelContext* c) override {
    const Tensor &E = c->input(0), &R = c->input(1), &mE = c->input(2), &vE = c->input(3), &mR = c->input(4), &vR = c->input(5),
                 &pos = c->input(6);
    OP_REQUIRES_OK(c, ctx_.Ensure(E, R, kind_, gamma_));
    const int64_t N = E.dim_size(0), Rn = R.dim_size(0), B = pos.dim_size(0);
    const int We = (int)E.dim_size(1), Wr = (int)R.dim_size(1);
    cudaStream_t st = StreamOf(c);
    if (!gE_) {   // all-zero between steps: the update kernel clears what it consumed
      OP_REQUIRES(c, cudaMalloc(&gE_, sizeof(float) * N * We) == cudaSuccess && cudaMalloc(&gR_, sizeof(float) * Rn * Wr) == cudaSuccess &&
                         cudaMalloc(&tE_, N) == cudaSuccess && cudaMalloc(&tR_, Rn) == cudaSuccess,
                  errors::Internal("KgebTrainStep: out of device memory"));
      cudaMemsetAsync(gE_, 0, sizeof(float) * N * We, st); cudaMemsetAsync(gR_, 0, sizeof(float) * Rn * Wr, st);
      cudaMemsetAsync(tE_, 0, N, st); cudaMemsetAsync(tR_, 0, Rn, st);
    }
    Tensor* loss = nullptr;
    OP_REQUIRES_OK(c, c->allocate_output(0, TensorShape({}), &loss));
    float* Ep = const_cast<float*>(E.flat<float>().data());
    float* Rp = const_cast<float*>(R.flat<float>().data());
    const int64_t step = ++step_;   // the step input orders the op after the previous one; the count itself lives here
    int rc = kgeb_train_fwd_bwd(ctx_.h, Ep, Rp, pos.flat<int32>().data(), B, nullptr, K_, loss_kind_, margin_, 1.0f, KGEB_CORRUPT_BOTH,
                                (uint64_t)seed_, (uint64_t)step, loss->flat<float>().data(), nullptr, nullptr, nullptr, gE_, gR_, tE_, tR_, st);
    OP_REQUIRES(c, rc == 0, errors::Internal("kgeb_train_fwd_bwd: ", kgeb_last_error(ctx_.h)));
    rc = kgeb_opt_update2(ctx_.h, Ep, gE_, tE_, N, We, We, const_cast<float*>(mE.flat<float>().data()), const_cast<float*>(vE.flat<float>().data()),
                          Rp, gR_, tR_, Rn, Wr, Wr, const_cast<float*>(mR.flat<float>().data()), const_cast<float*>(vR.flat<float>().data()),
                          KGEB_ADAM_DENSE, lr_, b1_, b2_, eps_, step, st);
    OP_REQUIRES(c, rc == 0, errors::Internal("kgeb_opt_update2: ", kgeb_last_error(ctx_.h)));
  }
 private:
  int kind_, K_, loss_kind_, seed_; float gamma_, margin_, lr_, b1_, b2_, eps_; int64_t step_ = 0; Ctx ctx_;
  float *gE_ = nullptr, *gR_ = nullptr; uint8_t *tE_ = nullptr, *tR_ = nullptr;
};

// filtered 1-N ranking of one side: counts[q] = #{c not filtered, c != target : score_c > score_target}; rank = 1 + counts
class KgebRank1NOp : public OpKernel {
 public:
  explicit KgebRank1NOp(OpKernelConstruction* c) : OpKernel(c) {
    OP_REQUIRES_OK(c, c->GetAttr("score_kind", &kind_));
    OP_REQUIRES_OK(c, c->GetAttr("gamma", &gamma_));
    OP_REQUIRES_OK(c, c->GetAttr("side", &side_));
  }
  void Compute(OpKernelContext* c) override {
    const Tensor &E = c->input(0), &R = c->input(1), &tri = c->input(2), &offs = c->input(3), &fids = c->input(4);
    OP_REQUIRES_OK(c, ctx_.Ensure(E, R, kind_, gamma_));
    const int64_t N = E.dim_size(0), Q = tri.dim_size(0);
    Tensor *counts = nullptr, qvec, tscore;
    OP_REQUIRES_OK(c, c->allocate_output(0, TensorShape({Q}), &counts));
    const int qw = kgeb_query_width(ctx_.h, side_);
    OP_REQUIRES(c, qw > 0, errors::Internal("kgeb_query_width: ", kgeb_last_error(ctx_.h)));
    OP_REQUIRES_OK(c, c->allocate_temp(DT_FLOAT, TensorShape({Q, qw}), &qvec));
    OP_REQUIRES_OK(c, c->allocate_temp(DT_FLOAT, TensorShape({Q}), &tscore));
    cudaStream_t st = StreamOf(c);
    cudaMemsetAsync(counts->flat<int32>().data(), 0, sizeof(int32_t) * Q, st);
    const float* Ep = E.flat<float>().data();
    int rc = kgeb_query_prep(ctx_.h, Ep, 0, N, R.flat<float>().data(), tri.flat<int32>().data(), Q, side_, qvec.flat<float>().data(), st);
    OP_REQUIRES(c, rc == 0, errors::Internal("kgeb_query_prep: ", kgeb_last_error(ctx_.h)));
    rc = kgeb_target_scores(ctx_.h, Ep, 0, N, qvec.flat<float>().data(), tri.flat<int32>().data(), Q, side_, tscore.flat<float>().data(), st);
    OP_REQUIRES(c, rc == 0, errors::Internal("kgeb_target_scores: ", kgeb_last_error(ctx_.h)));
    const int64_t* fo = offs.dim_size(0) > 1 && fids.dim_size(0) > 0 ? reinterpret_cast<const int64_t*>(offs.flat<int64>().data()) : nullptr;
    const int32_t* fi = fo ? fids.flat<int32>().data() : nullptr;
    const bool tc = kind_ == KGEB_DISTMULT || kind_ == KGEB_COMPLEX || kind_ == KGEB_HOLE || kind_ == KGEB_ROTATE;
    if (!tc) {
      rc = kgeb_rank_1n(ctx_.h, Ep, 0, N, qvec.flat<float>().data(), tscore.flat<float>().data(), tri.flat<int32>().data(), Q, side_, fo, fi,
                        counts->flat<int32>().data(), st);
      OP_REQUIRES(c, rc == 0, errors::Internal("kgeb_rank_1n: ", kgeb_last_error(ctx_.h)));
      return;
    }
    // tensor cores, exact mode.  The packed table image is rebuilt per call (the op cannot know whether E changed); a serving
    // graph would cache it per variable version.  The worklist is sized for 1 % undecided pairs; an overflow is reported.
    const int64_t cap = Q * (64 + N / 100) + (1 << 20);
    const int64_t pbytes = kgeb_rank_tc_packed_bytes(ctx_.h, N, KGEB_TC_F16X3_EXACT);
    const int64_t wbytes = kgeb_rank_tc_workspace_bytes(ctx_.h, Q, KGEB_TC_F16X3_EXACT, cap);
    OP_REQUIRES(c, pbytes > 0 && wbytes > 0, errors::Internal("kgeb_rank_tc sizes: ", kgeb_last_error(ctx_.h)));
    Tensor packed, ws;
    OP_REQUIRES_OK(c, c->allocate_temp(DT_UINT8, TensorShape({pbytes}), &packed));
    OP_REQUIRES_OK(c, c->allocate_temp(DT_UINT8, TensorShape({wbytes}), &ws));
    rc = kgeb_rank_tc_pack_entities(ctx_.h, Ep, N, KGEB_TC_F16X3_EXACT, packed.flat<uint8>().data(), pbytes, st);
    OP_REQUIRES(c, rc == 0, errors::Internal("kgeb_rank_tc_pack_entities: ", kgeb_last_error(ctx_.h)));
    rc = kgeb_rank_1n_tc(ctx_.h, Ep, 0, N, packed.flat<uint8>().data(), qvec.flat<float>().data(), tscore.flat<float>().data(),
                         tri.flat<int32>().data(), Q, side_, fo, fi, KGEB_TC_F16X3_EXACT, ws.flat<uint8>().data(), wbytes, cap,
                         counts->flat<int32>().data(), st);
    OP_REQUIRES(c, rc == 0, errors::Internal("kgeb_rank_1n_tc: ", kgeb_last_error(ctx_.h)));
    uint32_t hdr[2] = {0, 0};
    cudaMemcpyAsync(hdr, ws.flat<uint8>().data(), sizeof hdr, cudaMemcpyDeviceToHost, st);
    cudaStreamSynchronize(st);
    OP_REQUIRES(c, hdr[0] <= hdr[1], errors::Internal("KgebRank1N: undecided-pair worklist overflowed; raise its capacity"));
  }
 private:
  int kind_, side_; float gamma_; Ctx ctx_;
};

REGISTER_KERNEL_BUILDER(Name("KgebScore").Device(DEVICE_GPU), KgebScoreOp);
REGISTER_KERNEL_BUILDER(Name("KgebScoreGradSlices").Device(DEVICE_GPU), KgebScoreGradSlicesOp);
REGISTER_KERNEL_BUILDER(Name("KgebTrainStep").Device(DEVICE_GPU), KgebTrainStepOp);
REGISTER_KERNEL_BUILDER(Name("KgebRank1N").Device(DEVICE_GPU), KgebRank1NOp);
REGISTER_KERNEL_BUILDER(Name("KgebScoreGrad").Device(DEVICE_GPU), KgebScoreGradOp);
