// train_kernels_dp.cu -- the dropout flavour of the fused train-step kernel (a separate translation unit so both sets of
// template instantiations compile in parallel).
#include "train_ring.cuh"

int kgeb_internal_train_ring_dp(kgeb_ctx* ctx, const void* args, int nc, int mode, int D, int RB, size_t smem, int64_t B, cudaStream_t st) {
  return launch_train_ring<true>(ctx, *static_cast<const TrainArgs*>(args), nc, mode, D, RB, smem, B, st);
}
