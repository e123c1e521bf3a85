// Stream-ring versions of the memory-bound passes (kernels_stream.cu).  Each returns false without launching anything when the
// tensors do not fit the ring geometry (see stream_ring.cuh); SALT_EW_RING=0 turns the ring off (A/B measurements).
#pragma once
#include "kernels.h"

bool ring_enabled();
bool k_ring_bn_apply(cudaStream_t st, const Tensor& raw, const float* scale, const float* shift, const Tensor* res, const float* rscale,
                     const float* rshift, bool relu, const Tensor& out, const float* gate);
bool k_ring_bn_bwd_reduce(cudaStream_t st, const Tensor& g, const Tensor& raw, const BNRef& bn, bool self_mask, const float* gate,
                          const float* addc);
bool k_ring_bn_bwd_apply(cudaStream_t st, const Tensor& g, const Tensor& raw, const BNRef& bn, bool self_mask, const Tensor& graw,
                         const float* gate, const float* addc);
bool k_ring_relu_mask(cudaStream_t st, const Tensor& g, const Tensor& mask);
// per-pixel passes (bf16, C = 8 * 2^k <= 256): scSE gate forward / backward, final 1x1 convolution forward / backward
bool k_ring_scse_apply(cudaStream_t st, const Tensor& raw, const float* scale, const float* shift, const SERef& se, const Tensor& out);
bool k_ring_scse_bwd_apply(cudaStream_t st, const Tensor& gout, const Tensor& raw, const BNRef& bn, const SERef& se, const Tensor& gbn);
bool k_ring_final_fwd(cudaStream_t st, const Tensor& raw, const float* scale, const float* shift, const float* w, const float* b, int K,
                      float* logits);
bool k_ring_final_bwd(cudaStream_t st, const float* dlogits, const Tensor& raw, const BNRef& bn, const float* w, int K, float* dw, float* db,
                      const Tensor& gbn);
// squeeze pass of both SE kinds: part[n][slot][c] partial channel sums per image (g != NULL: the backward pre-pass, sum of g*z)
bool k_ring_se_pool(cudaStream_t st, const Tensor& raw, const void* g, bool relu, const float* scale, const float* shift, const SERef& se);
