// tcgen05 implicit-GEMM convolution (bf16 in, fp32 accumulate) - see conv_tc.cu
#pragma once
#include "kernels.h"
#include "tc_epilogue.h"

// can the tensor-core kernel run this geometry (forward, or stride-1 dgrad)?
bool tc_conv_supported(const ConvGeom& g, bool dgrad);

// out[n,y,x,k] (+)= sum_{r,s,c} A[n, y*stride+r-pad, x*stride+s-pad, c] * Wp[k][(r*S+s)*Ca + c]   (zero outside A)
// A: [B,Ha,Wa,Ca] bf16;  Wp: [Nout][R*S*Ca] bf16;  out: [B,Ho,Wo,Nout] bf16 (fp32 when out_f32: the split-bf16 parity mode);  stats: zeroed [SALT_STAT_SLOTS_CONV][2*Nout] float partial slots or NULL
void k_conv_tc(cudaStream_t st, const void* A, int B, int Ha, int Wa, int Ca, const void* Wp, int Nout, int R, int S, int stride,
               int pad, void* out, int Ho, int Wo, const float* bias, float* stats, bool accumulate, bool out_f32 = false,
               const EpiParams* ep = nullptr, int split_c = 0);

// 3x3 stride-1 variant with shared-memory row-halo reuse of the activation tile (conv_tc_rows.cu); env SALT_TC_ROWS=0 disables it
bool tc_conv_rows_supported(int Ca, int Nout, int R, int S, int stride, int Ho, int Wo);
void k_conv_tc_rows(cudaStream_t st, const void* A, int B, int Ha, int Wa, int Ca, const void* Wp, int Nout, int pad, void* out,
                    int Ho, int Wo, const float* bias, float* stats, bool accumulate, bool out_f32 = false, const EpiParams* ep = nullptr,
                    int split_c = 0);

// stride-2 dgrad as four parity phases of the same kernel; wpd = flipped-tap packing [Ci][R*S*Co]
void k_conv_tc_dgrad_s2(cudaStream_t st, const void* gout, int B, int Ho, int Wo, int Co, const void* wpd, int Ci, int R, int S,
                        int pad, void* gin, int Hi, int Wi, bool accumulate);

// tensor-core weight gradient (conv_wgrad_tc.cu): dw[k][c][r][s] += sum_pixels gout * in (fp32 accumulate, added into dw)
bool tc_wgrad_supported(const ConvGeom& g);
// dwp: zero-initialised fp32 scratch [R*S][ceil64(Ci)][Co] (k contiguous), partial sums are added with vector reductions
size_t tc_wgrad_scratch_floats(int Ci, int Co, int RS);
void k_conv_wgrad_tc(cudaStream_t st, const void* in, const void* gout, float* dwp, const ConvGeom& g);
// dw[k][c][r][s] (+)= dwp[tap][c][k]
void k_unpack_dw(cudaStream_t st, const float* dwp, float* dw, int Co, int Ci_real, int Ci_pad, int RS, bool accumulate);
