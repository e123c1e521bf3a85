// One-launch versions of per-layer helper kernels (kernels_batched.cu)
#pragma once
#include "common.cuh"

// groups > 1: a grouped convolution (SE-ResNeXt, groups = 32) runs as a DENSE convolution over block-diagonal packed weights - the
// master tensor is the reference's [Co][Ci_real/groups][RS], output channel k belongs to group k / (Co/groups) and sees the input
// channels [g*cpg, (g+1)*cpg), cpg = Ci_real/groups; everything outside the diagonal blocks is packed as zero and its gradient dropped.
struct PackDesc {            // master fp32 [Co][Ci_real/groups][RS] -> wp [Co][RS][Ci], wpd [Ci][RS(flipped)][Co]
    const float* w;
    void *wp, *wpd;
    int Co, Ci_real, Ci, RS, groups;
};
struct UnpackDesc {          // dw[k][c - g*cpg][t] = dwp[t][c][k]
    const float* dwp;
    float* dw;
    int Co, Ci_real, Ci_pad, RS, groups;
};
// channel-tile width of the pack kernel for a filter with RS taps (shared-memory tile = 32 k x pack_ct(RS) c x RS)
__host__ __device__ static inline int pack_ct(int RS) { return RS <= 9 ? 32 : 4; }
static inline int pack_blocks(int Co, int Ci, int RS) { return ((Co + 31) / 32) * ((Ci + pack_ct(RS) - 1) / pack_ct(RS)); }
void k_pack_all(cudaStream_t st, DType dt, const PackDesc* descs, const int* blk_start, int nlayers, int total_blocks, int max_rs);
void k_unpack_all(cudaStream_t st, const UnpackDesc* descs, const int* blk_start, int nlayers, int total_blocks, int max_rs);
