// One-launch versions of per-layer helper kernels (kernels_batched.cu)
#pragma once
#include "common.cuh"

struct PackDesc {            // master fp32 [Co][Ci_real][RS] -> wp [Co][RS][Ci], wpd [Ci][RS(flipped)][Co]
    const float* w;
    void *wp, *wpd;
    int Co, Ci_real, Ci, RS;
};
struct UnpackDesc {          // dw[k][c][t] = dwp[t][c][k]
    const float* dwp;
    float* dw;
    int Co, Ci_real, Ci_pad, RS;
};
void k_pack_all(cudaStream_t st, DType dt, const PackDesc* descs, const int* blk_start, int nlayers, int total_blocks);
void k_unpack_all(cudaStream_t st, const UnpackDesc* descs, const int* blk_start, int nlayers, int total_blocks, int max_rs);
