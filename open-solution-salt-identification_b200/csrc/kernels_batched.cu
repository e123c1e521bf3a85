// Whole-network ("batched") versions of the small per-layer kernels: one launch covers every convolution layer, the layer
// is found from a prefix table of block counts.  Removes ~100 tiny launches per training step.
#include "kernels.h"
#include "kernels_batched.h"

__device__ __forceinline__ int find_layer(const int* __restrict__ blk_start, int n, int b) {
    int lo = 0, hi = n - 1;                       // blk_start has n+1 entries; find l with blk_start[l] <= b < blk_start[l+1]
    while (lo < hi) {
        int mid = (lo + hi + 1) >> 1;
        if (blk_start[mid] <= b) lo = mid; else hi = mid - 1;
    }
    return lo;
}

// ------------------------------------------------------------------------------------------------ weight packing
// One block = a 32 (k) x CT (c) x RS tile of one layer, staged through shared memory so that the fp32 master weights are read
// as contiguous runs ([c][t] for one k) and both packed copies are written as contiguous runs (wp: c for one (k,t); wpd: k for
// one (c,t)).  The first version (one thread per element, strided reads, 2-byte scattered wpd stores) ran at 0.6 TB/s.
template <typename T>
__global__ void __launch_bounds__(256) pack_all_kernel(const PackDesc* __restrict__ descs, const int* __restrict__ blk_start, int nlayers) {
    extern __shared__ float tile[];                  // [32 k][CT * RS + 1]
    const int l = find_layer(blk_start, nlayers, blockIdx.x);
    const PackDesc d = descs[l];
    const int b = blockIdx.x - blk_start[l];
    const int RS = d.RS, CT = pack_ct(RS), pitch = CT * RS + 1;
    const int ktiles = (d.Co + 31) >> 5;
    const int k0 = (b % ktiles) * 32, c0 = (b / ktiles) * CT;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int creal = max(0, min(CT, d.Ci_real - c0));         // channels of this tile that exist in the master weights
    const int cmem = min(CT, d.Ci - c0);                       // channels of this tile that exist in the packed copies (zero pad)
    const int ncols = creal * RS;
    if (d.groups == 1) {
        for (int k = ty; k < 32; k += 8) {
            const float* src = d.w + ((size_t)(k0 + k) * d.Ci_real + c0) * RS;
            for (int j = tx; j < CT * RS; j += 32) tile[k * pitch + j] = (k0 + k < d.Co && j < ncols) ? src[j] : 0.f;
        }
    } else {                                           // block-diagonal: zero outside the output channel's own group
        const int cpg = d.Ci_real / d.groups, kpg = d.Co / d.groups;
        for (int k = ty; k < 32; k += 8) {
            const int kk = k0 + k, cg0 = kk < d.Co ? (kk / kpg) * cpg : 0;
            const float* src = d.w + (size_t)kk * cpg * RS;
            for (int j = tx; j < CT * RS; j += 32) {
                const int c = c0 + j / RS, t = j - (j / RS) * RS;
                tile[k * pitch + j] = (kk < d.Co && j < ncols && c >= cg0 && c < cg0 + cpg) ? src[(c - cg0) * RS + t] : 0.f;
            }
        }
    }
    __syncthreads();
    T* wp = (T*)d.wp;
    T* wpd = (T*)d.wpd;
    // wp[k][t][c]
    for (int kt = ty; kt < 32 * RS; kt += 8) {
        const int k = kt / RS, t = kt - k * RS;
        if (k0 + k >= d.Co) continue;
        for (int c = tx; c < cmem; c += 32) st1(wp + ((size_t)(k0 + k) * RS + t) * d.Ci + c0 + c, tile[k * pitch + c * RS + t]);
    }
    // wpd[c][RS-1-t][k]  (flipped taps, see kernels_conv_simt.cu)
    if (k0 + tx < d.Co) {
        for (int ct = ty; ct < cmem * RS; ct += 8) {
            const int c = ct / RS, t = ct - c * RS;
            st1(wpd + ((size_t)(c0 + c) * RS + (RS - 1 - t)) * d.Co + k0 + tx, tile[tx * pitch + c * RS + t]);
        }
    }
}
void k_pack_all(cudaStream_t st, DType dt, const PackDesc* descs, const int* blk_start, int nlayers, int total_blocks, int max_rs) {
    SALT_COUNT(1);
    const size_t smem = sizeof(float) * 32 * (pack_ct(max_rs) * max_rs + 1);
    const size_t smem9 = sizeof(float) * 32 * (pack_ct(9) * 9 + 1);
    const size_t need = smem > smem9 ? smem : smem9;
    SALT_DISPATCH(dt, T, {
        static size_t configured = 0;
        if (need > configured) {
            cudaFuncSetAttribute(pack_all_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)need);
            configured = need;
        }
        pack_all_kernel<T><<<total_blocks, 256, need, st>>>(descs, blk_start, nlayers);
    });
}

// ------------------------------------------------------------------------------------------------ wgrad unpack
// dw[k][c][t] = dwp[t][c][k]: 32(k) x 32(c) x RS tile transposed through shared memory (coalesced both sides)
__global__ void __launch_bounds__(256) unpack_all_kernel(const UnpackDesc* __restrict__ descs, const int* __restrict__ blk_start, int nlayers) {
    extern __shared__ float tile[];                  // [32 k][32 c * RS + 1]
    const int l = find_layer(blk_start, nlayers, blockIdx.x);
    const UnpackDesc d = descs[l];
    const int b = blockIdx.x - blk_start[l];
    const int ktiles = (d.Co + 31) >> 5;
    const int k0 = (b % ktiles) * 32, c0 = (b / ktiles) * 32, RS = d.RS, pitch = 32 * RS + 1;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    for (int t = 0; t < RS; ++t)
        for (int c = ty; c < 32; c += 8) {
            float v = 0.f;
            if (k0 + tx < d.Co && c0 + c < d.Ci_pad) v = d.dwp[((size_t)t * d.Ci_pad + c0 + c) * d.Co + k0 + tx];
            tile[tx * pitch + c * RS + t] = v;
        }
    __syncthreads();
    const int ncols = min(32, d.Ci_real - c0) * RS;  // contiguous run in dw for one k
    if (d.groups == 1) {
        for (int k = ty; k < 32; k += 8) {
            if (k0 + k >= d.Co) continue;
            float* o = d.dw + ((size_t)(k0 + k) * d.Ci_real + c0) * RS;
            for (int j = tx; j < ncols; j += 32) o[j] = tile[k * pitch + j];
        }
    } else {                                           // keep the diagonal blocks only
        const int cpg = d.Ci_real / d.groups, kpg = d.Co / d.groups;
        for (int k = ty; k < 32; k += 8) {
            const int kk = k0 + k;
            if (kk >= d.Co) continue;
            const int cg0 = (kk / kpg) * cpg;
            for (int j = tx; j < ncols; j += 32) {
                const int c = c0 + j / RS, t = j - (j / RS) * RS;
                if (c >= cg0 && c < cg0 + cpg) d.dw[((size_t)kk * cpg + (c - cg0)) * RS + t] = tile[k * pitch + j];
            }
        }
    }
}
void k_unpack_all(cudaStream_t st, const UnpackDesc* descs, const int* blk_start, int nlayers, int total_blocks, int max_rs) {
    SALT_COUNT(1);
    const size_t smem = sizeof(float) * 32 * (32 * max_rs + 1);
    static size_t configured = 0;
    if (smem > configured) {
        cudaFuncSetAttribute(unpack_all_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        configured = smem;
    }
    unpack_all_kernel<<<total_blocks, 256, smem, st>>>(descs, blk_start, nlayers);
}
