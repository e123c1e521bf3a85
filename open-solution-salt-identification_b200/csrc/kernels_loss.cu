// Loss, prediction and optimiser kernels.
//   * Lovasz hinge (ELU variant) - reference common_blocks/lovasz_losses.py:21-33,97-115: one CTA per image,
//     the whole image's C*H*W hinge errors sorted in shared memory (bitonic, key fp32 + 16-bit payload
//     = pixel index | label bit), block scan for the Jaccard gradient, loss and dL/dlogits in one launch.
//   * BCE + soft Dice - reference common_blocks/models.py:315-340,361-388.
//   * sigmoid + h-flip TTA mean + crop + threshold - utils.py:173, loaders.py:751-760, postprocessing.py:24-43.
//   * Adam with L2 (models.py:74-75,289-297), fused over the flat parameter buffer.
#include "kernels.h"
#include <math_constants.h>
#include <stdexcept>

// ------------------------------------------------------------------------------------------------
// Lovasz hinge
// ------------------------------------------------------------------------------------------------
#define LV_THREADS 1024

__device__ __forceinline__ bool lv_before(float ka, unsigned short ia, float kb, unsigned short ib) {
    // descending by key, ties by ascending pixel index (deterministic)
    return (ka > kb) || (ka == kb && (ia & 0x7fff) < (ib & 0x7fff));
}

__global__ void __launch_bounds__(LV_THREADS) lovasz_kernel(const float* __restrict__ logits, const float* __restrict__ target,
                                                            int P, int Ppad, float inv_b, float* __restrict__ per_image,
                                                            float* __restrict__ dlogits) {
    extern __shared__ __align__(16) unsigned char lv_smem[];
    float* key = reinterpret_cast<float*>(lv_smem);
    unsigned short* pay = reinterpret_cast<unsigned short*>(lv_smem + sizeof(float) * Ppad);
    __shared__ float warp_tot[32];
    __shared__ float warp_loss[32];

    const int tid = threadIdx.x, b = blockIdx.x;
    const float* lg = logits + (size_t)b * P;
    const float* tg = target + (size_t)b * P;
    for (int i = tid; i < Ppad; i += LV_THREADS) {
        if (i < P) {
            int lab = ((long long)tg[i]) != 0 ? 1 : 0;          // target.long() (models.py:327)
            float sign = lab ? 1.f : -1.f;
            key[i] = 1.f - lg[i] * sign;
            pay[i] = (unsigned short)(i | (lab << 15));
        } else {
            key[i] = -CUDART_INF_F;
            pay[i] = 0x7fff;
        }
    }
    __syncthreads();
    // bitonic sort, final order: lv_before(a, b) for a at the lower index
    for (int k = 2; k <= Ppad; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int t = tid; t < (Ppad >> 1); t += LV_THREADS) {
                int i = ((t & ~(j - 1)) << 1) | (t & (j - 1));
                int l = i | j;
                bool up = (i & k) == 0;
                float ka = key[i], kb = key[l];
                unsigned short ia = pay[i], ib = pay[l];
                bool in_order = lv_before(ka, ia, kb, ib);
                if (in_order != up) { key[i] = kb; key[l] = ka; pay[i] = ib; pay[l] = ia; }
            }
            __syncthreads();
        }
    }
    // each warp owns a contiguous chunk; pass 1: label totals
    const int warp = tid >> 5, lane = tid & 31;
    const int chunk = Ppad / 32;
    const int c0 = warp * chunk;
    float tot = 0.f;
    for (int i = c0 + lane; i < c0 + chunk; i += 32) tot += (float)(pay[i] >> 15);
    for (int o = 16; o > 0; o >>= 1) tot += __shfl_xor_sync(0xffffffffu, tot, o);
    if (lane == 0) warp_tot[warp] = tot;
    __syncthreads();
    float G = 0.f, carry = 0.f;
    for (int w = 0; w < 32; ++w) { float t = warp_tot[w]; G += t; if (w < warp) carry += t; }
    // pass 2: inclusive scan inside the chunk, Jaccard gradient, loss and d/dlogit
    float loss = 0.f;
    for (int base = c0; base < c0 + chunk; base += 32) {
        int i = base + lane;
        float gt = (float)(pay[i] >> 15);
        float inc = gt;
        for (int o = 1; o < 32; o <<= 1) {
            float n = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= o) inc += n;
        }
        float cs = carry + inc;                                  // inclusive cumsum of labels at rank i
        carry += __shfl_sync(0xffffffffu, inc, 31);
        if (i < P) {
            float inter = G - cs, uni = G + ((float)(i + 1) - cs);
            float jac = 1.f - inter / uni;
            float grad = jac;
            if (i > 0) {
                float csp = cs - gt;
                float interp = G - csp, unip = G + ((float)i - csp);
                grad = jac - (1.f - interp / unip);
            }
            float e = key[i];
            float elu = e > 0.f ? e : expm1f(e);
            float delu = e > 0.f ? 1.f : expf(e);
            loss = fmaf(elu, grad, loss);
            unsigned short pv = pay[i];
            float sign = (pv >> 15) ? 1.f : -1.f;
            dlogits[(size_t)b * P + (pv & 0x7fff)] = -sign * delu * grad * inv_b;
        }
    }
    for (int o = 16; o > 0; o >>= 1) loss += __shfl_xor_sync(0xffffffffu, loss, o);
    if (lane == 0) warp_loss[warp] = loss;
    __syncthreads();
    if (tid == 0) {
        float t = 0.f;
        for (int w = 0; w < 32; ++w) t += warp_loss[w];
        per_image[b] = t;
    }
}
__global__ void mean_kernel(const float* __restrict__ v, int n, float* __restrict__ out) {
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        float t = 0.f;
        for (int i = 0; i < n; ++i) t += v[i];
        out[0] = t / (float)n;
    }
}
// ---- images with more than 32768 logits (256x256 inputs, BASELINE config 4): the same algorithm with the sort in global memory.
// Keys / payloads (pixel index | label << 31) live in a [B][Ppad] scratch; chunks of LVB_CH elements are sorted in shared
// memory, the bitonic merge levels above the chunk size run their wide strides as global passes and their tail in shared memory.
#define LVB_CH 16384
#define LVB_SMEM (LVB_CH * 8)
__device__ __forceinline__ bool lvb_before(float ka, unsigned ia, float kb, unsigned ib) {
    return (ka > kb) || (ka == kb && (ia & 0x7fffffffu) < (ib & 0x7fffffffu));
}
// in-smem bitonic stages j = j0 .. 1 of merge level k on one chunk whose first element has global index g0
__device__ __forceinline__ void lvb_smem_stages(float* key, unsigned* pay, int g0, int k, int j0) {
    for (int j = j0; j > 0; j >>= 1) {
        for (int t = threadIdx.x; t < (LVB_CH >> 1); t += LV_THREADS) {
            const int i = ((t & ~(j - 1)) << 1) | (t & (j - 1)), l = i | j;
            const bool up = ((g0 + i) & k) == 0;
            const float ka = key[i], kb = key[l];
            const unsigned ia = pay[i], ib = pay[l];
            if (lvb_before(ka, ia, kb, ib) != up) { key[i] = kb; key[l] = ka; pay[i] = ib; pay[l] = ia; }
        }
        __syncthreads();
    }
}
// grid (Ppad / LVB_CH, B): hinge errors -> keys, then every merge level up to the chunk size
__global__ void __launch_bounds__(LV_THREADS) lvb_init_kernel(const float* __restrict__ logits, const float* __restrict__ target,
                                                              int P, int Ppad, float* __restrict__ gkey, unsigned* __restrict__ gpay) {
    extern __shared__ __align__(16) unsigned char lv_smem[];
    float* key = reinterpret_cast<float*>(lv_smem);
    unsigned* pay = reinterpret_cast<unsigned*>(lv_smem + sizeof(float) * LVB_CH);
    const int b = blockIdx.y, g0 = blockIdx.x * LVB_CH;
    const float* lg = logits + (size_t)b * P;
    const float* tg = target + (size_t)b * P;
    for (int i = threadIdx.x; i < LVB_CH; i += LV_THREADS) {
        const int g = g0 + i;
        if (g < P) {
            const unsigned lab = ((long long)tg[g]) != 0 ? 1u : 0u;
            key[i] = 1.f - lg[g] * (lab ? 1.f : -1.f);
            pay[i] = (unsigned)g | (lab << 31);
        } else {
            key[i] = -CUDART_INF_F;
            pay[i] = 0x7fffffffu;
        }
    }
    __syncthreads();
    for (int k = 2; k <= LVB_CH; k <<= 1) lvb_smem_stages(key, pay, g0, k, k >> 1);
    for (int i = threadIdx.x; i < LVB_CH; i += LV_THREADS) {
        gkey[(size_t)b * Ppad + g0 + i] = key[i];
        gpay[(size_t)b * Ppad + g0 + i] = pay[i];
    }
}
// one compare-exchange pass with stride j >= LVB_CH of merge level k.  grid (Ppad / 2 / 256, B)
__global__ void lvb_global_pass_kernel(float* __restrict__ gkey, unsigned* __restrict__ gpay, int Ppad, int k, int j) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (Ppad >> 1)) return;
    float* key = gkey + (size_t)blockIdx.y * Ppad;
    unsigned* pay = gpay + (size_t)blockIdx.y * Ppad;
    const int i = ((t & ~(j - 1)) << 1) | (t & (j - 1)), l = i | j;
    const bool up = (i & k) == 0;
    const float ka = key[i], kb = key[l];
    const unsigned ia = pay[i], ib = pay[l];
    if (lvb_before(ka, ia, kb, ib) != up) { key[i] = kb; key[l] = ka; pay[i] = ib; pay[l] = ia; }
}
// strides LVB_CH/2 .. 1 of merge level k, one chunk per CTA.  grid (Ppad / LVB_CH, B)
__global__ void __launch_bounds__(LV_THREADS) lvb_tail_kernel(float* __restrict__ gkey, unsigned* __restrict__ gpay, int Ppad, int k) {
    extern __shared__ __align__(16) unsigned char lv_smem[];
    float* key = reinterpret_cast<float*>(lv_smem);
    unsigned* pay = reinterpret_cast<unsigned*>(lv_smem + sizeof(float) * LVB_CH);
    const int g0 = blockIdx.x * LVB_CH;
    const size_t base = (size_t)blockIdx.y * Ppad + g0;
    for (int i = threadIdx.x; i < LVB_CH; i += LV_THREADS) { key[i] = gkey[base + i]; pay[i] = gpay[base + i]; }
    __syncthreads();
    lvb_smem_stages(key, pay, g0, k, LVB_CH >> 1);
    for (int i = threadIdx.x; i < LVB_CH; i += LV_THREADS) { gkey[base + i] = key[i]; gpay[base + i] = pay[i]; }
}
// one CTA per image over the sorted errors: label scan -> Jaccard gradient -> loss and d/dlogit (same arithmetic as lovasz_kernel)
__global__ void __launch_bounds__(LV_THREADS) lvb_scan_kernel(const float* __restrict__ gkey, const unsigned* __restrict__ gpay, int P,
                                                              int Ppad, float inv_b, float* __restrict__ per_image,
                                                              float* __restrict__ dlogits) {
    __shared__ float warp_tot[32];
    __shared__ float warp_loss[32];
    const int tid = threadIdx.x, b = blockIdx.x, warp = tid >> 5, lane = tid & 31;
    const float* key = gkey + (size_t)b * Ppad;
    const unsigned* pay = gpay + (size_t)b * Ppad;
    const int chunk = Ppad / 32, c0 = warp * chunk;
    float tot = 0.f;
    for (int i = c0 + lane; i < c0 + chunk; i += 32) tot += (float)(pay[i] >> 31);
    for (int o = 16; o > 0; o >>= 1) tot += __shfl_xor_sync(0xffffffffu, tot, o);
    if (lane == 0) warp_tot[warp] = tot;
    __syncthreads();
    float G = 0.f, carry = 0.f;
    for (int w = 0; w < 32; ++w) { float t = warp_tot[w]; G += t; if (w < warp) carry += t; }
    float loss = 0.f;
    for (int base = c0; base < c0 + chunk; base += 32) {
        const int i = base + lane;
        const unsigned pv = pay[i];
        const float gt = (float)(pv >> 31);
        float inc = gt;
        for (int o = 1; o < 32; o <<= 1) {
            float n = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= o) inc += n;
        }
        const float cs = carry + inc;
        carry += __shfl_sync(0xffffffffu, inc, 31);
        if (i < P) {
            const float inter = G - cs, uni = G + ((float)(i + 1) - cs);
            const float jac = 1.f - inter / uni;
            float grad = jac;
            if (i > 0) {
                const float csp = cs - gt;
                const float interp = G - csp, unip = G + ((float)i - csp);
                grad = jac - (1.f - interp / unip);
            }
            const float e = key[i];
            const float elu = e > 0.f ? e : expm1f(e);
            const float delu = e > 0.f ? 1.f : expf(e);
            loss = fmaf(elu, grad, loss);
            const float sign = (pv >> 31) ? 1.f : -1.f;
            dlogits[(size_t)b * P + (pv & 0x7fffffffu)] = -sign * delu * grad * inv_b;
        }
    }
    for (int o = 16; o > 0; o >>= 1) loss += __shfl_xor_sync(0xffffffffu, loss, o);
    if (lane == 0) warp_loss[warp] = loss;
    __syncthreads();
    if (tid == 0) {
        float t = 0.f;
        for (int w = 0; w < 32; ++w) t += warp_loss[w];
        per_image[b] = t;
    }
}
size_t lovasz_sort_scratch_bytes(int B, int P) {
    if (P <= 32768) return 0;
    size_t Ppad = LVB_CH;
    while (Ppad < (size_t)P) Ppad <<= 1;
    return (size_t)B * Ppad * 8;
}
void k_lovasz(cudaStream_t st, const float* logits, const float* target, int B, int P, float* per_image, float* loss_out,
              float* dlogits, void* sort_scratch) {
    if (P > 32768) {
        if (!sort_scratch) throw std::runtime_error("k_lovasz: images with more than 32768 logits need the sort scratch buffer");
        int Ppad = LVB_CH;
        while (Ppad < P) Ppad <<= 1;
        float* gkey = (float*)sort_scratch;
        unsigned* gpay = (unsigned*)(gkey + (size_t)B * Ppad);
        static bool configured = false;
        if (!configured) {
            cudaFuncSetAttribute(lvb_init_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, LVB_SMEM);
            cudaFuncSetAttribute(lvb_tail_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, LVB_SMEM);
            configured = true;
        }
        const dim3 gc(Ppad / LVB_CH, B), gp(cdiv(Ppad / 2, 256), B);
        int launches = 1;
        lvb_init_kernel<<<gc, LV_THREADS, LVB_SMEM, st>>>(logits, target, P, Ppad, gkey, gpay);
        for (int k = 2 * LVB_CH; k <= Ppad; k <<= 1) {
            for (int j = k >> 1; j >= LVB_CH; j >>= 1) { lvb_global_pass_kernel<<<gp, 256, 0, st>>>(gkey, gpay, Ppad, k, j); ++launches; }
            lvb_tail_kernel<<<gc, LV_THREADS, LVB_SMEM, st>>>(gkey, gpay, Ppad, k);
            ++launches;
        }
        lvb_scan_kernel<<<B, LV_THREADS, 0, st>>>(gkey, gpay, P, Ppad, 1.0f / B, per_image, dlogits);
        mean_kernel<<<1, 32, 0, st>>>(per_image, B, loss_out);
        SALT_COUNT(launches + 2);
        return;
    }
    SALT_COUNT(2);
    int Ppad = 1024;
    while (Ppad < P) Ppad <<= 1;
    size_t smem = (size_t)Ppad * (sizeof(float) + sizeof(unsigned short));
    static size_t configured = 0;
    if (smem > configured) {
        cudaFuncSetAttribute(lovasz_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        configured = smem;
    }
    lovasz_kernel<<<B, LV_THREADS, smem, st>>>(logits, target, P, Ppad, 1.0f / B, per_image, dlogits);
    mean_kernel<<<1, 32, 0, st>>>(per_image, B, loss_out);
}

// ------------------------------------------------------------------------------------------------
// BCE + Dice
// ------------------------------------------------------------------------------------------------
__global__ void bce_dice_reduce_kernel(const float* __restrict__ logits, const float* __restrict__ target, int K, int HW,
                                       double* __restrict__ sums) {
    __shared__ float red[4][256];
    const int plane = blockIdx.y;                 // n*K + k
    const int k = plane % K;
    const float* lg = logits + (size_t)plane * HW;
    const float* tg = target + (size_t)plane * HW;
    float s_pt = 0.f, s_p = 0.f, s_t = 0.f, s_b = 0.f;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < HW; i += gridDim.x * blockDim.x) {
        float x = lg[i];
        float t = ((long long)tg[i]) != 0 ? 1.f : 0.f;
        float p = 1.f / (1.f + expf(-x));
        s_pt += p * t; s_p += p; s_t += t;
        s_b += fmaxf(x, 0.f) - x * t + log1pf(expf(-fabsf(x)));
    }
    red[0][threadIdx.x] = s_pt; red[1][threadIdx.x] = s_p; red[2][threadIdx.x] = s_t; red[3][threadIdx.x] = s_b;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if (threadIdx.x < o)
            for (int q = 0; q < 4; ++q) red[q][threadIdx.x] += red[q][threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        atomicAdd(sums + k * 3 + 0, (double)red[0][0]);
        atomicAdd(sums + k * 3 + 1, (double)red[1][0]);
        atomicAdd(sums + k * 3 + 2, (double)red[2][0]);
        atomicAdd(sums + 3 * K, (double)red[3][0]);
    }
}
void k_bce_dice_reduce(cudaStream_t st, const float* logits, const float* target, int B, int K, int HW, double* sums) {
    SALT_COUNT(1);
    cudaMemsetAsync(sums, 0, sizeof(double) * (3 * K + 1), st);
    dim3 grid(max(1, min(cdiv(HW, 256 * 8), 16)), B * K);
    bce_dice_reduce_kernel<<<grid, 256, 0, st>>>(logits, target, K, HW, sums);
}
__global__ void bce_dice_finish_kernel(const float* __restrict__ logits, const float* __restrict__ target, int K, int HW,
                                       long long total, const double* __restrict__ sums, double total_count, float dice_w,
                                       float bce_w, float grad_scale, float* __restrict__ loss_out, float* __restrict__ dlogits) {
    const float eps = 1e-7f;
    long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx == 0) {
        double dice = 0;
        for (int k = 0; k < K; ++k) {
            float I = (float)sums[k * 3], U = (float)sums[k * 3 + 1] + (float)sums[k * 3 + 2] + eps;
            dice += 1.0 - 2.0 * I / U;
        }
        loss_out[0] = (float)(dice_w * dice / K + bce_w * sums[3 * K] / total_count);
    }
    if (idx >= total) return;
    int k = (int)((idx / HW) % K);
    float x = logits[idx];
    float t = ((long long)target[idx]) != 0 ? 1.f : 0.f;
    float p = 1.f / (1.f + expf(-x));
    float I = (float)sums[k * 3], U = (float)sums[k * 3 + 1] + (float)sums[k * 3 + 2] + eps;
    float ddice_dp = -2.f * (t * U - I) / (U * U);
    float g = bce_w * (p - t) / (float)total_count + dice_w / (float)K * ddice_dp * p * (1.f - p);
    dlogits[idx] = g * grad_scale;
}
void k_bce_dice_finish(cudaStream_t st, const float* logits, const float* target, int B, int K, int HW, const double* sums,
                       double total_count, float dice_w, float bce_w, float grad_scale, float* loss_out, float* dlogits) {
    SALT_COUNT(1);
    long long total = (long long)B * K * HW;
    bce_dice_finish_kernel<<<cdiv(total, 256), 256, 0, st>>>(logits, target, K, HW, total, sums, total_count, dice_w, bce_w,
                                                          grad_scale, loss_out, dlogits);
}

// ------------------------------------------------------------------------------------------------
// prediction: sigmoid, optional un-flipped h-flip copy mean, crop, threshold
// ------------------------------------------------------------------------------------------------
__global__ void predict_kernel(const float* __restrict__ logits, const float* __restrict__ logits_flip, int K, int S, int T,
                               int top, int left, float thr, long long total, float* __restrict__ probs,
                               uint8_t* __restrict__ mask) {
    long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= total) return;
    int x = (int)(idx % S), y = (int)((idx / S) % S);
    long long plane = idx / ((long long)S * S);
    int k = (int)(plane % K), n = (int)(plane / K);
    float p = 1.f / (1.f + expf(-logits[idx]));
    if (logits_flip) {
        float pf = 1.f / (1.f + expf(-logits_flip[(plane * S + y) * S + (S - 1 - x)]));
        p = (p + pf) / 2.f;
    }
    if (probs) probs[idx] = p;
    if (mask && k == 1) {
        int yy = y - top, xx = x - left;
        if (yy >= 0 && yy < T && xx >= 0 && xx < T) mask[((size_t)n * T + yy) * T + xx] = p > thr ? 1 : 0;
    }
}
void k_predict(cudaStream_t st, const float* logits, const float* logits_flip, int B, int K, int S, int T, float thr,
               float* probs, uint8_t* mask) {
    SALT_COUNT(1);
    int d = S - T;
    int top = d / 2, left = d - d / 2;           // utils.py:308-313 get_crop_pad_sequence
    long long total = (long long)B * K * S * S;
    predict_kernel<<<cdiv(total, 256), 256, 0, st>>>(logits, logits_flip, K, S, T, top, left, thr, total, probs, mask);
}

// ------------------------------------------------------------------------------------------------
// Adam + L2  (torch.optim.Adam semantics, weight_decay added to the gradient)
// ------------------------------------------------------------------------------------------------
__global__ void adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
                            size_t n, float lr, float wd, float b1, float b2, float eps, float bc1, float bc2_sqrt,
                            float grad_scale) {
    size_t i = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) * 4;
    if (i >= n) return;
    if (i + 4 <= n) {
        float4 pp = ld4(p + i), gg = ld4(g + i), mm = ld4(m + i), vv = ld4(v + i);
        float pr[4] = {pp.x, pp.y, pp.z, pp.w}, gr[4] = {gg.x, gg.y, gg.z, gg.w}, mr[4] = {mm.x, mm.y, mm.z, mm.w},
              vr[4] = {vv.x, vv.y, vv.z, vv.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            float gj = gr[j] * grad_scale + wd * pr[j];
            mr[j] = b1 * mr[j] + (1.f - b1) * gj;
            vr[j] = b2 * vr[j] + (1.f - b2) * gj * gj;
            float denom = sqrtf(vr[j]) / bc2_sqrt + eps;
            pr[j] -= (lr / bc1) * (mr[j] / denom);
        }
        st4(p + i, make_float4(pr[0], pr[1], pr[2], pr[3]));
        st4(m + i, make_float4(mr[0], mr[1], mr[2], mr[3]));
        st4(v + i, make_float4(vr[0], vr[1], vr[2], vr[3]));
    } else {
        for (size_t j = i; j < n; ++j) {
            float gj = g[j] * grad_scale + wd * p[j];
            m[j] = b1 * m[j] + (1.f - b1) * gj;
            v[j] = b2 * v[j] + (1.f - b2) * gj * gj;
            float denom = sqrtf(v[j]) / bc2_sqrt + eps;
            p[j] -= (lr / bc1) * (m[j] / denom);
        }
    }
}
void k_adam(cudaStream_t st, float* p, const float* g, float* m, float* v, size_t n, float lr, float wd, float b1, float b2,
            float eps, int step, float grad_scale) {
    SALT_COUNT(1);
    float bc1 = 1.f - powf(b1, (float)step);
    float bc2_sqrt = sqrtf(1.f - powf(b2, (float)step));
    adam_kernel<<<cdiv((long long)cdiv(n, 4), 256), 256, 0, st>>>(p, g, m, v, n, lr, wd, b1, b2, eps, bc1, bc2_sqrt, grad_scale);
}
