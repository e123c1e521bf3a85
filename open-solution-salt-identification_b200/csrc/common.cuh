// Shared definitions for the salt U-Net engine (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>

typedef __nv_bfloat16 bf16;

enum DType { DT_F32 = 0, DT_BF16 = 1 };

static inline size_t dtype_size(DType d) { return d == DT_F32 ? 4 : 2; }

// NHWC activation with optional physical border.  Logical pixel (y,x) lives at
// physical (y+pt, x+pl); physical extent is (H+pt+pb) x (W+pl+pr).
struct Tensor {
    void* p = nullptr;
    int B = 0, H = 0, W = 0, C = 0;
    int pt = 0, pb = 0, pl = 0, pr = 0;
    DType dt = DT_F32;
    int Hp() const { return H + pt + pb; }
    int Wp() const { return W + pl + pr; }
    size_t numel() const { return (size_t)B * Hp() * Wp() * C; }
    size_t bytes() const { return numel() * dtype_size(dt); }
};

// Launch-time dispatch on the activation storage type.
#define SALT_DISPATCH(dt, T, ...)                  \
    do {                                           \
        if ((dt) == DT_F32) { typedef float T; __VA_ARGS__; } \
        else { typedef bf16 T; __VA_ARGS__; }      \
    } while (0)

#ifdef __CUDACC__
__device__ __forceinline__ float4 ld4(const float* p) { return *reinterpret_cast<const float4*>(p); }
__device__ __forceinline__ float4 ld4(const bf16* p) {
    uint2 u = *reinterpret_cast<const uint2*>(p);
    __nv_bfloat162 a = *reinterpret_cast<__nv_bfloat162*>(&u.x);
    __nv_bfloat162 b = *reinterpret_cast<__nv_bfloat162*>(&u.y);
    float2 fa = __bfloat1622float2(a), fb = __bfloat1622float2(b);
    return make_float4(fa.x, fa.y, fb.x, fb.y);
}
__device__ __forceinline__ void st4(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }
__device__ __forceinline__ void st4(bf16* p, float4 v) {
    __nv_bfloat162 a = __floats2bfloat162_rn(v.x, v.y);
    __nv_bfloat162 b = __floats2bfloat162_rn(v.z, v.w);
    uint2 u;
    u.x = *reinterpret_cast<uint32_t*>(&a);
    u.y = *reinterpret_cast<uint32_t*>(&b);
    *reinterpret_cast<uint2*>(p) = u;
}
__device__ __forceinline__ float ld1(const float* p) { return *p; }
__device__ __forceinline__ float ld1(const bf16* p) { return __bfloat162float(*p); }
__device__ __forceinline__ void st1(float* p, float v) { *p = v; }
__device__ __forceinline__ void st1(bf16* p, float v) { *p = __float2bfloat16_rn(v); }

__device__ __forceinline__ float4 f4_add(float4 a, float4 b) { return make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w); }
__device__ __forceinline__ float4 f4_mul(float4 a, float4 b) { return make_float4(a.x * b.x, a.y * b.y, a.z * b.z, a.w * b.w); }
__device__ __forceinline__ float4 f4_fma(float4 a, float4 b, float4 c) {
    return make_float4(fmaf(a.x, b.x, c.x), fmaf(a.y, b.y, c.y), fmaf(a.z, b.z, c.z), fmaf(a.w, b.w, c.w));
}
__device__ __forceinline__ float4 f4_scale(float4 a, float s) { return make_float4(a.x * s, a.y * s, a.z * s, a.w * s); }
__device__ __forceinline__ float4 f4_relu(float4 a) { return make_float4(fmaxf(a.x, 0.f), fmaxf(a.y, 0.f), fmaxf(a.z, 0.f), fmaxf(a.w, 0.f)); }
__device__ __forceinline__ float4 f4_zero() { return make_float4(0.f, 0.f, 0.f, 0.f); }
// keep g where m > 0
__device__ __forceinline__ float4 f4_mask_pos(float4 g, float4 m) {
    return make_float4(m.x > 0.f ? g.x : 0.f, m.y > 0.f ? g.y : 0.f, m.z > 0.f ? g.z : 0.f, m.w > 0.f ? g.w : 0.f);
}
__device__ __forceinline__ float sigmoidf_(float x) { return 1.f / (1.f + __expf(-x)); }
#endif

extern unsigned long long g_salt_launches;   // kernels launched by this library (bench.py: gpu_launches)
#define SALT_COUNT(n) (g_salt_launches += (n))
extern unsigned long long g_salt_cluster_launches;   // of which: thread-block-cluster launches (TMA-multicast convolutions)

constexpr int SALT_STAT_SLOTS = 296;       // forward BatchNorm partial-sum slots per layer (kernels.h): >= the grid of every producer
constexpr int SALT_STAT_SLOTS_CONV = 148;  // slots a persistent tensor-core convolution can touch (its grid is capped at this)
constexpr int SALT_STAT_SLOTS_BWD = 1184;  // backward: up to 8 blocks per SM in the reduction passes, one slot per block

// ---- optional per-kernel-class device timing (Engine::profile_enable; CUDA events on the launch stream).  Classes 0-2 are the
// convolutions (engine.cu), the rest the memory-bound passes; `work` = algorithmic FLOPs (convolutions) or bytes (passes).
enum SaltProfClass { SALT_PROF_CONV_FWD = 0, SALT_PROF_CONV_DGRAD = 1, SALT_PROF_CONV_WGRAD = 2, SALT_PROF_BN_APPLY = 3,
                     SALT_PROF_BN_BWD_REDUCE = 4, SALT_PROF_BN_BWD_APPLY = 5, SALT_PROF_BN_FINALIZE = 6, SALT_PROF_GATHER = 7,
                     SALT_PROF_GATHER_BWD = 8, SALT_PROF_SCSE = 9, SALT_PROF_OTHER = 10, SALT_PROF_NCLASS = 11 };
extern void (*g_salt_prof_begin)(int cls, double work, cudaStream_t st);
extern void (*g_salt_prof_end)(cudaStream_t st);
struct SaltProfScope {
    cudaStream_t st; bool on;
    SaltProfScope(int cls, double work, cudaStream_t s) : st(s), on(g_salt_prof_begin != nullptr) { if (on) g_salt_prof_begin(cls, work, s); }
    ~SaltProfScope() { if (on && g_salt_prof_end) g_salt_prof_end(st); }
};

static inline int cdiv(long long a, long long b) { return (int)((a + b - 1) / b); }
