// Fused epilogue description of the tensor-core convolutions (host-visible; the device code is in tc_common.cuh).
#pragma once
struct EpiParams {
    const float* scale = nullptr;      // [Co]  y = (acc + bias) * scale + shift
    const float* shift = nullptr;
    const void* res = nullptr;         // residual, same type as the output, logical [B][Ho][Wo][Co]
    int relu = 0;
    int Hp = 0, Wp = 0, pt = 0, pl = 0;    // physical output tensor [B][Hp][Wp][Co]; logical pixel (y, x) lives at (y + pt, x + pl),
                                           // the border replicates the nearest logical pixel
};
