// Epilogue arithmetic of the tcgen05 convolution kernels, specialised at compile time on what the layer needs.
// The accumulators of 8 channels become  act(acc * scale + shift [+ residual])  in bf16.  A layer either adds a residual or
// not and applies ReLU before the add (GC-Net skips, gcnet.py:78-96), after it (PSMNet, stackhourglass.py:46-58) or never;
// evaluating that per element at run time costs two FMNMX and one FADD per element whether needed or not — 45 % of the
// epilogue's arithmetic instructions — so the kernels switch ONCE per accumulator block on a uniform mode instead.
#pragma once
#include "common.cuh"

// 0: affine   1: affine, ReLU   2: affine + residual   3: ReLU(affine + residual)   4: ReLU(affine) + residual
__host__ __device__ __forceinline__ int epi_mode(int relu, bool has_res) {
    return has_res ? (relu == 1 ? 3 : (relu == 2 ? 4 : 2)) : (relu ? 1 : 0);
}

template <int MODE>
__device__ __forceinline__ float epi_act(float a, float r) {
    if (MODE == 1) return fmaxf(a, 0.f);
    if (MODE == 2) return a + r;
    if (MODE == 3) return fmaxf(a + r, 0.f);
    if (MODE == 4) return fmaxf(a, 0.f) + r;
    return a;
}

// 8 accumulators v[0..7] (fp32 bits), their scale / shift (two float4 each) and 8 bf16 residuals (one uint4) -> 8 bf16 (one uint4)
template <int MODE>
__device__ __forceinline__ uint4 epi_pack8(const uint32_t* v, const float4& s0, const float4& s1, const float4& h0, const float4& h1,
                                           const uint4& rr) {
    float f[8];
    f[0] = epi_act<MODE>(fmaf(__uint_as_float(v[0]), s0.x, h0.x), bf16_lo(rr.x));
    f[1] = epi_act<MODE>(fmaf(__uint_as_float(v[1]), s0.y, h0.y), bf16_hi(rr.x));
    f[2] = epi_act<MODE>(fmaf(__uint_as_float(v[2]), s0.z, h0.z), bf16_lo(rr.y));
    f[3] = epi_act<MODE>(fmaf(__uint_as_float(v[3]), s0.w, h0.w), bf16_hi(rr.y));
    f[4] = epi_act<MODE>(fmaf(__uint_as_float(v[4]), s1.x, h1.x), bf16_lo(rr.z));
    f[5] = epi_act<MODE>(fmaf(__uint_as_float(v[5]), s1.y, h1.y), bf16_hi(rr.z));
    f[6] = epi_act<MODE>(fmaf(__uint_as_float(v[6]), s1.z, h1.z), bf16_lo(rr.w));
    f[7] = epi_act<MODE>(fmaf(__uint_as_float(v[7]), s1.w, h1.w), bf16_hi(rr.w));
    uint4 ov;
    ov.x = pack_bf16x2(f[0], f[1]); ov.y = pack_bf16x2(f[2], f[3]);
    ov.z = pack_bf16x2(f[4], f[5]); ov.w = pack_bf16x2(f[6], f[7]);
    return ov;
}

// NV groups of 8 channels (NV even): accumulators v[8*NV], scale / shift as float4 arrays starting at the block's first
// channel, residuals rv[NV]; 32-byte aligned output, one 256-bit store per 16 channels.
template <int MODE, int NV>
__device__ __forceinline__ void epi_store256(const uint32_t* v, const float4* sc4, const float4* sh4, const uint4* rv, uint4* out) {
#pragma unroll
    for (int c = 0; c < NV; c += 2) {
        const uint4 a = epi_pack8<MODE>(v + 8 * c, sc4[2 * c], sc4[2 * c + 1], sh4[2 * c], sh4[2 * c + 1], rv[c]);
        const uint4 b = epi_pack8<MODE>(v + 8 * c + 8, sc4[2 * c + 2], sc4[2 * c + 3], sh4[2 * c + 2], sh4[2 * c + 3], rv[c + 1]);
        st_v8(out + c, a, b);
    }
}

// the same with 128-bit loads of the residual (may be NULL) and 128-bit stores: 16-byte aligned buffers
template <int MODE, int NV>
__device__ __forceinline__ void epi_store128(const uint32_t* v, const float* sc, const float* sh, const uint4* res, uint4* out) {
    const float4* sc4 = reinterpret_cast<const float4*>(sc);
    const float4* sh4 = reinterpret_cast<const float4*>(sh);
#pragma unroll
    for (int c = 0; c < NV; ++c) {
        const uint4 rr = (MODE >= 2) ? __ldg(res + c) : make_uint4(0u, 0u, 0u, 0u);
        out[c] = epi_pack8<MODE>(v + 8 * c, sc4[2 * c], sc4[2 * c + 1], sh4[2 * c], sh4[2 * c + 1], rr);
    }
}

#define DSM_EPI_DISPATCH(mode, FN, NV, ...)                      \
    switch (mode) {                                              \
        case 1: FN<1, NV>(__VA_ARGS__); break;                   \
        case 2: FN<2, NV>(__VA_ARGS__); break;                   \
        case 3: FN<3, NV>(__VA_ARGS__); break;                   \
        case 4: FN<4, NV>(__VA_ARGS__); break;                   \
        default: FN<0, NV>(__VA_ARGS__); break;                  \
    }
