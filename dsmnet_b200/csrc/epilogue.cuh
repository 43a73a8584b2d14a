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

// two channels: acc (fp32 bits) * scale + shift, residual pair packed as bf16x2 -> one packed bf16x2
template <int MODE>
__device__ __forceinline__ uint32_t epi_pack2(uint32_t v0, uint32_t v1, float s0, float s1, float h0, float h1, uint32_t rr) {
    float a0, a1;
    ffma2(a0, a1, __uint_as_float(v0), __uint_as_float(v1), s0, s1, h0, h1);
    if (MODE == 4) { a0 = fmaxf(a0, 0.f); a1 = fmaxf(a1, 0.f); }
    if (MODE >= 2) fadd2(a0, a1, a0, a1, bf16_lo(rr), bf16_hi(rr));
    return (MODE == 1 || MODE == 3) ? pack_bf16x2_relu(a0, a1) : pack_bf16x2(a0, a1);   // a final ReLU rides on the conversion
}

// 8 accumulators v[0..7] (fp32 bits), their scale / shift (two float4 each) and 8 bf16 residuals (one uint4) -> 8 bf16 (one uint4)
template <int MODE>
__device__ __forceinline__ uint4 epi_pack8(const uint32_t* v, const float4& s0, const float4& s1, const float4& h0, const float4& h1,
                                           const uint4& rr) {
    uint4 ov;
    ov.x = epi_pack2<MODE>(v[0], v[1], s0.x, s0.y, h0.x, h0.y, rr.x);
    ov.y = epi_pack2<MODE>(v[2], v[3], s0.z, s0.w, h0.z, h0.w, rr.y);
    ov.z = epi_pack2<MODE>(v[4], v[5], s1.x, s1.y, h1.x, h1.y, rr.z);
    ov.w = epi_pack2<MODE>(v[6], v[7], s1.z, s1.w, h1.z, h1.w, rr.w);
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
