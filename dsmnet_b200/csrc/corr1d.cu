// op 1 — 1-D correlation layer (DispNetC / iResNet).
// Replaces Corr1d.forward, reference models/util_conv.py:71-81 (simfun_default :68-69):
//   out[b,d,y,x] = sum_c fL[b,c,y,x] * fR[b,c,y,x-d*s]   for x >= d*s (and d < W), else 0.
//
// Forward kernel: one CTA per (32-wide x tile, row y, pair b).  The left strip [C][32] and
// the right strip [C][32 + halo] (halo = every shift the tile can see) are staged in shared
// memory once; warp g owns disparities 8g..8g+7; inside a warp the 32 lanes are 8 x-quads
// times 4 channel groups, every lane keeps an 8(d) x 4(x) register tile and walks its quarter
// of the channels with 128-bit shared loads (1 for the left quad, ceil((7s+4)/4) for the
// sliding right window -> 32 FMAs per 4-6 LDS.128).  The four channel-group partials are
// combined with two rounds of warp shuffles and each lane stores two of the eight d rows
// as 128-bit coalesced writes.  Algorithmic HBM bytes: 4*B*H*W*(2C + D).
#include "common.cuh"

namespace {

constexpr int TX = 32;   // x positions per CTA
constexpr int DT = 8;    // disparities per warp

template <int S>
__global__ void __launch_bounds__(512)
corr1d_fwd_kernel(const float* __restrict__ fL, const float* __restrict__ fR, float* __restrict__ out,
                  int C, int H, int W, int D, int CC /*channels per smem chunk*/, int HT, int RW) {
    extern __shared__ __align__(16) float smem[];
    float* sL = smem;                 // [CC][TX]
    float* sR = smem + CC * TX;       // [CC][RW], element e <-> x = x0 + e - HT

    const int x0 = blockIdx.x * TX;
    const int y = blockIdx.y;
    const int b = blockIdx.z;
    const int tid = threadIdx.x;
    const int nthreads = blockDim.x;
    const int warp = tid >> 5, lane = tid & 31;
    const int xq = lane & 7, cg = lane >> 3;
    const int d0 = warp * DT;

    constexpr int WL = 7 * S + 4;            // window length in floats
    constexpr int WV = (WL + 3) / 4;         // in float4

    float acc[DT][4];
#pragma unroll
    for (int i = 0; i < DT; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

    const size_t plane = (size_t)H * W;
    const float* baseL = fL + ((size_t)b * C * H + y) * W;
    const float* baseR = fR + ((size_t)b * C * H + y) * W;

    for (int c0 = 0; c0 < C; c0 += CC) {
        const int cc = min(CC, C - c0);
        if (c0 > 0) __syncthreads();
        // stage the strips (coalesced scalar loads; zero outside the row)
        for (int i = tid; i < cc * TX; i += nthreads) {
            int c = i / TX, e = i - c * TX;
            int x = x0 + e;
            sL[c * TX + e] = (x < W) ? __ldg(baseL + (size_t)(c0 + c) * plane + x) : 0.f;
        }
        for (int i = tid; i < cc * RW; i += nthreads) {
            int c = i / RW, e = i - c * RW;
            int x = x0 + e - HT;
            sR[c * RW + e] = (x >= 0 && x < W) ? __ldg(baseR + (size_t)(c0 + c) * plane + x) : 0.f;
        }
        __syncthreads();

        // window start for this lane: e0 = 4*xq + HT - (d0 + 7)*S   (multiple of 4 by construction)
        const int e0 = 4 * xq + HT - (d0 + DT - 1) * S;
        for (int c = cg; c < cc; c += 4) {
            const float4 l4 = *reinterpret_cast<const float4*>(sL + c * TX + 4 * xq);
            const float l[4] = {l4.x, l4.y, l4.z, l4.w};
            float win[WV * 4];
            const float4* wp = reinterpret_cast<const float4*>(sR + c * RW + e0);
#pragma unroll
            for (int v = 0; v < WV; ++v) {
                float4 t = wp[v];
                win[4 * v + 0] = t.x; win[4 * v + 1] = t.y; win[4 * v + 2] = t.z; win[4 * v + 3] = t.w;
            }
#pragma unroll
            for (int dd = 0; dd < DT; ++dd)
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    acc[dd][j] = fmaf(l[j], win[(DT - 1 - dd) * S + j], acc[dd][j]);
        }
    }

    // combine the 4 channel groups (lanes differing in bits 3,4)
#pragma unroll
    for (int dd = 0; dd < DT; ++dd)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            float v = acc[dd][j];
            v += __shfl_xor_sync(0xffffffffu, v, 8);
            v += __shfl_xor_sync(0xffffffffu, v, 16);
            acc[dd][j] = v;
        }

    // lane (cg) stores rows dd = 2cg, 2cg+1
    const int x = x0 + 4 * xq;
    const bool vec_ok = ((W & 3) == 0);
#pragma unroll
    for (int dd = 0; dd < DT; ++dd) {
        if ((dd >> 1) != cg) continue;
        const int d = d0 + dd;
        if (d >= D || x >= W) continue;
        float* o = out + (((size_t)b * D + d) * H + y) * W + x;
        if (vec_ok) {
            st_stream_f4(reinterpret_cast<float4*>(o), make_float4(acc[dd][0], acc[dd][1], acc[dd][2], acc[dd][3]));
        } else {
#pragma unroll
            for (int j = 0; j < 4; ++j)
                if (x + j < W) o[j] = acc[dd][j];
        }
    }
}

// Backward w.r.t. the left features:
//   gL[b,c,y,x] = sum_{d<D, d*s<=x} g[b,d,y,x] * fR[b,c,y,x-d*s]
// CTA = (32-wide x tile, y, b); 256 threads = 32 channels x 8 x-quads; channels in chunks of 32.
__global__ void __launch_bounds__(256)
corr1d_bwd_left_kernel(const float* __restrict__ g, const float* __restrict__ fR, float* __restrict__ gL,
                       int C, int H, int W, int D, int S, int HT, int RW) {
    extern __shared__ __align__(16) float smem[];
    float* sG = smem;                 // [D][TX]
    float* sR = smem + D * TX;        // [32][RW]
    const int x0 = blockIdx.x * TX, y = blockIdx.y, b = blockIdx.z;
    const int tid = threadIdx.x;
    const int cl = tid >> 3, xq = tid & 7;

    for (int i = tid; i < D * TX; i += 256) {
        int d = i / TX, e = i - d * TX;
        int x = x0 + e;
        sG[i] = (x < W) ? __ldg(g + (((size_t)b * D + d) * H + y) * W + x) : 0.f;
    }
    for (int c0 = 0; c0 < C; c0 += 32) {
        __syncthreads();
        for (int i = tid; i < 32 * RW; i += 256) {
            int c = i / RW, e = i - c * RW;
            int x = x0 + e - HT;
            sR[i] = (c0 + c < C && x >= 0 && x < W)
                        ? __ldg(fR + (((size_t)b * C + c0 + c) * H + y) * W + x) : 0.f;
        }
        __syncthreads();
        float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
        const float* r = sR + cl * RW + HT + 4 * xq;
        for (int d = 0; d < D; ++d) {
            const float4 g4 = *reinterpret_cast<const float4*>(sG + d * TX + 4 * xq);
            const float* rr = r - d * S;
            a0 = fmaf(g4.x, rr[0], a0); a1 = fmaf(g4.y, rr[1], a1);
            a2 = fmaf(g4.z, rr[2], a2); a3 = fmaf(g4.w, rr[3], a3);
        }
        const int c = c0 + cl, x = x0 + 4 * xq;
        if (c < C && x < W) {
            float* o = gL + (((size_t)b * C + c) * H + y) * W + x;
            const float a[4] = {a0, a1, a2, a3};
#pragma unroll
            for (int j = 0; j < 4; ++j) if (x + j < W) o[j] = a[j];
        }
    }
}

// Backward w.r.t. the right features:
//   gR[b,c,y,x'] = sum_{d<D, x'+d*s<W} g[b,d,y,x'+d*s] * fL[b,c,y,x'+d*s]
__global__ void __launch_bounds__(256)
corr1d_bwd_right_kernel(const float* __restrict__ g, const float* __restrict__ fL, float* __restrict__ gR,
                        int C, int H, int W, int D, int S, int RW) {
    extern __shared__ __align__(16) float smem[];
    float* sG = smem;                 // [D][RW] : element e <-> x = x0 + e
    float* sLt = smem + D * RW;       // [32][RW]
    const int x0 = blockIdx.x * TX, y = blockIdx.y, b = blockIdx.z;
    const int tid = threadIdx.x;
    const int cl = tid >> 3, xq = tid & 7;

    for (int i = tid; i < D * RW; i += 256) {
        int d = i / RW, e = i - d * RW;
        int x = x0 + e;
        sG[i] = (x < W) ? __ldg(g + (((size_t)b * D + d) * H + y) * W + x) : 0.f;
    }
    for (int c0 = 0; c0 < C; c0 += 32) {
        __syncthreads();
        for (int i = tid; i < 32 * RW; i += 256) {
            int c = i / RW, e = i - c * RW;
            int x = x0 + e;
            sLt[i] = (c0 + c < C && x < W) ? __ldg(fL + (((size_t)b * C + c0 + c) * H + y) * W + x) : 0.f;
        }
        __syncthreads();
        float a[4] = {0.f, 0.f, 0.f, 0.f};
        const float* l = sLt + cl * RW + 4 * xq;
        for (int d = 0; d < D; ++d) {
            const float* gg = sG + d * RW + 4 * xq + d * S;
            const float* ll = l + d * S;
#pragma unroll
            for (int j = 0; j < 4; ++j) a[j] = fmaf(gg[j], ll[j], a[j]);
        }
        const int c = c0 + cl, x = x0 + 4 * xq;
        if (c < C && x < W) {
            float* o = gR + (((size_t)b * C + c) * H + y) * W + x;
#pragma unroll
            for (int j = 0; j < 4; ++j) if (x + j < W) o[j] = a[j];
        }
    }
}

}  // namespace

extern "C" int dsm_corr1d_fwd(const float* fL, const float* fR, float* out,
                              int B, int C, int H, int W, int D, int stride, void* stream) {
    if (!fL || !fR || !out || B <= 0 || C <= 0 || H <= 0 || W <= 0 || D <= 0) return DSM_EINVAL;
    if (stride != 1 && stride != 2) return DSM_EUNSUPPORTED;
    if (H > 65535 || B > 65535) return DSM_EUNSUPPORTED;
    if (!dsm_aligned16(out)) return DSM_EALIGN;
    const int NW = dsm_ceil_div(D, DT);
    if (NW > 16) return DSM_EUNSUPPORTED;          // D <= 128 (one warp per 8 disparities)
    const int HT = (DT * NW - 1) * stride;         // left halo; (HT - 7s) % 4 == 0
    const int RW = (HT + TX + 4 + 3) & ~3;
    int CC = (100 * 1024 / 4) / (TX + RW);
    CC &= ~3;
    if (CC < 4) return DSM_EUNSUPPORTED;
    if (CC > C) CC = (C + 3) & ~3;
    const size_t smem = (size_t)CC * (TX + RW) * sizeof(float);
    dim3 grid(dsm_ceil_div(W, TX), H, B), block(NW * 32);
    cudaStream_t st = (cudaStream_t)stream;
    cudaError_t e;
    if (stride == 1) {
        e = cudaFuncSetAttribute(corr1d_fwd_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return (int)e;
        corr1d_fwd_kernel<1><<<grid, block, smem, st>>>(fL, fR, out, C, H, W, D, CC, HT, RW);
    } else {
        e = cudaFuncSetAttribute(corr1d_fwd_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return (int)e;
        corr1d_fwd_kernel<2><<<grid, block, smem, st>>>(fL, fR, out, C, H, W, D, CC, HT, RW);
    }
    return dsm_launch_status();
}

extern "C" int dsm_corr1d_bwd(const float* gout, const float* fL, const float* fR, float* gL, float* gR,
                              int B, int C, int H, int W, int D, int stride, void* stream) {
    if (!gout || !fL || !fR || !gL || !gR || B <= 0 || C <= 0 || H <= 0 || W <= 0 || D <= 0) return DSM_EINVAL;
    if (stride < 1 || stride > 4) return DSM_EUNSUPPORTED;
    if (H > 65535 || B > 65535) return DSM_EUNSUPPORTED;
    cudaStream_t st = (cudaStream_t)stream;
    const int HT = (D - 1) * stride;
    const int RW = (HT + TX + 3) & ~3;
    dim3 grid(dsm_ceil_div(W, TX), H, B), block(256);
    {
        const size_t smem = ((size_t)D * TX + 32 * (size_t)RW) * sizeof(float);
        if (smem > 200 * 1024) return DSM_EUNSUPPORTED;
        cudaError_t e = cudaFuncSetAttribute(corr1d_bwd_left_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return (int)e;
        corr1d_bwd_left_kernel<<<grid, block, smem, st>>>(gout, fR, gL, C, H, W, D, stride, HT, RW);
    }
    {
        const size_t smem = ((size_t)D * RW + 32 * (size_t)RW) * sizeof(float);
        if (smem > 200 * 1024) return DSM_EUNSUPPORTED;
        cudaError_t e = cudaFuncSetAttribute(corr1d_bwd_right_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return (int)e;
        corr1d_bwd_right_kernel<<<grid, block, smem, st>>>(gout, fL, gR, C, H, W, D, stride, RW);
    }
    return dsm_launch_status();
}
