// Fused SSIM map of the self-supervised photometric loss (reference losses/SSIM.py:24-42 `_ssim` as called from
// losses/loss.py:196-236): the channel-averaged 11x11 Gaussian (sigma 1.5) SSIM of two [B][C][H][W] images,
//   mu1 = G*mean_c(a), mu2 = G*mean_c(b), E11 = G*mean_c(a^2), E22 = G*mean_c(b^2), E12 = G*mean_c(a*b)   (zero padding)
//   s   = (2 mu1 mu2 + C1)(2 (E12 - mu1 mu2) + C2) / ((mu1^2 + mu2^2 + C1)(E11 - mu1^2 + E22 - mu2^2 + C2))
// The reference runs five F.conv2d (11x11, C -> 1) plus ~15 elementwise kernels and keeps every intermediate for autograd;
// here one kernel reads a and b once per tile (+ halo), filters the five channel means separably in shared memory and
// writes s.  Backward (the gradient flows to b = the warped image only; a is the real image): one kernel recomputes the five
// filtered means and writes the three per-pixel partials gs*ds/d(mu2, E22, E12), a second one filters those (the Gaussian is
// symmetric, so the adjoint of the zero-padded filter is the filter itself) and combines
//   gb_c = ( G*P_mu + 2 b_c G*P_E22 + a_c G*P_E12 ) / C.
// HBM-bound streams: fwd 4*(2C+1) bytes per pixel, bwd 4*(2C+1+3) + 4*(3+2C+C).
#include "common.cuh"

namespace {

constexpr int TX = 32, TY = 16, R = 5, K = 11;
constexpr int IW = TX + 2 * R, IH = TY + 2 * R;          // 42 x 26 input region
constexpr int THREADS = 256;

struct Gauss { float g[K]; };

__host__ Gauss make_gauss() {
    Gauss G; double s = 0, v[K];
    for (int i = 0; i < K; ++i) { v[i] = exp(-(double)((i - K / 2) * (i - K / 2)) / (2.0 * 1.5 * 1.5)); s += v[i]; }
    // the reference builds the window in fp32: gauss/gauss.sum() (SSIM.py:6-8), then the outer product (:11-12)
    float f[K], fs = 0.f;
    for (int i = 0; i < K; ++i) { f[i] = (float)v[i]; fs += f[i]; }
    for (int i = 0; i < K; ++i) G.g[i] = f[i] / fs;
    (void)s;
    return G;
}

// separable filter of NQ quantities held in s_in[NQ][IH][IW]; result for output pixel (ox, oy) of the tile -> out[NQ]
template <int NQ>
__device__ __forceinline__ void hpass(const float (*s_in)[IH][IW], float (*s_h)[IH][TX], const Gauss& G) {
    for (int i = threadIdx.x; i < NQ * IH * TX; i += THREADS) {
        const int x = i % TX, y = (i / TX) % IH, q = i / (TX * IH);
        float acc = 0.f;
#pragma unroll
        for (int k = 0; k < K; ++k) acc = fmaf(G.g[k], s_in[q][y][x + k], acc);
        s_h[q][y][x] = acc;
    }
}
template <int NQ>
__device__ __forceinline__ void vpass(const float (*s_h)[IH][TX], int x, int y, const Gauss& G, float* out) {
#pragma unroll
    for (int q = 0; q < NQ; ++q) {
        float acc = 0.f;
#pragma unroll
        for (int k = 0; k < K; ++k) acc = fmaf(G.g[k], s_h[q][y + k][x], acc);
        out[q] = acc;
    }
}

__device__ __forceinline__ void load_means(const float* __restrict__ a, const float* __restrict__ b, float (*s_in)[IH][IW],
                                           int C, int H, int W, int x0, int y0, size_t img) {
    const float invC = 1.f / (float)C;
    const size_t hw = (size_t)H * W;
    for (int i = threadIdx.x; i < IH * IW; i += THREADS) {
        const int lx = i % IW, ly = i / IW;
        const int x = x0 + lx - R, y = y0 + ly - R;
        float m1 = 0.f, m2 = 0.f, e11 = 0.f, e22 = 0.f, e12 = 0.f;
        if (x >= 0 && x < W && y >= 0 && y < H) {
            const size_t o = img + (size_t)y * W + x;
            for (int c = 0; c < C; ++c) {
                const float va = __ldg(a + o + c * hw), vb = __ldg(b + o + c * hw);
                m1 += va; m2 += vb; e11 = fmaf(va, va, e11); e22 = fmaf(vb, vb, e22); e12 = fmaf(va, vb, e12);
            }
        }
        s_in[0][ly][lx] = m1 * invC; s_in[1][ly][lx] = m2 * invC; s_in[2][ly][lx] = e11 * invC;
        s_in[3][ly][lx] = e22 * invC; s_in[4][ly][lx] = e12 * invC;
    }
}

constexpr float kC1 = 0.01f * 0.01f, kC2 = 0.03f * 0.03f;

// MODE 0: s map.  MODE 1: the three partials (times gs) for the backward pass.
template <int MODE>
__global__ void __launch_bounds__(THREADS)
ssim_kernel(const float* __restrict__ a, const float* __restrict__ b, const float* __restrict__ gs,
            float* __restrict__ out, int C, int H, int W, Gauss G) {
    __shared__ float s_in[5][IH][IW];
    __shared__ float s_h[5][IH][TX];
    const int x0 = blockIdx.x * TX, y0 = blockIdx.y * TY, bz = blockIdx.z;
    const size_t hw = (size_t)H * W;
    load_means(a, b, s_in, C, H, W, x0, y0, (size_t)bz * C * hw);
    __syncthreads();
    hpass<5>(s_in, s_h, G);
    __syncthreads();
    for (int i = threadIdx.x; i < TX * TY; i += THREADS) {
        const int lx = i % TX, ly = i / TX;
        const int x = x0 + lx, y = y0 + ly;
        if (x >= W || y >= H) continue;
        float v[5];
        vpass<5>(s_h, lx, ly, G, v);
        const float mu1 = v[0], mu2 = v[1];
        const float mu12 = mu1 * mu2, mu1sq = mu1 * mu1, mu2sq = mu2 * mu2;
        const float A1 = 2.f * mu12 + kC1, A2 = 2.f * (v[4] - mu12) + kC2;
        const float B1 = mu1sq + mu2sq + kC1, B2 = (v[2] - mu1sq) + (v[3] - mu2sq) + kC2;
        const size_t o = (size_t)bz * hw + (size_t)y * W + x;
        if (MODE == 0) {
            out[o] = (A1 * A2) / (B1 * B2);
        } else {
            const float g = __ldg(gs + o);
            const float inv = 1.f / (B1 * B2);
            const float s = A1 * A2 * inv;
            const size_t n = (size_t)gridDim.z * hw;
            out[o]         = g * (2.f * mu1 * (A2 - A1) * inv - s * 2.f * mu2 * (B2 - B1) * inv);   // d s / d mu2
            out[n + o]     = g * (-s / B2);                                                          // d s / d E22
            out[2 * n + o] = g * (2.f * A1 * inv);                                                   // d s / d E12
        }
    }
}

__global__ void __launch_bounds__(THREADS)
ssim_bwd_combine_kernel(const float* __restrict__ a, const float* __restrict__ b, const float* __restrict__ part,
                        float* __restrict__ gb, int C, int H, int W, Gauss G) {
    __shared__ float s_in[3][IH][IW];
    __shared__ float s_h[3][IH][TX];
    const int x0 = blockIdx.x * TX, y0 = blockIdx.y * TY, bz = blockIdx.z;
    const size_t hw = (size_t)H * W, n = (size_t)gridDim.z * hw;
    for (int i = threadIdx.x; i < IH * IW; i += THREADS) {
        const int lx = i % IW, ly = i / IW;
        const int x = x0 + lx - R, y = y0 + ly - R;
        const bool in = x >= 0 && x < W && y >= 0 && y < H;
        const size_t o = (size_t)bz * hw + (size_t)y * W + x;
#pragma unroll
        for (int q = 0; q < 3; ++q) s_in[q][ly][lx] = in ? __ldg(part + q * n + o) : 0.f;
    }
    __syncthreads();
    hpass<3>(s_in, s_h, G);
    __syncthreads();
    const float invC = 1.f / (float)C;
    for (int i = threadIdx.x; i < TX * TY; i += THREADS) {
        const int lx = i % TX, ly = i / TX;
        const int x = x0 + lx, y = y0 + ly;
        if (x >= W || y >= H) continue;
        float v[3];
        vpass<3>(s_h, lx, ly, G, v);
        const size_t o = (size_t)bz * C * hw + (size_t)y * W + x;
        for (int c = 0; c < C; ++c) {
            const float va = __ldg(a + o + c * hw), vb = __ldg(b + o + c * hw);
            gb[o + c * hw] = (v[0] + 2.f * vb * v[1] + va * v[2]) * invC;
        }
    }
}

}  // namespace

extern "C" int dsm_ssim_fwd(const float* a, const float* b, float* s, int B, int C, int H, int W, void* stream) {
    DsmDeviceGuard dsm_guard_(a);
    if (!a || !b || !s || B < 1 || C < 1 || H < 1 || W < 1) return DSM_EINVAL;
    if (B > 65535 || dsm_ceil_div(H, TY) > 65535) return DSM_EUNSUPPORTED;
    static const Gauss G = make_gauss();
    ssim_kernel<0><<<dim3(dsm_ceil_div(W, TX), dsm_ceil_div(H, TY), B), THREADS, 0, (cudaStream_t)stream>>>(a, b, nullptr, s, C, H, W, G);
    return dsm_launch_status();
}

extern "C" size_t dsm_ssim_bwd_workspace_bytes(int B, int H, int W) { return (size_t)3 * B * H * W * sizeof(float); }

extern "C" int dsm_ssim_bwd(const float* a, const float* b, const float* gs, float* gb, int B, int C, int H, int W,
                            void* ws, size_t ws_bytes, void* stream) {
    DsmDeviceGuard dsm_guard_(a);
    if (!a || !b || !gs || !gb || !ws || B < 1 || C < 1 || H < 1 || W < 1) return DSM_EINVAL;
    if (ws_bytes < dsm_ssim_bwd_workspace_bytes(B, H, W)) return DSM_EINVAL;
    if (B > 65535 || dsm_ceil_div(H, TY) > 65535) return DSM_EUNSUPPORTED;
    static const Gauss G = make_gauss();
    const dim3 grid(dsm_ceil_div(W, TX), dsm_ceil_div(H, TY), B);
    cudaStream_t st = (cudaStream_t)stream;
    float* part = static_cast<float*>(ws);
    ssim_kernel<1><<<grid, THREADS, 0, st>>>(a, b, gs, part, C, H, W, G);
    ssim_bwd_combine_kernel<<<grid, THREADS, 0, st>>>(a, b, part, gb, C, H, W, G);
    return dsm_launch_status();
}
