// op 3, backward of the single-output-channel layers — PSMNet classif*.2 (Conv3d 32 -> 1, stride 1,
// models/psmnet/stackhourglass.py:99-109) and GC-Net l37 (ConvTranspose3d 32 -> 1, stride 2, models/gcnet.py:60-62).
// With one output channel there is no contraction worth a tensor core: per input voxel v and input channel ci
//     gx[v][ci]  = sum_t gy[s*v + sg*(t-1)] * w[ci][t]          (27 taps t = (kd,kh,kw), per axis)
//     dw[ci][t] += x[v][ci] * gy[s*v + sg*(t-1)]
// with (s, sg) = (1, -1) for the convolution and (2, +1) for the transposed convolution.  Both sums use the same 27
// gathered gy values, so ONE kernel produces the input gradient and the weight gradient: 54 FMAs per (v, ci), fp32.
// (Padding gy to 32 zero channels and reusing the tensor-core kernels — the previous route — moves 32x the bytes.)
//
// x / gx: padded NDHWC bf16 [B][D+2][H+2][W+2][32];  gy: fp32 [B][Do][Ho][Wo] (the layer's output layout);
// w: fp32 [32][27];  dw: fp32 [32][27] (per-CTA partials in `ws`, summed in a fixed order by a second kernel).
// CTA = 64 consecutive voxels of one x row: the 9 gy rows it touches are staged in shared memory (zero outside the
// tensor), lane = channel, a warp walks 16 voxels in pairs (the two voxels' kw windows overlap, so one 128/64-bit
// shared load per row feeds both).
#include "common.cuh"

namespace {

constexpr int C1_TW = 64;          // voxels per tile
constexpr int C1_THREADS = 128;
constexpr int C1_GROW = 2 * C1_TW + 8;   // floats per staged gy row (s = 2: 2*64 + 1 columns), padded

struct C1Geom {
    int B, D, H, W;                // x extent
    int Do, Ho, Wo;                // gy extent
    int s, sg;                     // gy index = s*v + sg*(t-1)
    int tiles_w, ntiles;
};

template <int S>
__global__ void __launch_bounds__(C1_THREADS)
conv3d_c1_bwd_kernel(const __nv_bfloat16* __restrict__ x, const float* __restrict__ gy, const float* __restrict__ w,
                     __nv_bfloat16* __restrict__ gx, float* __restrict__ partial, C1Geom g) {
    __shared__ __align__(16) float s_g[9][C1_GROW];
    __shared__ __align__(16) __nv_bfloat16 s_x[C1_TW][32];
    __shared__ __align__(16) __nv_bfloat16 s_o[C1_TW][32];
    __shared__ float s_dw[32][28];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    constexpr int SG = (S == 2) ? 1 : -1;
    float wr[27], acc[27];
#pragma unroll
    for (int t = 0; t < 27; ++t) { wr[t] = w[lane * 27 + t]; acc[t] = 0.f; }
    const int Wq = g.W + 2, Hq = g.H + 2;

    // Software pipeline over tiles: the global loads of tile i+1 (its 9 gy rows and its x row segment) are issued into
    // registers before tile i is computed and land in shared memory after it, so their latency hides behind the FMAs.
    constexpr int NCOL = S * C1_TW + 2;
    constexpr int NG = (9 * NCOL + C1_THREADS - 1) / C1_THREADS;      // staged gy elements per thread
    float pg[NG];
    uint4 px[2];
    int pnv = 0; size_t pxrow = 0;
    auto decode = [&](int tile, int& b, int& d, int& h, int& w0) {
        int r = tile;
        const int tw = r % g.tiles_w; r /= g.tiles_w;
        h = r % g.H; r /= g.H;
        d = r % g.D; b = r / g.D;
        w0 = tw * C1_TW;
    };
    auto prefetch = [&](int tile) {
        int b, d, h, w0;
        decode(tile, b, d, h, w0);
        pnv = min(C1_TW, g.W - w0);
        const int c0 = S * w0 - 1;             // s_g[row][j] = gy[b][S*d + SG*(kd-1)][S*h + SG*(kh-1)][c0 + j]   (0 outside)
#pragma unroll
        for (int k = 0; k < NG; ++k) {
            const int i = threadIdx.x + k * C1_THREADS;
            const int row = i / NCOL, j = i - row * NCOL;
            const int kd = row / 3, kh = row - kd * 3;
            const int zd = S * d + SG * (kd - 1), zh = S * h + SG * (kh - 1), zw = c0 + j;
            float v = 0.f;
            if (i < 9 * NCOL && zd >= 0 && zd < g.Do && zh >= 0 && zh < g.Ho && zw >= 0 && zw < g.Wo)
                v = __ldg(gy + (((size_t)b * g.Do + zd) * g.Ho + zh) * g.Wo + zw);
            pg[k] = v;
        }
        pxrow = ((((size_t)b * (g.D + 2) + d + 1) * Hq + h + 1) * Wq + w0 + 1) * 32;
        const uint4* src = reinterpret_cast<const uint4*>(x + pxrow);
#pragma unroll
        for (int k = 0; k < 2; ++k) {
            const int i = threadIdx.x + k * C1_THREADS;       // 64 voxels x 4 chunks = 256 chunks of 16 bytes
            px[k] = (i < pnv * 4) ? __ldg(src + i) : make_uint4(0u, 0u, 0u, 0u);
        }
    };
    if ((int)blockIdx.x < g.ntiles) prefetch(blockIdx.x);
    for (int tile = blockIdx.x; tile < g.ntiles; tile += gridDim.x) {
        const int nv = pnv;
        const size_t xrow = pxrow;
#pragma unroll
        for (int k = 0; k < NG; ++k) {
            const int i = threadIdx.x + k * C1_THREADS;
            if (i < 9 * NCOL) { const int row = i / NCOL; s_g[row][i - row * NCOL] = pg[k]; }
        }
#pragma unroll
        for (int k = 0; k < 2; ++k) reinterpret_cast<uint4*>(&s_x[0][0])[threadIdx.x + k * C1_THREADS] = px[k];
        __syncthreads();
        if (tile + (int)gridDim.x < g.ntiles) prefetch(tile + gridDim.x);
        // a warp owns voxels [16*warp, 16*warp+16) of the tile, two at a time
#pragma unroll 1
        for (int p = 0; p < 8; ++p) {
            const int va = warp * 16 + 2 * p;
            if (va >= nv) break;
            const float xa = __bfloat162float(s_x[va][lane]);
            const float xb = (va + 1 < nv) ? __bfloat162float(s_x[va + 1][lane]) : 0.f;
            float oa = 0.f, ob = 0.f;
#pragma unroll
            for (int row = 0; row < 9; ++row) {
                float ga[3], gb[3];            // tap kw = 0,1,2 of voxel a / b
                if (S == 2) {                  // columns 2*va-1 .. 2*va+3  <->  j = 2*va .. 2*va+4
                    const float4 q = *reinterpret_cast<const float4*>(&s_g[row][2 * va]);
                    const float e = s_g[row][2 * va + 4];
                    ga[0] = q.x; ga[1] = q.y; ga[2] = q.z; gb[0] = q.z; gb[1] = q.w; gb[2] = e;
                } else {                       // tap kw reads column v - (kw-1): columns va-1 .. va+2 <-> j = va .. va+3
                    const float2 q0 = *reinterpret_cast<const float2*>(&s_g[row][va]);
                    const float2 q1 = *reinterpret_cast<const float2*>(&s_g[row][va + 2]);
                    ga[0] = q1.x; ga[1] = q0.y; ga[2] = q0.x; gb[0] = q1.y; gb[1] = q1.x; gb[2] = q0.y;
                }
#pragma unroll
                for (int kw = 0; kw < 3; ++kw) {
                    const int t = row * 3 + kw;
                    oa = fmaf(ga[kw], wr[t], oa); ob = fmaf(gb[kw], wr[t], ob);
                    acc[t] = fmaf(xa, ga[kw], fmaf(xb, gb[kw], acc[t]));
                }
            }
            s_o[va][lane] = __float2bfloat16(oa);
            if (va + 1 < nv) s_o[va + 1][lane] = __float2bfloat16(ob);
        }
        __syncthreads();
        {
            uint4* dst = reinterpret_cast<uint4*>(gx + xrow);
            const uint4* src = reinterpret_cast<const uint4*>(&s_o[0][0]);
            for (int i = threadIdx.x; i < nv * 4; i += C1_THREADS) dst[i] = src[i];
        }
        // the next iteration rewrites s_g / s_x before its barrier and s_o after it: every thread has finished its reads
        // of s_g / s_x at the barrier above and finishes this copy-out before it arrives at the next one
    }
    // CTA partial of dw: warps are summed in a fixed order
    for (int wi = 0; wi < C1_THREADS / 32; ++wi) {
        if (warp == wi) {
#pragma unroll
            for (int t = 0; t < 27; ++t) s_dw[lane][t] = (wi == 0 ? 0.f : s_dw[lane][t]) + acc[t];
        }
        __syncthreads();
    }
    for (int i = threadIdx.x; i < 32 * 27; i += C1_THREADS) partial[(size_t)blockIdx.x * 864 + i] = s_dw[i / 27][i % 27];
}

__global__ void __launch_bounds__(256)
c1_reduce_kernel(const float* __restrict__ partial, float* __restrict__ dw, int ncta) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= 864) return;
    float sum = 0.f;
    for (int c = 0; c < ncta; ++c) sum += partial[(size_t)c * 864 + i];
    dw[i] = sum;
}

constexpr int C1_MAX_CTAS = DSM_NUM_SMS_B200 * 8;

}  // namespace

extern "C" size_t dsm_conv3d_c1_bwd_workspace_bytes(void) { return (size_t)C1_MAX_CTAS * 864 * sizeof(float); }

extern "C" int dsm_conv3d_c1_bwd(const void* x, const float* gy, const float* w, void* gx, float* dw,
                                 int B, int D, int H, int W, int Do, int Ho, int Wo, int transposed,
                                 void* ws, size_t ws_bytes, void* stream) {
    DsmDeviceGuard dsm_guard_(x);
    if (!x || !gy || !w || !gx || !dw || !ws || B < 1 || D < 1 || H < 1 || W < 1 || Do < 1 || Ho < 1 || Wo < 1) return DSM_EINVAL;
    if (ws_bytes < dsm_conv3d_c1_bwd_workspace_bytes()) return DSM_EINVAL;
    if (!dsm_aligned16(x) || !dsm_aligned16(gx)) return DSM_EALIGN;
    if (transposed) { if (Do > 2 * D || Ho > 2 * H || Wo > 2 * W) return DSM_EINVAL; }
    else if (Do != D || Ho != H || Wo != W) return DSM_EINVAL;
    C1Geom g;
    g.B = B; g.D = D; g.H = H; g.W = W; g.Do = Do; g.Ho = Ho; g.Wo = Wo;
    g.s = transposed ? 2 : 1; g.sg = transposed ? 1 : -1;
    g.tiles_w = dsm_ceil_div(W, C1_TW);
    const long long nt = (long long)B * D * H * g.tiles_w;
    if (nt > 0x7fffffffLL) return DSM_EUNSUPPORTED;
    g.ntiles = (int)nt;
    const int ncta = (int)(nt < C1_MAX_CTAS ? nt : C1_MAX_CTAS);
    cudaStream_t st = (cudaStream_t)stream;
    float* partial = static_cast<float*>(ws);
    if (transposed)
        conv3d_c1_bwd_kernel<2><<<ncta, C1_THREADS, 0, st>>>(static_cast<const __nv_bfloat16*>(x), gy, w,
                                                             static_cast<__nv_bfloat16*>(gx), partial, g);
    else
        conv3d_c1_bwd_kernel<1><<<ncta, C1_THREADS, 0, st>>>(static_cast<const __nv_bfloat16*>(x), gy, w,
                                                             static_cast<__nv_bfloat16*>(gx), partial, g);
    int rc = dsm_launch_status();
    if (rc != 0) return rc;
    c1_reduce_kernel<<<dsm_ceil_div(864, 256), 256, 0, st>>>(partial, dw, ncta);
    return dsm_launch_status();
}
