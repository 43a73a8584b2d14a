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
#include "ptx.cuh"
#include "tma_host.cuh"

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

// ---------------------------------------------------------------------------------------------
// Fast forward path (H*W % 16 == 0, 16-byte aligned pointers, D <= 112).
//
// At the reference's sizes the op is FMA-bound, not HBM-bound (C=128, D=41: 2*D*C flops per
// 4*(2C+D) bytes = 8.8 flop/B, the fp32 ridge of a B200 is ~11), and it is small: 96x312x41 outputs
// are ~1.2 M accumulators.  So the design goal is one FMA-dense warp per SM sub-partition:
//   * pixels are FLATTENED over (y,x): a channel plane is contiguous, a tile is 256 consecutive
//     pixels and its right-feature window [p0-128, p0+256) is contiguous too; positions that wrap into
//     the previous row are exactly the ones the reference zero-fills (x < d*s), masked at the store;
//   * a thread owns 8 consecutive pixels x 14 disparities (112 accumulators): per channel it reads
//     8 left values and a sliding window of 21 (stride 2: 34) right values with 128-bit shared loads
//     -> 112 FMAs per 8 (12) LDS.128, i.e. 1.1 B of shared traffic per FMA (the SM sustains 1.0);
//     warp w of the CTA owns disparities 14w..14w+13;
//   * a producer lane fills a 4-stage ring of 8-channel chunks with TWO TMA loads per stage: the
//     planes are viewed as [B*C][H*W/16][16 floats] and loaded with SWIZZLE_64B, which makes the
//     32-byte lane stride of the 128-bit loads bank-conflict free; out-of-range rows are zero-filled.
// ---------------------------------------------------------------------------------------------
constexpr int FT = 256;     // pixels per CTA
constexpr int FHT = 128;    // right-window halo (>= (8*14-1)*... see dispatch), keeps channel blocks multiples of 512 B
constexpr int FRW = FT + FHT;
constexpr int FDT = 14;     // disparities per warp
constexpr int FCC = 8;      // channels per stage
constexpr int FSTAGES = 4;
constexpr int FSTAGE_FLOATS = FCC * (FT + FRW);

// SWIZZLE_64B on rows of 16 floats: 16-byte chunk index ^= bits [7,9) of the byte address
__device__ __forceinline__ int fswz(int e) { return e ^ (((e >> 5) & 3) << 2); }

// bounded mbarrier wait: a broken pipeline must fault (trap), never hang the GPU
__device__ __forceinline__ void wait_or_trap(uint32_t bar, uint32_t parity) {
    for (uint32_t spins = 0; !ptx::mbar_try_wait(bar, parity); ++spins)
        if (spins > (1u << 24)) __trap();
}

template <int S, int MIS>
__device__ __forceinline__ void corr_flat_chunk(const float* sL, const float* sR, int cc, int lane, int ea, float (&acc)[FDT][8]) {
    constexpr int WL = 8 + (FDT - 1) * S + 3;       // floats that may be needed from the aligned start
    constexpr int NV = (WL + 3) / 4;
    int qoff[NV];
#pragma unroll
    for (int k = 0; k < NV; ++k) qoff[k] = fswz(ea + 4 * k);
    const int loff0 = fswz(8 * lane), loff1 = fswz(8 * lane + 4);
#pragma unroll 2
    for (int c = 0; c < cc; ++c) {
        const float4 l0 = *reinterpret_cast<const float4*>(sL + c * FT + loff0);
        const float4 l1 = *reinterpret_cast<const float4*>(sL + c * FT + loff1);
        const float l[8] = {l0.x, l0.y, l0.z, l0.w, l1.x, l1.y, l1.z, l1.w};
        float win[NV * 4];
#pragma unroll
        for (int k = 0; k < NV; ++k) {
            const float4 t = *reinterpret_cast<const float4*>(sR + c * FRW + qoff[k]);
            win[4 * k] = t.x; win[4 * k + 1] = t.y; win[4 * k + 2] = t.z; win[4 * k + 3] = t.w;
        }
#pragma unroll
        for (int dd = 0; dd < FDT; ++dd)
#pragma unroll
            for (int j = 0; j < 8; ++j)
                acc[dd][j] = fmaf(l[j], win[MIS + (FDT - 1 - dd) * S + j], acc[dd][j]);
    }
}

template <int S>
__global__ void __launch_bounds__(288, 1)
corr1d_flat_kernel(const __grid_constant__ CUtensorMap mapL, const __grid_constant__ CUtensorMap mapR,
                   float* __restrict__ out, int C, int HW, int W, int D, int NW) {
    extern __shared__ uint8_t smem_raw[];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t raw = ptx::smem_u32(smem_raw);
    const uint32_t base = (raw + 1023u) & ~1023u;
    const float* ring = reinterpret_cast<const float*>(smem_raw + (base - raw));
    const uint32_t bars = base + FSTAGES * FSTAGE_FLOATS * 4;     // full[S], empty[S]
    auto full_bar = [&](int s) { return bars + 8u * s; };
    auto empty_bar = [&](int s) { return bars + 8u * (FSTAGES + s); };
    const int b = blockIdx.y;
    const int p0 = blockIdx.x * FT;
    const int nchunks = (C + FCC - 1) / FCC;

    if (tid == 0) {
        ptx::prefetch_tensormap(&mapL); ptx::prefetch_tensormap(&mapR);
        for (int s = 0; s < FSTAGES; ++s) { ptx::mbar_init(full_bar(s), 1); ptx::mbar_init(empty_bar(s), NW); }
        ptx::fence_mbar_init();
    }
    __syncthreads();

    if (warp == NW) {
        // ================= producer: two TMA loads per 8-channel stage =================
        if (lane == 0) {
            int s = 0; uint32_t ph = 0;
            for (int ch = 0; ch < nchunks; ++ch) {
                wait_or_trap(empty_bar(s), ph ^ 1u);
                ptx::mbar_arrive_expect_tx(full_bar(s), FSTAGE_FLOATS * 4);
                const uint32_t dst = base + s * FSTAGE_FLOATS * 4;
                ptx::tma_load_3d(dst, &mapL, full_bar(s), 0, p0 / 16, b * C + ch * FCC);
                ptx::tma_load_3d(dst + FCC * FT * 4, &mapR, full_bar(s), 0, (p0 - FHT) / 16, b * C + ch * FCC);
                if (++s == FSTAGES) { s = 0; ph ^= 1u; }
            }
        }
    } else {
        // ================= compute warps: warp w owns disparities 14w .. 14w+13 =================
        const int d0 = warp * FDT;
        float acc[FDT][8];
#pragma unroll
        for (int i = 0; i < FDT; ++i)
#pragma unroll
            for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;
        const int e0 = 8 * lane + FHT - (d0 + FDT - 1) * S;  // >= 0: the dispatcher guarantees (NW*14-1)*S <= FHT
        const int mis = e0 & 3, ea = e0 - mis;               // warp-uniform misalignment
        int s = 0; uint32_t ph = 0;
        for (int ch = 0; ch < nchunks; ++ch) {
            const int cc = min(FCC, C - ch * FCC);
            wait_or_trap(full_bar(s), ph);
            const float* sL = ring + s * FSTAGE_FLOATS;
            const float* sR = sL + FCC * FT;
            switch (mis) {
                case 0: corr_flat_chunk<S, 0>(sL, sR, cc, lane, ea, acc); break;
                case 1: corr_flat_chunk<S, 1>(sL, sR, cc, lane, ea, acc); break;
                case 2: corr_flat_chunk<S, 2>(sL, sR, cc, lane, ea, acc); break;
                default: corr_flat_chunk<S, 3>(sL, sR, cc, lane, ea, acc); break;
            }
            __syncwarp();
            if (lane == 0) ptx::mbar_arrive(empty_bar(s));
            if (++s == FSTAGES) { s = 0; ph ^= 1u; }
        }
        // masked store: out[b,d,p] = acc if (p % W) >= d*S else 0 (the reference leaves those at zero)
        const int p = p0 + 8 * lane;
        if (p < HW) {
            int xr[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) xr[j] = (p + j) % W;
#pragma unroll
            for (int dd = 0; dd < FDT; ++dd) {
                const int d = d0 + dd;
                if (d >= D) break;
                float v[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) v[j] = (xr[j] >= d * S) ? acc[dd][j] : 0.f;
                float* o = out + ((size_t)b * D + d) * HW + p;
                st_stream_f4(reinterpret_cast<float4*>(o), make_float4(v[0], v[1], v[2], v[3]));
                if (p + 4 < HW) st_stream_f4(reinterpret_cast<float4*>(o + 4), make_float4(v[4], v[5], v[6], v[7]));
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------
// Fast backward path, same decomposition as the forward one (flattened pixels, 256-pixel tiles, a
// thread = 8 pixels x 14 disparities, warp w = disparities 14w..14w+13):
//   RIGHT = 0 : gL[c,p]  = sum_d g[d,p]      * fR[c, p  - d*s]   (g masked where x(p) < d*s)
//   RIGHT = 1 : gR[c,p'] = sum_d g[d,p'+d*s] * fL[c, p' + d*s]   (same mask, on the shifted pixel)
// The masked gradient tile (14 x 8 per thread) is loaded ONCE into registers; per channel a thread
// reads a sliding window of the feature row (6 / 10 LDS.128) and does 112 FMAs into 8 partial
// sums, which the NW warps of the CTA (the d chunks) combine through a double-buffered shared
// buffer in a fixed order (deterministic) before one coalesced 128-bit store per 4 pixels.
// ---------------------------------------------------------------------------------------------
template <int S, int RIGHT, int MIS>
__device__ __forceinline__ void corr_bwd_chunk(const float* sF, float* red, int cc, int lane, int ea, const float (&g)[FDT][8]) {
    constexpr int WL = 8 + (FDT - 1) * S + 3;
    constexpr int NV = (WL + 3) / 4;
    int qoff[NV];
#pragma unroll
    for (int k = 0; k < NV; ++k) qoff[k] = fswz(ea + 4 * k);
    for (int c = 0; c < cc; ++c) {
        float win[NV * 4];
#pragma unroll
        for (int k = 0; k < NV; ++k) {
            const float4 t = *reinterpret_cast<const float4*>(sF + c * FRW + qoff[k]);
            win[4 * k] = t.x; win[4 * k + 1] = t.y; win[4 * k + 2] = t.z; win[4 * k + 3] = t.w;
        }
        float a[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int dd = 0; dd < FDT; ++dd)
#pragma unroll
            for (int j = 0; j < 8; ++j)
                a[j] = fmaf(g[dd][j], win[MIS + (RIGHT ? dd : (FDT - 1 - dd)) * S + j], a[j]);
        float4* r = reinterpret_cast<float4*>(red + c * FT + 8 * lane);
        r[0] = make_float4(a[0], a[1], a[2], a[3]);
        r[1] = make_float4(a[4], a[5], a[6], a[7]);
    }
}

template <int S, int RIGHT>
__device__ __forceinline__ void corr1d_bwd_flat_body(const CUtensorMap& mapF, const float* __restrict__ gout, float* __restrict__ gfeat,
                                                     int C, int HW, int W, int D, int NW) {
    extern __shared__ uint8_t smem_raw[];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t raw = ptx::smem_u32(smem_raw);
    const uint32_t base = (raw + 1023u) & ~1023u;
    float* ring = reinterpret_cast<float*>(smem_raw + (base - raw));        // [FSTAGES][FCC][FRW]
    float* red = ring + FSTAGES * FCC * FRW;                                // [2][NW][FCC][FT]
    const uint32_t bars = base + (FSTAGES * FCC * FRW + 2 * NW * FCC * FT) * 4;
    auto full_bar = [&](int s) { return bars + 8u * s; };
    auto empty_bar = [&](int s) { return bars + 8u * (FSTAGES + s); };
    const int b = blockIdx.y;
    const int p0 = blockIdx.x * FT;
    const int nchunks = (C + FCC - 1) / FCC;

    if (tid == 0) {
        ptx::prefetch_tensormap(&mapF);
        for (int s = 0; s < FSTAGES; ++s) { ptx::mbar_init(full_bar(s), 1); ptx::mbar_init(empty_bar(s), NW); }
        ptx::fence_mbar_init();
    }
    __syncthreads();

    if (warp == NW) {
        if (lane == 0) {
            int s = 0; uint32_t ph = 0;
            for (int ch = 0; ch < nchunks; ++ch) {
                wait_or_trap(empty_bar(s), ph ^ 1u);
                ptx::mbar_arrive_expect_tx(full_bar(s), FCC * FRW * 4);
                ptx::tma_load_3d(base + s * FCC * FRW * 4, &mapF, full_bar(s), 0, (RIGHT ? p0 : p0 - FHT) / 16, b * C + ch * FCC);
                if (++s == FSTAGES) { s = 0; ph ^= 1u; }
            }
        }
    } else {
        const int d0 = warp * FDT;
        const int p = p0 + 8 * lane;
        // the masked gradient tile, once
        float g[FDT][8];
#pragma unroll
        for (int dd = 0; dd < FDT; ++dd) {
            const int d = d0 + dd;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const int q = RIGHT ? p + j + d * S : p + j;          // pixel whose gradient / mask applies
                const bool ok = d < D && q < HW && (q % W) >= d * S;
                g[dd][j] = ok ? __ldg(gout + ((size_t)b * D + d) * HW + q) : 0.f;
            }
        }
        const int e0 = RIGHT ? 8 * lane + d0 * S : 8 * lane + FHT - (d0 + FDT - 1) * S;
        const int mis = e0 & 3, ea = e0 - mis;
        int s = 0; uint32_t ph = 0;
        for (int ch = 0; ch < nchunks; ++ch) {
            const int cc = min(FCC, C - ch * FCC);
            float* my_red = red + ((ch & 1) * NW + warp) * FCC * FT;
            wait_or_trap(full_bar(s), ph);
            const float* sF = ring + s * FCC * FRW;
            switch (mis) {
                case 0: corr_bwd_chunk<S, RIGHT, 0>(sF, my_red, cc, lane, ea, g); break;
                case 1: corr_bwd_chunk<S, RIGHT, 1>(sF, my_red, cc, lane, ea, g); break;
                case 2: corr_bwd_chunk<S, RIGHT, 2>(sF, my_red, cc, lane, ea, g); break;
                default: corr_bwd_chunk<S, RIGHT, 3>(sF, my_red, cc, lane, ea, g); break;
            }
            __syncwarp();
            if (lane == 0) ptx::mbar_arrive(empty_bar(s));
            if (++s == FSTAGES) { s = 0; ph ^= 1u; }
            // combine the d chunks (fixed order) and store: compute warps only (named barrier 1)
            asm volatile("bar.sync 1, %0;" :: "r"(NW * 32) : "memory");
            const float* rbase = red + (ch & 1) * NW * FCC * FT;
            for (int i = tid; i < cc * (FT / 4); i += NW * 32) {
                const int c = i / (FT / 4), x4 = i - c * (FT / 4);
                float4 a = *reinterpret_cast<const float4*>(rbase + c * FT + 4 * x4);
                for (int w = 1; w < NW; ++w) {
                    const float4 t = *reinterpret_cast<const float4*>(rbase + (w * FCC + c) * FT + 4 * x4);
                    a.x += t.x; a.y += t.y; a.z += t.z; a.w += t.w;
                }
                const int pp = p0 + 4 * x4;
                if (pp < HW) *reinterpret_cast<float4*>(gfeat + ((size_t)b * C + ch * FCC + c) * HW + pp) = a;
            }
        }
    }
}

// Both gradients in ONE launch: blockIdx.z = 0 computes gL (reads fR), 1 computes gR (reads fL).  A CTA of the backward
// is latency-bound (three compute warps, one per sub-partition, IPC ~0.36 in the r01 capture) and at B = 1 each gradient
// alone fills only 117 of 148 SMs; the 234 CTAs of the pair fit the machine at two per SM (99 KB shared memory, 168
// registers x 128 threads each), so the two problems hide each other's latencies instead of running back to back.
template <int S>
__global__ void __launch_bounds__(288, 1)
corr1d_bwd_flat_kernel(const __grid_constant__ CUtensorMap mapR, const __grid_constant__ CUtensorMap mapL,
                       const float* __restrict__ gout, float* __restrict__ gL, float* __restrict__ gR,
                       int C, int HW, int W, int D, int NW) {
    if (blockIdx.z == 0) corr1d_bwd_flat_body<S, 0>(mapR, gout, gL, C, HW, W, D, NW);
    else                 corr1d_bwd_flat_body<S, 1>(mapL, gout, gR, C, HW, W, D, NW);
}

template <int S>
int launch_corr_bwd_flat(const CUtensorMap& mapR, const CUtensorMap& mapL, const float* gout, float* gL, float* gR,
                         int B, int C, long long HW, int W, int D, int NW, cudaStream_t st) {
    const size_t smem = (size_t)(FSTAGES * FCC * FRW + 2 * NW * FCC * FT) * sizeof(float) + 1024 + 2 * FSTAGES * 8;
    cudaError_t e = cudaFuncSetAttribute(corr1d_bwd_flat_kernel<S>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    dim3 grid((unsigned)dsm_ceil_div_ll(HW, FT), B, 2), block((NW + 1) * 32);
    corr1d_bwd_flat_kernel<S><<<grid, block, smem, st>>>(mapR, mapL, gout, gL, gR, C, (int)HW, W, D, NW);
    return dsm_launch_status();
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
    DsmDeviceGuard dsm_guard_(fL);
    if (!fL || !fR || !out || B <= 0 || C <= 0 || H <= 0 || W <= 0 || D <= 0) return DSM_EINVAL;
    if (stride != 1 && stride != 2) return DSM_EUNSUPPORTED;
    if (H > 65535 || B > 65535) return DSM_EUNSUPPORTED;
    if (!dsm_aligned16(out)) return DSM_EALIGN;
    {   // fast path: flattened-pixel kernel
        const long long HW = (long long)H * W;
        const int NWF = dsm_ceil_div(D, FDT);
        if ((HW & 15) == 0 && HW < (1LL << 30) && NWF <= 8 && (NWF * FDT - 1) * stride <= FHT && (long long)B * C < (1LL << 31) &&
            dsm_aligned16(fL) && dsm_aligned16(fR)) {
            CUtensorMap mapL, mapR;
            cuuint64_t dims[3] = {16, (cuuint64_t)(HW / 16), (cuuint64_t)B * C};
            cuuint64_t strides[2] = {64, (cuuint64_t)HW * 4};
            cuuint32_t boxL[3] = {16, FT / 16, FCC}, boxR[3] = {16, FRW / 16, FCC};
            if (!tma_host::encode(&mapL, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, fL, 3, dims, strides, boxL, CU_TENSOR_MAP_SWIZZLE_64B) ||
                !tma_host::encode(&mapR, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, fR, 3, dims, strides, boxR, CU_TENSOR_MAP_SWIZZLE_64B))
                return DSM_EDRIVER;
            const size_t smem = (size_t)FSTAGES * FSTAGE_FLOATS * sizeof(float) + 1024 + 2 * FSTAGES * 8;
            dim3 grid((unsigned)dsm_ceil_div_ll(HW, FT), B), block((NWF + 1) * 32);
            cudaStream_t st = (cudaStream_t)stream;
            cudaError_t e;
            if (stride == 1) {
                e = cudaFuncSetAttribute(corr1d_flat_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
                if (e != cudaSuccess) return (int)e;
                corr1d_flat_kernel<1><<<grid, block, smem, st>>>(mapL, mapR, out, C, (int)HW, W, D, NWF);
            } else {
                e = cudaFuncSetAttribute(corr1d_flat_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
                if (e != cudaSuccess) return (int)e;
                corr1d_flat_kernel<2><<<grid, block, smem, st>>>(mapL, mapR, out, C, (int)HW, W, D, NWF);
            }
            return dsm_launch_status();
        }
    }
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
    DsmDeviceGuard dsm_guard_(gout);
    if (!gout || !fL || !fR || !gL || !gR || B <= 0 || C <= 0 || H <= 0 || W <= 0 || D <= 0) return DSM_EINVAL;
    if (stride < 1 || stride > 4) return DSM_EUNSUPPORTED;
    if (H > 65535 || B > 65535) return DSM_EUNSUPPORTED;
    cudaStream_t st = (cudaStream_t)stream;
    {   // fast path: flattened-pixel kernels
        const long long HW = (long long)H * W;
        const int NWF = dsm_ceil_div(D, FDT);
        if ((stride == 1 || stride == 2) && (HW & 15) == 0 && HW < (1LL << 30) && NWF <= 8 && (NWF * FDT - 1) * stride <= FHT &&
            (long long)B * C < (1LL << 31) && dsm_aligned16(fL) && dsm_aligned16(fR) && dsm_aligned16(gL) && dsm_aligned16(gR)) {
            CUtensorMap mapL, mapR;
            cuuint64_t dims[3] = {16, (cuuint64_t)(HW / 16), (cuuint64_t)B * C};
            cuuint64_t strides[2] = {64, (cuuint64_t)HW * 4};
            cuuint32_t box[3] = {16, FRW / 16, FCC};
            if (!tma_host::encode(&mapL, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, fL, 3, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_64B) ||
                !tma_host::encode(&mapR, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, fR, 3, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_64B))
                return DSM_EDRIVER;
            return stride == 1 ? launch_corr_bwd_flat<1>(mapR, mapL, gout, gL, gR, B, C, HW, W, D, NWF, st)
                               : launch_corr_bwd_flat<2>(mapR, mapL, gout, gL, gR, B, C, HW, W, D, NWF, st);
        }
    }
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
