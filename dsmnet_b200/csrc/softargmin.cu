// op 4 — soft-argmin disparity regression.
// Replaces F.softmax (legacy implicit dim -> 1) + disparityregression of reference
// models/psmnet/submodule.py:56-63 as used at models/psmnet/stackhourglass.py:155-166, the
// GC-Net head models/gcnet.py:104-111 (softmax of MINUS the cost), and — fused — the trilinear
// upsample that precedes it at stackhourglass.py:152-153,163.
//   disp[b,y,x] = sum_d d * softmax_d(sign * cost[b,:,y,x])
// Streaming kernels: every thread owns 4 consecutive pixels (128-bit loads), walks D in chunks
// of 8 planes with a chunked online softmax (one rescale per chunk), fp32 throughout.
// Algorithmic HBM bytes: 4*B*H*W*(D+1).
#include "common.cuh"

namespace {

constexpr int DCH = 8;

struct Online {   // running max m, sum s = sum e^(v-m), t = sum d*e^(v-m)
    float m, s, t;
    __device__ __forceinline__ void init() { m = -INFINITY; s = 0.f; t = 0.f; }
    __device__ __forceinline__ void chunk(const float* v, int n, int d0) {
        float cm = v[0];
#pragma unroll
        for (int i = 1; i < DCH; ++i) if (i < n) cm = fmaxf(cm, v[i]);
        const float nm = fmaxf(m, cm);
        const float sc = __expf(m - nm);     // exp(-inf) = 0 on the first chunk
        s *= sc; t *= sc;
#pragma unroll
        for (int i = 0; i < DCH; ++i) if (i < n) {
            const float e = __expf(v[i] - nm);
            s += e; t = fmaf((float)(d0 + i), e, t);
        }
        m = nm;
    }
};

template <bool VEC>
__global__ void __launch_bounds__(128)
softargmin_fwd_kernel(const float* __restrict__ cost, float* __restrict__ disp,
                      int D, long long HW, long long nq /*pixel groups per batch item*/, float sign) {
    const long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= nq) return;
    const int b = blockIdx.y;
    constexpr int PV = VEC ? 4 : 1;
    const float* c = cost + (size_t)b * D * HW + q * PV;
    Online o[PV];
#pragma unroll
    for (int p = 0; p < PV; ++p) o[p].init();
    for (int d0 = 0; d0 < D; d0 += DCH) {
        const int n = min(DCH, D - d0);
        float v[PV][DCH];
#pragma unroll
        for (int i = 0; i < DCH; ++i) if (i < n) {
            if (VEC) {
                const float4 t = ld_stream_f4(reinterpret_cast<const float4*>(c + (size_t)(d0 + i) * HW));
                v[0][i] = sign * t.x; v[1 % PV][i] = sign * t.y; v[2 % PV][i] = sign * t.z; v[3 % PV][i] = sign * t.w;
            } else {
                v[0][i] = sign * ld_stream_f1(c + (size_t)(d0 + i) * HW);
            }
        }
#pragma unroll
        for (int p = 0; p < PV; ++p) o[p].chunk(v[p], n, d0);
    }
    float* out = disp + (size_t)b * HW + q * PV;
    if (VEC) {
        *reinterpret_cast<float4*>(out) = make_float4(o[0].t / o[0].s, o[1 % PV].t / o[1 % PV].s,
                                                      o[2 % PV].t / o[2 % PV].s, o[3 % PV].t / o[3 % PV].s);
    } else {
        out[0] = o[0].t / o[0].s;
    }
}

// D-split variant for small images: a CTA is 32 pixel-quads x DSPLIT warps; the 8-plane chunks are dealt
// round-robin to the warps (warp w takes chunks w, w+DSPLIT, ...: all warps of the GPU still sweep the
// volume front to back together, which keeps DRAM pages open — a contiguous D range per warp measured 35%
// slower), each with the same chunked online softmax, and the DSPLIT partial (m, s, t) triples are merged
// through shared memory (m = max; s, t rescaled by e^(m_i - m)).  DSPLIT x more threads put DSPLIT x more
// independent 128-bit loads in flight: one thread per pixel-quad only reaches ~16% of HBM bandwidth at
// GC-Net's 256x512 (0.13 M pixels are too few threads).
constexpr int DSPLIT = 4;

__global__ void __launch_bounds__(32 * DSPLIT)
softargmin_fwd_split_kernel(const float* __restrict__ cost, float* __restrict__ disp,
                            int D, long long HW, long long nq, float sign) {
    __shared__ float4 s_m[DSPLIT][32], s_s[DSPLIT][32], s_t[DSPLIT][32];
    const int lane = threadIdx.x, w = threadIdx.y;
    const long long q = (long long)blockIdx.x * 32 + lane;
    const int b = blockIdx.y;
    Online o[4];
#pragma unroll
    for (int p = 0; p < 4; ++p) o[p].init();
    if (q < nq) {
        const float* c = cost + (size_t)b * D * HW + q * 4;
        // software pipeline: the next chunk's eight 128-bit loads are issued before the current chunk is consumed
        float4 nxt[DCH];
        auto fetch = [&](int d0) {
#pragma unroll
            for (int i = 0; i < DCH; ++i)
                nxt[i] = (d0 + i < D) ? ld_stream_f4(reinterpret_cast<const float4*>(c + (size_t)(d0 + i) * HW)) : make_float4(0.f, 0.f, 0.f, 0.f);
        };
        int d0 = w * DCH;
        if (d0 < D) fetch(d0);
        for (; d0 < D; d0 += DSPLIT * DCH) {
            const int n = min(DCH, D - d0);
            float v[4][DCH];
#pragma unroll
            for (int i = 0; i < DCH; ++i) {
                v[0][i] = sign * nxt[i].x; v[1][i] = sign * nxt[i].y; v[2][i] = sign * nxt[i].z; v[3][i] = sign * nxt[i].w;
            }
            if (d0 + DSPLIT * DCH < D) fetch(d0 + DSPLIT * DCH);
#pragma unroll
            for (int p = 0; p < 4; ++p) o[p].chunk(v[p], n, d0);
        }
    }
    s_m[w][lane] = make_float4(o[0].m, o[1].m, o[2].m, o[3].m);
    s_s[w][lane] = make_float4(o[0].s, o[1].s, o[2].s, o[3].s);
    s_t[w][lane] = make_float4(o[0].t, o[1].t, o[2].t, o[3].t);
    __syncthreads();
    if (w == 0 && q < nq) {
        float m[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll
        for (int i = 0; i < DSPLIT; ++i) {
            const float4 a = s_m[i][lane];
            m[0] = fmaxf(m[0], a.x); m[1] = fmaxf(m[1], a.y); m[2] = fmaxf(m[2], a.z); m[3] = fmaxf(m[3], a.w);
        }
        float ss[4] = {0.f, 0.f, 0.f, 0.f}, tt[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int i = 0; i < DSPLIT; ++i) {
            const float4 a = s_m[i][lane], sv = s_s[i][lane], tv = s_t[i][lane];
            const float e0 = __expf(a.x - m[0]), e1 = __expf(a.y - m[1]), e2 = __expf(a.z - m[2]), e3 = __expf(a.w - m[3]);
            ss[0] = fmaf(sv.x, e0, ss[0]); ss[1] = fmaf(sv.y, e1, ss[1]); ss[2] = fmaf(sv.z, e2, ss[2]); ss[3] = fmaf(sv.w, e3, ss[3]);
            tt[0] = fmaf(tv.x, e0, tt[0]); tt[1] = fmaf(tv.y, e1, tt[1]); tt[2] = fmaf(tv.z, e2, tt[2]); tt[3] = fmaf(tv.w, e3, tt[3]);
        }
        *reinterpret_cast<float4*>(disp + (size_t)b * HW + q * 4) = make_float4(tt[0] / ss[0], tt[1] / ss[1], tt[2] / ss[2], tt[3] / ss[3]);
    }
}

// backward: gcost[b,d,y,x] = sign * p_d * (d - disp) * gdisp,  p = softmax_d(sign*cost)
// pass 1 recomputes (m, s); pass 2 writes.  One thread per pixel (scalar, coalesced along x).
__global__ void __launch_bounds__(256)
softargmin_bwd_kernel(const float* __restrict__ cost, const float* __restrict__ disp,
                      const float* __restrict__ gdisp, float* __restrict__ gcost,
                      int D, long long HW, float sign) {
    const long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= HW) return;
    const int b = blockIdx.y;
    const float* c = cost + (size_t)b * D * HW + p;
    float* gc = gcost + (size_t)b * D * HW + p;
    float m = -INFINITY;
    for (int d = 0; d < D; ++d) m = fmaxf(m, sign * __ldg(c + (size_t)d * HW));
    float s = 0.f;
    for (int d = 0; d < D; ++d) s += __expf(sign * __ldg(c + (size_t)d * HW) - m);
    const float out = disp[(size_t)b * HW + p];
    const float g = gdisp[(size_t)b * HW + p] * sign / s;
    for (int d = 0; d < D; ++d) {
        const float e = __expf(sign * __ldg(c + (size_t)d * HW) - m);
        gc[(size_t)d * HW] = g * e * ((float)d - out);
    }
}

__global__ void __launch_bounds__(256)
dispreg_fwd_kernel(const float* __restrict__ prob, float* __restrict__ disp, int D, long long HW) {
    const long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= HW) return;
    const int b = blockIdx.y;
    const float* c = prob + (size_t)b * D * HW + p;
    float t = 0.f;
#pragma unroll 8
    for (int d = 0; d < D; ++d) t = fmaf((float)d, ld_stream_f1(c + (size_t)d * HW), t);
    disp[(size_t)b * HW + p] = t;
}

__global__ void __launch_bounds__(256)
dispreg_bwd_kernel(const float* __restrict__ gdisp, float* __restrict__ gprob, int D, long long HW) {
    const long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= HW) return;
    const int b = blockIdx.y;
    const float g = gdisp[(size_t)b * HW + p];
    float* o = gprob + (size_t)b * D * HW + p;
    for (int d = 0; d < D; ++d) o[(size_t)d * HW] = (float)d * g;
}

// ---------------------------------------------------------------------------------------
// Fused head: trilinear x-upsample of the low-res cost + softmax over D + regression.
// The (h,w) interpolation weights do not depend on d, so each thread keeps the two bilinearly
// interpolated low-res planes that bracket the current d and refreshes one of them whenever
// the low-res index advances (warp-uniform: it depends on d only).  Reads 4*B*Dl*Hl*Wl bytes
// (L1/L2 resident), writes 4*B*H*W; bound by the exp (SFU) rate, not HBM.
// Index/weight arithmetic follows ATen's area_pixel_compute_source_index (fp32).
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ void src_index(int dst, float scale, int in_size, int align_corners, int& i0, int& i1, float& l1) {
    float real;
    if (align_corners) real = scale * (float)dst;
    else { real = scale * ((float)dst + 0.5f) - 0.5f; if (real < 0.f) real = 0.f; }
    i0 = min((int)floorf(real), in_size - 1);
    l1 = fminf(fmaxf(real - (float)i0, 0.f), 1.f);
    i1 = min(i0 + 1, in_size - 1);
}

// 128-bit vector reduction to global memory (sm_90+): four adjacent floats, 16-byte aligned, one L2 operation
__device__ __forceinline__ void red_add_v4(float* p, float a, float b, float c, float d) {
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" :: "l"(p), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

__device__ __forceinline__ float ex2_approx(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

// One CTA = 256 consecutive output pixels of one output row (two per thread).
//  1. tables (shared): l1[d] = D-axis interpolation weight of output plane d, cnt[dl] = how many
//     output planes have low-res source plane dl (they are consecutive in d);
//  2. the low-res row pair (h0,h1) the output row falls between is blended ONCE per CTA into shared
//     memory: row[dl][j] for the ~0.25*128+2 low-res columns the CTA touches (warp = a group of planes,
//     lanes = columns: no index division, eight independent loads in flight per thread);
//  3. each thread blends its two columns per plane on the fly: a first sweep over the Dl planes finds the
//     maximum (a linear interpolation never exceeds its end points, so max_d of the upsampled column
//     <= max_dl c[dl] and the softmax is a single pass), a second sweep walks the D planes: one FMA + one
//     ex2 + two accumulations per plane.  The loops are real loops (compact code: the first, fully
//     unrolled version of this kernel spent a third of its time on instruction-cache misses).
__global__ void __launch_bounds__(128)
upsample_softargmin_kernel(const float* __restrict__ cost, float* __restrict__ disp, float* __restrict__ lse2,
                           int Dl, int Hl, int Wl, int D, int H, int W,
                           float sd, float sh, float sw, int align_corners, int rw) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float* s_l1 = reinterpret_cast<float*>(smem_raw);        // [D]
    int* s_cnt = reinterpret_cast<int*>(s_l1 + D);           // [Dl]
    float* s_row = reinterpret_cast<float*>(s_cnt + Dl);     // [Dl + 1][rw]   (one extra, zero, row: the clamped top interval)
    const int tid = threadIdx.x, lane = tid & 31, wrp = tid >> 5;
    for (int i = tid; i < Dl; i += blockDim.x) s_cnt[i] = 0;
    __syncthreads();
    for (int d = tid; d < D; d += blockDim.x) {
        int d0, d1; float l1;
        src_index(d, sd, Dl, align_corners, d0, d1, l1);
        s_l1[d] = (d1 == d0) ? 0.f : l1;                     // clamped top: both taps are the same plane
        atomicAdd(&s_cnt[d0], 1);
    }
    const int x0 = blockIdx.x * (2 * blockDim.x);            // a thread owns pixels x0+tid and x0+128+tid (shares the d tables)
    const int y = blockIdx.y, b = blockIdx.z;
    int h0, h1, wb, wb1; float lh1, lwb;
    src_index(y, sh, Hl, align_corners, h0, h1, lh1);
    src_index(x0, sw, Wl, align_corners, wb, wb1, lwb);      // first low-res column of this CTA
    const float lh0 = 1.f - lh1;
    const float* base = cost + (size_t)b * Dl * Hl * Wl;
    const size_t pl = (size_t)Hl * Wl;
    for (int j = lane; j < rw; j += 32) {
        const int wl = min(wb + j, Wl - 1);
        const float* p0 = base + (size_t)h0 * Wl + wl;
        const float* p1 = base + (size_t)h1 * Wl + wl;
        for (int dl = wrp; dl < Dl; dl += 16) {              // 4 warps x 4 planes per trip: 8 loads in flight
            float a[4], c[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int dd = min(dl + 4 * u, Dl - 1);
                a[u] = __ldg(p0 + (size_t)dd * pl); c[u] = __ldg(p1 + (size_t)dd * pl);
            }
#pragma unroll
            for (int u = 0; u < 4; ++u)
                if (dl + 4 * u < Dl) s_row[(dl + 4 * u) * rw + j] = lh0 * a[u] + lh1 * c[u];
        }
        if (wrp == 0) s_row[Dl * rw + j] = 0.f;
    }
    __syncthreads();
    const int xa = x0 + tid, xb = min(xa + (int)blockDim.x, W - 1);      // xb clamped: computed, stored only if in range
    if (xa >= W) return;
    int w0, w1; float lwa1, lwb1;
    src_index(xa, sw, Wl, align_corners, w0, w1, lwa1);
    const float* ra0 = s_row + (w0 - wb); const float* ra1 = s_row + (w1 - wb);
    src_index(xb, sw, Wl, align_corners, w0, w1, lwb1);
    const float* rb0 = s_row + (w0 - wb); const float* rb1 = s_row + (w1 - wb);
    const float lwa0 = 1.f - lwa1, lwb0 = 1.f - lwb1;
    float ma = -INFINITY, mb = -INFINITY;
    for (int dl = 0; dl < Dl; ++dl) {
        ma = fmaxf(ma, lwa0 * ra0[dl * rw] + lwa1 * ra1[dl * rw]);
        mb = fmaxf(mb, lwb0 * rb0[dl * rw] + lwb1 * rb1[dl * rw]);
    }
    constexpr float LOG2E = 1.4426950408889634f;
    const float mla = ma * LOG2E, mlb = mb * LOG2E;
    float sa = 0.f, ta = 0.f, sb = 0.f, tb = 0.f, fd = 0.f;
    float a0 = fmaf(lwa0 * ra0[0] + lwa1 * ra1[0], LOG2E, -mla);
    float b0 = fmaf(lwb0 * rb0[0] + lwb1 * rb1[0], LOG2E, -mlb);
    int d = 0;
    for (int dl = 0; dl < Dl; ++dl) {
        const int o = (dl + 1) * rw;                           // row Dl is zero and so is its weight (clamped top)
        const float a1 = fmaf(lwa0 * ra0[o] + lwa1 * ra1[o], LOG2E, -mla);
        const float b1 = fmaf(lwb0 * rb0[o] + lwb1 * rb1[o], LOG2E, -mlb);
        const float da = a1 - a0, db = b1 - b0;
        const int n = s_cnt[dl];
        for (int k = 0; k < n; ++k, ++d) {
            // the two pixels of the thread as packed fp32 pairs (FFMA2 / FADD2: same results, 7 instead of 10 instructions)
            const float l = s_l1[d];
            float xa, xb;
            ffma2(xa, xb, l, l, da, db, a0, b0);
            const float ea = ex2_approx(xa), eb = ex2_approx(xb);
            fadd2(sa, sb, sa, sb, ea, eb);
            ffma2(ta, tb, fd, fd, ea, eb, ta, tb);
            fd += 1.f;
        }
        a0 = a1; b0 = b1;
    }
    float* o = disp + ((size_t)b * H + y) * W;
    o[xa] = ta / sa;
    if (xa + (int)blockDim.x < W) o[xa + blockDim.x] = tb / sb;
    if (lse2) {                                              // log2 of the softmax denominator, for the backward pass
        float* q = lse2 + ((size_t)b * H + y) * W;
        q[xa] = mla + log2f(sa);
        if (xa + (int)blockDim.x < W) q[xa + blockDim.x] = mlb + log2f(sb);
    }
}

// Backward of the fused head: gcost_lr += J^T gdisp.  Same CTA shape and shared tables as the forward kernel; the
// upsampled volume is never materialised.  Per pixel and output plane d:
//     g_d = gdisp * p_d * (d - disp),  p_d = 2^(c_d*log2e - lse2)          (one ex2 per point)
// g_d is split onto its two source planes with the D-axis weights while the thread walks the planes, the per-plane
// totals are split onto the two source columns with shared-memory atomics (s_g[dl][j], the gradient of the blended
// row), and once per CTA s_g is split onto the two source rows with global atomics.
__global__ void __launch_bounds__(128)
upsample_softargmin_bwd_kernel(const float* __restrict__ cost, const float* __restrict__ disp, const float* __restrict__ lse2,
                               const float* __restrict__ gdisp, float* __restrict__ gcost,
                               int Dl, int Hl, int Wl, int D, int H, int W,
                               float sd, float sh, float sw, int align_corners, int rw) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float* s_l1 = reinterpret_cast<float*>(smem_raw);        // [D]
    int* s_cnt = reinterpret_cast<int*>(s_l1 + D);           // [Dl]
    float* s_row = reinterpret_cast<float*>(s_cnt + Dl);     // [Dl + 1][rw]
    float* s_g = s_row + (size_t)(Dl + 1) * rw;              // [Dl][rw]
    const int tid = threadIdx.x, lane = tid & 31, wrp = tid >> 5;
    for (int i = tid; i < Dl; i += blockDim.x) s_cnt[i] = 0;
    for (int i = tid; i < Dl * rw; i += blockDim.x) s_g[i] = 0.f;
    __syncthreads();
    for (int d = tid; d < D; d += blockDim.x) {
        int d0, d1; float l1;
        src_index(d, sd, Dl, align_corners, d0, d1, l1);
        s_l1[d] = (d1 == d0) ? 0.f : l1;
        atomicAdd(&s_cnt[d0], 1);
    }
    const int x0 = blockIdx.x * (2 * blockDim.x);
    const int y = blockIdx.y, b = blockIdx.z;
    int h0, h1, wb, wb1; float lh1, lwb;
    src_index(y, sh, Hl, align_corners, h0, h1, lh1);
    src_index(x0, sw, Wl, align_corners, wb, wb1, lwb);
    const float lh0 = 1.f - lh1;
    const float* base = cost + (size_t)b * Dl * Hl * Wl;
    const size_t pl = (size_t)Hl * Wl;
    for (int j = lane; j < rw; j += 32) {
        const int wl = min(wb + j, Wl - 1);
        const float* p0 = base + (size_t)h0 * Wl + wl;
        const float* p1 = base + (size_t)h1 * Wl + wl;
        for (int dl = wrp; dl < Dl; dl += 16) {
            float a[4], c[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int dd = min(dl + 4 * u, Dl - 1);
                a[u] = __ldg(p0 + (size_t)dd * pl); c[u] = __ldg(p1 + (size_t)dd * pl);
            }
#pragma unroll
            for (int u = 0; u < 4; ++u)
                if (dl + 4 * u < Dl) s_row[(dl + 4 * u) * rw + j] = lh0 * a[u] + lh1 * c[u];
        }
        if (wrp == 0) s_row[Dl * rw + j] = 0.f;
    }
    __syncthreads();
    const int xa = x0 + tid;
    if (xa < W) {
        const bool has_b = xa + (int)blockDim.x < W;
        const int xb = has_b ? xa + (int)blockDim.x : xa;
        int wa0, wa1, wb0, wb1i; float lwa1, lwb1;
        src_index(xa, sw, Wl, align_corners, wa0, wa1, lwa1);
        src_index(xb, sw, Wl, align_corners, wb0, wb1i, lwb1);
        wa0 -= wb; wa1 -= wb; wb0 -= wb; wb1i -= wb;
        const float lwa0 = 1.f - lwa1, lwb0 = 1.f - lwb1;
        const size_t po = ((size_t)b * H + y) * W;
        constexpr float LOG2E = 1.4426950408889634f;
        const float la = lse2[po + xa], lb = lse2[po + xb];
        const float dpa = disp[po + xa], dpb = disp[po + xb];
        const float ga = gdisp[po + xa], gb = has_b ? gdisp[po + xb] : 0.f;
        float a0 = fmaf(lwa0 * s_row[wa0] + lwa1 * s_row[wa1], LOG2E, -la);
        float b0 = fmaf(lwb0 * s_row[wb0] + lwb1 * s_row[wb1i], LOG2E, -lb);
        float carry_a = 0.f, carry_b = 0.f, fd = 0.f;
        int d = 0;
        for (int dl = 0; dl < Dl; ++dl) {
            const int o = (dl + 1) * rw;
            const float a1 = fmaf(lwa0 * s_row[o + wa0] + lwa1 * s_row[o + wa1], LOG2E, -la);
            const float b1 = fmaf(lwb0 * s_row[o + wb0] + lwb1 * s_row[o + wb1i], LOG2E, -lb);
            const float da = a1 - a0, db = b1 - b0;
            const int n = s_cnt[dl];
            float lo_a = carry_a, hi_a = 0.f, lo_b = carry_b, hi_b = 0.f;
            for (int k = 0; k < n; ++k, ++d) {
                const float l = s_l1[d];
                const float va = ex2_approx(fmaf(l, da, a0)) * (fd - dpa);
                const float vb = ex2_approx(fmaf(l, db, b0)) * (fd - dpb);
                hi_a = fmaf(l, va, hi_a); lo_a += va;           // lo accumulates the full term, corrected below
                hi_b = fmaf(l, vb, hi_b); lo_b += vb;
                fd += 1.f;
            }
            lo_a -= hi_a; lo_b -= hi_b;                            // (1-l)*v = v - l*v
            const float ta = ga * lo_a, tb = gb * lo_b;
            const int r = dl * rw;
            atomicAdd(&s_g[r + wa0], lwa0 * ta); atomicAdd(&s_g[r + wa1], lwa1 * ta);
            atomicAdd(&s_g[r + wb0], lwb0 * tb); atomicAdd(&s_g[r + wb1i], lwb1 * tb);
            carry_a = hi_a; carry_b = hi_b;
            a0 = a1; b0 = b1;
        }
    }
    __syncthreads();
    float* gbase = gcost + (size_t)b * Dl * Hl * Wl;
    // flush: the CTA's window is rw consecutive columns of two low-res rows per plane.  Aligned groups of four columns go
    // out as ONE 128-bit vector reduction (red.global.add.v4.f32) — the kernel is bound by the number of L2 atomics
    // (36 M scalar ones for the three heads), not by their bytes; the window's ragged ends stay scalar.
    const bool vec_ok = ((Wl & 3) == 0) && ((reinterpret_cast<uintptr_t>(gcost) & 15u) == 0);
    const int lead = vec_ok ? ((4 - (wb & 3)) & 3) : rw;               // scalar columns before the first aligned group
    const int ngrp = vec_ok ? (rw - lead) / 4 : 0;
    const int per = lead + ngrp + (rw - lead - 4 * ngrp);              // work items per plane: leading scalars, groups, trailing scalars
    for (int i = tid; i < Dl * per; i += blockDim.x) {
        const int dl = i / per, k = i - dl * per;
        const float* sg = s_g + dl * rw;
        float* q0 = gbase + (size_t)dl * pl + wb;
        if (k >= lead && k < lead + ngrp) {
            const int j = lead + 4 * (k - lead);
            if (wb + j + 3 < Wl) {
                const float4 v = make_float4(sg[j], sg[j + 1], sg[j + 2], sg[j + 3]);
                if (v.x != 0.f || v.y != 0.f || v.z != 0.f || v.w != 0.f) {
                    red_add_v4(q0 + (size_t)h0 * Wl + j, lh0 * v.x, lh0 * v.y, lh0 * v.z, lh0 * v.w);
                    red_add_v4(q0 + (size_t)h1 * Wl + j, lh1 * v.x, lh1 * v.y, lh1 * v.z, lh1 * v.w);
                }
                continue;
            }
            for (int jj = j; jj < j + 4; ++jj) {                      // the group straddles the right edge of the tensor
                if (wb + jj >= Wl || sg[jj] == 0.f) continue;
                atomicAdd(q0 + (size_t)h0 * Wl + jj, lh0 * sg[jj]);
                atomicAdd(q0 + (size_t)h1 * Wl + jj, lh1 * sg[jj]);
            }
            continue;
        }
        const int j = (k < lead) ? k : lead + 4 * ngrp + (k - lead - ngrp);
        if (wb + j >= Wl) continue;
        const float v = sg[j];
        if (v == 0.f) continue;
        atomicAdd(q0 + (size_t)h0 * Wl + j, lh0 * v);
        atomicAdd(q0 + (size_t)h1 * Wl + j, lh1 * v);
    }
}

}  // namespace

extern "C" int dsm_softargmin_fwd(const float* cost, float* disp, int B, int D, int H, int W, float sign, void* stream) {
    DsmDeviceGuard dsm_guard_(cost);
    if (!cost || !disp || B <= 0 || D <= 0 || H <= 0 || W <= 0) return DSM_EINVAL;
    if (B > 65535) return DSM_EUNSUPPORTED;
    const long long HW = (long long)H * W;
    cudaStream_t st = (cudaStream_t)stream;
    if ((HW & 3) == 0 && dsm_aligned16(cost) && dsm_aligned16(disp)) {
        const long long nq = HW / 4;
        if (D >= 2 * DSPLIT * DCH)
            softargmin_fwd_split_kernel<<<dim3((unsigned)dsm_ceil_div_ll(nq, 32), B), dim3(32, DSPLIT), 0, st>>>(cost, disp, D, HW, nq, sign);
        else
            softargmin_fwd_kernel<true><<<dim3((unsigned)dsm_ceil_div_ll(nq, 128), B), 128, 0, st>>>(cost, disp, D, HW, nq, sign);
    } else {
        softargmin_fwd_kernel<false><<<dim3((unsigned)dsm_ceil_div_ll(HW, 128), B), 128, 0, st>>>(cost, disp, D, HW, HW, sign);
    }
    return dsm_launch_status();
}

extern "C" int dsm_softargmin_bwd(const float* cost, const float* disp, const float* gdisp, float* gcost,
                                  int B, int D, int H, int W, float sign, void* stream) {
    DsmDeviceGuard dsm_guard_(cost);
    if (!cost || !disp || !gdisp || !gcost || B <= 0 || D <= 0 || H <= 0 || W <= 0) return DSM_EINVAL;
    if (B > 65535) return DSM_EUNSUPPORTED;
    const long long HW = (long long)H * W;
    softargmin_bwd_kernel<<<dim3((unsigned)dsm_ceil_div_ll(HW, 256), B), 256, 0, (cudaStream_t)stream>>>(cost, disp, gdisp, gcost, D, HW, sign);
    return dsm_launch_status();
}

extern "C" int dsm_disparity_regression_fwd(const float* prob, float* disp, int B, int D, int H, int W, void* stream) {
    DsmDeviceGuard dsm_guard_(prob);
    if (!prob || !disp || B <= 0 || D <= 0 || H <= 0 || W <= 0) return DSM_EINVAL;
    if (B > 65535) return DSM_EUNSUPPORTED;
    const long long HW = (long long)H * W;
    dispreg_fwd_kernel<<<dim3((unsigned)dsm_ceil_div_ll(HW, 256), B), 256, 0, (cudaStream_t)stream>>>(prob, disp, D, HW);
    return dsm_launch_status();
}

extern "C" int dsm_disparity_regression_bwd(const float* gdisp, float* gprob, int B, int D, int H, int W, void* stream) {
    DsmDeviceGuard dsm_guard_(gdisp);
    if (!gdisp || !gprob || B <= 0 || D <= 0 || H <= 0 || W <= 0) return DSM_EINVAL;
    if (B > 65535) return DSM_EUNSUPPORTED;
    const long long HW = (long long)H * W;
    dispreg_bwd_kernel<<<dim3((unsigned)dsm_ceil_div_ll(HW, 256), B), 256, 0, (cudaStream_t)stream>>>(gdisp, gprob, D, HW);
    return dsm_launch_status();
}

static int upsample_softargmin_fwd_impl(const float* cost_lr, float* disp, float* lse2, int B, int Dl, int Hl, int Wl,
                                        int D, int H, int W, int align_corners, void* stream) {
    if (!cost_lr || !disp || B <= 0 || Dl <= 0 || Hl <= 0 || Wl <= 0 || D <= 0 || H <= 0 || W <= 0) return DSM_EINVAL;
    if (H > 65535 || B > 65535) return DSM_EUNSUPPORTED;
    // ATen: align_corners -> (in-1)/(out-1) (0 when out==1); otherwise in/out; all in fp32
    auto scale = [&](int in, int out) -> float {
        if (align_corners) return out > 1 ? (float)(in - 1) / (float)(out - 1) : 0.f;
        return (float)in / (float)out;
    };
    const float sw = scale(Wl, W);
    const int rw = (int)(sw * 255.f) + 3;                    // low-res columns one 256-pixel CTA can touch
    const size_t smem = (size_t)D * sizeof(float) + (size_t)Dl * sizeof(int) + (size_t)(Dl + 1) * rw * sizeof(float);
    if (smem > 48 * 1024) return DSM_EUNSUPPORTED;
    const dim3 grid(dsm_ceil_div(W, 256), H, B);
    upsample_softargmin_kernel<<<grid, 128, smem, (cudaStream_t)stream>>>(
        cost_lr, disp, lse2, Dl, Hl, Wl, D, H, W, scale(Dl, D), scale(Hl, H), sw, align_corners, rw);
    return dsm_launch_status();
}

extern "C" int dsm_upsample_softargmin_fwd(const float* cost_lr, float* disp, int B, int Dl, int Hl, int Wl,
                                           int D, int H, int W, int align_corners, void* stream) {
    DsmDeviceGuard dsm_guard_(cost_lr);
    return upsample_softargmin_fwd_impl(cost_lr, disp, nullptr, B, Dl, Hl, Wl, D, H, W, align_corners, stream);
}

extern "C" int dsm_upsample_softargmin_fwd_lse(const float* cost_lr, float* disp, float* lse2, int B, int Dl, int Hl, int Wl,
                                               int D, int H, int W, int align_corners, void* stream) {
    DsmDeviceGuard dsm_guard_(cost_lr);
    if (!lse2) return DSM_EINVAL;
    return upsample_softargmin_fwd_impl(cost_lr, disp, lse2, B, Dl, Hl, Wl, D, H, W, align_corners, stream);
}

extern "C" int dsm_upsample_softargmin_bwd(const float* cost_lr, const float* disp, const float* lse2, const float* gdisp,
                                           float* gcost_lr, int B, int Dl, int Hl, int Wl,
                                           int D, int H, int W, int align_corners, void* stream) {
    DsmDeviceGuard dsm_guard_(cost_lr);
    if (!cost_lr || !disp || !lse2 || !gdisp || !gcost_lr || B <= 0 || Dl <= 0 || Hl <= 0 || Wl <= 0 || D <= 0 || H <= 0 || W <= 0)
        return DSM_EINVAL;
    if (H > 65535 || B > 65535) return DSM_EUNSUPPORTED;
    auto scale = [&](int in, int out) -> float {
        if (align_corners) return out > 1 ? (float)(in - 1) / (float)(out - 1) : 0.f;
        return (float)in / (float)out;
    };
    const float sw = scale(Wl, W);
    const int rw = (int)(sw * 255.f) + 3;
    const size_t smem = (size_t)D * sizeof(float) + (size_t)Dl * sizeof(int) + (size_t)(2 * Dl + 1) * rw * sizeof(float);
    if (smem > 48 * 1024) return DSM_EUNSUPPORTED;
    cudaStream_t st = (cudaStream_t)stream;
    cudaError_t ce = cudaMemsetAsync(gcost_lr, 0, sizeof(float) * (size_t)B * Dl * Hl * Wl, st);
    if (ce != cudaSuccess) return (int)ce;
    const dim3 grid(dsm_ceil_div(W, 256), H, B);
    upsample_softargmin_bwd_kernel<<<grid, 128, smem, st>>>(
        cost_lr, disp, lse2, gdisp, gcost_lr, Dl, Hl, Wl, D, H, W, scale(Dl, D), scale(Hl, H), sw, align_corners, rw);
    return dsm_launch_status();
}
