// op 2 — concatenation cost volume (PSMNet / GC-Net), plus the layout converters that sit at
// the reference boundary of the 3-D stack.
// Replaces the Python loops at reference models/psmnet/stackhourglass.py:124-133 (PSM),
// models/gcnet.py:131-135 (GC) and models/gcnet.py:157-164 (GC right-reference volume):
//   PSM      : vol[b,c,d,y,x]   = fL[b,c,y,x]   (x>=d else 0)   vol[b,C+c,d,y,x] = fR[b,c,y,x-d] (x>=d else 0)
//   GC       : vol[b,c,d,y,x]   = fL[b,c,y,x]   (all x)         vol[b,C+c,d,y,x] = fR[b,c,y,x-d] (x>=d else 0)
//   GC_RIGHT : vol[b,c,d,y,x]   = fR[b,c,y,x]   (all x)         vol[b,C+c,d,y,x] = fL[b,c,y,x+d] (x+d<W else 0)
// Pure copies, so results are bit-exact (the bf16 variant is RNE of the same values).
// HBM-write bound: algorithmic bytes = 4*B*H*W*2C (read) + e*B*2C*D*H*W (write), e = 4 or 2.
#include "common.cuh"

namespace {

// ---------------------------------------------------------------------------------------
// NCDHW fp32 (the reference's layout).  CTA = (y, channel ch of 2C, b): the source row is
// staged once in shared memory, then D rows of W floats are streamed out with 128-bit stores.
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128)
concat_ncdhw_kernel(const float* __restrict__ fL, const float* __restrict__ fR, float* __restrict__ out,
                    int C, int D, int H, int W, int mode) {
    extern __shared__ __align__(16) float row[];       // [W]
    const int y = blockIdx.x, ch = blockIdx.y, b = blockIdx.z;
    const bool second = ch >= C;
    const int c = second ? ch - C : ch;
    const float* src;
    if (mode == DSM_VOL_GC_RIGHT) src = second ? fL : fR; else src = second ? fR : fL;
    src += (((size_t)b * C + c) * H + y) * W;
    for (int x = threadIdx.x; x < W; x += blockDim.x) row[x] = __ldg(src + x);
    __syncthreads();

    float* o = out + ((((size_t)b * 2 * C + ch) * D) * H + y) * W;
    const size_t dstride = (size_t)H * W;
    // shift of the source index and zero region per disparity
    //   first half : PSM -> zero for x<d, no shift; GC/GC_RIGHT -> plain copy
    //   second half: PSM/GC -> src[x-d], zero for x<d; GC_RIGHT -> src[x+d], zero for x+d>=W
    const bool vec = ((W & 3) == 0);
    for (int d = 0; d < D; ++d) {
        int shift = 0, lo = 0, hi = W;                  // out[x] = row[x+shift] for lo<=x<hi else 0
        if (!second) { if (mode == DSM_VOL_PSM) lo = d; }
        else if (mode == DSM_VOL_GC_RIGHT) { shift = d; hi = W - d; }
        else { shift = -d; lo = d; }
        if (lo > W) lo = W;
        if (hi < 0) hi = 0;
        float* od = o + d * dstride;
        if (vec) {
            for (int xq = threadIdx.x; xq < (W >> 2); xq += blockDim.x) {
                const int x = xq << 2;
                float v[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const int xx = x + j;
                    v[j] = (xx >= lo && xx < hi) ? row[xx + shift] : 0.f;
                }
                st_stream_f4(reinterpret_cast<float4*>(od + x), make_float4(v[0], v[1], v[2], v[3]));
            }
        } else {
            for (int x = threadIdx.x; x < W; x += blockDim.x)
                od[x] = (x >= lo && x < hi) ? row[x + shift] : 0.f;
        }
    }
}

// ---------------------------------------------------------------------------------------
// Padded NDHWC bf16 (the 3-D stack's layout): [B][D+2][H+2][W+2][2C], zero rim included.
// CTA = (padded row y', slab of d', b).  The two feature rows are transposed once into shared
// memory as bf16 [x][C]; every voxel is then 2C*2 bytes = a run of 16-byte chunks, and a padded
// row of the volume is one contiguous (W+2)*2C*2-byte stream written with 128-bit stores.
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
concat_ndhwc_bf16_kernel(const float* __restrict__ fL, const float* __restrict__ fR, uint4* __restrict__ out,
                         int C, int D, int H, int W, int mode, int dslab) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __nv_bfloat16* sA = reinterpret_cast<__nv_bfloat16*>(smem_raw);          // first half source  [W][C]
    __nv_bfloat16* sB = sA + (size_t)W * C;                                    // second half source [W][C]
    const int yp = blockIdx.x, b = blockIdx.z;
    const int dp0 = blockIdx.y * dslab;
    const int dp1 = min(dp0 + dslab, D + 2);
    const int Wp = W + 2, Hp = H + 2;
    const int cpv = (2 * C) / 8;          // 16-byte chunks per voxel
    const int half = cpv / 2;
    const bool yrim = (yp == 0 || yp == H + 1);

    // 16-byte chunk k of voxel x lives at chunk (k ^ swz(x)): conflict-free 128-bit shared stores by 8 consecutive x
    const int swz_mask = (half >= 4 && (half & (half - 1)) == 0) ? 3 : 0;
    uint4* wA = reinterpret_cast<uint4*>(sA);
    uint4* wB = reinterpret_cast<uint4*>(sB);
    if (!yrim) {
        const int y = yp - 1;
        const float* a = (mode == DSM_VOL_GC_RIGHT ? fR : fL) + ((size_t)b * C * H + y) * W;
        const float* r = (mode == DSM_VOL_GC_RIGHT ? fL : fR) + ((size_t)b * C * H + y) * W;
        const size_t plane = (size_t)H * W;
        // thread = (x, group of 8 channels): 8 coalesced-along-x loads per source, one 128-bit shared store each
        for (int i = threadIdx.x; i < W * half; i += blockDim.x) {
            const int kq = i / W, x = i - kq * W;
            const float* pa = a + (size_t)(kq * 8) * plane + x;
            const float* pr = r + (size_t)(kq * 8) * plane + x;
            float fa[8], fr[8];
#pragma unroll
            for (int c = 0; c < 8; ++c) { fa[c] = __ldg(pa + c * plane); fr[c] = __ldg(pr + c * plane); }
            const int dst = x * half + (kq ^ ((x >> 1) & swz_mask));
            wA[dst] = make_uint4(pack_bf16x2(fa[0], fa[1]), pack_bf16x2(fa[2], fa[3]), pack_bf16x2(fa[4], fa[5]), pack_bf16x2(fa[6], fa[7]));
            wB[dst] = make_uint4(pack_bf16x2(fr[0], fr[1]), pack_bf16x2(fr[2], fr[3]), pack_bf16x2(fr[4], fr[5]), pack_bf16x2(fr[6], fr[7]));
        }
    }
    __syncthreads();

    const uint4 zero = make_uint4(0u, 0u, 0u, 0u);
    const uint4* vA = reinterpret_cast<const uint4*>(sA);
    const uint4* vB = reinterpret_cast<const uint4*>(sB);
    const int row_chunks = Wp * cpv;
    const size_t plane_chunks = (size_t)Hp * row_chunks;
    uint4* obase = out + (((size_t)b * (D + 2) + dp0) * Hp + yp) * (size_t)row_chunks;
    // a thread keeps its chunk positions and walks the planes of the slab: the (voxel, chunk) decode is done once
    for (int i = threadIdx.x; i < row_chunks; i += blockDim.x) {
        const int xp = i / cpv, k = i - xp * cpv;
        const int x = xp - 1;
        const bool inside = !yrim && xp >= 1 && xp <= W;
        const bool first = k < half;
        const int kk = first ? k : k - half;
        uint4 va = zero;
        if (inside && first) va = vA[x * half + (kk ^ ((x >> 1) & swz_mask))];
        uint4* o = obase + i;
        for (int dp = dp0; dp < dp1; ++dp, o += plane_chunks) {
            uint4 v = zero;
            if (inside && dp >= 1 && dp <= D) {
                const int d = dp - 1;
                if (first) {
                    if (mode != DSM_VOL_PSM || x >= d) v = va;
                } else if (mode == DSM_VOL_GC_RIGHT) {
                    const int xs = x + d;
                    if (xs < W) v = vB[xs * half + (kk ^ ((xs >> 1) & swz_mask))];
                } else {
                    const int xs = x - d;
                    if (xs >= 0) v = vB[xs * half + (kk ^ ((xs >> 1) & swz_mask))];
                }
            }
            st_stream_u4(o, v);
        }
    }
}

// ---------------------------------------------------------------------------------------
// backward, NCDHW fp32:  gA[b,c,y,x] = sum_d g[b,c,d,y,x] (PSM: only d<=x)
//                        gB[b,c,y,x'] = sum_d g[b,C+c,d,y,x'+d] (GC_RIGHT: x'-d), where it exists
// one thread per (b,c,y,x); reads are coalesced along x for every d.
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
concat_bwd_ncdhw_kernel(const float* __restrict__ g, float* __restrict__ gL, float* __restrict__ gR,
                        int B, int C, int D, int H, int W, int mode) {
    const long long n = (long long)B * C * H * W;
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int x = (int)(i % W);
    long long t = i / W;
    const int y = (int)(t % H); t /= H;
    const int c = (int)(t % C);
    const int b = (int)(t / C);
    const size_t dstride = (size_t)H * W;
    const float* g1 = g + ((((size_t)b * 2 * C + c) * D) * H + y) * W;
    const float* g2 = g + ((((size_t)b * 2 * C + C + c) * D) * H + y) * W;
    float s1 = 0.f, s2 = 0.f;
    const int d1 = (mode == DSM_VOL_PSM) ? min(D, x + 1) : D;
    for (int d = 0; d < d1; ++d) s1 += ld_stream_f1(g1 + d * dstride + x);
    if (mode == DSM_VOL_GC_RIGHT) {
        const int d2 = min(D, x + 1);                 // source index x' = x, taken at volume x'-d >= 0
        for (int d = 0; d < d2; ++d) s2 += ld_stream_f1(g2 + d * dstride + x - d);
    } else {
        const int d2 = min(D, W - x);                 // volume position x+d < W
        for (int d = 0; d < d2; ++d) s2 += ld_stream_f1(g2 + d * dstride + x + d);
    }
    // first half belongs to fL (fR for GC_RIGHT); second half to the other one
    if (mode == DSM_VOL_GC_RIGHT) { gR[i] = s1; gL[i] = s2; }
    else                          { gL[i] = s1; gR[i] = s2; }
}

// ---------------------------------------------------------------------------------------
// Backward from the 3-D stack's own layout: g is padded NDHWC bf16 [B][D+2][H+2][W+2][2C].
// CTA = one image row (b, y).  A thread owns output chunks (x, j): 8 channels of one pixel, j < C/8 from the
// first half of the voxel, j >= C/8 from the second; for every d it reads the 16-byte chunk of the voxel that
// pixel was copied to — (d, y, x) for the first half, (d, y, x +- d) for the second — so a warp's loads are
// contiguous runs of the row and nothing needs an atomic.  The fp32 sums are transposed through shared
// memory into the NCHW gradients of the two feature maps.
// ---------------------------------------------------------------------------------------
constexpr int CB_THREADS = 512;
constexpr int CB_MAXK = 8;                               // chunks per thread: W * 2C/8 <= 4096

__global__ void __launch_bounds__(CB_THREADS)
concat_bwd_ndhwc_kernel(const uint4* __restrict__ g, float* __restrict__ gL, float* __restrict__ gR,
                        int C, int D, int H, int W, int mode) {
    extern __shared__ __align__(16) float s_out[];        // [2C][W + 1]
    const int y = blockIdx.x, b = blockIdx.y;
    const int cpv = (2 * C) / 8, half = cpv / 2;
    // blockIdx.z = x-chunk: a CTA owns at most CB_THREADS*CB_MAXK/cpv pixels of the row, so any width works
    const int wchunk = (CB_THREADS * CB_MAXK) / cpv;
    const int x0 = blockIdx.z * wchunk;
    const int wc = min(wchunk, W - x0);                     // pixels of this CTA
    const int nchunks = wc * cpv;
    const int Wp = W + 2, Hp = H + 2;
    const size_t plane = (size_t)Hp * Wp * cpv;            // 16-byte chunks per padded plane
    const uint4* row = g + (((size_t)b * (D + 2) + 1) * Hp + (y + 1)) * Wp * cpv + cpv;   // voxel (d=0, y, x=0)
    float acc[CB_MAXK][8];
    int xk[CB_MAXK], offk[CB_MAXK];                        // pixel and chunk offset of each owned output chunk; xk < 0: none
#pragma unroll
    for (int k = 0; k < CB_MAXK; ++k) {
#pragma unroll
        for (int i = 0; i < 8; ++i) acc[k][i] = 0.f;
        const int idx = threadIdx.x + k * CB_THREADS;
        xk[k] = idx < nchunks ? x0 + idx / cpv : -1;
        offk[k] = x0 * cpv + idx;                          // = x * cpv + j
    }
    const bool second = (threadIdx.x % cpv) >= half;       // CB_THREADS % cpv == 0: the same half for every k
    const int step = second ? ((mode == DSM_VOL_GC_RIGHT) ? -cpv : cpv) : 0;      // chunk offset per unit of d
    for (int d = 0; d < D; ++d) {
        const uint4* rd = row + (size_t)d * plane;
#pragma unroll
        for (int k = 0; k < CB_MAXK; ++k) {
            const int x = xk[k];
            if (x < 0) continue;
            bool ok;
            if (!second) ok = (mode != DSM_VOL_PSM) || x >= d;
            else if (mode == DSM_VOL_GC_RIGHT) ok = x - d >= 0;
            else ok = x + d < W;
            if (ok) {
                const uint4 v = __ldg(rd + offk[k] + d * step);
                const uint32_t w4[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
                for (int i = 0; i < 4; ++i) { acc[k][2 * i] += bf16_lo(w4[i]); acc[k][2 * i + 1] += bf16_hi(w4[i]); }
            }
        }
    }
    const int pitch = wc + 1;
#pragma unroll
    for (int k = 0; k < CB_MAXK; ++k) {
        const int idx = threadIdx.x + k * CB_THREADS;
        if (idx >= nchunks) break;
        const int x = idx / cpv, j = idx - x * cpv;          // x local to the chunk
#pragma unroll
        for (int i = 0; i < 8; ++i) s_out[(j * 8 + i) * pitch + x] = acc[k][i];
    }
    __syncthreads();
    // channel ch < C is the first half: fL (fR for GC_RIGHT); ch >= C the other map
    for (int i = threadIdx.x; i < 2 * C * wc; i += CB_THREADS) {
        const int ch = i / wc, x = i - ch * wc;
        const bool first = ch < C;
        const int c = first ? ch : ch - C;
        float* dst = (first != (mode == DSM_VOL_GC_RIGHT)) ? gL : gR;
        dst[(((size_t)b * C + c) * H + y) * W + x0 + x] = s_out[ch * pitch + x];
    }
}

// ---------------------------------------------------------------------------------------
// NCDHW fp32 <-> padded NDHWC bf16.  CTA = (y', d', b); transposes a [C][W] slab through smem.
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
pack_ndhwc_kernel(const float* __restrict__ x, uint4* __restrict__ y, int C, int D, int H, int W) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __nv_bfloat16* s = reinterpret_cast<__nv_bfloat16*>(smem_raw);            // [W][C]
    const int yp = blockIdx.x, dp = blockIdx.y, b = blockIdx.z;
    const int Wp = W + 2, Hp = H + 2;
    const int cpv = C / 8;
    const bool rim = (yp == 0 || yp == H + 1 || dp == 0 || dp == D + 1);
    if (!rim) {
        const float* src = x + (((size_t)b * C * D + (dp - 1)) * H + (yp - 1)) * W;
        const size_t cstride = (size_t)D * H * W;
        for (int i = threadIdx.x; i < C * W; i += blockDim.x) {
            const int c = i / W, xx = i - c * W;
            s[xx * C + c] = __float2bfloat16_rn(__ldg(src + c * cstride + xx));
        }
    }
    __syncthreads();
    const uint4* v = reinterpret_cast<const uint4*>(s);
    uint4* orow = y + (((size_t)b * (D + 2) + dp) * Hp + yp) * (size_t)(Wp * cpv);
    for (int i = threadIdx.x; i < Wp * cpv; i += blockDim.x) {
        const int xp = i / cpv, k = i - xp * cpv;
        uint4 val = make_uint4(0u, 0u, 0u, 0u);
        if (!rim && xp >= 1 && xp <= W) val = v[(xp - 1) * cpv + k];
        st_stream_u4(orow + i, val);
    }
}

__global__ void __launch_bounds__(256)
unpack_ndhwc_kernel(const __nv_bfloat16* __restrict__ x, float* __restrict__ y, int C, int D, int H, int W) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __nv_bfloat16* s = reinterpret_cast<__nv_bfloat16*>(smem_raw);            // [W][C+2] (padded against conflicts)
    const int yy = blockIdx.x, d = blockIdx.y, b = blockIdx.z;
    const int Wp = W + 2, Hp = H + 2;
    const int CP = C + 2;
    const __nv_bfloat16* src = x + ((((size_t)b * (D + 2) + d + 1) * Hp + yy + 1) * Wp + 1) * (size_t)C;
    for (int i = threadIdx.x; i < W * C; i += blockDim.x) {
        const int xx = i / C, c = i - xx * C;
        s[xx * CP + c] = src[i];
    }
    __syncthreads();
    float* dst = y + (((size_t)b * C * D + d) * H + yy) * W;
    const size_t cstride = (size_t)D * H * W;
    for (int i = threadIdx.x; i < C * W; i += blockDim.x) {
        const int c = i / W, xx = i - c * W;
        dst[c * cstride + xx] = __bfloat162float(s[xx * CP + c]);
    }
}

// NCHW fp32 -> NHWC bf16 with a zero rim of `rim` pixels, [B][H+2r][W+2r][C] (the feature-map layout of the fused volume
// convolution and of the 2-D trunk).  One thread per PADDED pixel: a warp's 32 neighbouring pixels make every per-channel load
// a coalesced 128-byte run, and the thread's C channels leave as consecutive 16-byte stores (a warp fills a contiguous
// 32 * 2C-byte span) — no shared-memory transpose, all of a thread's loads in flight at once.  (The first version staged a row
// through shared memory with a division per element and 16-way conflicted stores: 24.6 us for 11.5 MB at PSMNet's size.)
__global__ void __launch_bounds__(256)
pack_nhwc_kernel(const float* __restrict__ x, uint4* __restrict__ y, const float* __restrict__ x2, uint4* __restrict__ y2,
                 int C, int H, int W, int rim, long long npix) {
    if (blockIdx.z == 1) { x = x2; y = y2; }                                  // the second map of a pair (one launch for both)
    const long long i = (long long)blockIdx.x * 256 + threadIdx.x;            // padded pixel index over [B][H+2r][W+2r]
    if (i >= npix) return;
    const int Hp = H + 2 * rim, Wp = W + 2 * rim, cpv = C / 8;
    const int xp = (int)(i % Wp);
    const long long t = i / Wp;
    const int yp = (int)(t % Hp), b = (int)(t / Hp);
    const int xx = xp - rim, yy = yp - rim;
    uint4* o = y + (size_t)i * cpv;
    if (xx < 0 || xx >= W || yy < 0 || yy >= H) {
        for (int k = 0; k < cpv; ++k) o[k] = make_uint4(0u, 0u, 0u, 0u);
        return;
    }
    const size_t cstride = (size_t)H * W;
    const float* src = x + ((size_t)b * C * H + yy) * W + xx;
#pragma unroll 4
    for (int k = 0; k < cpv; ++k) {
        float f[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) f[e] = __ldg(src + (size_t)(8 * k + e) * cstride);
        uint4 v;
        v.x = pack_bf16x2(f[0], f[1]); v.y = pack_bf16x2(f[2], f[3]); v.z = pack_bf16x2(f[4], f[5]); v.w = pack_bf16x2(f[6], f[7]);
        o[k] = v;
    }
}

}  // namespace

extern "C" int dsm_pack_nhwc_bf16(const float* x, void* y, int B, int C, int H, int W, int rim, void* stream) {
    DsmDeviceGuard dsm_guard_(x);
    if (!x || !y || B <= 0 || C <= 0 || H <= 0 || W <= 0 || rim < 0 || rim > 8) return DSM_EINVAL;
    if (C % 8 != 0 || H > 65535 * 32 || B > 65535) return DSM_EUNSUPPORTED;
    if (!dsm_aligned16(y)) return DSM_EALIGN;
    const long long npix = (long long)B * (H + 2 * rim) * (W + 2 * rim);
    const long long blocks = dsm_ceil_div_ll(npix, 256);
    if (blocks > 0x7fffffffLL) return DSM_EUNSUPPORTED;
    pack_nhwc_kernel<<<dim3((unsigned)blocks, 1, 1), 256, 0, (cudaStream_t)stream>>>(x, (uint4*)y, nullptr, nullptr, C, H, W, rim, npix);
    return dsm_launch_status();
}

extern "C" int dsm_pack_nhwc_bf16_pair(const float* xL, const float* xR, void* yL, void* yR, int B, int C, int H, int W, int rim, void* stream) {
    DsmDeviceGuard dsm_guard_(xL);
    if (!xL || !xR || !yL || !yR || B <= 0 || C <= 0 || H <= 0 || W <= 0 || rim < 0 || rim > 8) return DSM_EINVAL;
    if (C % 8 != 0 || H > 65535 * 32 || B > 65535) return DSM_EUNSUPPORTED;
    if (!dsm_aligned16(yL) || !dsm_aligned16(yR)) return DSM_EALIGN;
    const long long npix = (long long)B * (H + 2 * rim) * (W + 2 * rim);
    const long long blocks = dsm_ceil_div_ll(npix, 256);
    if (blocks > 0x7fffffffLL) return DSM_EUNSUPPORTED;
    pack_nhwc_kernel<<<dim3((unsigned)blocks, 1, 2), 256, 0, (cudaStream_t)stream>>>(xL, (uint4*)yL, xR, (uint4*)yR, C, H, W, rim, npix);
    return dsm_launch_status();
}

extern "C" int dsm_concat_volume_fwd(const float* fL, const float* fR, void* out,
                                     int B, int C, int D, int H, int W,
                                     int mode, int out_dtype, int out_layout, void* stream) {
    DsmDeviceGuard dsm_guard_(fL);
    if (!fL || !fR || !out || B <= 0 || C <= 0 || D <= 0 || H <= 0 || W <= 0) return DSM_EINVAL;
    if (mode < DSM_VOL_PSM || mode > DSM_VOL_GC_RIGHT) return DSM_EINVAL;
    if (B > 65535 || 2 * C > 65535) return DSM_EUNSUPPORTED;
    if (!dsm_aligned16(out)) return DSM_EALIGN;
    cudaStream_t st = (cudaStream_t)stream;
    if (out_dtype == DSM_F32 && out_layout == DSM_NCDHW) {
        const size_t smem = (size_t)W * sizeof(float);
        if (smem > 48 * 1024) return DSM_EUNSUPPORTED;
        concat_ncdhw_kernel<<<dim3(H, 2 * C, B), 128, smem, st>>>(fL, fR, (float*)out, C, D, H, W, mode);
        return dsm_launch_status();
    }
    if (out_dtype == DSM_BF16 && out_layout == DSM_NDHWC_PADDED) {
        if (C % 8 != 0) return DSM_EUNSUPPORTED;
        const size_t smem = 2 * (size_t)W * C * sizeof(__nv_bfloat16);
        if (smem > 200 * 1024) return DSM_EUNSUPPORTED;
        cudaError_t e = cudaFuncSetAttribute(concat_ndhwc_bf16_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return (int)e;
        // enough CTAs for >= 4 waves over 148 SMs: split the padded disparity range into slabs
        int slabs = dsm_ceil_div(4 * DSM_NUM_SMS_B200, (H + 2) * B);
        if (slabs < 1) slabs = 1;
        if (slabs > D + 2) slabs = D + 2;
        const int dslab = dsm_ceil_div(D + 2, slabs);
        slabs = dsm_ceil_div(D + 2, dslab);
        concat_ndhwc_bf16_kernel<<<dim3(H + 2, slabs, B), 256, smem, st>>>(fL, fR, (uint4*)out, C, D, H, W, mode, dslab);
        return dsm_launch_status();
    }
    return DSM_EUNSUPPORTED;
}

extern "C" int dsm_concat_volume_bwd(const void* gout, float* gL, float* gR,
                                     int B, int C, int D, int H, int W,
                                     int mode, int dtype, int layout, void* stream) {
    DsmDeviceGuard dsm_guard_(gout);
    if (!gout || !gL || !gR || B <= 0 || C <= 0 || D <= 0 || H <= 0 || W <= 0) return DSM_EINVAL;
    if (mode < DSM_VOL_PSM || mode > DSM_VOL_GC_RIGHT) return DSM_EINVAL;
    if (dtype == DSM_BF16 && layout == DSM_NDHWC_PADDED) {
        if (C % 8 != 0 || CB_THREADS % (2 * C / 8) != 0 || H > 65535 || B > 65535)
            return DSM_EUNSUPPORTED;
        if (!dsm_aligned16(gout)) return DSM_EALIGN;
        const int wchunk = (CB_THREADS * CB_MAXK) / (2 * C / 8);       // pixels per CTA; wider rows take several CTAs
        const int xchunks = dsm_ceil_div(W, wchunk);
        const size_t smem = (size_t)2 * C * ((W < wchunk ? W : wchunk) + 1) * sizeof(float);
        if (smem > 200 * 1024) return DSM_EUNSUPPORTED;
        cudaError_t e = cudaFuncSetAttribute(concat_bwd_ndhwc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return (int)e;
        concat_bwd_ndhwc_kernel<<<dim3(H, B, xchunks), CB_THREADS, smem, (cudaStream_t)stream>>>((const uint4*)gout, gL, gR, C, D, H, W, mode);
        return dsm_launch_status();
    }
    if (dtype != DSM_F32 || layout != DSM_NCDHW) return DSM_EUNSUPPORTED;
    const long long n = (long long)B * C * H * W;
    const long long blocks = dsm_ceil_div_ll(n, 256);
    if (blocks > 0x7fffffffLL) return DSM_EUNSUPPORTED;
    concat_bwd_ncdhw_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>((const float*)gout, gL, gR, B, C, D, H, W, mode);
    return dsm_launch_status();
}

// Filter repacking: PyTorch fp32 weights -> the kernels' bf16 [27][CoutP][Cin] (tap-major, K-major rows), one launch.
//   mode 0: w is [Cout][Cin][27]  (nn.Conv3d)                      out[t][co][ci] = w[co][ci][t]
//   mode 1: w is [Cin][Cout][27]  (nn.ConvTranspose3d)             out[t][co][ci] = w[ci][co][t]
//   mode 2: w is [Cin][Cout][27] read as the stride-1 dgrad filter out[t][co][ci] = w[ci][co][26-t]   (flipped taps)
//   mode 3: w is [1][Cin][27] (nn.Conv3d with ONE output channel), kw folded into the output columns for the plane-sharing
//           kernel's kw-fold mode (dsm_conv3d_fwd_ex variant bit 8): out[(kd,kh,0)][c'][ci] = w[0][ci][(kd,kh,c')] for c' < 3, else 0
// Rows co >= Cout (CoutP = max(16, Cout)) are zero.
__global__ void __launch_bounds__(256)
pack_weight_kernel(const float* __restrict__ w, __nv_bfloat16* __restrict__ out, int Cout, int CoutP, int Cin, int mode) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int n = 27 * CoutP * Cin;
    if (i >= n) return;
    const int ci = i % Cin; int r = i / Cin;
    const int co = r % CoutP; const int t = r / CoutP;
    float v = 0.f;
    if (mode == 3) {
        // single-output-channel Conv3d, kw folded into the output columns: tile (kd, kh, kw = 0) row c' < 3 = w[0][ci][kd][kh][c']
        if (t % 3 == 0 && co < 3) v = w[(size_t)ci * 27 + t + co];
    } else if (co < Cout) {
        if (mode == 0) v = w[((size_t)co * Cin + ci) * 27 + t];
        else v = w[((size_t)ci * Cout + co) * 27 + (mode == 2 ? 26 - t : t)];
    }
    out[i] = __float2bfloat16_rn(v);
}

extern "C" int dsm_pack_weight(const float* w, void* out, int Cout, int Cin, int mode, void* stream) {
    DsmDeviceGuard dsm_guard_(w);
    if (!w || !out || Cout < 1 || Cin < 1 || mode < 0 || mode > 3 || (mode == 3 && Cout != 1)) return DSM_EINVAL;
    const int CoutP = Cout < 16 ? 16 : Cout;
    const long long n = 27LL * CoutP * Cin;
    if (n > 0x7fffffffLL) return DSM_EUNSUPPORTED;
    pack_weight_kernel<<<(unsigned)dsm_ceil_div_ll(n, 256), 256, 0, (cudaStream_t)stream>>>(w, static_cast<__nv_bfloat16*>(out), Cout, CoutP, Cin, mode);
    return dsm_launch_status();
}

extern "C" int dsm_pack_ndhwc(const float* x, void* y, int B, int C, int D, int H, int W, void* stream) {
    DsmDeviceGuard dsm_guard_(x);
    if (!x || !y || B <= 0 || C <= 0 || D <= 0 || H <= 0 || W <= 0) return DSM_EINVAL;
    if (C % 8 != 0 || D + 2 > 65535 || B > 65535) return DSM_EUNSUPPORTED;
    if (!dsm_aligned16(y)) return DSM_EALIGN;
    const size_t smem = (size_t)W * C * sizeof(__nv_bfloat16);
    if (smem > 200 * 1024) return DSM_EUNSUPPORTED;
    cudaError_t e = cudaFuncSetAttribute(pack_ndhwc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    pack_ndhwc_kernel<<<dim3(H + 2, D + 2, B), 256, smem, (cudaStream_t)stream>>>(x, (uint4*)y, C, D, H, W);
    return dsm_launch_status();
}

extern "C" int dsm_unpack_ndhwc(const void* x, float* y, int B, int C, int D, int H, int W, void* stream) {
    DsmDeviceGuard dsm_guard_(x);
    if (!x || !y || B <= 0 || C <= 0 || D <= 0 || H <= 0 || W <= 0) return DSM_EINVAL;
    if (D > 65535 || B > 65535) return DSM_EUNSUPPORTED;
    const size_t smem = (size_t)W * (C + 2) * sizeof(__nv_bfloat16);
    if (smem > 200 * 1024) return DSM_EUNSUPPORTED;
    cudaError_t e = cudaFuncSetAttribute(unpack_ndhwc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    unpack_ndhwc_kernel<<<dim3(H, D, B), 256, smem, (cudaStream_t)stream>>>((const __nv_bfloat16*)x, y, C, D, H, W);
    return dsm_launch_status();
}
