// CUDA-core companions of the 2-D trunk convolutions (the tensor-core layers are dsm_conv2d_fwd in conv3d.cu):
//   * dsm_conv2d_first_fwd — the image-facing convolution: Conv2d(3 -> 32, k 3|5, stride 2, pad k/2) + folded BatchNorm + ReLU
//     from the reference's NCHW fp32 image to the trunk's padded NHWC bf16 layout (models/psmnet/submodule.py:68,
//     models/gcnet.py:21).  K = 27 / 75 is no tensor-core shape and the layer is 0.1 % of the trunk's flops: it is an HBM
//     stream (read 12 B, write 64 B per output pixel) with 864 / 2400 FMAs per pixel on the way.
//   * dsm_spp_fwd — PSMNet's spatial-pyramid branches (submodule.py:84-98,126-136): AvgPool 64/32/16/8 of the 128-channel
//     skip tensor, 1x1 convolution 128 -> 32 + BatchNorm + ReLU (with the reference's padding quirk: convbn pads by
//     `dilation` = 1 even for k = 1, so the pooled map grows by a ring of relu(BatchNorm(0))), bilinear upsampling
//     (align_corners as given) back to the skip tensor's size, written as bf16 into the branch slices of the 320-channel
//     concatenation buffer.  Three small launches; the 8x8 cell sums are shared by all four levels (16 = 2x2, 32 = 4x4,
//     64 = 8x8 cells: the windows start at multiples of their size and never overlap).
#include "common.cuh"

namespace {

template <int K>
__global__ void __launch_bounds__(128)
conv_first_kernel(const float* __restrict__ img, const float* __restrict__ w, const float* __restrict__ scale,
                  const float* __restrict__ shift, uint4* __restrict__ out, int B, int H, int W, int Ho, int Wo, int ro, int relu) {
    // w: [32][3][K][K] (PyTorch layout) -> shared [c][kh][kw][32]
    __shared__ float sw[3 * K * K * 32];
    __shared__ float ssc[32], ssh[32];
    for (int i = threadIdx.x; i < 3 * K * K * 32; i += 128) {
        const int co = i & 31, t = i >> 5;                       // t = (c*K + kh)*K + kw
        sw[i] = w[(size_t)co * 3 * K * K + t];
    }
    if (threadIdx.x < 32) { ssc[threadIdx.x] = scale ? scale[threadIdx.x] : 1.f; ssh[threadIdx.x] = shift ? shift[threadIdx.x] : 0.f; }
    __syncthreads();
    const long long n = (long long)B * Ho * Wo;
    const long long i = (long long)blockIdx.x * 128 + threadIdx.x;
    if (i >= n) return;
    const int ox = (int)(i % Wo); long long t = i / Wo;
    const int oy = (int)(t % Ho); const int b = (int)(t / Ho);
    float acc[32];
#pragma unroll
    for (int c = 0; c < 32; ++c) acc[c] = 0.f;
    constexpr int P = K / 2;
    for (int c = 0; c < 3; ++c) {
        const float* plane = img + ((size_t)b * 3 + c) * H * W;
#pragma unroll
        for (int kh = 0; kh < K; ++kh) {
            const int iy = 2 * oy + kh - P;
            if (iy < 0 || iy >= H) continue;
#pragma unroll
            for (int kw = 0; kw < K; ++kw) {
                const int ix = 2 * ox + kw - P;
                if (ix < 0 || ix >= W) continue;
                const float v = __ldg(plane + (size_t)iy * W + ix);
                const float* wr = sw + ((c * K + kh) * K + kw) * 32;
#pragma unroll
                for (int co = 0; co < 32; ++co) acc[co] = fmaf(v, wr[co], acc[co]);
            }
        }
    }
    uint32_t pk[16];
#pragma unroll
    for (int c = 0; c < 16; ++c) {
        float a0 = fmaf(acc[2 * c], ssc[2 * c], ssh[2 * c]), a1 = fmaf(acc[2 * c + 1], ssc[2 * c + 1], ssh[2 * c + 1]);
        if (relu) { a0 = fmaxf(a0, 0.f); a1 = fmaxf(a1, 0.f); }
        pk[c] = pack_bf16x2(a0, a1);
    }
    uint4* o = out + ((((size_t)b * (Ho + 2 * ro) + oy + ro) * (Wo + 2 * ro)) + ox + ro) * 4;     // 32 bf16 = 4 x 16 B
#pragma unroll
    for (int q = 0; q < 4; ++q) o[q] = make_uint4(pk[4 * q], pk[4 * q + 1], pk[4 * q + 2], pk[4 * q + 3]);
}

// ---- SPP ------------------------------------------------------------------------------------------------------
// cells: [B][ch][cw][128] fp32 sums over 8x8 pixel cells (ch = H/8, cw = W/8, floor).  CTA = one cell: thread (dx = t / 16,
// cg = t % 16) reads the 8 channels [8 cg, 8 cg + 8) of the cell's column dx for all 8 rows as 16-byte loads (all in flight),
// then the 8 columns are added through shared memory.
__global__ void __launch_bounds__(128)
spp_cells_kernel(const __nv_bfloat16* __restrict__ skip, float* __restrict__ cells, int H, int W, int rim, int ld, int ch, int cw) {
    const int cx = blockIdx.x, cy = blockIdx.y, b = blockIdx.z, t = threadIdx.x;
    const int dx = t >> 4, cg = t & 15;
    const int Wp = W + 2 * rim, Hp = H + 2 * rim;
    __shared__ float part[8][128];
    const __nv_bfloat16* p0 = skip + (((size_t)b * Hp + cy * 8 + rim) * Wp + cx * 8 + dx + rim) * ld + cg * 8;
    uint4 v[8];
#pragma unroll
    for (int dy = 0; dy < 8; ++dy) v[dy] = __ldg(reinterpret_cast<const uint4*>(p0 + (size_t)dy * Wp * ld));
    float s[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) s[e] = 0.f;
#pragma unroll
    for (int dy = 0; dy < 8; ++dy) {                             // rows in order, as the reference's pooling window is summed
        s[0] += bf16_lo(v[dy].x); s[1] += bf16_hi(v[dy].x); s[2] += bf16_lo(v[dy].y); s[3] += bf16_hi(v[dy].y);
        s[4] += bf16_lo(v[dy].z); s[5] += bf16_hi(v[dy].z); s[6] += bf16_lo(v[dy].w); s[7] += bf16_hi(v[dy].w);
    }
#pragma unroll
    for (int e = 0; e < 8; ++e) part[dx][cg * 8 + e] = s[e];
    __syncthreads();
    float a = 0.f;
#pragma unroll
    for (int x = 0; x < 8; ++x) a += part[x][t];
    cells[(((size_t)b * ch + cy) * cw + cx) * 128 + t] = a;
}

struct SppLevel { int k8, ph, pw, off; };        // window = k8 x k8 cells; pooled map ph x pw; `off` = first element of this level in `br`
struct SppGeom { int B, H, W, ch, cw; SppLevel lv[4]; };

// br: for level l, [B][ph+2][pw+2][32] fp32 = relu(BN(conv1x1(avgpool))) with the ring of relu(shift); one CTA of 128 threads
// per output pixel: thread = channel for the pooling (up to 64 cells, four independent partial sums so that the loads overlap)
// and for its share of the 128 -> 32 dot products (see below).
__global__ void __launch_bounds__(128)
spp_branch_kernel(const float* __restrict__ cells, const float* __restrict__ w /*[4][32][128]*/, const float* __restrict__ scale /*[4][32]*/,
                  const float* __restrict__ shift, float* __restrict__ br, SppGeom g) {
    const int l = blockIdx.z & 3, b = blockIdx.z >> 2;
    const SppLevel L = g.lv[l];
    const int px = blockIdx.x, py = blockIdx.y;
    if (px >= L.pw + 2 || py >= L.ph + 2) return;
    const int t = threadIdx.x, co = t & 31, q = t >> 5;
    __shared__ float part[4][32];
    float acc = 0.f;
    if (px >= 1 && px <= L.pw && py >= 1 && py <= L.ph) {          // uniform over the CTA
        const float inv = 1.f / (float)(L.k8 * L.k8 * 64);
        const float* cb = cells + (((size_t)b * g.ch + (py - 1) * L.k8) * g.cw + (px - 1) * L.k8) * 128 + t;
        float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
        if (L.k8 >= 4) {
            for (int yy = 0; yy < L.k8; ++yy) {
                const float* r = cb + (size_t)yy * g.cw * 128;
                for (int xx = 0; xx < L.k8; xx += 4) {
                    s0 += r[(xx + 0) * 128]; s1 += r[(xx + 1) * 128]; s2 += r[(xx + 2) * 128]; s3 += r[(xx + 3) * 128];
                }
            }
        } else {
            for (int yy = 0; yy < L.k8; ++yy)
                for (int xx = 0; xx < L.k8; ++xx) s0 += cb[((size_t)yy * g.cw + xx) * 128];
        }
        // 128 -> 32 dot products with the thread still owning channel t: lanes read 32 consecutive weights of output co (the
        // PyTorch layout [co][c]; one lane per OUTPUT read 32 different 128-byte lines per load), a warp reduces its 32
        // channels with shuffles, the four warps' partial sums meet in shared memory
        const float pc = ((s0 + s1) + (s2 + s3)) * inv;
        const float* wl = w + (size_t)l * 32 * 128 + t;
#pragma unroll 4
        for (int o = 0; o < 32; ++o) {
            float p = pc * __ldg(wl + o * 128);
            p += __shfl_xor_sync(0xffffffffu, p, 16); p += __shfl_xor_sync(0xffffffffu, p, 8);
            p += __shfl_xor_sync(0xffffffffu, p, 4);  p += __shfl_xor_sync(0xffffffffu, p, 2);
            p += __shfl_xor_sync(0xffffffffu, p, 1);
            if (co == 0) part[q][o] = p;
        }
        __syncthreads();
        acc = (part[0][co] + part[1][co]) + (part[2][co] + part[3][co]);
    }
    if (t < 32) {
        const float v = fmaxf(fmaf(acc, scale[l * 32 + co], shift[l * 32 + co]), 0.f);
        br[L.off + (((size_t)b * (L.ph + 2) + py) * (L.pw + 2) + px) * 32 + co] = v;
    }
}

// bilinear upsampling of the four branch maps to H x W, bf16 into channels [coff + 32*slot, +32) of the concat buffer;
// slot order in the concatenation is (branch4, branch3, branch2, branch1) = levels (3, 2, 1, 0) (submodule.py:138)
__global__ void __launch_bounds__(256)
spp_upsample_kernel(const float* __restrict__ br, __nv_bfloat16* __restrict__ cat, SppGeom g, int rim, int ld, int coff, int align_corners) {
    // grid = (ceil(16 W / 256), B * H): a thread is (pixel x of row blockIdx.y, level, 8-channel group).  (The first version
    // decoded a flat 64-bit index with four 64-bit divisions per thread: most of the SPP's 48 us.)
    const int i = blockIdx.x * 256 + threadIdx.x;
    if (i >= g.W * 16) return;
    const int grp = i & 3, l = (i >> 2) & 3;
    const int x = i >> 4;
    const int b = (int)blockIdx.y / g.H, y = (int)blockIdx.y - b * g.H;
    const SppLevel L = g.lv[l];
    const int ih = L.ph + 2, iw = L.pw + 2;
    float sy, sx;
    if (align_corners) {
        sy = (g.H > 1) ? (float)(ih - 1) / (float)(g.H - 1) * (float)y : 0.f;
        sx = (g.W > 1) ? (float)(iw - 1) / (float)(g.W - 1) * (float)x : 0.f;
    } else {
        sy = fmaxf(((float)y + 0.5f) * ((float)ih / (float)g.H) - 0.5f, 0.f);
        sx = fmaxf(((float)x + 0.5f) * ((float)iw / (float)g.W) - 0.5f, 0.f);
    }
    const int y0 = min((int)sy, ih - 1), x0 = min((int)sx, iw - 1);
    const int y1 = min(y0 + 1, ih - 1), x1 = min(x0 + 1, iw - 1);
    const float ly = sy - (float)y0, lx = sx - (float)x0, hy = 1.f - ly, hx = 1.f - lx;
    const float* base = br + L.off + (size_t)b * ih * iw * 32 + grp * 8;
    const float4* p00 = reinterpret_cast<const float4*>(base + ((size_t)y0 * iw + x0) * 32);
    const float4* p01 = reinterpret_cast<const float4*>(base + ((size_t)y0 * iw + x1) * 32);
    const float4* p10 = reinterpret_cast<const float4*>(base + ((size_t)y1 * iw + x0) * 32);
    const float4* p11 = reinterpret_cast<const float4*>(base + ((size_t)y1 * iw + x1) * 32);
    float v[8];
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        const float4 a = p00[h], bq = p01[h], c = p10[h], d = p11[h];
        // PyTorch's upsample_bilinear2d: h0lambda * (w0lambda * a + w1lambda * b) + h1lambda * (w0lambda * c + w1lambda * d)
        v[4 * h + 0] = hy * (hx * a.x + lx * bq.x) + ly * (hx * c.x + lx * d.x);
        v[4 * h + 1] = hy * (hx * a.y + lx * bq.y) + ly * (hx * c.y + lx * d.y);
        v[4 * h + 2] = hy * (hx * a.z + lx * bq.z) + ly * (hx * c.z + lx * d.z);
        v[4 * h + 3] = hy * (hx * a.w + lx * bq.w) + ly * (hx * c.w + lx * d.w);
    }
    const int slot = 3 - l;
    __nv_bfloat16* o = cat + ((((size_t)b * (g.H + 2 * rim) + y + rim) * (g.W + 2 * rim)) + x + rim) * ld + coff + slot * 32 + grp * 8;
    *reinterpret_cast<uint4*>(o) = make_uint4(pack_bf16x2(v[0], v[1]), pack_bf16x2(v[2], v[3]), pack_bf16x2(v[4], v[5]), pack_bf16x2(v[6], v[7]));
}

bool spp_geom(SppGeom& g, int B, int H, int W) {
    g.B = B; g.H = H; g.W = W; g.ch = H / 8; g.cw = W / 8;
    const int ks[4] = {64, 32, 16, 8};
    int off = 0;
    for (int l = 0; l < 4; ++l) {
        g.lv[l].k8 = ks[l] / 8; g.lv[l].ph = H / ks[l]; g.lv[l].pw = W / ks[l]; g.lv[l].off = off;
        if (g.lv[l].ph < 1 || g.lv[l].pw < 1) return false;          // nn.AvgPool2d raises on an empty output too
        off += B * (g.lv[l].ph + 2) * (g.lv[l].pw + 2) * 32;
    }
    return true;
}

}  // namespace

extern "C" int dsm_conv2d_first_fwd(const float* img, const float* w, const float* scale, const float* shift, void* out,
                                    int B, int H, int W, int ksize, int relu, int rim_out, void* stream) {
    DsmDeviceGuard dsm_guard_(img);
    if (!img || !w || !out || B < 1 || H < 1 || W < 1 || rim_out < 0 || rim_out > 2) return DSM_EINVAL;
    if (ksize != 3 && ksize != 5) return DSM_EUNSUPPORTED;
    if (!dsm_aligned16(out)) return DSM_EALIGN;
    const int Ho = (H - 1) / 2 + 1, Wo = (W - 1) / 2 + 1;           // floor((H + 2*(k/2) - k) / 2) + 1
    const long long n = (long long)B * Ho * Wo;
    const long long blocks = dsm_ceil_div_ll(n, 128);
    if (blocks > 0x7fffffffLL) return DSM_EUNSUPPORTED;
    cudaStream_t st = (cudaStream_t)stream;
    if (ksize == 3) conv_first_kernel<3><<<(unsigned)blocks, 128, 0, st>>>(img, w, scale, shift, (uint4*)out, B, H, W, Ho, Wo, rim_out, relu);
    else            conv_first_kernel<5><<<(unsigned)blocks, 128, 0, st>>>(img, w, scale, shift, (uint4*)out, B, H, W, Ho, Wo, rim_out, relu);
    return dsm_launch_status();
}

extern "C" size_t dsm_spp_workspace_bytes(int B, int H, int W) {
    SppGeom g;
    if (B < 1 || H < 1 || W < 1 || !spp_geom(g, B, H, W)) return 0;
    const size_t cells = (size_t)B * g.ch * g.cw * 128;
    const size_t br = (size_t)g.lv[3].off + (size_t)B * (g.lv[3].ph + 2) * (g.lv[3].pw + 2) * 32;
    return (cells + br) * sizeof(float);
}

extern "C" int dsm_spp_fwd(const void* skip, const float* w, const float* scale, const float* shift, void* cat,
                           int B, int H, int W, int rim, int ld, int coff, int align_corners,
                           void* ws, size_t ws_bytes, void* stream) {
    DsmDeviceGuard dsm_guard_(skip);
    if (!skip || !w || !scale || !shift || !cat || !ws || B < 1 || H < 1 || W < 1 || rim < 0 || ld < 128 || coff < 0) return DSM_EINVAL;
    if ((ld & 7) || (coff & 7) || coff + 128 > ld) return DSM_EINVAL;
    SppGeom g;
    if (!spp_geom(g, B, H, W)) return DSM_EINVAL;                   // the skip tensor must be at least 64 x 64 (AvgPool2d(64))
    if (ws_bytes < dsm_spp_workspace_bytes(B, H, W)) return DSM_EINVAL;
    if (!dsm_aligned16(cat) || !dsm_aligned16(ws) || !dsm_aligned16(skip)) return DSM_EALIGN;
    if (B > 16383 || g.ch > 65535 || (long long)B * H > 65535 || W > (1 << 24)) return DSM_EUNSUPPORTED;
    cudaStream_t st = (cudaStream_t)stream;
    float* cells = static_cast<float*>(ws);
    float* br = cells + (size_t)B * g.ch * g.cw * 128;
    spp_cells_kernel<<<dim3(g.cw, g.ch, B), 128, 0, st>>>(static_cast<const __nv_bfloat16*>(skip), cells, H, W, rim, ld, g.ch, g.cw);
    spp_branch_kernel<<<dim3(g.lv[3].pw + 2, g.lv[3].ph + 2, B * 4), 128, 0, st>>>(cells, w, scale, shift, br, g);
    spp_upsample_kernel<<<dim3((unsigned)dsm_ceil_div(W * 16, 256), (unsigned)(B * H)), 256, 0, st>>>(br, static_cast<__nv_bfloat16*>(cat), g, rim, ld, coff, align_corners);
    return dsm_launch_status();
}
