// op 5 — imwrap bilinear warp (self-supervised photometric/SSIM loss, iResNet feature constancy).
// Replaces the device part of imwrap_BCHW, reference utils/imwrap.py:59-71:
//   grid.x = k*(row[j] - disp[b,i,j]*2.0/(W0-1)), grid.y = col[i]          (imwrap.py:62-67)
//   out    = F.grid_sample(src + delt, grid)   bilinear, zero padding, align_corners=True (:70-71)
// The sampling coordinate is computed with exactly the fp32 operation sequence of the reference
// (mul, IEEE div, sub, mul, then ATen's ((g+1)/2)*(size-1) un-normalisation), written with
// non-contractable intrinsics so that floor() lands on the same source pixel (bit-exact indexing).
// One thread per output pixel computes the coordinate once and loops over channels; reads of
// neighbouring lanes hit neighbouring source pixels, writes are coalesced.
// Algorithmic HBM bytes fwd: 4*B*(C*H0*W0 + H*W + C*H*W).
#include "common.cuh"

namespace {

struct Tap {
    int x0, y0;            // north-west source pixel
    float nw, ne, sw, se;  // bilinear weights
    float tx, ty;          // ix - x0, iy - y0 (for the coordinate gradient)
    bool vx0, vx1, vy0, vy1;
};

__device__ __forceinline__ Tap make_tap(float d, float rowv, float colv, float k, int H0, int W0) {
    const float wm1 = (float)(W0 - 1), hm1 = (float)(H0 - 1);
    const float q  = __fdiv_rn(__fmul_rn(d, 2.0f), wm1);          // disp*2.0/(w0-1)
    const float gx = __fmul_rn(k, __fsub_rn(rowv, q));            // k*(grid.x - q)
    const float gy = colv;
    const float ix = __fmul_rn(__fdiv_rn(__fadd_rn(gx, 1.0f), 2.0f), wm1);   // ((g+1)/2)*(size-1)
    const float iy = __fmul_rn(__fdiv_rn(__fadd_rn(gy, 1.0f), 2.0f), hm1);
    const float fx = floorf(ix), fy = floorf(iy);
    Tap t;
    t.x0 = (int)fx; t.y0 = (int)fy;
    // weights as ATen's CPU kernel forms them: w = x - floor(x), e = 1 - w, n = y - floor(y), s = 1 - n
    t.tx = __fsub_rn(ix, fx); t.ty = __fsub_rn(iy, fy);
    const float ex = __fsub_rn(1.0f, t.tx);
    const float ey = __fsub_rn(1.0f, t.ty);
    t.nw = __fmul_rn(ex, ey); t.ne = __fmul_rn(t.tx, ey);
    t.sw = __fmul_rn(ex, t.ty); t.se = __fmul_rn(t.tx, t.ty);
    t.vx0 = (t.x0 >= 0 && t.x0 < W0);  t.vx1 = (t.x0 + 1 >= 0 && t.x0 + 1 < W0);
    t.vy0 = (t.y0 >= 0 && t.y0 < H0);  t.vy1 = (t.y0 + 1 >= 0 && t.y0 + 1 < H0);
    // a non-finite coordinate (inf/nan disparity) samples nothing, as in ATen
    if (!(ix > -2.0f && ix < (float)W0 + 1.0f) || !(iy > -2.0f && iy < (float)H0 + 1.0f)) {
        t.vx0 = t.vx1 = t.vy0 = t.vy1 = false; t.x0 = 0; t.y0 = 0;
    }
    return t;
}

__device__ __forceinline__ void warp_fwd_body(const float* __restrict__ src, const float* __restrict__ disp,
                                              const float* __restrict__ row, const float* __restrict__ col,
                                              float delt, float k, float* __restrict__ out,
                                              int C, int H0, int W0, int H, int W, int j, int i, int b) {
    if (j >= W) return;
    const Tap t = make_tap(__ldg(disp + ((size_t)b * H + i) * W + j), __ldg(row + j), __ldg(col + i), k, H0, W0);
    const size_t sp = (size_t)H0 * W0, op = (size_t)H * W;
    const float* s = src + (size_t)b * C * sp;
    float* o = out + (size_t)b * C * op + (size_t)i * W + j;
    const int a00 = t.y0 * W0 + t.x0;
    const bool v00 = t.vy0 && t.vx0, v01 = t.vy0 && t.vx1, v10 = t.vy1 && t.vx0, v11 = t.vy1 && t.vx1;
#pragma unroll 4
    for (int c = 0; c < C; ++c) {
        const float* p = s + c * sp + a00;
        float r = 0.f;
        if (v00) r = fmaf(__ldg(p) + delt, t.nw, r);
        if (v01) r = fmaf(__ldg(p + 1) + delt, t.ne, r);
        if (v10) r = fmaf(__ldg(p + W0) + delt, t.sw, r);
        if (v11) r = fmaf(__ldg(p + W0 + 1) + delt, t.se, r);
        o[c * op] = r;
    }
}

__global__ void __launch_bounds__(128)
warp_fwd_kernel(const float* __restrict__ src, const float* __restrict__ disp,
                const float* __restrict__ row, const float* __restrict__ col,
                float delt, float k, float* __restrict__ out,
                int C, int H0, int W0, int H, int W) {
    warp_fwd_body(src, disp, row, col, delt, k, out, C, H0, W0, H, W, blockIdx.x * blockDim.x + threadIdx.x, blockIdx.y, blockIdx.z);
}

// backward: gsrc[b,c,tap] += w_tap * g   (scatter, fp32 red.global.add)
//           gdisp[b,i,j]   = -k*(2/(W0-1))*((W0-1)/2) * sum_c g * d(out)/d(ix)
// with d(out)/d(ix) = (ne_v - nw_v)*(y0+1-iy) + (se_v - sw_v)*(iy-y0) over in-range tap values
// (tap value = src + delt).  The chain through the normalise/un-normalise pair is applied as the
// two separate factors ATen and autograd use ((W0-1)/2, then 2.0/(W0-1)).
__device__ __forceinline__ void warp_bwd_body(const float* __restrict__ gout, const float* __restrict__ src, const float* __restrict__ disp,
                                              const float* __restrict__ row, const float* __restrict__ col,
                                              float delt, float k, float* __restrict__ gsrc, float* __restrict__ gdisp,
                                              int C, int H0, int W0, int H, int W, int j, int i, int b) {
    if (j >= W) return;
    const Tap t = make_tap(__ldg(disp + ((size_t)b * H + i) * W + j), __ldg(row + j), __ldg(col + i), k, H0, W0);
    const size_t sp = (size_t)H0 * W0, op = (size_t)H * W;
    const float* s = src + (size_t)b * C * sp;
    float* gs = gsrc ? gsrc + (size_t)b * C * sp : nullptr;       // NULL: the source needs no gradient (an input image)
    const float* g = gout + (size_t)b * C * op + (size_t)i * W + j;
    const int a00 = t.y0 * W0 + t.x0;
    const bool v00 = t.vy0 && t.vx0, v01 = t.vy0 && t.vx1, v10 = t.vy1 && t.vx0, v11 = t.vy1 && t.vx1;
    const float ey = 1.0f - t.ty;
    float gix = 0.f;
    for (int c = 0; c < C; ++c) {
        const float gv = __ldg(g + c * op);
        const float* p = s + c * sp + a00;
        float* q = gs + c * sp + a00;
        float nwv = 0.f, nev = 0.f, swv = 0.f, sev = 0.f;
        if (v00) { nwv = __ldg(p) + delt;          if (gs) atomicAdd(q, t.nw * gv); }
        if (v01) { nev = __ldg(p + 1) + delt;      if (gs) atomicAdd(q + 1, t.ne * gv); }
        if (v10) { swv = __ldg(p + W0) + delt;     if (gs) atomicAdd(q + W0, t.sw * gv); }
        if (v11) { sev = __ldg(p + W0 + 1) + delt; if (gs) atomicAdd(q + W0 + 1, t.se * gv); }
        gix = fmaf(gv, (nev - nwv) * ey + (sev - swv) * t.ty, gix);
    }
    const float wm1 = (float)(W0 - 1);
    // d(ix)/d(gx) = (W0-1)/2 ; d(gx)/d(disp) = -k*2/(W0-1)
    gdisp[((size_t)b * H + i) * W + j] = -k * (2.0f / wm1) * (gix * (wm1 * 0.5f));
}

__global__ void __launch_bounds__(128)
warp_bwd_kernel(const float* __restrict__ gout, const float* __restrict__ src, const float* __restrict__ disp,
                const float* __restrict__ row, const float* __restrict__ col,
                float delt, float k, float* __restrict__ gsrc, float* __restrict__ gdisp,
                int C, int H0, int W0, int H, int W) {
    warp_bwd_body(gout, src, disp, row, col, delt, k, gsrc, gdisp, C, H0, W0, H, W, blockIdx.x * blockDim.x + threadIdx.x, blockIdx.y, blockIdx.z);
}

// ---- batched form: all the warps of one training step (losses/loss.py:449-452: 4 warps x 7 pyramid levels) in ONE launch.
// The jobs are independent (no warp reads another warp's output); a block finds its job by scanning the block-count prefix.
struct WarpJobs {
    DsmWarpJob job[DSM_WARP_MAX_JOBS];
    int first_block[DSM_WARP_MAX_JOBS + 1];
    int n;
};

template <bool BWD>
__global__ void __launch_bounds__(128)
warp_batched_kernel(const __grid_constant__ WarpJobs J) {
    int q = 0;
    while (q + 1 < J.n && (int)blockIdx.x >= J.first_block[q + 1]) ++q;
    const DsmWarpJob& w = J.job[q];
    int r = (int)blockIdx.x - J.first_block[q];
    const int xb = dsm_ceil_div(w.W, 128);
    const int bx = r % xb; r /= xb;
    const int i = r % w.H, b = r / w.H;
    const int j = bx * 128 + (int)threadIdx.x;
    const float k = w.fliplr ? -1.0f : 1.0f;
    if (BWD) warp_bwd_body(w.gout, w.src, w.disp, w.row, w.col, w.delt, k, w.gsrc, w.gdisp, w.C, w.H0, w.W0, w.H, w.W, j, i, b);
    else     warp_fwd_body(w.src, w.disp, w.row, w.col, w.delt, k, w.out, w.C, w.H0, w.W0, w.H, w.W, j, i, b);
}

int launch_batched(const DsmWarpJob* jobs, int n, bool bwd, void* stream) {
    if (!jobs || n < 1 || n > DSM_WARP_MAX_JOBS) return DSM_EINVAL;
    WarpJobs J;
    long long total = 0;
    for (int q = 0; q < n; ++q) {
        const DsmWarpJob& w = jobs[q];
        if (!w.src || !w.disp || !w.row || !w.col || w.B <= 0 || w.C <= 0) return DSM_EINVAL;
        if (w.H0 < 2 || w.W0 < 2 || w.H < 2 || w.W < 2) return DSM_EINVAL;
        if (bwd ? (!w.gout || !w.gdisp) : !w.out) return DSM_EINVAL;
        J.job[q] = w;
        J.first_block[q] = (int)total;
        total += (long long)dsm_ceil_div(w.W, 128) * w.H * w.B;
        if (total > 0x7fffffffLL) return DSM_EUNSUPPORTED;
    }
    J.first_block[n] = (int)total; J.n = n;
    if (bwd) warp_batched_kernel<true><<<(unsigned)total, 128, 0, (cudaStream_t)stream>>>(J);
    else     warp_batched_kernel<false><<<(unsigned)total, 128, 0, (cudaStream_t)stream>>>(J);
    return dsm_launch_status();
}

__global__ void __launch_bounds__(128)
warp_indices_kernel(const float* __restrict__ disp, const float* __restrict__ row, const float* __restrict__ col,
                    float k, int* __restrict__ x0, int* __restrict__ y0, int H0, int W0, int H, int W) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    const int i = blockIdx.y, b = blockIdx.z;
    if (j >= W) return;
    const size_t o = ((size_t)b * H + i) * W + j;
    const float wm1 = (float)(W0 - 1), hm1 = (float)(H0 - 1);
    const float q  = __fdiv_rn(__fmul_rn(__ldg(disp + o), 2.0f), wm1);
    const float gx = __fmul_rn(k, __fsub_rn(__ldg(row + j), q));
    const float ix = __fmul_rn(__fdiv_rn(__fadd_rn(gx, 1.0f), 2.0f), wm1);
    const float iy = __fmul_rn(__fdiv_rn(__fadd_rn(__ldg(col + i), 1.0f), 2.0f), hm1);
    x0[o] = (int)floorf(ix); y0[o] = (int)floorf(iy);
}

}  // namespace

// test hook for the "bit-exact warp indexing" requirement: the north-west source pixel (x0,y0)
// [B][H][W] int32 that dsm_warp_fwd/bwd use for every output pixel (same arithmetic as make_tap).
extern "C" int dsm_warp_indices(const float* disp, const float* row, const float* col, int fliplr,
                                int* x0, int* y0, int B, int H0, int W0, int H, int W, void* stream) {
    DsmDeviceGuard dsm_guard_(disp);
    if (!disp || !row || !col || !x0 || !y0 || B <= 0) return DSM_EINVAL;
    if (H0 < 2 || W0 < 2 || H < 2 || W < 2) return DSM_EINVAL;
    if (H > 65535 || B > 65535) return DSM_EUNSUPPORTED;
    warp_indices_kernel<<<dim3(dsm_ceil_div(W, 128), H, B), 128, 0, (cudaStream_t)stream>>>(
        disp, row, col, fliplr ? -1.0f : 1.0f, x0, y0, H0, W0, H, W);
    return dsm_launch_status();
}

extern "C" int dsm_warp_fwd(const float* src, const float* disp, const float* row, const float* col,
                            float delt, int fliplr, float* out,
                            int B, int C, int H0, int W0, int H, int W, void* stream) {
    DsmDeviceGuard dsm_guard_(src);
    if (!src || !disp || !row || !col || !out || B <= 0 || C <= 0) return DSM_EINVAL;
    if (H0 < 2 || W0 < 2 || H < 2 || W < 2) return DSM_EINVAL;            // imwrap.py:48
    if (H > 65535 || B > 65535) return DSM_EUNSUPPORTED;
    warp_fwd_kernel<<<dim3(dsm_ceil_div(W, 128), H, B), 128, 0, (cudaStream_t)stream>>>(
        src, disp, row, col, delt, fliplr ? -1.0f : 1.0f, out, C, H0, W0, H, W);
    return dsm_launch_status();
}

extern "C" int dsm_warp_fwd_batched(const DsmWarpJob* jobs, int n, void* stream) {
    DsmDeviceGuard dsm_guard_(jobs && n > 0 ? jobs[0].src : nullptr);
    return launch_batched(jobs, n, false, stream);
}

extern "C" int dsm_warp_bwd_batched(const DsmWarpJob* jobs, int n, void* stream) {
    DsmDeviceGuard dsm_guard_(jobs && n > 0 ? jobs[0].src : nullptr);
    return launch_batched(jobs, n, true, stream);
}

extern "C" int dsm_warp_bwd(const float* gout, const float* src, const float* disp, const float* row,
                            const float* col, float delt, int fliplr, float* gsrc, float* gdisp,
                            int B, int C, int H0, int W0, int H, int W, void* stream) {
    DsmDeviceGuard dsm_guard_(gout);
    if (!gout || !src || !disp || !row || !col || !gsrc || !gdisp || B <= 0 || C <= 0) return DSM_EINVAL;
    if (H0 < 2 || W0 < 2 || H < 2 || W < 2) return DSM_EINVAL;
    if (H > 65535 || B > 65535) return DSM_EUNSUPPORTED;
    cudaStream_t st = (cudaStream_t)stream;
    cudaError_t e = cudaMemsetAsync(gsrc, 0, (size_t)B * C * H0 * W0 * sizeof(float), st);
    if (e != cudaSuccess) return (int)e;
    warp_bwd_kernel<<<dim3(dsm_ceil_div(W, 128), H, B), 128, 0, st>>>(
        gout, src, disp, row, col, delt, fliplr ? -1.0f : 1.0f, gsrc, gdisp, C, H0, W0, H, W);
    return dsm_launch_status();
}
