// Training-mode BatchNorm3d (batch statistics) + ReLU + skip add on padded NDHWC bf16 volumes, forward and backward.
// In the reference these are stock modules between the 3-D convolutions (models/psmnet/submodule.py:16-19,
// models/psmnet/stackhourglass.py:43-62, models/util_conv.py:160-178 under model.train()); here they are four
// HBM-streaming kernels so that a training step of the 3-D stack touches every activation a minimal number of times:
//
//   forward : y (raw conv output) --stats--> sum, sumsq per channel --finalize--> scale, shift (+ running stats)
//             z = act(y*scale + shift [+ res])                                              (dsm_bn_act_fwd)
//   backward: g = gz * mask;  sum g, sum g*y per channel --finalize--> dgamma, dbeta, a, b, c
//             dy = a*g + b*y + c (interior voxels; the rim stays zero), gres = g            (dsm_bn_act_bwd)
//
// relu: 0 none, 1 after the skip add (PSMNet; mask = z > 0), 2 before it (GC-Net; mask = y*scale+shift > 0).
// All volumes are [B][D+2][H+2][W+2][C] bf16 with a zero rim; the rim contributes nothing to the sums, so the
// reductions stream the whole array.  Algorithmic bytes per voxel-channel: stats 2, fwd 4 (+2 with a skip),
// bwd reduce 4 (+2), bwd apply 6 (+2, +2).
#include "common.cuh"

namespace {

constexpr int BN_THREADS = 256;

__device__ __forceinline__ void unpack16(const uint4& a, const uint4& b, float* v) {
    const uint32_t w[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
#pragma unroll
    for (int i = 0; i < 8; ++i) { v[2 * i] = bf16_lo(w[i]); v[2 * i + 1] = bf16_hi(w[i]); }
}
__device__ __forceinline__ void pack16(const float* v, uint4& a, uint4& b) {
    a = make_uint4(pack_bf16x2(v[0], v[1]), pack_bf16x2(v[2], v[3]), pack_bf16x2(v[4], v[5]), pack_bf16x2(v[6], v[7]));
    b = make_uint4(pack_bf16x2(v[8], v[9]), pack_bf16x2(v[10], v[11]), pack_bf16x2(v[12], v[13]), pack_bf16x2(v[14], v[15]));
}

// Per-channel reduction of two quantities over the whole padded array.  A thread walks 32-byte vectors (16
// channels) with a stride that is a multiple of C/16, so its channel group never changes and the partial sums
// stay in registers; lanes of equal group are folded with shuffles, warps through shared memory, CTAs through
// double atomics (a few hundred per address per launch).
//   MODE 0: (y, y*y)                    forward statistics
//   MODE 1: (g, g*y), g = gz            relu 0
//   MODE 2: (g, g*y), g = gz * (z > 0)  relu 1
//   MODE 3: (g, g*y), g = gz * (y*scale+shift > 0)   relu 2
template <int MODE>
__global__ void __launch_bounds__(BN_THREADS)
bn_reduce_kernel(const uint4* __restrict__ y, const uint4* __restrict__ gz, const uint4* __restrict__ z,
                 const float* __restrict__ scale, const float* __restrict__ shift,
                 long long nvec, int C, double* __restrict__ sums) {
    __shared__ float s_acc[2][128];
    const int groups = C >> 4;                                   // 16-channel groups per voxel: 2, 4 or 8
    const int grp = threadIdx.x & (groups - 1);
    for (int i = threadIdx.x; i < 2 * 128; i += BN_THREADS) (&s_acc[0][0])[i] = 0.f;
    float s0[16], s1[16], sc[16], sh[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) { s0[i] = 0.f; s1[i] = 0.f; }
    if (MODE == 3) {
#pragma unroll
        for (int i = 0; i < 16; ++i) { sc[i] = scale[grp * 16 + i]; sh[i] = shift[grp * 16 + i]; }
    }
    const long long stride = (long long)gridDim.x * BN_THREADS;
    constexpr int U = (MODE == 0) ? 4 : 2;                       // independent 32-byte loads in flight per stream and thread
    for (long long v0 = (long long)blockIdx.x * BN_THREADS + threadIdx.x; v0 < nvec; v0 += stride * U) {
        uint4 ya[U], yb[U], ga[U], gb[U], za[U], zb[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const long long v = v0 + u * stride;
            if (v < nvec) {
                ld_nc_v8(y + 2 * v, ya[u], yb[u]);
                if (MODE != 0) ld_nc_v8(gz + 2 * v, ga[u], gb[u]);
                if (MODE == 2) ld_nc_v8(z + 2 * v, za[u], zb[u]);
            }
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            if (v0 + u * stride >= nvec) break;
            float fy[16];
            unpack16(ya[u], yb[u], fy);
            if (MODE == 0) {
#pragma unroll
                for (int i = 0; i < 16; ++i) { s0[i] += fy[i]; s1[i] = fmaf(fy[i], fy[i], s1[i]); }
            } else {
                float fg[16];
                unpack16(ga[u], gb[u], fg);
                if (MODE == 2) {
                    float fz[16];
                    unpack16(za[u], zb[u], fz);
#pragma unroll
                    for (int i = 0; i < 16; ++i) fg[i] = fz[i] > 0.f ? fg[i] : 0.f;
                } else if (MODE == 3) {
#pragma unroll
                    for (int i = 0; i < 16; ++i) fg[i] = fmaf(fy[i], sc[i], sh[i]) > 0.f ? fg[i] : 0.f;
                }
#pragma unroll
                for (int i = 0; i < 16; ++i) { s0[i] += fg[i]; s1[i] = fmaf(fg[i], fy[i], s1[i]); }
            }
        }
    }
    // lanes l and l ^ off share a channel group when off >= groups
    for (int off = 16; off >= groups; off >>= 1) {
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            s0[i] += __shfl_xor_sync(0xffffffffu, s0[i], off);
            s1[i] += __shfl_xor_sync(0xffffffffu, s1[i], off);
        }
    }
    __syncthreads();
    if ((threadIdx.x & 31) < groups) {
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            atomicAdd(&s_acc[0][grp * 16 + i], s0[i]);
            atomicAdd(&s_acc[1][grp * 16 + i], s1[i]);
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 2 * C; i += BN_THREADS) {
        const int which = i / C, c = i - which * C;
        atomicAdd(&sums[which * C + c], (double)s_acc[which][c]);
    }
}

// sums -> per-channel affine of the forward pass, and the running-statistics update of nn.BatchNorm3d
// (torch: running = (1-m)*running + m*batch, unbiased variance for the running estimate).
__global__ void bn_finalize_fwd_kernel(const double* __restrict__ sums, const float* __restrict__ gamma,
                                       const float* __restrict__ beta, const float* __restrict__ conv_bias,
                                       double count, float eps, float momentum,
                                       float* running_mean, float* running_var,
                                       float* __restrict__ scale, float* __restrict__ shift,
                                       float* __restrict__ mean_out, float* __restrict__ rstd_out, int C) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    const double mean = sums[c] / count;
    double var = sums[C + c] / count - mean * mean;
    if (var < 0.0) var = 0.0;
    const float rstd = (float)(1.0 / sqrt(var + (double)eps));
    const float g = gamma ? gamma[c] : 1.f, b = beta ? beta[c] : 0.f;
    const float sc = g * rstd;
    scale[c] = sc;
    shift[c] = b - (float)mean * sc;                       // a conv bias cancels against the batch mean
    mean_out[c] = (float)mean;
    rstd_out[c] = rstd;
    if (running_mean) {
        const float m = (float)mean + (conv_bias ? conv_bias[c] : 0.f);
        running_mean[c] = (1.f - momentum) * running_mean[c] + momentum * m;
    }
    if (running_var) {
        const double unbiased = count > 1.0 ? var * count / (count - 1.0) : var;
        running_var[c] = (1.f - momentum) * running_var[c] + momentum * (float)unbiased;
    }
}

// sums (sum g, sum g*y) -> dgamma, dbeta and the coefficients of dy = a*g + b*y + c
__global__ void bn_finalize_bwd_kernel(const double* __restrict__ sums, const float* __restrict__ gamma,
                                       const float* __restrict__ mean, const float* __restrict__ rstd,
                                       double count, float* __restrict__ dgamma, float* __restrict__ dbeta,
                                       float* __restrict__ coef, int C) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    const double sg = sums[c], sgy = sums[C + c];
    const double mu = mean[c], rs = rstd[c];
    const double dgh = rs * (sgy - mu * sg);               // sum g * xhat
    const double g = gamma ? gamma[c] : 1.f;
    const double sc = g * rs;
    dgamma[c] = (float)dgh;
    dbeta[c] = (float)sg;
    coef[c] = (float)sc;                                                   // a
    coef[C + c] = (float)(-sc * rs * dgh / count);                         // b
    coef[2 * C + c] = (float)(-sc * sg / count + sc * rs * mu * dgh / count);   // c
}

struct BnGeom { int B, C, D, H, W; };

// One 32-byte vector per thread and iteration; blockIdx.y = padded plane (b, d'), so the rim test costs one
// division per vector.  RELU as above; HAS_RES adds the skip tensor (same geometry).
template <int RELU, bool HAS_RES>
__global__ void __launch_bounds__(BN_THREADS)
bn_act_fwd_kernel(const uint4* __restrict__ y, const uint4* __restrict__ res, uint4* __restrict__ z,
                  const float* __restrict__ scale, const float* __restrict__ shift, BnGeom g) {
    const int groups = g.C >> 4;
    const int grp = threadIdx.x & (groups - 1);
    const int rowvec = (g.W + 2) * groups;
    const int planevec = (g.H + 2) * rowvec;
    const int dp = blockIdx.y % (g.D + 2);
    const bool rim_plane = dp == 0 || dp == g.D + 1;
    float sc[16], sh[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) { sc[i] = scale[grp * 16 + i]; sh[i] = shift[grp * 16 + i]; }
    const size_t base = (size_t)blockIdx.y * planevec;
    constexpr int U = 2;                                         // vectors per thread and iteration, loads issued first
    const int step = gridDim.x * BN_THREADS;
    for (int v0 = blockIdx.x * BN_THREADS + threadIdx.x; v0 < planevec; v0 += step * U) {
        uint4 ya[U], yb[U], ra[U], rb[U];
        bool rim[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int v = v0 + u * step;
            const int hp = v / rowvec;
            const int wp = (v - hp * rowvec) / groups;
            rim[u] = rim_plane || hp == 0 || hp >= g.H + 1 || wp == 0 || wp == g.W + 1;
            if (!rim[u] && v < planevec) {
                ld_nc_v8(y + 2 * (base + v), ya[u], yb[u]);
                if (HAS_RES) ld_nc_v8(res + 2 * (base + v), ra[u], rb[u]);
            }
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int v = v0 + u * step;
            if (v >= planevec) break;
            uint4 a = make_uint4(0, 0, 0, 0), b = a;
            if (!rim[u]) {
                float f[16];
                unpack16(ya[u], yb[u], f);
#pragma unroll
                for (int i = 0; i < 16; ++i) f[i] = fmaf(f[i], sc[i], sh[i]);
                if (RELU == 2) {
#pragma unroll
                    for (int i = 0; i < 16; ++i) f[i] = fmaxf(f[i], 0.f);
                }
                if (HAS_RES) {
                    float r[16];
                    unpack16(ra[u], rb[u], r);
#pragma unroll
                    for (int i = 0; i < 16; ++i) f[i] += r[i];
                }
                if (RELU == 1) {
#pragma unroll
                    for (int i = 0; i < 16; ++i) f[i] = fmaxf(f[i], 0.f);
                }
                pack16(f, a, b);
            }
            st_v8(z + 2 * (base + v), a, b);
        }
    }
}

// dy = a*g + b*y + c on interior voxels, g = gz * mask; optionally gres = g (relu 1 with a skip tensor).
template <int RELU, bool WRITE_GRES>
__global__ void __launch_bounds__(BN_THREADS)
bn_act_bwd_kernel(const uint4* __restrict__ gz, const uint4* __restrict__ y, const uint4* __restrict__ z,
                  const float* __restrict__ scale, const float* __restrict__ shift, const float* __restrict__ coef,
                  uint4* __restrict__ dy, uint4* __restrict__ gres, BnGeom g) {
    const int groups = g.C >> 4;
    const int grp = threadIdx.x & (groups - 1);
    const int rowvec = (g.W + 2) * groups;
    const int planevec = (g.H + 2) * rowvec;
    const int dp = blockIdx.y % (g.D + 2);
    const bool rim_plane = dp == 0 || dp == g.D + 1;
    float ca[16], cb[16], cc[16], sc[16], sh[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) {
        ca[i] = coef[grp * 16 + i]; cb[i] = coef[g.C + grp * 16 + i]; cc[i] = coef[2 * g.C + grp * 16 + i];
        if (RELU == 2) { sc[i] = scale[grp * 16 + i]; sh[i] = shift[grp * 16 + i]; }
    }
    const size_t base = (size_t)blockIdx.y * planevec;
    constexpr int U = 2;
    const int step = gridDim.x * BN_THREADS;
    for (int v0 = blockIdx.x * BN_THREADS + threadIdx.x; v0 < planevec; v0 += step * U) {
        uint4 qa[U], qb[U], ya[U], yb[U], za[U], zb[U];
        bool rim[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int v = v0 + u * step;
            const int hp = v / rowvec;
            const int wp = (v - hp * rowvec) / groups;
            rim[u] = rim_plane || hp == 0 || hp >= g.H + 1 || wp == 0 || wp == g.W + 1;
            if (!rim[u] && v < planevec) {
                ld_nc_v8(gz + 2 * (base + v), qa[u], qb[u]);
                ld_nc_v8(y + 2 * (base + v), ya[u], yb[u]);
                if (RELU == 1) ld_nc_v8(z + 2 * (base + v), za[u], zb[u]);
            }
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int v = v0 + u * step;
            if (v >= planevec) break;
            uint4 a = make_uint4(0, 0, 0, 0), b = a, ga = a, gb = a;
            if (!rim[u]) {
                float fg[16], fy[16];
                unpack16(qa[u], qb[u], fg);
                unpack16(ya[u], yb[u], fy);
                if (RELU == 1) {
                    float fz[16];
                    unpack16(za[u], zb[u], fz);
#pragma unroll
                    for (int i = 0; i < 16; ++i) fg[i] = fz[i] > 0.f ? fg[i] : 0.f;
                } else if (RELU == 2) {
#pragma unroll
                    for (int i = 0; i < 16; ++i) fg[i] = fmaf(fy[i], sc[i], sh[i]) > 0.f ? fg[i] : 0.f;
                }
                if (WRITE_GRES) pack16(fg, ga, gb);
#pragma unroll
                for (int i = 0; i < 16; ++i) fy[i] = fmaf(ca[i], fg[i], fmaf(cb[i], fy[i], cc[i]));
                pack16(fy, a, b);
            }
            st_v8(dy + 2 * (base + v), a, b);
            if (WRITE_GRES) st_v8(gres + 2 * (base + v), ga, gb);
        }
    }
}

// Zero rim of a padded volume whose interior is (or will be) written by a convolution kernel: rim planes entirely,
// rim rows of the other planes, and the first / last voxel of every interior row.  ~6 % of the bytes of a memset.
__global__ void __launch_bounds__(BN_THREADS)
zero_rim_kernel(uint4* __restrict__ data, int C8, int D, int H, int W) {
    const int dp = blockIdx.y % (D + 2);
    const int rowvec = (W + 2) * C8;                               // 16-byte vectors per padded row
    uint4* plane = data + (size_t)blockIdx.y * (H + 2) * rowvec;
    const uint4 zero = make_uint4(0, 0, 0, 0);
    if (dp == 0 || dp == D + 1) {
        const int n = (H + 2) * rowvec;
        for (int v = blockIdx.x * BN_THREADS + threadIdx.x; v < n; v += gridDim.x * BN_THREADS) plane[v] = zero;
        return;
    }
    for (int hp = blockIdx.x; hp < H + 2; hp += gridDim.x) {
        uint4* row = plane + (size_t)hp * rowvec;
        if (hp == 0 || hp == H + 1) {
            for (int v = threadIdx.x; v < rowvec; v += BN_THREADS) row[v] = zero;
        } else if (threadIdx.x < 2 * C8) {
            const int side = threadIdx.x / C8, k = threadIdx.x - side * C8;
            row[(side ? (W + 1) * C8 : 0) + k] = zero;
        }
    }
}


// Cropped skip add (the reference's myadd_3d / myAdd3d crop-to-min, stackhourglass.py:10-20, util_fun.py:41-51) for the
// training path with odd sizes: `full` is the BatchNorm'ed deconv output at its NATURAL extent (Dn,Hn,Wn) — batch
// statistics are taken over all of it, as the reference does — and the skip tensor / result have the smaller extent
// (Do,Ho,Wo).  One 16-byte chunk (8 channels) per thread.
//   fwd: z = act(crop(full) + res)           relu 0 none, 1 after the add
//   bwd: g = gz * (z > 0 | 1); gres = g (interior; its rim is zeroed by the caller); gfull = g inside the crop, 0 elsewhere
struct CropGeom { int B, C8, Dn, Hn, Wn, Do, Ho, Wo; };

__device__ __forceinline__ uint4 relu_bf16x8_mask(const uint4& g, const uint4& z) {
    uint4 r;
    const uint32_t gw[4] = {g.x, g.y, g.z, g.w}, zw[4] = {z.x, z.y, z.z, z.w};
    uint32_t o[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const uint32_t lo = bf16_lo(zw[i]) > 0.f ? (gw[i] & 0xffffu) : 0u;
        const uint32_t hi = bf16_hi(zw[i]) > 0.f ? (gw[i] & 0xffff0000u) : 0u;
        o[i] = lo | hi;
    }
    r.x = o[0]; r.y = o[1]; r.z = o[2]; r.w = o[3];
    return r;
}

template <int RELU>
__global__ void __launch_bounds__(BN_THREADS)
crop_add_fwd_kernel(const uint4* __restrict__ full, const uint4* __restrict__ res, uint4* __restrict__ z, CropGeom g) {
    const long long n = (long long)g.B * (g.Do + 2) * (g.Ho + 2) * (g.Wo + 2) * g.C8;
    for (long long i = (long long)blockIdx.x * BN_THREADS + threadIdx.x; i < n; i += (long long)gridDim.x * BN_THREADS) {
        const int k = (int)(i % g.C8); long long t = i / g.C8;
        const int wp = (int)(t % (g.Wo + 2)); t /= (g.Wo + 2);
        const int hp = (int)(t % (g.Ho + 2)); t /= (g.Ho + 2);
        const int dp = (int)(t % (g.Do + 2)); const int b = (int)(t / (g.Do + 2));
        uint4 o = make_uint4(0u, 0u, 0u, 0u);
        if (dp >= 1 && dp <= g.Do && hp >= 1 && hp <= g.Ho && wp >= 1 && wp <= g.Wo) {
            const size_t fi = ((((size_t)b * (g.Dn + 2) + dp) * (g.Hn + 2) + hp) * (g.Wn + 2) + wp) * g.C8 + k;
            const uint4 a = __ldg(full + fi), r = __ldg(res + i);
            const uint32_t aw[4] = {a.x, a.y, a.z, a.w}, rw[4] = {r.x, r.y, r.z, r.w};
            uint32_t ow[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                float lo = bf16_lo(aw[j]) + bf16_lo(rw[j]), hi = bf16_hi(aw[j]) + bf16_hi(rw[j]);
                if (RELU == 1) { lo = fmaxf(lo, 0.f); hi = fmaxf(hi, 0.f); }
                ow[j] = pack_bf16x2(lo, hi);
            }
            o = make_uint4(ow[0], ow[1], ow[2], ow[3]);
        }
        z[i] = o;
    }
}

template <int RELU>
__global__ void __launch_bounds__(BN_THREADS)
crop_add_bwd_kernel(const uint4* __restrict__ gz, const uint4* __restrict__ z, uint4* __restrict__ gfull,
                    uint4* __restrict__ gres, CropGeom g) {
    const long long n = (long long)g.B * (g.Dn + 2) * (g.Hn + 2) * (g.Wn + 2) * g.C8;
    for (long long i = (long long)blockIdx.x * BN_THREADS + threadIdx.x; i < n; i += (long long)gridDim.x * BN_THREADS) {
        const int k = (int)(i % g.C8); long long t = i / g.C8;
        const int wp = (int)(t % (g.Wn + 2)); t /= (g.Wn + 2);
        const int hp = (int)(t % (g.Hn + 2)); t /= (g.Hn + 2);
        const int dp = (int)(t % (g.Dn + 2)); const int b = (int)(t / (g.Dn + 2));
        uint4 o = make_uint4(0u, 0u, 0u, 0u);
        if (dp >= 1 && dp <= g.Do && hp >= 1 && hp <= g.Ho && wp >= 1 && wp <= g.Wo) {
            const size_t ci = ((((size_t)b * (g.Do + 2) + dp) * (g.Ho + 2) + hp) * (g.Wo + 2) + wp) * g.C8 + k;
            o = __ldg(gz + ci);
            if (RELU == 1) o = relu_bf16x8_mask(o, __ldg(z + ci));
            if (gres) gres[ci] = o;
        }
        gfull[i] = o;
    }
}

int check_geom(int B, int C, int D, int H, int W) {
    if (B < 1 || D < 1 || H < 1 || W < 1) return DSM_EINVAL;
    if (C != 32 && C != 64 && C != 128) return DSM_EUNSUPPORTED;
    return 0;
}

long long padded_vec32(int B, int C, int D, int H, int W) {
    return (long long)B * (D + 2) * (H + 2) * (W + 2) * C / 16;
}

int reduce_grid(long long nvec) {
    long long want = dsm_ceil_div_ll(nvec, BN_THREADS * 4);
    const long long cap = DSM_NUM_SMS_B200 * 8;
    return (int)(want < 1 ? 1 : (want > cap ? cap : want));
}

dim3 apply_grid(int B, int C, int D, int H, int W) {
    const int planevec = (H + 2) * (W + 2) * (C / 16);
    int gx = dsm_ceil_div(planevec, BN_THREADS * 2);
    if (gx < 1) gx = 1;
    return dim3(gx, B * (D + 2), 1);
}

}  // namespace

extern "C" int dsm_zero_rim(void* data, int B, int C, int D, int H, int W, void* stream_) {
    DsmDeviceGuard dsm_guard_(data);
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    if (!data || B < 1 || D < 1 || H < 1 || W < 1 || C < 8 || (C & 7) || C > 1024) return DSM_EINVAL;
    if (!dsm_aligned16(data)) return DSM_EALIGN;
    int gx = (H + 2) < 32 ? (H + 2) : 32;
    zero_rim_kernel<<<dim3(gx, B * (D + 2)), BN_THREADS, 0, stream>>>(static_cast<uint4*>(data), C / 8, D, H, W);
    return dsm_launch_status();
}

extern "C" int dsm_bn_stats(const void* y, int B, int C, int D, int H, int W, double* sums, void* stream_) {
    DsmDeviceGuard dsm_guard_(y);
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    if (int e = check_geom(B, C, D, H, W)) return e;
    if (!y || !sums) return DSM_EINVAL;
    if (!dsm_aligned32(y)) return DSM_EALIGN;
    cudaError_t ce = cudaMemsetAsync(sums, 0, sizeof(double) * 2 * C, stream);
    if (ce != cudaSuccess) return (int)ce;
    const long long nvec = padded_vec32(B, C, D, H, W);
    bn_reduce_kernel<0><<<reduce_grid(nvec), BN_THREADS, 0, stream>>>(
        static_cast<const uint4*>(y), nullptr, nullptr, nullptr, nullptr, nvec, C, sums);
    return dsm_launch_status();
}

extern "C" int dsm_bn_finalize_fwd(const double* sums, const float* gamma, const float* beta, const float* conv_bias,
                                   int C, long long count, float eps, float momentum,
                                   float* running_mean, float* running_var,
                                   float* scale, float* shift, float* mean, float* rstd, void* stream_) {
    DsmDeviceGuard dsm_guard_(sums);
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    if (!sums || !scale || !shift || !mean || !rstd || C < 1 || count < 1) return DSM_EINVAL;
    bn_finalize_fwd_kernel<<<dsm_ceil_div(C, 128), 128, 0, stream>>>(sums, gamma, beta, conv_bias, (double)count, eps, momentum,
                                                                     running_mean, running_var, scale, shift, mean, rstd, C);
    return dsm_launch_status();
}

extern "C" int dsm_bn_act_fwd(const void* y, const float* scale, const float* shift, const void* residual, int relu,
                              void* z, int B, int C, int D, int H, int W, void* stream_) {
    DsmDeviceGuard dsm_guard_(y);
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    if (int e = check_geom(B, C, D, H, W)) return e;
    if (!y || !z || !scale || !shift || relu < 0 || relu > 2) return DSM_EINVAL;
    if (!dsm_aligned32(y) || !dsm_aligned32(z) || (residual && !dsm_aligned32(residual))) return DSM_EALIGN;
    const BnGeom g{B, C, D, H, W};
    const dim3 grid = apply_grid(B, C, D, H, W);
    const uint4 *py = static_cast<const uint4*>(y), *pr = static_cast<const uint4*>(residual);
    uint4* pz = static_cast<uint4*>(z);
#define DSM_BN_FWD(R, HR) bn_act_fwd_kernel<R, HR><<<grid, BN_THREADS, 0, stream>>>(py, pr, pz, scale, shift, g)
    if (residual) { if (relu == 0) DSM_BN_FWD(0, true); else if (relu == 1) DSM_BN_FWD(1, true); else DSM_BN_FWD(2, true); }
    else          { if (relu == 0) DSM_BN_FWD(0, false); else if (relu == 1) DSM_BN_FWD(1, false); else DSM_BN_FWD(2, false); }
#undef DSM_BN_FWD
    return dsm_launch_status();
}

extern "C" int dsm_bn_act_bwd_reduce(const void* gz, const void* y, const void* z, const float* scale, const float* shift,
                                     int relu, double* sums, int B, int C, int D, int H, int W, void* stream_) {
    DsmDeviceGuard dsm_guard_(gz);
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    if (int e = check_geom(B, C, D, H, W)) return e;
    if (!gz || !y || !sums || relu < 0 || relu > 2) return DSM_EINVAL;
    if (relu == 1 && !z) return DSM_EINVAL;
    if (relu == 2 && (!scale || !shift)) return DSM_EINVAL;
    if (!dsm_aligned32(gz) || !dsm_aligned32(y) || (z && !dsm_aligned32(z))) return DSM_EALIGN;
    cudaError_t ce = cudaMemsetAsync(sums, 0, sizeof(double) * 2 * C, stream);
    if (ce != cudaSuccess) return (int)ce;
    const long long nvec = padded_vec32(B, C, D, H, W);
    const int grid = reduce_grid(nvec);
    const uint4 *pg = static_cast<const uint4*>(gz), *py = static_cast<const uint4*>(y), *pz = static_cast<const uint4*>(z);
    if (relu == 0) bn_reduce_kernel<1><<<grid, BN_THREADS, 0, stream>>>(py, pg, pz, scale, shift, nvec, C, sums);
    else if (relu == 1) bn_reduce_kernel<2><<<grid, BN_THREADS, 0, stream>>>(py, pg, pz, scale, shift, nvec, C, sums);
    else bn_reduce_kernel<3><<<grid, BN_THREADS, 0, stream>>>(py, pg, pz, scale, shift, nvec, C, sums);
    return dsm_launch_status();
}

extern "C" int dsm_bn_finalize_bwd(const double* sums, const float* gamma, const float* mean, const float* rstd,
                                   int C, long long count, float* dgamma, float* dbeta, float* coef, void* stream_) {
    DsmDeviceGuard dsm_guard_(sums);
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    if (!sums || !mean || !rstd || !dgamma || !dbeta || !coef || C < 1 || count < 1) return DSM_EINVAL;
    bn_finalize_bwd_kernel<<<dsm_ceil_div(C, 128), 128, 0, stream>>>(sums, gamma, mean, rstd, (double)count, dgamma, dbeta, coef, C);
    return dsm_launch_status();
}

extern "C" int dsm_bn_act_bwd(const void* gz, const void* y, const void* z, const float* scale, const float* shift,
                              const float* coef, int relu, void* dy, void* gres,
                              int B, int C, int D, int H, int W, void* stream_) {
    DsmDeviceGuard dsm_guard_(gz);
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    if (int e = check_geom(B, C, D, H, W)) return e;
    if (!gz || !y || !dy || !coef || relu < 0 || relu > 2) return DSM_EINVAL;
    if (relu == 1 && !z) return DSM_EINVAL;
    if (relu == 2 && (!scale || !shift)) return DSM_EINVAL;
    if (!dsm_aligned32(gz) || !dsm_aligned32(y) || !dsm_aligned32(dy) || (z && !dsm_aligned32(z)) || (gres && !dsm_aligned32(gres)))
        return DSM_EALIGN;
    const BnGeom g{B, C, D, H, W};
    const dim3 grid = apply_grid(B, C, D, H, W);
    const uint4 *pg = static_cast<const uint4*>(gz), *py = static_cast<const uint4*>(y), *pz = static_cast<const uint4*>(z);
    uint4 *pd = static_cast<uint4*>(dy), *pr = static_cast<uint4*>(gres);
#define DSM_BN_BWD(R, G) bn_act_bwd_kernel<R, G><<<grid, BN_THREADS, 0, stream>>>(pg, py, pz, scale, shift, coef, pd, pr, g)
    if (gres) { if (relu == 0) DSM_BN_BWD(0, true); else if (relu == 1) DSM_BN_BWD(1, true); else DSM_BN_BWD(2, true); }
    else      { if (relu == 0) DSM_BN_BWD(0, false); else if (relu == 1) DSM_BN_BWD(1, false); else DSM_BN_BWD(2, false); }
#undef DSM_BN_BWD
    return dsm_launch_status();
}

extern "C" int dsm_crop_add_fwd(const void* full, const void* residual, void* z, int B, int C,
                                int Dn, int Hn, int Wn, int Do, int Ho, int Wo, int relu, void* stream_) {
    DsmDeviceGuard dsm_guard_(full);
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    if (!full || !residual || !z || B < 1 || C < 8 || (C & 7) || relu < 0 || relu > 1) return DSM_EINVAL;
    if (Do < 1 || Ho < 1 || Wo < 1 || Do > Dn || Ho > Hn || Wo > Wn) return DSM_EINVAL;
    if (!dsm_aligned16(full) || !dsm_aligned16(residual) || !dsm_aligned16(z)) return DSM_EALIGN;
    const CropGeom g{B, C / 8, Dn, Hn, Wn, Do, Ho, Wo};
    const long long n = (long long)B * (Do + 2) * (Ho + 2) * (Wo + 2) * (C / 8);
    long long blocks = dsm_ceil_div_ll(n, BN_THREADS);
    if (blocks > DSM_NUM_SMS_B200 * 16) blocks = DSM_NUM_SMS_B200 * 16;
    const uint4 *pf = static_cast<const uint4*>(full), *pr = static_cast<const uint4*>(residual);
    if (relu) crop_add_fwd_kernel<1><<<(unsigned)blocks, BN_THREADS, 0, stream>>>(pf, pr, static_cast<uint4*>(z), g);
    else      crop_add_fwd_kernel<0><<<(unsigned)blocks, BN_THREADS, 0, stream>>>(pf, pr, static_cast<uint4*>(z), g);
    return dsm_launch_status();
}

extern "C" int dsm_crop_add_bwd(const void* gz, const void* z, void* gfull, void* gres, int B, int C,
                                int Dn, int Hn, int Wn, int Do, int Ho, int Wo, int relu, void* stream_) {
    DsmDeviceGuard dsm_guard_(gz);
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    if (!gz || !gfull || B < 1 || C < 8 || (C & 7) || relu < 0 || relu > 1 || (relu == 1 && !z)) return DSM_EINVAL;
    if (Do < 1 || Ho < 1 || Wo < 1 || Do > Dn || Ho > Hn || Wo > Wn) return DSM_EINVAL;
    if (!dsm_aligned16(gz) || !dsm_aligned16(gfull) || (z && !dsm_aligned16(z)) || (gres && !dsm_aligned16(gres))) return DSM_EALIGN;
    const CropGeom g{B, C / 8, Dn, Hn, Wn, Do, Ho, Wo};
    const long long n = (long long)B * (Dn + 2) * (Hn + 2) * (Wn + 2) * (C / 8);
    long long blocks = dsm_ceil_div_ll(n, BN_THREADS);
    if (blocks > DSM_NUM_SMS_B200 * 16) blocks = DSM_NUM_SMS_B200 * 16;
    const uint4 *pg = static_cast<const uint4*>(gz), *pz = static_cast<const uint4*>(z);
    if (relu) crop_add_bwd_kernel<1><<<(unsigned)blocks, BN_THREADS, 0, stream>>>(pg, pz, static_cast<uint4*>(gfull), static_cast<uint4*>(gres), g);
    else      crop_add_bwd_kernel<0><<<(unsigned)blocks, BN_THREADS, 0, stream>>>(pg, pz, static_cast<uint4*>(gfull), static_cast<uint4*>(gres), g);
    return dsm_launch_status();
}
