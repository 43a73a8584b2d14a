// op 3, backward w.r.t. the filter ("wgrad") of the k=3 3-D convolutions — training path.
// Autograd of nn.Conv3d / nn.ConvTranspose3d in reference models/psmnet/submodule.py:16-19,
// models/psmnet/stackhourglass.py:26-41, models/util_conv.py:150-179.
//
// One formulation covers all three layer types.  With "anchor" = the coarser of (x, gy) and
// "partner" = the other one,
//     G[tap][a][b] = sum over anchor voxels v of  A[v][a] * P[s*v + tap - 1][b]        (per axis)
//   conv, stride s:  anchor = gy (channels a = co), partner = x  (b = ci)  -> dW [co][ci][tap]
//   transposed, s=2: anchor = x  (channels a = ci), partner = gy (b = co)  -> dWt[ci][co][tap]
// i.e. in both cases the result is written as dW[a][b][kd][kh][kw], PyTorch's weight layout.
// Tensors are the padded NDHWC bf16 volumes of the forward path (zero rims: out-of-range partner
// voxels contribute nothing, exactly like the zero padding / the missing taps of the reference).
//
// The reduction dimension (voxels) is the ROW index of both operands, so the MMA operands are the
// transposes of the stored tiles: this kernel uses the warp-level path (mma.sync m16n8k16 bf16 with
// ldmatrix.trans, fp32 accumulate in registers).  A tcgen05 version needs voxel-contiguous
// (transposed) copies of both tensors or MN-major descriptors and is left for the next round; wgrad
// is 1/3 of a training step's flops and does not exist on the inference path.
// CTA = 64 anchor voxels of one row; per (kd,kh) one contiguous partner segment of 64*s+2 voxels
// (the three kw taps are row offsets into it, stride s) is staged with cp.async; a warp owns a set
// of 16x8 output tiles and keeps the accumulators of ALL its taps in registers across the
// persistent tile loop; CTA partials go to a workspace and a second kernel reduces them in a fixed
// order (deterministic, no atomics) and applies the optional per-channel scale (folded BatchNorm).
#include "common.cuh"

namespace {

constexpr int TW = 64;        // anchor voxels per tile
constexpr int PADE = 8;       // bf16 elements of row padding in shared memory (16 B): conflict-free ldmatrix

struct WgGeom {
    int Ca, Cb;               // full channel counts of anchor / partner (row strides); a CTA handles a CA x CB slice
    int B, Da, Ha, Wa;        // anchor extent (unpadded)
    int Dp, Hp, Wp;           // partner extent (unpadded)
    int s;                    // partner index = s*anchor + tap - 1
    int tiles_w, ntiles;
};

__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" :: "r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

__device__ __forceinline__ void ldsm_x4_t(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
__device__ __forceinline__ void ldsm_x2_t(uint32_t addr, uint32_t& r0, uint32_t& r1) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x2.trans.shared.b16 {%0,%1}, [%2];" : "=r"(r0), "=r"(r1) : "r"(addr));
}
__device__ __forceinline__ void mma_bf16(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// NKD = number of kd planes this CTA handles (3: all 27 taps; 1: the 9 taps of kd = blockIdx.y)
template <int CA, int CB, int NKD>
__global__ void __launch_bounds__(256)
conv3d_wgrad_kernel(const __nv_bfloat16* __restrict__ anchor, const __nv_bfloat16* __restrict__ partner,
                    float* __restrict__ partial, const __grid_constant__ WgGeom g) {
    constexpr int NT = 9 * NKD;                        // taps of this CTA
    constexpr int NSEG = 3 * NKD;                      // (kd, kh) partner segments
    constexpr int MT = CA / 16, NTL = CB / 8;          // 16x8 output tiles
    constexpr int TPW = (MT * NTL) / 8;                // tiles per warp (8 warps)
    static_assert((MT * NTL) % 8 == 0 && TPW >= 1, "tile split");
    constexpr int PA = CA + PADE, PB = CB + PADE;      // shared row pitches (elements)
    constexpr int SEGMAX = 2 * TW + 2;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __nv_bfloat16* As = reinterpret_cast<__nv_bfloat16*>(smem_raw);                 // [TW][PA]
    __nv_bfloat16* Ps = As + TW * PA;                                               // [NSEG][SEGMAX][PB]
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int kd0 = (NKD == 3) ? 0 : (int)blockIdx.y;
    const int nbs = g.Cb / CB;                         // channel slices: blockIdx.z = a_slice * nbs + b_slice
    const int a0 = ((int)blockIdx.z / nbs) * CA, b0 = ((int)blockIdx.z % nbs) * CB;
    const int s = g.s;
    const int seglen = s * TW + 2;
    const int Hap = g.Ha + 2, Wap = g.Wa + 2, Dpp = g.Dp + 2, Hpp = g.Hp + 2, Wpp = g.Wp + 2;

    float acc[NT][TPW][4];
#pragma unroll
    for (int t = 0; t < NT; ++t)
#pragma unroll
        for (int i = 0; i < TPW; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) acc[t][i][j] = 0.f;

    for (int tile = blockIdx.x; tile < g.ntiles; tile += gridDim.x) {
        int t = tile;
        const int wt = t % g.tiles_w; t /= g.tiles_w;
        const int h = t % g.Ha; t /= g.Ha;
        const int d = t % g.Da; const int b = t / g.Da;
        const int w0 = wt * TW;
        const int nv = min(TW, g.Wa - w0);
        __syncthreads();                                // the previous tile's fragments have been read
        // ---- stage the anchor rows: voxels (b, d, h, w0 .. w0+nv) of the padded volume ----
        {
            const __nv_bfloat16* src = anchor + ((((size_t)b * (g.Da + 2) + d + 1) * Hap + h + 1) * Wap + w0 + 1) * g.Ca + a0;
            constexpr int CH = CA / 8;                  // 16-byte chunks per row
            for (int i = tid; i < TW * CH; i += 256) {
                const int v = i / CH, c = i - v * CH;
                const uint32_t dst = (uint32_t)__cvta_generic_to_shared(As + v * PA + c * 8);
                if (v < nv) cp_async16(dst, src + (size_t)v * g.Ca + c * 8);
                else *reinterpret_cast<uint4*>(As + v * PA + c * 8) = make_uint4(0u, 0u, 0u, 0u);
            }
        }
        // ---- stage the partner segments: padded rows (s*d + kd, s*h + kh), columns s*w0 .. s*w0 + seglen ----
        {
            constexpr int CH = CB / 8;
            for (int i = tid; i < NSEG * seglen * CH; i += 256) {
                const int c = i % CH; int r = i / CH;
                const int j = r % seglen; const int sg = r / seglen;
                const int kd = kd0 + sg / 3, kh = sg % 3;
                const int pd = s * d + kd, ph = s * h + kh, pw = s * w0 + j;          // padded partner coordinates
                __nv_bfloat16* dstp = Ps + ((size_t)sg * SEGMAX + j) * PB + c * 8;
                if (pd < Dpp && ph < Hpp && pw < Wpp)
                    cp_async16((uint32_t)__cvta_generic_to_shared(dstp),
                               partner + ((((size_t)b * Dpp + pd) * Hpp + ph) * Wpp + pw) * g.Cb + b0 + c * 8);
                else *reinterpret_cast<uint4*>(dstp) = make_uint4(0u, 0u, 0u, 0u);
            }
        }
        cp_async_wait_all();
        __syncthreads();

        // ---- G[tap] += A^T (CA x 64) * P_tap (64 x CB), K = 64 anchor voxels in 4 steps of 16 ----
#pragma unroll
        for (int k0 = 0; k0 < TW; k0 += 16) {
            uint32_t af[TPW][4];
            int mt_of[TPW], nt_of[TPW];
#pragma unroll
            for (int i = 0; i < TPW; ++i) {
                const int id = warp * TPW + i;
                mt_of[i] = id / NTL; nt_of[i] = id % NTL;
            }
#pragma unroll
            for (int i = 0; i < TPW; ++i) {
                // A[m][k] = As[k][m]: four stored 8x8 blocks (k 0-7 / 8-15) x (m 0-7 / 8-15), transposed on load
                if (i == 0 || mt_of[i] != mt_of[i - 1]) {
                    const __nv_bfloat16* p = As + (k0 + (lane & 7) + ((lane >> 4) & 1) * 8) * PA + mt_of[i] * 16 + ((lane >> 3) & 1) * 8;
                    ldsm_x4_t((uint32_t)__cvta_generic_to_shared(p), af[i][0], af[i][1], af[i][2], af[i][3]);
                } else {
                    af[i][0] = af[i - 1][0]; af[i][1] = af[i - 1][1]; af[i][2] = af[i - 1][2]; af[i][3] = af[i - 1][3];
                }
            }
#pragma unroll
            for (int sg = 0; sg < NSEG; ++sg) {
#pragma unroll
                for (int kw = 0; kw < 3; ++kw) {
                    const int tap = sg * 3 + kw;
#pragma unroll
                    for (int i = 0; i < TPW; ++i) {
                        // B[k][n] = P[kw + s*k][n]: two stored 8x8 blocks (k 0-7, 8-15), transposed on load
                        const __nv_bfloat16* p = Ps + ((size_t)sg * SEGMAX + kw + s * (k0 + (lane & 15))) * PB + nt_of[i] * 8;
                        uint32_t b0, b1;
                        ldsm_x2_t((uint32_t)__cvta_generic_to_shared(p), b0, b1);
                        mma_bf16(acc[tap][i], af[i], b0, b1);
                    }
                }
            }
        }
    }

    // ---- CTA partial -> workspace [cta][27][CA][CB] (taps of other kd planes are written by other CTAs) ----
    float* out = partial + ((size_t)((blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x)) * (NT * CA * CB);
#pragma unroll
    for (int t = 0; t < NT; ++t)
#pragma unroll
        for (int i = 0; i < TPW; ++i) {
            const int id = warp * TPW + i;
            const int m0 = (id / NTL) * 16 + (lane >> 2), n0 = (id % NTL) * 8 + 2 * (lane & 3);
            float* o = out + ((size_t)t * CA + m0) * CB + n0;
            *reinterpret_cast<float2*>(o) = make_float2(acc[t][i][0], acc[t][i][1]);
            *reinterpret_cast<float2*>(o + 8 * CB) = make_float2(acc[t][i][2], acc[t][i][3]);
        }
}

// dW[a][b][tap] = scale_a[a] * scale_b[b] * sum_cta partial[slice][kd group][cta][tap][a][b]   (fixed summation order)
__global__ void __launch_bounds__(256)
wgrad_reduce_kernel(const float* __restrict__ partial, float* __restrict__ dw, int nkdgroups, int ncta, int Ca, int Cb, int CA, int CB,
                    int CAo, int CBo, const float* __restrict__ scale_a, const float* __restrict__ scale_b, int accumulate) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= 27 * Ca * Cb) return;
    const int b = i % Cb; int r = i / Cb;
    const int a = r % Ca; const int tap = r / Ca;
    const int nt = 27 / nkdgroups;                     // taps per kd group
    const int grp = tap / nt, tl = tap - grp * nt;
    const int z = (a / CA) * (Cb / CB) + (b / CB), al = a % CA, bl = b % CB;
    const size_t per_cta = (size_t)nt * CA * CB;
    const float* p = partial + ((size_t)(z * nkdgroups + grp) * ncta) * per_cta + ((size_t)tl * CA + al) * CB + bl;
    float sum = 0.f;
    for (int c = 0; c < ncta; ++c) sum += p[(size_t)c * per_cta];
    if (a >= CAo || b >= CBo) return;                  // channels that only exist as zero padding
    if (scale_a) sum *= scale_a[a];
    if (scale_b) sum *= scale_b[b];
    float* o = dw + ((size_t)a * CBo + b) * 27 + tap;
    *o = accumulate ? *o + sum : sum;
}

template <int CA, int CB, int NKD>
int launch_wgrad(const void* anchor, const void* partner, float* partial, const WgGeom& g, int ncta, cudaStream_t st) {
    const size_t smem = ((size_t)TW * (CA + PADE) + (size_t)3 * NKD * (2 * TW + 2) * (CB + PADE)) * sizeof(__nv_bfloat16);
    auto kern = conv3d_wgrad_kernel<CA, CB, NKD>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    kern<<<dim3(ncta, NKD == 3 ? 1 : 3, (g.Ca / CA) * (g.Cb / CB)), 256, smem, st>>>(reinterpret_cast<const __nv_bfloat16*>(anchor),
                                                            reinterpret_cast<const __nv_bfloat16*>(partner), partial, g);
    return dsm_launch_status();
}

int wgrad_ncta(long long ntiles) {
    int nsm = DSM_NUM_SMS_B200, dev = 0;
    if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, dev);
    return (int)(ntiles < nsm ? ntiles : nsm);
}

}  // namespace

extern "C" size_t dsm_conv3d_wgrad_workspace_bytes(int Ca, int Cb) {
    return (size_t)DSM_NUM_SMS_B200 * 2 * 27 * (size_t)Ca * Cb * sizeof(float);      // one partial per CTA (<= #SMs), 2x margin
}

extern "C" int dsm_conv3d_wgrad(const void* anchor, const void* partner, float* dw,
                                int B, int Ca, int Cb, int Da, int Ha, int Wa, int Dp, int Hp, int Wp, int stride,
                                int Ca_out, int Cb_out, const float* scale_a, const float* scale_b, int accumulate,
                                void* ws, size_t ws_bytes, void* stream) {
    if (!anchor || !partner || !dw || !ws || B <= 0 || Da <= 0 || Ha <= 0 || Wa <= 0 || Dp <= 0 || Hp <= 0 || Wp <= 0) return DSM_EINVAL;
    if (stride != 1 && stride != 2) return DSM_EINVAL;
    if (Ca_out <= 0 || Ca_out > Ca || Cb_out <= 0 || Cb_out > Cb) return DSM_EINVAL;
    if (!dsm_aligned16(anchor) || !dsm_aligned16(partner) || !dsm_aligned16(ws)) return DSM_EALIGN;
    if (ws_bytes < dsm_conv3d_wgrad_workspace_bytes(Ca, Cb)) return DSM_EINVAL;
    if ((Ca != 32 && Ca != 64 && Ca != 128) || (Cb != 32 && Cb != 64 && Cb != 128)) return DSM_EUNSUPPORTED;
    WgGeom g;
    g.Ca = Ca; g.Cb = Cb;
    g.B = B; g.Da = Da; g.Ha = Ha; g.Wa = Wa; g.Dp = Dp; g.Hp = Hp; g.Wp = Wp; g.s = stride;
    g.tiles_w = dsm_ceil_div(Wa, TW);
    const long long nt = (long long)B * Da * Ha * g.tiles_w;
    if (nt > 0x7fffffffLL) return DSM_EUNSUPPORTED;
    g.ntiles = (int)nt;
    const int ncta = wgrad_ncta(nt);
    cudaStream_t st = (cudaStream_t)stream;
    float* partial = reinterpret_cast<float*>(ws);
    const int CA = Ca > 64 ? 64 : Ca, CB = Cb > 64 ? 64 : Cb;      // slice widths; 128-channel tensors run as 2 slices
    int rc, groups;
    if (CA == 32 && CB == 32)      { rc = launch_wgrad<32, 32, 3>(anchor, partner, partial, g, ncta, st); groups = 1; }
    else if (CA == 64 && CB == 32) { rc = launch_wgrad<64, 32, 1>(anchor, partner, partial, g, ncta, st); groups = 3; }
    else if (CA == 32 && CB == 64) { rc = launch_wgrad<32, 64, 1>(anchor, partner, partial, g, ncta, st); groups = 3; }
    else                           { rc = launch_wgrad<64, 64, 1>(anchor, partner, partial, g, ncta, st); groups = 3; }
    if (rc != 0) return rc;
    const int per = 27 * Ca * Cb;
    wgrad_reduce_kernel<<<dsm_ceil_div(per, 256), 256, 0, st>>>(partial, dw, groups, ncta, Ca, Cb, CA, CB, Ca_out, Cb_out, scale_a, scale_b, accumulate);
    return dsm_launch_status();
}
