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
// Two kernels:
//  * stride 1 (every Conv3d of the stacks except the down-sampling ones): `wgrad_tc_kernel`, tcgen05 with
//    MN-major operand descriptors — the reduction index (voxels) is the ROW index of the NDHWC tiles, which is
//    exactly what an MN-major UMMA operand is, so TMA tiles of both tensors feed the MMA untransposed;
//  * stride 2 / transposed: `conv3d_wgrad_kernel`, the warp-level path (mma.sync m16n8k16 bf16 with
//    ldmatrix.trans, fp32 accumulate in registers), described next.
// CTA = 64 anchor voxels of one row; per (kd,kh) one contiguous partner segment of 64*s+2 voxels
// (the three kw taps are row offsets into it, stride s) is staged with cp.async; a warp owns a set
// of 16x8 output tiles and keeps the accumulators of ALL its taps in registers across the
// persistent tile loop; CTA partials go to a workspace and a second kernel reduces them in a fixed
// order (deterministic, no atomics) and applies the optional per-channel scale (folded BatchNorm).
#include "common.cuh"
#include "ptx.cuh"
#include "tma_host.cuh"

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


// ---------------------------------------------------------------------------------------------
// tcgen05 weight gradient, stride 1, one 32 x 32 channel block (anchor channels a0.., partner channels b0..).
//
// Per padded plane the volume is a flat list of voxels p (row pitch Wq = W+2), and for an interior anchor voxel the
// partner voxel of tap (kd,kh,kw) is at plane d+kd-1, position p + (kh-1)*Wq + (kw-1): dW is a 1-D correlation of
// the two flat lists.  One step = 128 consecutive positions of one anchor plane d:
//     A (M = 128) = gy tile [136 rows][32 ch], MN-major SWIZZLE_64B, read as FOUR OVERLAPPING atoms along M
//                   (descriptor LBO = 64 B = one voxel row): M-atom j is the tile shifted by j voxels  -> kw = 2 - j
//     B (N = 96)  = x plane d+kd-1, three TMA copies of [128 rows][32 ch] starting at p0 + (kh-1)*Wq + 1  -> kh
//     D_kd[(j,co)][(kh,ci)] += A^T B   for kd = 0,1,2  (three accumulators, 288 TMEM columns), K = 128 = 8 MMAs each
// (verified on the GPU by tests/cuda/mnmajor_test.cu).  M-atom 3 is unused: 25 % of the tensor work buys kw for free.
// A CTA owns a contiguous range of steps in (chunk, batch, plane) order, so consecutive steps share two of their
// three x planes (ring of NB plane slots, each loaded once per sweep), accumulates its whole range in TMEM and
// writes one [27][32][32] partial; wgrad_reduce_kernel sums the partials in a fixed order.
// Zero rims make every out-of-range product vanish (gy is 0 there); TMA zero-fills out-of-plane rows.
// Warp 0 = TMA producer, warp 1 = MMA issuer, all four warps drain TMEM at the end.
// ---------------------------------------------------------------------------------------------
namespace tc {

constexpr int KB = 128;                         // positions per step
constexpr int A_ROWS = KB + 8;                  // + the kw shifts, rounded to a swizzle atom
constexpr int A_SLOT = 9216;                    // A_ROWS * 64 rounded to 1024
constexpr int B_COPY = KB * 64;                 // 8192
constexpr int B_SLOT = 3 * B_COPY;
constexpr int NA = 3, NB = 5;
constexpr int SMEM = NA * A_SLOT + NB * B_SLOT + 1024;
constexpr int TMEM_COLS = 512;                  // 3 x 96 used

struct Geom {
    int D, BD;                                   // interior planes per batch item; B * D
    int Wq;                                      // padded row pitch (voxels)
    int nsteps;                                  // nchunks * BD
    int cls_blocks;                              // 0: plain stride-1 problem.  > 0: the partner is the parity-gathered fine tensor of a
                                                 // stride-2 layer (8 classes x cls_blocks 32-channel blocks): class bit = 1 on an axis needs
                                                 // only the centre tap there, bit = 0 the centre and the +1 tap (see s2d_gather_kernel)
    int pad_;
};

__device__ unsigned int g_wgrad_timeouts = 0;
__device__ int g_wgrad_trap = 1;                 // see conv3d.cu: a timed-out wait is fatal unless bring-up mode is on

__device__ __forceinline__ unsigned long long gtimer() {
    unsigned long long t; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t)); return t;
}
// bounded wait (a wedged pipeline must not hang the GPU): after limit_ns the kernel traps (sticky CUDA error, nothing
// can be silently wrong); in bring-up mode it gives up instead, and so does every later wait
__device__ __forceinline__ bool wait_bar(uint32_t bar, uint32_t parity, unsigned long long limit_ns = 2000000000ULL) {
    if (ptx::mbar_try_wait(bar, parity)) return true;
    unsigned long long t0 = 0;
    for (uint32_t spins = 1; ; ++spins) {
        if (ptx::mbar_try_wait(bar, parity)) return true;
        if ((spins & 4095u) == 0u) {
            if (t0 == 0) t0 = gtimer();
            const bool trap = *reinterpret_cast<volatile int*>(&g_wgrad_trap) != 0;
            if (gtimer() - t0 > limit_ns || (!trap && *reinterpret_cast<volatile unsigned int*>(&g_wgrad_timeouts))) {
                atomicAdd(&g_wgrad_timeouts, 1u);
                if (trap) __trap();
                return false;
            }
        }
    }
}
__device__ __forceinline__ uint64_t mn_desc_sw64(uint32_t saddr, uint32_t lbo_bytes) {
    return (uint64_t)((saddr >> 4) & 0x3FFFu) | ((uint64_t)(lbo_bytes >> 4) << 16) | ((uint64_t)(512u >> 4) << 32) |
           (1ull << 46) | (4ull << 61);
}

__global__ void __launch_bounds__(128, 1)
wgrad_tc_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
                float* __restrict__ partial, Geom g, int nslice_b) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (ptx::smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t sA = base, sB = base + NA * A_SLOT;
    __shared__ __align__(8) uint64_t bars[2 * NA + 2 * NB + 1];
    __shared__ uint32_t slot;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    auto full_a = [&](int i) { return ptx::smem_u32(&bars[i]); };
    auto empty_a = [&](int i) { return ptx::smem_u32(&bars[NA + i]); };
    auto full_b = [&](int i) { return ptx::smem_u32(&bars[2 * NA + i]); };
    auto empty_b = [&](int i) { return ptx::smem_u32(&bars[2 * NA + NB + i]); };
    const uint32_t done_bar = ptx::smem_u32(&bars[2 * NA + 2 * NB]);
    if (threadIdx.x == 0) {
        for (int i = 0; i < 2 * NA + 2 * NB + 1; ++i) ptx::mbar_init(ptx::smem_u32(&bars[i]), 1);
        ptx::fence_mbar_init();
        ptx::prefetch_tensormap(&map_a); ptx::prefetch_tensormap(&map_b);
    }
    if (warp == 1) ptx::tmem_alloc(ptx::smem_u32(&slot), TMEM_COLS);
    ptx::tc_fence_before(); __syncthreads(); ptx::tc_fence_after();
    const uint32_t tmem = slot;

    // this CTA's channel block and step range
    const int z = blockIdx.y;
    const int ca0 = (z / nslice_b) * 32, cb0 = (z % nslice_b) * 32;
    const long long t0 = (long long)g.nsteps * blockIdx.x / gridDim.x, t1 = (long long)g.nsteps * (blockIdx.x + 1) / gridDim.x;
    // tap ranges this channel block needs (all 27 for a stride-1 problem)
    int kd_lo = 0, kd_hi = 2, kh_lo = 0, kh_hi = 2, kw_lo = 0, kw_hi = 2;
    if (g.cls_blocks > 0) {
        const int cls = (z % nslice_b) / g.cls_blocks;           // (pd << 2) | (ph << 1) | pw
        kd_lo = kh_lo = kw_lo = 1;
        kd_hi = (cls & 4) ? 1 : 2; kh_hi = (cls & 2) ? 1 : 2; kw_hi = (cls & 1) ? 1 : 2;
    }

    if (warp == 0) {
        if (ptx::elect_one_sync()) {
            uint32_t xn = 0, an = 0;
            bool ok = true;
            auto load_x = [&](int p0, int plane) {
                const int s = xn % NB;
                ok = ok && wait_bar(empty_b(s), ((xn / NB) & 1u) ^ 1u);
                ptx::mbar_arrive_expect_tx(full_b(s), (kh_hi - kh_lo + 1) * B_COPY);     // only the kh copies this class uses
#pragma unroll
                for (int c = 0; c < 3; ++c)
                    if (c >= kh_lo && c <= kh_hi)
                        ptx::tma_load_3d(sB + s * B_SLOT + c * B_COPY, &map_b, full_b(s), cb0, p0 + (c - 1) * g.Wq + 1, plane);
                ++xn;
            };
            for (long long t = t0; t < t1 && ok;) {
                const int q = (int)(t / g.BD), r = (int)(t % g.BD);
                const int b = r / g.D, dl = r % g.D;
                const int len = (int)min((long long)(g.D - dl), t1 - t);
                const int p0 = q * KB, pb = b * (g.D + 2);
                load_x(p0, pb + dl); load_x(p0, pb + dl + 1);
                for (int i = 0; i < len && ok; ++i) {
                    load_x(p0, pb + dl + i + 2);
                    const int s = an % NA;
                    ok = ok && wait_bar(empty_a(s), ((an / NA) & 1u) ^ 1u);
                    ptx::mbar_arrive_expect_tx(full_a(s), A_ROWS * 64);
                    ptx::tma_load_3d(sA + s * A_SLOT, &map_a, full_a(s), ca0, p0, pb + dl + i + 1);
                    ++an;
                }
                t += len;
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        if (ptx::elect_one_sync()) {
            const uint32_t idesc_n = ptx::make_idesc_bf16(32 * (kh_hi - kh_lo + 1)) | (1u << 15) | (1u << 16);     // both operands MN-major
            uint32_t xn = 0, an = 0;
            bool ok = true, first = true;
            for (long long t = t0; t < t1 && ok;) {
                const int r = (int)(t % g.BD);
                const int dl = r % g.D;
                const int len = (int)min((long long)(g.D - dl), t1 - t);
                ok = ok && wait_bar(full_b(xn % NB), (xn / NB) & 1u);
                ok = ok && wait_bar(full_b((xn + 1) % NB), ((xn + 1) / NB) & 1u);
                for (int i = 0; i < len && ok; ++i) {
                    const uint32_t xnew = xn + i + 2;
                    ok = ok && wait_bar(full_b(xnew % NB), (xnew / NB) & 1u);
                    const int sa = an % NA;
                    ok = ok && wait_bar(full_a(sa), (an / NA) & 1u);
                    ptx::tc_fence_after();
                    const uint64_t ad0 = mn_desc_sw64(sA + sa * A_SLOT, 64u);
#pragma unroll
                    for (int kd = 0; kd < 3; ++kd) {
                        if (kd < kd_lo || kd > kd_hi) continue;
                        const uint32_t sb = (xn + i + kd) % NB;
                        const uint64_t bd0 = mn_desc_sw64(sB + sb * B_SLOT + kh_lo * B_COPY, B_COPY);
#pragma unroll
                        for (int ks = 0; ks < KB / 16; ++ks)
                            ptx::umma_bf16(tmem + kd * 96 + kh_lo * 32, ptx::desc_advance(ad0, ks * 1024), ptx::desc_advance(bd0, ks * 1024),
                                           idesc_n, (first && ks == 0) ? 0u : 1u);
                    }
                    first = false;
                    ptx::umma_commit(empty_a(sa));
                    ptx::umma_commit(empty_b((xn + i) % NB));
                    ++an;
                }
                ptx::umma_commit(empty_b((xn + len) % NB));
                ptx::umma_commit(empty_b((xn + len + 1) % NB));
                xn += len + 2;
                t += len;
            }
            ptx::umma_commit(done_bar);
        }
        __syncwarp();
    }
    // epilogue: lane quarter w <-> M-atom j = w <-> kw = 2 - w (quarter 3 carries nothing)
    const bool fin = wait_bar(done_bar, 0, 4000000000ULL);     // the whole sweep: seconds, not the 0.2 s of a pipeline stage
    __syncwarp();
    ptx::tc_fence_after();
    if (fin && warp < 3 && t1 > t0) {
        const int kw = 2 - warp;
        float* out = partial + ((size_t)z * gridDim.x + blockIdx.x) * (27 * 32 * 32);
#pragma unroll 1
        for (int kd = 0; kd < 3; ++kd)
#pragma unroll 1
            for (int kh = 0; kh < 3; ++kh) {
                uint32_t v[32];
                ptx::tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16) + kd * 96 + kh * 32, v);
                ptx::tc_wait_ld();
                if (kd < kd_lo || kd > kd_hi || kh < kh_lo || kh > kh_hi || kw < kw_lo || kw > kw_hi) {   // tap not computed for this class
#pragma unroll
                    for (int i = 0; i < 32; ++i) v[i] = 0u;
                }
                float4* o = reinterpret_cast<float4*>(out + ((size_t)((kd * 3 + kh) * 3 + kw) * 32 + lane) * 32);
#pragma unroll
                for (int i = 0; i < 8; ++i)
                    o[i] = make_float4(__uint_as_float(v[4 * i]), __uint_as_float(v[4 * i + 1]), __uint_as_float(v[4 * i + 2]),
                                       __uint_as_float(v[4 * i + 3]));
            }
    }
    ptx::tc_fence_before();
    __syncthreads();
    if (warp == 1) ptx::tmem_dealloc(tmem, TMEM_COLS);
}

}  // namespace tc

// Parity gather for the stride-2 layers.  With anchor voxel v and tap t the partner voxel is (padded) f = 2v + t per axis:
// f even -> t = 0 (coarse offset 0) or t = 2 (offset +1), f odd -> t = 1 (offset 0).  Q holds the eight parity classes of the
// fine tensor as channel blocks of a volume in the ANCHOR's padded geometry,
//     Q[b][u'][pi*Cb + c] = P[b][2(u'-1) + pi][c]   (per axis; zero where the fine index leaves the padded extent, and at u' = 0),
// so that class pi is a stride-1 weight-gradient problem against the anchor with coarse offsets {0, +1} — taps k = 1, 2 of the
// tcgen05 kernel; k = 2 only exists on axes with pi = 0.  One thread per 16-byte chunk of Q.
__global__ void __launch_bounds__(256)
s2d_gather_kernel(const uint4* __restrict__ P, uint4* __restrict__ Q, int B, int Cb8, int Da, int Ha, int Wa, int Dp, int Hp, int Wp) {
    const long long n = (long long)B * (Da + 2) * (Ha + 2) * (Wa + 2) * 8 * Cb8;
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int c = (int)(i % Cb8); long long r = i / Cb8;
    const int cls = (int)(r % 8); r /= 8;
    const int uw = (int)(r % (Wa + 2)); r /= (Wa + 2);
    const int uh = (int)(r % (Ha + 2)); r /= (Ha + 2);
    const int ud = (int)(r % (Da + 2)); const int b = (int)(r / (Da + 2));
    const int fd = 2 * (ud - 1) + ((cls >> 2) & 1), fh = 2 * (uh - 1) + ((cls >> 1) & 1), fw = 2 * (uw - 1) + (cls & 1);
    uint4 v = make_uint4(0u, 0u, 0u, 0u);
    if (ud >= 1 && uh >= 1 && uw >= 1 && fd <= Dp + 1 && fh <= Hp + 1 && fw <= Wp + 1)
        v = __ldg(P + ((((size_t)b * (Dp + 2) + fd) * (Hp + 2) + fh) * (Wp + 2) + fw) * Cb8 + c);
    Q[i] = v;
}

// Reduction for the gathered form: partial is [a-block][class*nb + b-block][cta][27][32][32] over (kd,kh,kw) in {0,1,2};
// original tap t per axis <- (class bit, k): t = 0 <- (0, 1), t = 1 <- (1, 1), t = 2 <- (0, 2).
__global__ void __launch_bounds__(256)
wgrad_reduce_s2_kernel(const float* __restrict__ partial, float* __restrict__ dw, int ncta, int Ca, int Cb, int CAo, int CBo,
                       const float* __restrict__ scale_a, const float* __restrict__ scale_b, int accumulate) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= 27 * Ca * Cb) return;
    const int b = i % Cb; int r = i / Cb;
    const int a = r % Ca; const int tap = r / Ca;
    const int td = tap / 9, th = (tap / 3) % 3, tw = tap % 3;
    const int cls = ((td == 1) << 2) | ((th == 1) << 1) | (tw == 1);
    const int k = ((td == 2 ? 2 : 1) * 3 + (th == 2 ? 2 : 1)) * 3 + (tw == 2 ? 2 : 1);
    const int nb = Cb / 32;
    const int z = (a / 32) * (8 * nb) + cls * nb + (b / 32);
    const size_t per_cta = (size_t)27 * 32 * 32;
    const float* p = partial + ((size_t)z * ncta) * per_cta + ((size_t)k * 32 + (a % 32)) * 32 + (b % 32);
    float sum = 0.f;
    for (int c = 0; c < ncta; ++c) sum += p[(size_t)c * per_cta];
    if (a >= CAo || b >= CBo) return;
    if (scale_a) sum *= scale_a[a];
    if (scale_b) sum *= scale_b[b];
    float* o = dw + ((size_t)a * CBo + b) * 27 + tap;
    *o = accumulate ? *o + sum : sum;
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


int g_wgrad_mode = 0;          // 0 = tcgen05 kernel where it applies, 1 = warp-level kernel everywhere (A/B timing, tests)

// stride-1 path: returns DSM_EUNSUPPORTED (negative) when the shape does not fit the tcgen05 kernel
int launch_wgrad_tc(const void* anchor, const void* partner, float* partial, int B, int Ca, int Cb, int D, int H, int W,
                    int* ncta_out, cudaStream_t st, int cls_blocks = 0) {
    const long long HpWp = (long long)(H + 2) * (W + 2);
    const long long planes = (long long)B * (D + 2);
    if (HpWp > 0x7fffffffLL - 4096 || planes > 0x7fffffffLL) return DSM_EUNSUPPORTED;
    const int nchunks = (int)dsm_ceil_div_ll(HpWp, tc::KB);
    const long long nsteps = (long long)nchunks * B * D;
    if (nsteps > 0x7fffffffLL) return DSM_EUNSUPPORTED;
    const int nslices = (Ca / 32) * (Cb / 32);
    int nsm = DSM_NUM_SMS_B200, dev = 0;
    if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, dev);
    if (nsm > DSM_NUM_SMS_B200) nsm = DSM_NUM_SMS_B200;      // the workspace is sized for <= 148 partials per channel block... x2
    int ncta = nsm / nslices;
    if (ncta < 1) ncta = 1;
    if (ncta > nsteps) ncta = (int)nsteps;
    CUtensorMap ma, mb;
    cuuint64_t dims_a[3] = {(cuuint64_t)Ca, (cuuint64_t)HpWp, (cuuint64_t)planes};
    cuuint64_t str_a[2] = {(cuuint64_t)Ca * 2, (cuuint64_t)HpWp * Ca * 2};
    cuuint32_t box_a[3] = {32, (cuuint32_t)tc::A_ROWS, 1};
    cuuint64_t dims_b[3] = {(cuuint64_t)Cb, (cuuint64_t)HpWp, (cuuint64_t)planes};
    cuuint64_t str_b[2] = {(cuuint64_t)Cb * 2, (cuuint64_t)HpWp * Cb * 2};
    cuuint32_t box_b[3] = {32, (cuuint32_t)tc::KB, 1};
    if (!tma_host::encode(&ma, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, anchor, 3, dims_a, str_a, box_a, CU_TENSOR_MAP_SWIZZLE_64B) ||
        !tma_host::encode(&mb, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, partner, 3, dims_b, str_b, box_b, CU_TENSOR_MAP_SWIZZLE_64B))
        return DSM_EDRIVER;
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(tc::wgrad_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, tc::SMEM);
        if (e != cudaSuccess) return (int)e;
        attr_set = true;
    }
    tc::Geom g;
    g.D = D; g.BD = B * D; g.Wq = W + 2; g.nsteps = (int)nsteps; g.cls_blocks = cls_blocks; g.pad_ = 0;
    tc::wgrad_tc_kernel<<<dim3(ncta, nslices), 128, tc::SMEM, st>>>(ma, mb, partial, g, Cb / 32);
    *ncta_out = ncta;
    return dsm_launch_status();
}

// bytes of the parity-gathered partner of a stride-2 problem (anchor geometry, 8*Cb channels)
size_t s2_gather_bytes(int B, int Cb, int Da, int Ha, int Wa) {
    return (size_t)B * (Da + 2) * (Ha + 2) * (Wa + 2) * 8 * (size_t)Cb * sizeof(__nv_bfloat16);
}
size_t s2_partial_bytes(int Ca, int Cb) {
    return (size_t)DSM_NUM_SMS_B200 * 27 * 32 * 32 * sizeof(float) + (size_t)(Ca / 32) * (Cb / 32) * 8 * 27 * 32 * 32 * sizeof(float);
}

}  // namespace

// workspace for a given problem: the stride-2 / transposed layers can use the tcgen05 kernel through a parity gather of the
// fine tensor when the caller provides room for it (otherwise they run on the warp-level kernel)
extern "C" size_t dsm_conv3d_wgrad_workspace_bytes_ex(int B, int Ca, int Cb, int Da, int Ha, int Wa, int stride) {
    size_t n = (size_t)DSM_NUM_SMS_B200 * 2 * 27 * (size_t)Ca * Cb * sizeof(float);
    if (stride == 2 && Ca % 32 == 0 && Cb % 32 == 0) {
        const size_t m = ((s2_partial_bytes(Ca, Cb) + 1023) / 1024) * 1024 + s2_gather_bytes(B, Cb, Da, Ha, Wa);
        if (m > n) n = m;
    }
    return n;
}

extern "C" size_t dsm_conv3d_wgrad_workspace_bytes(int Ca, int Cb) {
    return (size_t)DSM_NUM_SMS_B200 * 2 * 27 * (size_t)Ca * Cb * sizeof(float);      // one partial per CTA (<= #SMs), 2x margin
}

extern "C" int dsm_conv3d_wgrad(const void* anchor, const void* partner, float* dw,
                                int B, int Ca, int Cb, int Da, int Ha, int Wa, int Dp, int Hp, int Wp, int stride,
                                int Ca_out, int Cb_out, const float* scale_a, const float* scale_b, int accumulate,
                                void* ws, size_t ws_bytes, void* stream) {
    DsmDeviceGuard dsm_guard_(anchor);
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
    if (stride == 1 && g_wgrad_mode == 0 && Da == Dp && Ha == Hp && Wa == Wp) {
        int ncta_tc = 0;
        const int rc_tc = launch_wgrad_tc(anchor, partner, partial, B, Ca, Cb, Da, Ha, Wa, &ncta_tc, st);
        if (rc_tc != DSM_EUNSUPPORTED) {
            if (rc_tc != 0) return rc_tc;
            const int per_tc = 27 * Ca * Cb;
            wgrad_reduce_kernel<<<dsm_ceil_div(per_tc, 256), 256, 0, st>>>(partial, dw, 1, ncta_tc, Ca, Cb, 32, 32, Ca_out, Cb_out,
                                                                          scale_a, scale_b, accumulate);
            return dsm_launch_status();
        }
    }
    if (stride == 2 && g_wgrad_mode == 0 && Ca % 32 == 0 && Cb % 32 == 0 &&
        ws_bytes >= dsm_conv3d_wgrad_workspace_bytes_ex(B, Ca, Cb, Da, Ha, Wa, 2) &&
        (long long)B * (Da + 2) * (Ha + 2) * (Wa + 2) * 8 * (Cb / 8) < 0x7fffffffLL * 256LL) {
        // parity gather of the fine tensor, then the tcgen05 kernel with 8 x (Cb/32) partner blocks and per-class tap masks
        const size_t poff = ((s2_partial_bytes(Ca, Cb) + 1023) / 1024) * 1024;
        void* q = static_cast<char*>(ws) + poff;
        const long long nchunk = (long long)B * (Da + 2) * (Ha + 2) * (Wa + 2) * 8 * (Cb / 8);
        s2d_gather_kernel<<<(unsigned)dsm_ceil_div_ll(nchunk, 256), 256, 0, st>>>(static_cast<const uint4*>(partner), static_cast<uint4*>(q),
                                                                                 B, Cb / 8, Da, Ha, Wa, Dp, Hp, Wp);
        int rc_g = dsm_launch_status();
        if (rc_g != 0) return rc_g;
        int ncta_tc = 0;
        const int rc_tc = launch_wgrad_tc(anchor, q, partial, B, Ca, 8 * Cb, Da, Ha, Wa, &ncta_tc, st, Cb / 32);
        if (rc_tc != DSM_EUNSUPPORTED) {
            if (rc_tc != 0) return rc_tc;
            const int per_tc = 27 * Ca * Cb;
            wgrad_reduce_s2_kernel<<<dsm_ceil_div(per_tc, 256), 256, 0, st>>>(partial, dw, ncta_tc, Ca, Cb, Ca_out, Cb_out, scale_a, scale_b, accumulate);
            return dsm_launch_status();
        }
    }
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

// 0 = tcgen05 kernel for stride-1 layers (default), 1 = warp-level kernel everywhere; returns the previous mode
extern "C" int dsm_debug_wgrad_mode(int mode) {
    const int prev = g_wgrad_mode;
    if (mode == 0 || mode == 1) g_wgrad_mode = mode;
    return prev;
}
extern "C" int dsm_debug_wgrad_set_trap(int on) {
    on = on ? 1 : 0;
    return (int)cudaMemcpyToSymbol(tc::g_wgrad_trap, &on, sizeof(int));
}

extern "C" int dsm_debug_wgrad_timeouts(void) {
    unsigned int v = 0;
    cudaMemcpyFromSymbol(&v, tc::g_wgrad_timeouts, sizeof(v));
    return (int)v;
}
