// op 3 — 3-D convolution block (k=3, pad=1) as a tcgen05/TMEM implicit GEMM, bf16 x bf16 -> fp32.
// Replaces Conv3d / ConvTranspose3d + BatchNorm3d (+ReLU, +residual) of reference
// models/psmnet/submodule.py:16-19, models/psmnet/stackhourglass.py:26-41,46-60,73-98,135-149,
// models/util_conv.py:150-179 and models/gcnet.py:38-61.
//
// Data layout: activations are bf16 "padded NDHWC" [B][D+2][H+2][W+2][C] with a zero rim, so
//   * the GEMM row of a voxel is its C contiguous channels (K-major operand, 64 or 128 bytes),
//   * for stride-1 convolutions (and each output-parity class of a k3/s2 transposed conv) a
//     filter tap is a CONSTANT row offset in the flattened voxel index, the zero rim supplies
//     the padding, and an M tile is simply 128 consecutive voxels: one 2-D TMA box per tap;
//   * stride-2 convolutions read 16x8 output patches through eight "parity sub-lattice" 5-D
//     tensor maps (tap k along a dim -> parity k&1, half-index o+(k>>1)).
// GEMM: M = 128 voxels (TMEM lanes), N = Cout (TMEM columns, fp32), K = taps * Cin walked in
// (tap, 64- or 32-channel chunk) pipeline stages: TMA -> swizzled smem ring -> tcgen05.mma
// issued by one thread -> tcgen05.commit frees the stage; the accumulator is read back with
// tcgen05.ld and the epilogue applies the folded BatchNorm affine, residual add and ReLU and
// writes bf16 NDHWC (or fp32 for the Cout=1 classifiers) straight from registers.
// MODE_SHIFT loads one (kd,kh) row segment of 130 voxels and issues the three kw taps from it
// with row-shifted matrix descriptors (3x less L2->smem traffic).
//
// Roofline: tensor pipe; algorithmic flops = 2*27*Cin*Cout*voxels (transposed: input voxels).
#include "common.cuh"
#include "ptx.cuh"
#include "epilogue.cuh"
#include "tma_host.cuh"
#include <string.h>

namespace {

enum { MODE_FLAT = 0, MODE_SHIFT = 1, MODE_BOX = 2 };

struct ConvMaps {
    CUtensorMap a[8];    // FLAT/SHIFT use a[0]; BOX uses the 8 parity sub-lattices
    CUtensorMap w;       // packed weights [27*CoutP][Cin]
};

struct ConvGeom {
    int B, Di, Hi, Wi;       // input extent (unpadded)
    int Do, Ho, Wo;          // output buffer extent (unpadded; may be a crop of the natural size)
    int Cout;                // channels actually stored
    int transposed, relu, y_f32;
    int nchunks;             // Cin / KC
    long long P;             // rows of the flattened padded input
    int tiles_w, tiles_h;    // BOX tiling of the output plane (16 x 8 patches)
    int ntiles, mtiles;      // total tiles (all classes); FLAT/SHIFT: 128-row tiles per class
    int desc_variant;        // 0: base_offset 0; 1: base_offset from address bits [7,10)
    int cls_begin[9];
    int a_off[27];           // FLAT/SHIFT: row offset of the tap; BOX: map | ow<<4 | oh<<5 | od<<6
    int w_row[27];           // first row of the tap in the packed weights
    // generalisation to 2-D images (the feature-extraction trunks run on the same kernel): a 3-D volume has
    // dpad = ri = ro = dil = 1, ldy = ldr = Cout, y_mode = y_f32
    int dpad;                // rim planes along D: 1 for volumes, 0 for images (Di = Do = 1)
    int ri, ro;              // rim width along H and W of the input / of the output and residual buffers
    int dil;                 // dilation (stride-1 kernels): tap offsets and the kw row shift scale with it
    int ldy, ldr;            // channels per voxel of the y / residual buffers (>= Cout: a channel slice of a wider tensor)
    int y_mode;              // 0 padded NDHWC bf16; 1 fp32 single channel [B][Do][Ho][Wo]; 2 fp32 NCHW [B][Cout][Ho][Wo] (images)
    int tx_bytes;            // bytes one pipeline stage receives (A rows * row bytes + the weight tiles)
};

__device__ int g_conv_timeouts = 0;
__device__ int g_conv_trap = 1;              // 1 (default): a timed-out pipeline wait is FATAL (__trap -> sticky CUDA error)
__device__ int* g_conv_progress = nullptr;   // bring-up only: host-mapped int[4], one slot per warp of CTA 1

// ptxas lowers tcgen05.wait::ld to nothing and relies on the register scoreboard of the first
// consumer.  A thread that never reads its tcgen05.ld result (an invalid rim voxel) would then run
// on to the CTA barrier / TMEM dealloc / EXIT with the load still in flight — observed on B200 as
// a hung kernel.  This forces every thread to consume one loaded register (one ISETP, and a
// shared store that can only fire for a single NaN payload, which changes nothing).
__device__ __forceinline__ void consume_tmem_load(uint32_t v0, uint32_t scratch_smem) {
    asm volatile(
        "{\n\t.reg .pred q;\n\t"
        "setp.eq.u32 q, %0, 0xFFF0DEAD;\n\t"
        "@q st.shared.u32 [%1], %0;\n\t}"
        :: "r"(v0), "r"(scratch_smem) : "memory");
}

// epilogue activation: relu 0 = none, 1 = ReLU AFTER the residual add (PSMNet: stackhourglass.py:46-58),
// 2 = ReLU BEFORE it (GC-Net: myAdd3d(l33(x), x29) adds to an already-activated deconv, gcnet.py:78-96)
__device__ __forceinline__ float fuse_act(float a, float r, int relu) {
    if (relu == 2) a = fmaxf(a, 0.f);
    a += r;
    if (relu == 1) a = fmaxf(a, 0.f);
    return a;
}

__device__ __forceinline__ unsigned long long globaltimer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

// Bounded wait: a broken pipeline must never hang the GPU, and it must never go unnoticed either.  A wait that
// is not satisfied within 2 s of wall clock (far beyond any legitimate stall: time slicing, a PDL dependent waiting
// for a long primary kernel) counts a timeout and TRAPS: the kernel dies, the context reports a sticky launch
// failure at the next CUDA call and no later result can be silently wrong.  Bring-up mode
// (dsm_debug_conv_set_trap(0)): the wait returns false instead, every further wait in the process gives up at its
// first check, the kernel runs to its end on garbage and dsm_debug_conv_timeouts() tells — only for
// debugging a new kernel without losing the context.
// The fast path is a bare try_wait spin (the instruction itself suspends the warp for a while); the wall clock is
// only consulted every 4096 failed polls.
__device__ __forceinline__ bool wait_bar(uint32_t bar, uint32_t parity) {
    if (ptx::mbar_try_wait(bar, parity)) return true;
    unsigned long long t0 = 0;
    for (uint32_t spins = 1; ; ++spins) {
        if (ptx::mbar_try_wait(bar, parity)) return true;
        if ((spins & 4095u) == 0u) {
            if (t0 == 0) t0 = globaltimer_ns();
            const bool trap = *reinterpret_cast<volatile int*>(&g_conv_trap) != 0;
            if (globaltimer_ns() - t0 > 2000000000ULL /*2 s*/ ||
                (!trap && *reinterpret_cast<volatile int*>(&g_conv_timeouts))) {
                atomicAdd(&g_conv_timeouts, 1);
                if (trap) __trap();
                return false;
            }
        }
    }
}

// WRES (MODE_BOX, one K chunk, 27 taps): all weight tiles stay resident in shared memory and only activation boxes stream
// through the ring — a third less TMA traffic per tile for the stride-2 32 -> 64 layer (hourglass conv1).
template <int KC, int NP, int MODE, int WRES = 0>
struct Cfg {
    static constexpr int ROWB = KC * 2;
    static constexpr int A_ROWS = (MODE == MODE_SHIFT) ? 130 : 128;
    static constexpr int A_BYTES = ((A_ROWS * ROWB + 1023) / 1024) * 1024;
    static constexpr int NB = (MODE == MODE_SHIFT) ? 3 : 1;
    static constexpr int B_TILE = NP * ROWB;
    static constexpr int B_BYTES = ((NB * B_TILE + 1023) / 1024) * 1024;
    static constexpr int W_RES = WRES ? 27 * B_TILE : 0;
    static constexpr int TPS = WRES ? 3 : 1;                  // taps per pipeline stage (WRES: the three kw taps of one (kd, kh))
    static constexpr int STAGE = WRES ? TPS * A_BYTES : A_BYTES + B_BYTES;
    // two CTAs per SM when a useful ring (>= 4 stages) fits in ~100 KB, otherwise one CTA with up to 200 KB
    static constexpr int CTAS_PER_SM = (!WRES && 3 * STAGE <= 100 * 1024) ? 2 : 1;
    static constexpr int BUDGET = WRES ? 224 * 1024 - W_RES : (CTAS_PER_SM == 2 ? 100 : 200) * 1024;
    static constexpr int S_RAW = BUDGET / STAGE;
    static constexpr int S_CAP = 8;
    static constexpr int STAGES = S_RAW > S_CAP ? S_CAP : (S_RAW < 2 ? 2 : S_RAW);
    static constexpr int ACC_COLS = NP < 32 ? 32 : NP;        // TMEM columns of one accumulator
    static constexpr int NUM_ACC = 2;                         // double-buffered: MMA of tile i+1 overlaps epilogue of tile i
    static constexpr int TMEM_COLS = NUM_ACC * ACC_COLS;      // 64 / 128 / 256: a power of two
    static constexpr int BAR_BYTES = 512;
    static constexpr int SMEM = W_RES + STAGES * STAGE + 1024 /*align slack*/ + BAR_BYTES + 2 * NP * 4;
    static constexpr int THREADS = 192;                       // warp 0 TMA, warp 1 MMA, warps 2-5 epilogue
    static_assert(!WRES || (MODE == MODE_BOX && (B_TILE % 1024) == 0), "resident weights: box mode, swizzle-aligned tiles");
};

// Which tile is it, and does it produce anything?
struct TileInfo {
    long long p0;      // FLAT/SHIFT: first row of the tile in the flattened padded input
    int cls;           // transposed conv: output parity class
    int b, d, h, w;    // BOX: batch, output plane, 8-row and 16-column patch index
    bool skip;
};

// Flat padded position -> (b, d', offset inside the plane).  The single-warp roles (TMA producer, MMA issuer) run
// this once per tile on their critical path: 64-bit divisions (~120 dependent instructions each) made the bare
// pipeline of the small-tile kernels cost >2000 cycles per tile, so positions that fit 32 bits use 32-bit divisions.
struct FlatPos { int b, dp, off; };
__device__ __forceinline__ FlatPos flat_decode(long long p, long long plane, long long vol) {
    FlatPos f;
    if (p < 0x7fffffffLL && vol < 0x7fffffffLL) {
        const unsigned up = (unsigned)p, uv = (unsigned)vol, upl = (unsigned)plane;
        const unsigned b = up / uv, rem = up - b * uv, dp = rem / upl;
        f.b = (int)b; f.dp = (int)dp; f.off = (int)(rem - dp * upl);
    } else {
        const long long b = p / vol, rem = p - b * vol, dp = rem / plane;
        f.b = (int)b; f.dp = (int)dp; f.off = (int)(rem - dp * plane);
    }
    return f;
}
// a 128-position tile that lies completely inside a rim plane (d' = 0 or D+1) reads only zeros and stores nothing
__device__ __forceinline__ bool rim_tile(long long p0, long long P, long long plane, long long vol, int Dp, int dpad = 1) {
    if (!dpad) return false;                                                      // an image has no rim planes
    const FlatPos f = flat_decode(p0, plane, vol);
    const long long span = (p0 + 127 < P ? 127 : P - 1 - p0);                    // the tail tile is cut at P
    const bool one_plane = (long long)f.off + span < plane;
    return one_plane && (f.dp == 0 || f.dp == Dp - 1);
}

template <int MODE>
__device__ __forceinline__ TileInfo decode_tile(const ConvGeom& g, int t, long long plane, long long vol, int Dp) {
    TileInfo ti;
    ti.p0 = 0; ti.cls = 0; ti.b = ti.d = ti.h = ti.w = 0; ti.skip = false;
    if (MODE == MODE_BOX) {
        ti.w = t % g.tiles_w; t /= g.tiles_w;
        ti.h = t % g.tiles_h; t /= g.tiles_h;
        ti.d = t % g.Do;      ti.b = t / g.Do;
    } else {
        ti.cls = t / g.mtiles;
        ti.p0 = (long long)(t - ti.cls * g.mtiles) * 128;
        ti.skip = rim_tile(ti.p0, g.P, plane, vol, Dp, g.dpad);
    }
    return ti;
}

// Persistent, warp-specialised implicit-GEMM kernel.  grid = #SMs x CTAS_PER_SM (or fewer tiles),
// every CTA walks tiles blockIdx.x, +gridDim.x, ...:
//   warp 0 (one lane) : TMA producer — keeps the smem ring full, running ahead across tiles
//   warp 1 (one lane) : tcgen05.mma issuer — accumulates a tile into one of two TMEM buffers
//   warps 2..5        : epilogue — tcgen05.ld the finished buffer (lane quadrant = warp % 4),
//                       hand it back, then affine + residual + ReLU + store while the next tile's
//                       MMAs already run into the other buffer.
template <int KC, int NP, int MODE, int WRES>
__global__ void __launch_bounds__(192)
conv3d_igemm_kernel(const __grid_constant__ ConvMaps maps, const __grid_constant__ ConvGeom g,
                    const float* __restrict__ scale, const float* __restrict__ shift,
                    const void* __restrict__ residual, void* __restrict__ y) {
    using C = Cfg<KC, NP, MODE, WRES>;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int Dp = g.Di + 2 * g.dpad, Hp = g.Hi + 2 * g.ri, Wp = g.Wi + 2 * g.ri;
    const long long plane = (long long)Hp * Wp, vol = plane * Dp;

    // ---- shared memory carve-up ---------------------------------------------------------
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw = ptx::smem_u32(smem_raw);
    const uint32_t wres = (raw + 1023u) & ~1023u;          // WRES: the 27 resident weight tiles, then the ring
    const uint32_t base = wres + C::W_RES;
    uint8_t* base_ptr = smem_raw + (base - raw);
    const uint32_t bars = base + C::STAGES * C::STAGE;     // full[S], empty[S], tmem_full[2], tmem_empty[2], [wfull], slot, scratch
    auto full_bar = [&](int s) { return bars + 8u * s; };
    auto empty_bar = [&](int s) { return bars + 8u * (C::STAGES + s); };
    auto tfull_bar = [&](int a) { return bars + 8u * (2 * C::STAGES + a); };
    auto tempty_bar = [&](int a) { return bars + 8u * (2 * C::STAGES + C::NUM_ACC + a); };
    const uint32_t wfull_bar = bars + 8u * (2 * C::STAGES + 2 * C::NUM_ACC);
    constexpr int NBARS = 2 * C::STAGES + 2 * C::NUM_ACC + 1;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(base_ptr + C::STAGES * C::STAGE + 8 * NBARS);
    const uint32_t scratch_smem = bars + 8u * NBARS + 8u;
    float* s_scale = reinterpret_cast<float*>(base_ptr + C::STAGES * C::STAGE + C::BAR_BYTES);
    float* s_shift = s_scale + NP;

    if (tid < NP) {
        s_scale[tid] = scale ? __ldg(scale + tid) : 1.f;
        s_shift[tid] = shift ? __ldg(shift + tid) : 0.f;
    }
    if (warp == 0 && lane == 0) {
        ptx::prefetch_tensormap(&maps.w);
        ptx::prefetch_tensormap(&maps.a[0]);
        for (int s = 0; s < C::STAGES; ++s) { ptx::mbar_init(full_bar(s), 1); ptx::mbar_init(empty_bar(s), 1); }
        for (int a = 0; a < C::NUM_ACC; ++a) { ptx::mbar_init(tfull_bar(a), 1); ptx::mbar_init(tempty_bar(a), 4); }
        ptx::mbar_init(wfull_bar, 1);
        ptx::fence_mbar_init();
    }
    if (warp == 1) ptx::tmem_alloc(ptx::smem_u32(tmem_slot), C::TMEM_COLS);
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem = *tmem_slot;

    ptx::griddep_launch_dependents();
    if (warp == 0) {
        // ================= TMA producer (whole warp walks the loop, one elected lane issues) =================
        if (WRES) {                                    // the weights are parameters: before the dependency wait
            if (ptx::elect_one_sync()) {
                const int nt = g.cls_begin[1] - g.cls_begin[0];
                ptx::mbar_arrive_expect_tx(wfull_bar, (uint32_t)(nt * C::B_TILE));
                for (int tp = 0; tp < nt; ++tp)
                    ptx::tma_load_2d(wres + tp * C::B_TILE, &maps.w, wfull_bar, 0, g.w_row[g.cls_begin[0] + tp]);
            }
            __syncwarp();
        }
        ptx::griddep_wait();                           // the activations are the previous kernel's output
        int it = 0;                                    // ring position, runs on across tiles
        for (int t = blockIdx.x; t < g.ntiles; t += gridDim.x) {
            const TileInfo ti = decode_tile<MODE>(g, t, plane, vol, Dp);
            if (ti.skip) continue;
            const int tap0 = g.cls_begin[ti.cls];
            const int ntaps = g.cls_begin[ti.cls + 1] - tap0;
            if (WRES) {
                // one stage = the three kw taps of a (kd, kh): one barrier round trip per three boxes
                for (int grp = 0; grp < ntaps / 3; ++grp, ++it) {
                    const int s = it % C::STAGES;
                    const uint32_t ph = (uint32_t)(it / C::STAGES) & 1u;
                    wait_bar(empty_bar(s), ph ^ 1u);
                    if (ptx::elect_one_sync()) {
                        ptx::mbar_arrive_expect_tx(full_bar(s), (uint32_t)(3 * 128 * C::ROWB));
#pragma unroll
                        for (int j = 0; j < 3; ++j) {
                            const int code = g.a_off[tap0 + 3 * grp + j];
                            ptx::tma_load_5d(base + s * C::STAGE + j * C::A_BYTES, &maps.a[code & 7], full_bar(s), 0,
                                             ti.w * 16 + ((code >> 4) & 1), ti.h * 8 + ((code >> 5) & 1),
                                             ti.d + ((code >> 6) & 1), ti.b);
                        }
                    }
                    __syncwarp();
                }
                continue;
            }
            const int ngroups = (MODE == MODE_SHIFT) ? ntaps / 3 : ntaps;
            for (int grp = 0; grp < ngroups; ++grp) {
                const int tp = tap0 + ((MODE == MODE_SHIFT) ? 3 * grp : grp);
                for (int kc = 0; kc < g.nchunks; ++kc, ++it) {
                    const int s = it % C::STAGES;
                    const uint32_t ph = (uint32_t)(it / C::STAGES) & 1u;
                    wait_bar(empty_bar(s), ph ^ 1u);
                    const uint32_t sa = base + s * C::STAGE, sb = sa + C::A_BYTES;
                    if (ptx::elect_one_sync()) {
                        ptx::mbar_arrive_expect_tx(full_bar(s), (uint32_t)g.tx_bytes);
                        if (MODE == MODE_BOX) {
                            const int code = g.a_off[tp];
                            ptx::tma_load_5d(sa, &maps.a[code & 7], full_bar(s), kc * KC,
                                             ti.w * 16 + ((code >> 4) & 1), ti.h * 8 + ((code >> 5) & 1),
                                             ti.d + ((code >> 6) & 1), ti.b);
                        } else {
                            ptx::tma_load_2d(sa, &maps.a[0], full_bar(s), kc * KC, (int)(ti.p0 + g.a_off[tp]));
                        }
#pragma unroll
                        for (int j = 0; j < C::NB; ++j)
                            ptx::tma_load_2d(sb + j * C::B_TILE, &maps.w, full_bar(s), kc * KC, g.w_row[tp + j]);
                    }
                    __syncwarp();
                }
            }
        }
    } else if (warp == 1) {
        // ================= MMA issuer (whole warp walks the loop, one elected lane issues) =================
        constexpr uint32_t idesc = ptx::make_idesc_bf16(NP);
        int it = 0, tcount = 0;
        if (WRES) { wait_bar(wfull_bar, 0); ptx::tc_fence_after(); }
        for (int t = blockIdx.x; t < g.ntiles; t += gridDim.x) {
            const TileInfo ti = decode_tile<MODE>(g, t, plane, vol, Dp);
            if (ti.skip) continue;
            const int ntaps = g.cls_begin[ti.cls + 1] - g.cls_begin[ti.cls];
            const int n_it = ((MODE == MODE_SHIFT || WRES) ? ntaps / 3 : ntaps) * g.nchunks;
            const int acc = tcount & 1;
            const uint32_t acc_ph = (uint32_t)(tcount >> 1) & 1u;
            wait_bar(tempty_bar(acc), acc_ph ^ 1u);        // epilogue has drained this buffer
            ptx::tc_fence_after();
            const uint32_t d_tmem = tmem + acc * C::ACC_COLS;
            for (int i = 0; i < n_it; ++i, ++it) {
                const int s = it % C::STAGES;
                const uint32_t ph = (uint32_t)(it / C::STAGES) & 1u;
                wait_bar(full_bar(s), ph);
                ptx::tc_fence_after();
                const uint32_t sa = base + s * C::STAGE, sb = WRES ? wres + (uint32_t)(3 * i) * C::B_TILE : sa + C::A_BYTES;
                if (ptx::elect_one_sync()) {
                    const uint64_t ad0 = ptx::make_kmajor_desc(sa, C::ROWB, 0u);
                    const uint64_t bd0 = ptx::make_kmajor_desc(sb, C::ROWB, 0u);
                    if (WRES) {
#pragma unroll
                        for (int j = 0; j < 3; ++j) {
#pragma unroll
                            for (int k = 0; k < KC / 16; ++k)
                                ptx::umma_bf16(d_tmem, ptx::desc_advance(ad0, j * C::A_BYTES + k * 32),
                                               ptx::desc_advance(bd0, j * C::B_TILE + k * 32), idesc, (i | j | k) ? 1u : 0u);
                        }
                    } else
#pragma unroll
                    for (int j = 0; j < C::NB; ++j) {
#pragma unroll
                        for (int k = 0; k < KC / 16; ++k) {
                            ptx::umma_bf16(d_tmem, ptx::desc_advance(ad0, j * g.dil * C::ROWB + k * 32),
                                           ptx::desc_advance(bd0, j * C::B_TILE + k * 32), idesc, (i | j | k) ? 1u : 0u);
                        }
                    }
                    ptx::umma_commit(empty_bar(s));        // frees the stage when these MMAs retire
                    if (i == n_it - 1) ptx::umma_commit(tfull_bar(acc));   // accumulator complete
                }
                __syncwarp();
            }
            ++tcount;
        }
    } else {
        // ================= epilogue warps (one TMEM lane = one voxel per thread) =================
        ptx::griddep_wait();                               // residual reads / output writes must follow the previous kernel
        const int q = warp & 3;                            // TMEM lane quadrant this warp may read
        const int r = q * 32 + lane;                       // row of the tile
        int tcount = 0;
        const int emode = epi_mode(g.relu, residual != nullptr);
        for (int t = blockIdx.x; t < g.ntiles; t += gridDim.x) {
            const TileInfo ti = decode_tile<MODE>(g, t, plane, vol, Dp);
            if (ti.skip) continue;
            // ---- where does this row go? ----
            bool valid = false;
            int ob = 0, od = 0, oh = 0, ow = 0;
            if (MODE == MODE_BOX) {
                ob = ti.b; od = ti.d; oh = ti.h * 8 + (r >> 4); ow = ti.w * 16 + (r & 15);
                valid = (oh < g.Ho) && (ow < g.Wo);
            } else {
                const long long p = ti.p0 + r;
                if (p < g.P) {
                    const FlatPos fp = flat_decode(p, plane, vol);
                    const int b = fp.b, dp = fp.dp, rem2 = fp.off;
                    const int hp = (int)((unsigned)rem2 / (unsigned)Wp), wp = rem2 - hp * Wp;
                    if (dp >= g.dpad && dp < g.Di + g.dpad && hp >= g.ri && hp < g.Hi + g.ri && wp >= g.ri && wp < g.Wi + g.ri) {
                        ob = (int)b;
                        if (g.transposed) {
                            od = 2 * (dp - 1) + ((ti.cls >> 2) & 1); oh = 2 * (hp - 1) + ((ti.cls >> 1) & 1); ow = 2 * (wp - 1) + (ti.cls & 1);
                        } else { od = dp - g.dpad; oh = hp - g.ri; ow = wp - g.ri; }
                        valid = (od < g.Do) && (oh < g.Ho) && (ow < g.Wo);
                    }
                }
            }
            const int acc = tcount & 1;
            const uint32_t acc_ph = (uint32_t)(tcount >> 1) & 1u;
            ++tcount;
            wait_bar(tfull_bar(acc), acc_ph);
            __syncwarp();
            ptx::tc_fence_after();
            const uint32_t taddr = tmem + ((uint32_t)(q * 32) << 16) + acc * C::ACC_COLS;

            if (g.y_f32) {
                // single output channel (classifier / GC-Net l37): fp32, unpadded [B][Do][Ho][Wo]
                uint32_t v[16];
                ptx::tmem_ld16(taddr, v);
                ptx::tc_wait_ld();
                consume_tmem_load(v[0], scratch_smem);
                ptx::tc_fence_before();
                __syncwarp();
                if (lane == 0) ptx::mbar_arrive(tempty_bar(acc));
                if (valid) {
                    const size_t o = (((size_t)ob * g.Do + od) * g.Ho + oh) * g.Wo + ow;
                    const float a = fuse_act(fmaf(__uint_as_float(v[0]), s_scale[0], s_shift[0]),
                                             residual ? __ldg(reinterpret_cast<const float*>(residual) + o) : 0.f, g.relu);
                    reinterpret_cast<float*>(y)[o] = a;
                }
            } else if (g.y_mode == 2) {
                // images only: fp32 NCHW [B][Cout][Ho][Wo] (the layout the cost-volume / correlation ops take); a warp's 32
                // consecutive pixels make every per-channel store a 128-byte run
                constexpr int CH = NP >= 32 ? 32 : 16;
                const size_t hw = (size_t)g.Ho * g.Wo;
                float* outp = reinterpret_cast<float*>(y) + (size_t)ob * g.Cout * hw + (size_t)oh * g.Wo + ow;
#pragma unroll
                for (int c0 = 0; c0 < NP; c0 += CH) {
                    uint32_t v[CH];
                    if (CH == 32) ptx::tmem_ld32(taddr + c0, v); else ptx::tmem_ld16(taddr + c0, v);
                    ptx::tc_wait_ld();
                    consume_tmem_load(v[0], scratch_smem);
                    if (c0 + CH >= NP) {
                        ptx::tc_fence_before();
                        __syncwarp();
                        if (lane == 0) ptx::mbar_arrive(tempty_bar(acc));
                    }
                    if (valid) {
#pragma unroll
                        for (int i = 0; i < CH; ++i) {
                            if (c0 + i < g.Cout) {
                                float a = fmaf(__uint_as_float(v[i]), s_scale[c0 + i], s_shift[c0 + i]);
                                if (g.relu) a = fmaxf(a, 0.f);
                                outp[(size_t)(c0 + i) * hw] = a;
                            }
                        }
                    }
                }
            } else {
                constexpr int CH = NP >= 32 ? 32 : 16;
                const size_t ov = (((size_t)ob * (g.Do + 2 * g.dpad) + od + g.dpad) * (g.Ho + 2 * g.ro) + oh + g.ro) * (g.Wo + 2 * g.ro) + ow + g.ro;
                const uint4* res = residual ? reinterpret_cast<const uint4*>(reinterpret_cast<const __nv_bfloat16*>(residual) + ov * (size_t)g.ldr) : nullptr;
                uint4* out = reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(y) + ov * (size_t)g.ldy);
#pragma unroll
                for (int c0 = 0; c0 < NP; c0 += CH) {
                    uint32_t v[CH];
                    if (CH == 32) ptx::tmem_ld32(taddr + c0, v); else ptx::tmem_ld16(taddr + c0, v);
                    ptx::tc_wait_ld();
                    consume_tmem_load(v[0], scratch_smem);
                    if (c0 + CH >= NP) {                   // last chunk is in registers: hand the buffer back
                        ptx::tc_fence_before();
                        __syncwarp();
                        if (lane == 0) ptx::mbar_arrive(tempty_bar(acc));
                    }
                    if (valid) {
                        DSM_EPI_DISPATCH(emode, epi_store128, CH / 8, v, s_scale + c0, s_shift + c0, res ? res + c0 / 8 : nullptr, out + c0 / 8)
                    }
                }
            }
        }
    }

    ptx::tc_fence_before();
    __syncthreads();
    if (warp == 1) ptx::tmem_dealloc(tmem, C::TMEM_COLS);
}

// ---------------------------------------------------------------------------------------------
// Plane-sharing kernel for stride-1 convolutions with few output channels (Cout <= 32).
//
// Measured on B200 (tests/cuda/mma_bench.cu): an SS-mode tcgen05.mma M=128,K=16 costs
// max(N/2, 32 + N/4) cycles — the 4 KB A slab is re-read from shared memory by every instruction, so
// N=32 tops out at 40% of the tensor pipe, N=96 at 86%.  This kernel therefore widens N by letting
// ONE activation tile feed the three kd taps at once: a CTA owns a band of R=8 consecutive output
// planes of one 128-voxel flat (h,w) tile, each plane with its own NP accumulator columns
// (R*NP columns per buffer, two buffers = all 512 TMEM columns for NP=32).  The A tile of input
// plane i contributes to output planes i-1, i, i+1 through taps kd=2,1,0, whose weight rows are laid
// out back to back in shared memory, so one MMA with N = 3*NP updates three adjacent accumulator
// blocks.  All 27 weight tiles stay resident in shared memory for the whole kernel (55 KB for
// 32->32, 110 KB for 64->32); only 130-row activation tiles (the three kw taps are row-shifted
// descriptors into the same tile) stream through the TMA ring: (R+2)*3 tile loads per R*128
// output voxels instead of 27 (or 9) per 128.
// Issue rate: these MMAs are short (56 cycles at N=96), and one warp cannot prepare descriptors and
// issue faster than ~76 cycles per MMA (measured), so THREE warps issue, one per kh tap.  MMAs from
// different threads are not ordered against each other; that is harmless because every MMA
// accumulates — the epilogue warps hand each accumulator block back zero-filled (tcgen05.st), so
// there is no "first MMA overwrites" case and the sum commutes.
// ---------------------------------------------------------------------------------------------
struct RsGeom {
    int B, D, H, W;          // extent (stride 1: input == natural output extent)
    int Do, Ho, Wo;          // output buffer extent (may be a crop)
    int Cout, relu, y_f32;
    int plane_tiles;         // ceil(Hp*Wp / 128)
    int nbands;              // ceil(Do / R)
    int nitems;              // B * nbands * plane_tiles
    int dbg;                 // timing experiments only (variant bits 4..6): 1 = no epilogue work, 2 = no TMA traffic, 4 = no MMAs
    int ngroups;             // output-channel groups of NP channels (Cout = ngroups*NP when > 1): CTA c owns group c % ngroups
    int kwfold;              // single-output-channel layers (KC = 32, fp32 output): the kw taps are output COLUMNS (see below)
    int w_row[27];           // first row of tap (kd*3+kh)*3+kw in the packed weights
    // fused concat volume (FUSED instantiation only): the A operand is built in shared memory from the two feature maps
    const __nv_bfloat16* featL;   // bf16 NHWC [B][H][W][32]
    const __nv_bfloat16* featR;
    int vol_mode;                 // DSM_VOL_PSM / DSM_VOL_GC / DSM_VOL_GC_RIGHT
};

template <int KC, int NP>
struct RsCfg {
    static constexpr int R = 8;
    static constexpr int ROWB = KC * 2;
    static constexpr int A_ROWS = 130;
    static constexpr int A_BYTES = ((A_ROWS * ROWB + 1023) / 1024) * 1024;
    static constexpr int KHS = (KC <= 32) ? 3 : 1;               // kh tiles per pipeline stage (a whole input plane when they fit)
    static constexpr int STAGE_BYTES = KHS * A_BYTES;
    static constexpr int W_TILE = 3 * NP * ROWB;                  // the three kd taps of one (kh,kw), kd = 2,1,0
    static constexpr int W_BYTES = 9 * W_TILE;
    static constexpr int BAR_BYTES = 512;
    static constexpr int BUDGET = 225 * 1024 - W_BYTES - 1024 - BAR_BYTES - 2 * NP * 4;
    static constexpr int S_RAW = BUDGET / STAGE_BYTES;
    // A multiple of the three MMA issuers: stage s then always belongs to issuer s % 3, which observes EVERY phase of its
    // barrier.  (With a depth of 5, 7 or 8 a warp would look at a barrier only every third use, and a parity wait cannot
    // tell "two phases behind" from "done")
    static constexpr int STAGES = S_RAW >= 9 ? 9 : (S_RAW >= 6 ? 6 : S_RAW);
    static constexpr int TX_BYTES = KHS * A_ROWS * ROWB;
    static constexpr int ACC_COLS = R * NP;
    static constexpr int TMEM_COLS = 2 * ACC_COLS;                // 256 (NP=16) or 512 (NP=32)
    static constexpr int SMEM = W_BYTES + STAGES * STAGE_BYTES + 1024 + BAR_BYTES + 2 * NP * 4;
    static constexpr int THREADS = 384;                           // warp 0 TMA, warps 1-3 MMA (kh = 0,1,2), warps 4-11 epilogue
    static_assert(STAGES >= 4 && STAGES % 3 == 0, "activation ring: at least 4 stages, a multiple of the issuer count");
    static_assert((W_TILE % 1024) == 0 && ((NP * ROWB) % 1024) == 0, "weight sub-tiles must keep the swizzle phase");
};

struct RsItem { int b, z0, nb, tile; };

__device__ __forceinline__ RsItem rs_decode(const RsGeom& g, int t, int R) {
    RsItem it;
    it.tile = t % g.plane_tiles; t /= g.plane_tiles;
    const int band = t % g.nbands; it.b = t / g.nbands;
    it.z0 = band * R;
    it.nb = min(R, g.Do - it.z0);
    return it;
}

// FUSED (KC = 64 only): the input IS the concatenation cost volume of two 32-channel feature maps
// (stackhourglass.py:124-133, gcnet.py:131-135,156-164) and is never materialised.  A voxel's 64 channels are two K chunks:
// first half = map A at the voxel's (y, x), second half = map B at (y, x -/+ d).  With the feature maps stored like every
// other activation (bf16 NHWC with a zero rim, the pitch of the volume's planes) each half of a tap-row tile is a plain
// TMA box of 130 consecutive padded pixels — the second one read d pixels further left (right) — landing as a 64-byte-row
// SWIZZLE_64B tile; TMA stays the data mover.  What TMA cannot express is the reference's zero region (x < d: both
// halves for PSMNet, the second half for GC-Net; rim positions of the shifted half): four "masking" warps wait for the
// stage's loads, zero those rows (a few per tile, most tiles none) and only then release the stage to the MMA issuers,
// which run the K steps of the first half from one tile and those of the second half from the other, against the same
// resident 128-byte-row weight tiles.  (A first version built the whole tile with cp.async from four warps: correct, but
// bound by their instruction stream — profiles/r02g_fused_volume_ab.txt.)
// kw-fold mode (classif*.2 of PSMNet: Conv3d 32 -> 1).  Run as a 32 -> 16 layer, 15 of the 16 accumulator columns and two
// thirds of the MMAs are wasted on zeros.  Here the three kw taps become three output COLUMNS of a 3x3x1 convolution —
// Z[v][kw] = sum over (kd, kh, c) of x[v + (kd,kh) offset][c] * w[kd][kh][kw][c] — which needs one MMA per (kh, K step)
// instead of three, and the epilogue adds the shifted columns, out[v] = Z[v-1][0] + Z[v][1] + Z[v+1][2], with two warp
// shuffles.  So that no lane needs a neighbour from another warp, each 32-lane quadrant of a tile covers 32 consecutive padded
// positions of which the inner 30 are outputs: a tile advances by 120 positions and its A operand is four 32-row TMA boxes.
template <int KC, int NP, bool FUSED = false>
__global__ void __launch_bounds__(FUSED ? 512 : 384, 1)
conv3d_rs_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_a2,
                 const __grid_constant__ CUtensorMap map_w,
                 const __grid_constant__ RsGeom g, const float* __restrict__ scale, const float* __restrict__ shift,
                 const void* __restrict__ residual, void* __restrict__ y) {
    using C = RsCfg<KC, NP>;
    constexpr int HALF_BYTES = ((C::A_ROWS * 64 + 511) / 512) * 512;       // FUSED: one 130-row x 64-byte (SWIZZLE_64B) half tile
    static_assert(!FUSED || (KC == 64 && 2 * HALF_BYTES <= C::STAGE_BYTES), "two half tiles must fit one stage");
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int Dp = g.D + 2, Hp = g.H + 2, Wp = g.W + 2;
    const int plane = Hp * Wp;
    long long dbg_c0 = 0; unsigned long long dbg_t0 = 0;
    if (g_conv_progress && blockIdx.x == 0 && tid == 0) { dbg_c0 = clock64(); dbg_t0 = globaltimer_ns(); }
    // Cout > NP (64->64 layers): the output channels are split into groups of NP; a CTA keeps the weights of ONE group
    // resident and walks the items with the CTAs of its group (the activation tiles are read once per group, from L2)
    const int grp = (int)blockIdx.x % g.ngroups;
    const int item0 = (int)blockIdx.x / g.ngroups, item_step = (int)gridDim.x / g.ngroups;
    const int ch0 = grp * NP;                                  // first output channel of this CTA

    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw = ptx::smem_u32(smem_raw);
    const uint32_t base = (raw + 1023u) & ~1023u;
    uint8_t* base_ptr = smem_raw + (base - raw);
    const uint32_t wsm = base;                                  // resident weights
    const uint32_t ring = base + C::W_BYTES;                    // activation ring
    constexpr int RING_END = C::W_BYTES + C::STAGES * C::STAGE_BYTES;
    const uint32_t bars = base + RING_END;                      // full[S], empty[S], tfull[2], tempty[2], wfull, slot, scratch
    auto full_bar = [&](int s) { return bars + 8u * s; };
    auto empty_bar = [&](int s) { return bars + 8u * (C::STAGES + s); };
    auto tfull_bar = [&](int a) { return bars + 8u * (2 * C::STAGES + a); };
    auto tempty_bar = [&](int a) { return bars + 8u * (2 * C::STAGES + 2 + a); };
    const uint32_t wfull_bar = bars + 8u * (2 * C::STAGES + 4);
    constexpr int NBARS = 2 * C::STAGES + 5;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(base_ptr + RING_END + 8 * NBARS);
    const uint32_t scratch_smem = bars + 8u * NBARS + 8u;
    auto ready_bar = [&](int s) { return bars + 8u * (NBARS + 2 + s); };        // FUSED: stage masked, MMAs may read it
    static_assert(8 * (NBARS + 2 + C::STAGES) <= C::BAR_BYTES, "barrier block too small");
    float* s_scale = reinterpret_cast<float*>(base_ptr + RING_END + C::BAR_BYTES);
    float* s_shift = s_scale + NP;

    if (tid < NP) {
        s_scale[tid] = scale ? __ldg(scale + ch0 + tid) : 1.f;
        s_shift[tid] = shift ? __ldg(shift + ch0 + tid) : 0.f;
    }
    if (warp == 0 && lane == 0) {
        ptx::prefetch_tensormap(&map_w);
        ptx::prefetch_tensormap(&map_a);
        if (FUSED) ptx::prefetch_tensormap(&map_a2);
        for (int s = 0; s < C::STAGES; ++s) { ptx::mbar_init(full_bar(s), 1); ptx::mbar_init(empty_bar(s), 1); }
        if (FUSED) for (int s = 0; s < C::STAGES; ++s) ptx::mbar_init(ready_bar(s), 4);
        for (int a = 0; a < 2; ++a) { ptx::mbar_init(tfull_bar(a), 3); ptx::mbar_init(tempty_bar(a), 8); }
        ptx::mbar_init(wfull_bar, 1);
        ptx::fence_mbar_init();
    }
    if (warp == 1) ptx::tmem_alloc(ptx::smem_u32(tmem_slot), C::TMEM_COLS);
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem = *tmem_slot;

    ptx::griddep_launch_dependents();
    if (FUSED) {
        // 512 threads cap the kernel at 128 registers per thread, but the epilogue wants ~170 (the unfused build has them):
        // the two light warpgroups (TMA + MMA issuers, masking warps) hand registers to the two epilogue warpgroups
        if (warp < 4 || warp >= 12) asm volatile("setmaxnreg.dec.sync.aligned.u32 56;" ::: "memory");
        else                        asm volatile("setmaxnreg.inc.sync.aligned.u32 192;" ::: "memory");
    }
    if (warp == 0) {
        // ================= TMA producer =================
        if (ptx::elect_one_sync()) {                            // all 27 weight tiles, once
            ptx::mbar_arrive_expect_tx(wfull_bar, 27 * NP * C::ROWB);
            for (int kd = 0; kd < 3; ++kd)
                for (int kh = 0; kh < 3; ++kh)
                    for (int kw = 0; kw < 3; ++kw)
                        ptx::tma_load_2d(wsm + (kh * 3 + kw) * C::W_TILE + (2 - kd) * NP * C::ROWB, &map_w, wfull_bar, 0,
                                         g.w_row[(kd * 3 + kh) * 3 + kw] + ch0);
        }
        __syncwarp();
        ptx::griddep_wait();                                    // weights are parameters; the activations are the previous kernel's output
        int s = 0; uint32_t ph = 0;                              // ring slot and its phase
        for (int t = item0; t < g.nitems; t += item_step) {
            const RsItem item = rs_decode(g, t, C::R);
            for (int i = -1; i <= item.nb; ++i) {
                const int zp = item.z0 + 1 + i;                  // padded input plane
                if (zp <= 0 || zp >= Dp - 1) continue;           // zero rim plane: contributes nothing
                const int row0 = FUSED ? item.b * plane + item.tile * 128 - Wp - 1        // feature maps have no D dimension
                                       : (item.b * Dp + zp) * plane + item.tile * 128 - Wp - 1;
                for (int khs = 0; khs < 3 / C::KHS; ++khs) {
                    wait_bar(empty_bar(s), ph ^ 1u);
                    if (ptx::elect_one_sync()) {
                        if (g.dbg & 2) {
                            ptx::mbar_arrive(full_bar(s));
                        } else if (FUSED) {
                            const int dshift = (g.vol_mode == DSM_VOL_GC_RIGHT) ? (zp - 1) : -(zp - 1);
                            ptx::mbar_arrive_expect_tx(full_bar(s), 2 * C::A_ROWS * 64);
                            ptx::tma_load_2d(ring + s * C::STAGE_BYTES, &map_a, full_bar(s), 0, row0 + khs * Wp);
                            ptx::tma_load_2d(ring + s * C::STAGE_BYTES + HALF_BYTES, &map_a2, full_bar(s), 0, row0 + khs * Wp + dshift);
                        } else if (KC == 32 && g.kwfold) {
                            // lane l of quadrant q <-> padded position tile*120 + 30q - 1 + l: four 32-row boxes per tap row
                            const int r0f = (item.b * Dp + zp) * plane + item.tile * 120 - 1 - Wp;
                            ptx::mbar_arrive_expect_tx(full_bar(s), C::KHS * 128 * C::ROWB);
#pragma unroll
                            for (int kk = 0; kk < C::KHS; ++kk)
#pragma unroll
                                for (int qd = 0; qd < 4; ++qd)
                                    ptx::tma_load_2d(ring + s * C::STAGE_BYTES + kk * C::A_BYTES + qd * 32 * C::ROWB, &map_a2, full_bar(s), 0,
                                                     r0f + (khs * C::KHS + kk) * Wp + qd * 30);
                        } else {
                            ptx::mbar_arrive_expect_tx(full_bar(s), C::TX_BYTES);
#pragma unroll
                            for (int kk = 0; kk < C::KHS; ++kk)
                                ptx::tma_load_2d(ring + s * C::STAGE_BYTES + kk * C::A_BYTES, &map_a, full_bar(s), 0,
                                                 row0 + (khs * C::KHS + kk) * Wp);
                        }
                    }
                    __syncwarp();
                    if (++s == C::STAGES) { s = 0; ph ^= 1u; }
                }
            }
        }
    } else if (warp <= 3) {
        // ================= MMA issuers: warp w issues the taps with kh = w-1 =================
        const int my_kh = warp - 1;
        constexpr uint32_t idesc0 = ptx::make_idesc_bf16(0);
        wait_bar(wfull_bar, 0);
        ptx::tc_fence_after();
        // descriptor words: only the start-address field of the low word changes from MMA to MMA
        const uint64_t dsc = ptx::make_kmajor_desc(0u, C::ROWB, 0u);
        const uint32_t desc_hi = (uint32_t)(dsc >> 32);
        const uint32_t ring_lo = (uint32_t)dsc | (ring >> 4);
        const uint32_t w_lo = ((uint32_t)dsc | (wsm >> 4)) + (uint32_t)((my_kh * 3 * C::W_TILE) >> 4);
        const uint32_t a_tile = 0u;
        const uint32_t w_base = (uint32_t)dsc | (wsm >> 4);
        int s = 0; uint32_t ph = 0;
        int tcount = 0, rot = 0;
        for (int t = item0; t < g.nitems; t += item_step) {
            const RsItem item = rs_decode(g, t, C::R);
            const int acc = tcount & 1;
            const uint32_t use_ph = (uint32_t)(tcount >> 1) & 1u;
            ++tcount;
            wait_bar(tempty_bar(acc), use_ph);                   // drained AND zero-filled by the epilogue warps
            ptx::tc_fence_after();
            const uint32_t d_tmem = tmem + acc * C::ACC_COLS;
            int last_i = item.nb;
            if (item.z0 + 1 + last_i >= Dp - 1) --last_i;        // trailing rim plane is skipped
            for (int i = -1; i <= item.nb; ++i) {
                const int zp = item.z0 + 1 + i;
                if (zp <= 0 || zp >= Dp - 1) continue;
                const int jlo = max(i - 1, 0), jhi = min(i + 1, item.nb - 1);
                const int brow = (2 - (i - jlo + 1)) * NP;        // weight row of block jlo's tap (kd = i-jlo+1)
                const uint32_t d_lo = d_tmem + jlo * NP;
                const uint32_t idesc = idesc0 | ((uint32_t)((jhi - jlo + 1) * NP >> 3) << 17);
                const uint32_t b_lo0 = w_lo + ((uint32_t)(brow * C::ROWB) >> 4);
                if (C::KHS == 3) {
                    // one stage = one plane (all three kh tiles): the issuers take the planes in rotation and issue all
                    // 18 MMAs of theirs, so the per-stage bookkeeping is paid once per 18 MMAs instead of once per 6
                    if (rot == my_kh) {
                        wait_bar(full_bar(s), ph);
                        ptx::tc_fence_after();
                        if (ptx::elect_one_sync()) {
                            if (!(g.dbg & 4)) {
#pragma unroll
                                for (int kh = 0; kh < 3; ++kh) {
                                    const uint32_t a_lo0 = ring_lo + (uint32_t)s * (C::STAGE_BYTES >> 4) + (uint32_t)((kh * C::A_BYTES) >> 4);
                                    const uint32_t b_kh = w_base + (uint32_t)((kh * 3 * C::W_TILE + brow * C::ROWB) >> 4);
#pragma unroll
                                    for (int kw = 0; kw < 3; ++kw) {
                                        if (kw > 0 && g.kwfold) break;       // kw-fold: the three kw taps are columns of ONE MMA
#pragma unroll
                                        for (int k = 0; k < KC / 16; ++k)
                                            ptx::umma_bf16_lohi(d_lo, a_lo0 + ((kw * C::ROWB + k * 32) >> 4),
                                                                b_kh + ((kw * C::W_TILE + k * 32) >> 4), desc_hi, idesc, 1u);
                                    }
                                }
                            }
                            ptx::umma_commit(empty_bar(s));
                        }
                        __syncwarp();
                    }
                    if (++rot == 3) rot = 0;
                    if (++s == C::STAGES) { s = 0; ph ^= 1u; }
                } else {
#pragma unroll
                    for (int khs = 0; khs < 3; ++khs) {
                        if (khs == my_kh) {
                            wait_bar(FUSED ? ready_bar(s) : full_bar(s), ph);
                            if (FUSED) ptx::fence_proxy_async();   // the masking warps zeroed rows with generic stores
                            ptx::tc_fence_after();
                            if (ptx::elect_one_sync()) {
                                const uint32_t a_lo0 = ring_lo + (uint32_t)s * (C::STAGE_BYTES >> 4) + a_tile;
                                if (!(g.dbg & 4)) {
                                    if (FUSED) {
                                        // K steps 0,1: first-half tile; 2,3: second-half tile (64-byte rows, SWIZZLE_64B);
                                        // the weight rows stay 128 bytes (SWIZZLE_128B): two different descriptor high words
                                        const uint64_t adsc = ptx::make_kmajor_desc(0u, 64, 0u);
                                        const uint64_t bdsc = ((uint64_t)desc_hi << 32);
                                        const uint32_t a_base = (uint32_t)adsc | ((ring + (uint32_t)s * C::STAGE_BYTES) >> 4);
#pragma unroll
                                        for (int kw = 0; kw < 3; ++kw) {
#pragma unroll
                                            for (int k = 0; k < 4; ++k) {
                                                const uint32_t a_lo = a_base + (uint32_t)(((k >> 1) * HALF_BYTES + kw * 64 + (k & 1) * 32) >> 4);
                                                ptx::umma_bf16(d_lo, (adsc & 0xffffffff00000000ull) | a_lo,
                                                               bdsc | (uint64_t)(b_lo0 + ((kw * C::W_TILE + k * 32) >> 4)), idesc, 1u);
                                            }
                                        }
                                    } else {
#pragma unroll
                                    for (int kw = 0; kw < 3; ++kw) {
#pragma unroll
                                        for (int k = 0; k < KC / 16; ++k)
                                            ptx::umma_bf16_lohi(d_lo, a_lo0 + ((kw * C::ROWB + k * 32) >> 4),
                                                                b_lo0 + ((kw * C::W_TILE + k * 32) >> 4), desc_hi, idesc, 1u);
                                    }
                                    }
                                }
                                ptx::umma_commit(empty_bar(s));
                            }
                            __syncwarp();
                        }
                        if (++s == C::STAGES) { s = 0; ph ^= 1u; }
                    }
                }
            }
            // all MMAs this warp issued for the item are in flight: its share of "accumulators complete"
            if (ptx::elect_one_sync()) ptx::umma_commit(tfull_bar(acc));
            __syncwarp();
        }
    } else if (FUSED && warp >= 12) {
        // ================= masking warps: zero what the reference leaves zero, then release the stage =================
        const int pt = tid - 12 * 32;                            // row pt of the tile (and row 128 + pt for pt < 2)
        const int sgn = (g.vol_mode == DSM_VOL_GC_RIGHT) ? 1 : -1;
        const bool mask_first = (g.vol_mode == DSM_VOL_PSM);
        int s = 0; uint32_t ph = 0;
        for (int t = item0; t < g.nitems; t += item_step) {
            const RsItem item = rs_decode(g, t, C::R);
            // padded position of this thread's rows for kh = 0 (the three tap rows differ by whole image rows)
            int hp0[2], wp0[2];
#pragma unroll
            for (int u = 0; u < 2; ++u) {
                const int qq = item.tile * 128 - Wp - 1 + pt + 128 * u + 2 * Wp;      // >= 0
                hp0[u] = qq / Wp - 2; wp0[u] = qq - (hp0[u] + 2) * Wp;
            }
            for (int i = -1; i <= item.nb; ++i) {
                const int zp = item.z0 + 1 + i;
                if (zp <= 0 || zp >= Dp - 1) continue;
                const int d = zp - 1;
#pragma unroll 1
                for (int kh = 0; kh < 3; ++kh) {
                    wait_bar(full_bar(s), ph);                   // both half tiles have landed
                    const uint32_t tile0 = ring + s * C::STAGE_BYTES;
#pragma unroll
                    for (int u = 0; u < 2; ++u) {
                        const int r = pt + 128 * u;
                        if (r < C::A_ROWS) {
                            const int hp = hp0[u] + kh, wp = wp0[u], x = wp - 1;
                            const bool interior = hp >= 1 && hp <= g.H && wp >= 1 && wp <= g.W;
                            const bool zero2 = !interior || (sgn < 0 ? x < d : x >= g.W - d);
                            const bool zero1 = mask_first && interior && x < d;
                            const uint32_t z = 0u;
                            if (zero1) {
                                const uint32_t a = tile0 + (uint32_t)r * 64u;
#pragma unroll
                                for (int c = 0; c < 4; ++c)
                                    asm volatile("st.shared.v4.b32 [%0], {%1, %1, %1, %1};" :: "r"(a + 16u * c), "r"(z) : "memory");
                            }
                            if (zero2) {
                                const uint32_t a = tile0 + HALF_BYTES + (uint32_t)r * 64u;
#pragma unroll
                                for (int c = 0; c < 4; ++c)
                                    asm volatile("st.shared.v4.b32 [%0], {%1, %1, %1, %1};" :: "r"(a + 16u * c), "r"(z) : "memory");
                            }
                        }
                    }
                    __syncwarp();
                    if (lane == 0) ptx::mbar_arrive(ready_bar(s));   // release: ordered before the issuer's acquire + proxy fence
                    if (++s == C::STAGES) { s = 0; ph ^= 1u; }
                }
            }
        }
    } else {
        // ================= epilogue: 8 warps, two per TMEM lane quadrant =================
        // warp w reads lanes 32*(w%4)..+31 (hardware rule); the two warps of a quadrant take the even / odd
        // output planes of the band.  Residual rows are fetched BEFORE waiting for the accumulator so
        // that their DRAM latency hides behind the MMAs of this item.
        ptx::griddep_wait();
        const int q = warp & 3;
        const int half = (warp - 4) >> 2;
        const int r = q * 32 + lane;
        constexpr int JJ = C::R / 2;                              // planes per warp
        {   // both accumulator buffers start zero-filled; arriving completes phase 0 of their "free" barriers
#pragma unroll
            for (int a2 = 0; a2 < 2; ++a2) {
#pragma unroll
                for (int jj = 0; jj < JJ; ++jj) {
                    const uint32_t ta = tmem + ((uint32_t)(q * 32) << 16) + a2 * C::ACC_COLS + (2 * jj + half) * NP;
                    if (NP == 32) ptx::tmem_zero32(ta); else ptx::tmem_zero16(ta);
                }
            }
            ptx::tc_wait_st();
            ptx::tc_fence_before();
            __syncwarp();
            if (lane == 0) { ptx::mbar_arrive(tempty_bar(0)); ptx::mbar_arrive(tempty_bar(1)); }
        }
        constexpr int NV = NP / 8;                                // uint4 per bf16 output row
        const float4* sc4 = reinterpret_cast<const float4*>(s_scale);
        const int emode = epi_mode(g.relu, residual != nullptr);
        const float4* sh4 = reinterpret_cast<const float4*>(s_shift);
        int tcount = 0;
        for (int t = item0; t < g.nitems; t += item_step) {
            const RsItem item = rs_decode(g, t, C::R);
            const bool fold = g.kwfold != 0;
            const int pq = fold ? item.tile * 120 + q * 30 - 1 + lane : item.tile * 128 + r;   // position in the padded (h,w) plane
            const int hp = pq / Wp, wp = pq - hp * Wp;
            const bool valid = (pq >= 0) && (pq < plane) && hp >= 1 && hp <= g.Ho && wp >= 1 && wp <= g.Wo &&
                               (!fold || (lane >= 1 && lane <= 30));
            const int acc = tcount & 1;
            const uint32_t acc_ph = (uint32_t)(tcount >> 1) & 1u;
            ++tcount;
            const uint32_t taddr0 = tmem + ((uint32_t)(q * 32) << 16) + acc * C::ACC_COLS;
            int my_last = -1;                                     // my last plane of this band (-1: none)
            for (int jj = 0; jj < JJ; ++jj) if (2 * jj + half < item.nb) my_last = 2 * jj + half;
            if (g.dbg & 1) {
                wait_bar(tfull_bar(acc), acc_ph);
                __syncwarp();
                if (lane == 0) ptx::mbar_arrive(tempty_bar(acc));
                continue;
            }
            // hand the buffer back: my blocks are in registers and zero-filled again
            auto release = [&]() {
                ptx::tc_wait_st();
                ptx::tc_fence_before();
                __syncwarp();
                if (lane == 0) ptx::mbar_arrive(tempty_bar(acc));
            };

            if (g.y_f32) {
                // single output channel: fp32, unpadded [B][Do][Ho][Wo]
                const size_t o0 = (((size_t)item.b * g.Do + item.z0) * g.Ho + (hp - 1)) * g.Wo + (wp - 1);
                const size_t ostep = (size_t)g.Ho * g.Wo;
                float rf[JJ];
#pragma unroll
                for (int jj = 0; jj < JJ; ++jj) {
                    const int j = 2 * jj + half;
                    rf[jj] = (residual && valid && j < item.nb) ? __ldg(reinterpret_cast<const float*>(residual) + o0 + j * ostep) : 0.f;
                }
                wait_bar(tfull_bar(acc), acc_ph);
                __syncwarp();
                ptx::tc_fence_after();
                if (my_last < 0) release();
#pragma unroll
                for (int jj = 0; jj < JJ; ++jj) {
                    const int j = 2 * jj + half;
                    if (j < item.nb) {
                        uint32_t v[16];
                        ptx::tmem_ld16(taddr0 + j * NP, v);
                        ptx::tc_wait_ld();
                        consume_tmem_load(v[0], scratch_smem);
                        ptx::tmem_zero16(taddr0 + j * NP);
                        if (j == my_last) release();
                        float acc0 = __uint_as_float(v[0]);
                        if (fold) {                                  // out[p] = Z[p-1][kw 0] + Z[p][kw 1] + Z[p+1][kw 2]
                            const float zl = __shfl_up_sync(0xffffffffu, __uint_as_float(v[0]), 1);
                            const float zr = __shfl_down_sync(0xffffffffu, __uint_as_float(v[2]), 1);
                            acc0 = (zl + __uint_as_float(v[1])) + zr;
                        }
                        if (valid) {
                            reinterpret_cast<float*>(y)[o0 + j * ostep] =
                                fuse_act(fmaf(acc0, s_scale[0], s_shift[0]), rf[jj], g.relu);
                        }
                    }
                }
            } else {
                const size_t ostep = (size_t)(g.Ho + 2) * (g.Wo + 2) * g.Cout;     // one output plane
                const size_t o0 = ((((size_t)item.b * (g.Do + 2) + item.z0 + 1) * (g.Ho + 2) + hp) * (g.Wo + 2) + wp) * (size_t)g.Cout + ch0;
                const __nv_bfloat16* resb = reinterpret_cast<const __nv_bfloat16*>(residual);
                uint4 rv[JJ][NV];
#pragma unroll
                for (int jj = 0; jj < JJ; ++jj) {
                    const int j = 2 * jj + half;
                    const bool ld = residual && valid && j < item.nb;
#pragma unroll
                    for (int c = 0; c < NV; c += 2) {
                        rv[jj][c] = rv[jj][c + 1] = make_uint4(0u, 0u, 0u, 0u);
                        if (ld) ld_nc_v8(resb + o0 + j * ostep + 8 * c, rv[jj][c], rv[jj][c + 1]);
                    }
                }
                wait_bar(tfull_bar(acc), acc_ph);
                __syncwarp();
                ptx::tc_fence_after();
                if (my_last < 0) release();
#pragma unroll
                for (int jj = 0; jj < JJ; ++jj) {
                    const int j = 2 * jj + half;
                    if (j < item.nb) {
                        uint32_t v[NP];
                        if (NP == 32) ptx::tmem_ld32(taddr0 + j * NP, v); else ptx::tmem_ld16(taddr0 + j * NP, v);
                        ptx::tc_wait_ld();
                        consume_tmem_load(v[0], scratch_smem);
                        if (NP == 32) ptx::tmem_zero32(taddr0 + j * NP); else ptx::tmem_zero16(taddr0 + j * NP);
                        if (j == my_last) release();
                        if (valid) {
                            uint4* out = reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(y) + o0 + j * ostep);
                            DSM_EPI_DISPATCH(emode, epi_store256, NV, v, sc4, sh4, rv[jj], out)   // one 256-bit store per 16 channels
                        }
                    }
                }
            }
        }
    }

    ptx::tc_fence_before();
    __syncthreads();
    if (warp == 1) ptx::tmem_dealloc(tmem, C::TMEM_COLS);
    if (g_conv_progress && blockIdx.x == 0 && tid == 0) {      // diagnostic: SM cycles and wall ns of CTA 0 -> MHz
        g_conv_progress[0] = (int)((clock64() - dbg_c0) >> 4);
        g_conv_progress[1] = (int)((globaltimer_ns() - dbg_t0) >> 4);
    }
}

// ---------------------------------------------------------------------------------------------
// Class-sharing kernel for transposed convolutions (k3, s2, p1, output_padding 1) with Cout <= 32.
//
// out[2j+p] along each axis takes (p=0) tap k=1 of input j, or (p=1) tap k=2 of input j and tap k=0
// of input j+1, so the 8 output parity classes are 8 small stride-1 convolutions over the SAME input
// voxels with 1/2/4/8 taps (27 (class,tap) pairs in all).  The per-class kernel above runs them as
// 27 N=Cout MMAs per K step, each re-reading a 4 KB A slab.  Here a CTA owns one 128-voxel input tile
// and ALL 8 classes (8 x NP accumulator columns, two buffers): the input tile at offset
// (od,oh,ow) in {0,1}^3 is the A operand of every class with p >= o component-wise, and with the
// classes laid out in Gray-code order (000,001,011,010,110,111,101,100) those sets are 1 or 2
// contiguous column runs, so the 27 pairs become 10 MMAs per K step with N = 256,128,128,64,64,
// 64,64,32,32,32 -> 568 cycles instead of 27 x 40.  The four (od,oh) tiles stream through TMA (ow is
// a row-shifted descriptor); all 27 weight blocks stay resident in shared memory in MMA order.
// ---------------------------------------------------------------------------------------------
struct DcOp { short tile, ow, brow, n, dcol, pad; };          // one MMA (per K step)

struct DcGeom {
    int B, D, H, W;          // INPUT extent
    int Do, Ho, Wo;          // output buffer extent (<= 2D, 2H, 2W)
    int Cout, relu, y_f32;
    int ntiles;              // ceil(P / 128), P = B*Dp*Hp*Wp
    long long P;
    int nops;
    short op_begin[6];       // MMAs of tile a are ops[op_begin[a] .. op_begin[a+1])
    DcOp ops[12];
    short w_dst[27], w_src[27];   // weight block -> first smem row, first packed-weight row (of channel group 0)
    int ngroups;                  // output-channel groups of NP channels (Cout = ngroups * NP)
    int dbg;                 // timing experiments only (variant bits 4..6): 1 = no epilogue global traffic, 2 = no TMA traffic, 4 = no MMAs
};

__constant__ const int kDcPosClass[8] = {0, 1, 3, 2, 6, 7, 5, 4};   // accumulator block -> parity class (pd<<2|ph<<1|pw)
__constant__ const int kDcClassPos[8] = {0, 1, 3, 2, 7, 6, 4, 5};   // parity class -> accumulator block

template <int KC, int NP>
struct DcCfg {
    static constexpr int ROWB = KC * 2;
    static constexpr int A_ROWS = 130;
    static constexpr int A_BYTES = ((A_ROWS * ROWB + 1023) / 1024) * 1024;
    static constexpr int W_BYTES = 27 * NP * ROWB;
    static constexpr int BAR_BYTES = 512;
    static constexpr int BUDGET = 225 * 1024 - W_BYTES - 1024 - BAR_BYTES - 2 * NP * 4;
    static constexpr int S_RAW = BUDGET / A_BYTES;
    static constexpr int STAGES = S_RAW > 8 ? 8 : S_RAW;
    static constexpr int TX_BYTES = A_ROWS * ROWB;
    static constexpr int ACC_COLS = 8 * NP;
    static constexpr int TMEM_COLS = 2 * ACC_COLS;
    static constexpr int SMEM = W_BYTES + STAGES * A_BYTES + 1024 + BAR_BYTES + 2 * NP * 4;
    static constexpr int THREADS = 576;                          // warp 0 TMA, warp 1 MMA, warps 2-17 epilogue
    static_assert(STAGES >= 4, "activation ring too shallow");
    static_assert(((NP * ROWB) % 1024) == 0, "weight blocks must keep the swizzle phase");
};

// tiles that lie completely inside a rim plane (d' = 0 or D+1) read only zeros and store nothing
__device__ __forceinline__ bool dc_skip(long long p0, long long P, long long plane, long long vol, int Dp) {
    return rim_tile(p0, P, plane, vol, Dp);
}

template <int KC, int NP>
__global__ void __launch_bounds__(576, 1)
conv3d_dc_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_w,
                 const __grid_constant__ DcGeom g, const float* __restrict__ scale, const float* __restrict__ shift,
                 const void* __restrict__ residual, void* __restrict__ y) {
    using C = DcCfg<KC, NP>;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int Dp = g.D + 2, Hp = g.H + 2, Wp = g.W + 2;
    const long long plane = (long long)Hp * Wp, vol = plane * Dp;
    // Cout > NP (hourglass conv5: 64 -> 64): output-channel groups of NP; a CTA keeps ONE group's 27 weight blocks resident
    // and walks the tiles with the CTAs of its group (the input tiles are read once per group, from L2)
    const int grp = (int)blockIdx.x % g.ngroups;
    const int tile0 = (int)blockIdx.x / g.ngroups, tile_step = (int)gridDim.x / g.ngroups;
    const int ch0 = grp * NP;

    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw = ptx::smem_u32(smem_raw);
    const uint32_t base = (raw + 1023u) & ~1023u;
    uint8_t* base_ptr = smem_raw + (base - raw);
    const uint32_t wsm = base;
    const uint32_t ring = base + C::W_BYTES;
    constexpr int RING_END = C::W_BYTES + C::STAGES * C::A_BYTES;
    const uint32_t bars = base + RING_END;
    auto full_bar = [&](int s) { return bars + 8u * s; };
    auto empty_bar = [&](int s) { return bars + 8u * (C::STAGES + s); };
    auto tfull_bar = [&](int a) { return bars + 8u * (2 * C::STAGES + a); };
    auto tempty_bar = [&](int a) { return bars + 8u * (2 * C::STAGES + 2 + a); };
    const uint32_t wfull_bar = bars + 8u * (2 * C::STAGES + 4);
    constexpr int NBARS = 2 * C::STAGES + 5;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(base_ptr + RING_END + 8 * NBARS);
    const uint32_t scratch_smem = bars + 8u * NBARS + 8u;
    float* s_scale = reinterpret_cast<float*>(base_ptr + RING_END + C::BAR_BYTES);
    float* s_shift = s_scale + NP;

    if (tid < NP) {
        s_scale[tid] = scale ? __ldg(scale + ch0 + tid) : 1.f;
        s_shift[tid] = shift ? __ldg(shift + ch0 + tid) : 0.f;
    }
    if (warp == 0 && lane == 0) {
        ptx::prefetch_tensormap(&map_w);
        ptx::prefetch_tensormap(&map_a);
        for (int s = 0; s < C::STAGES; ++s) { ptx::mbar_init(full_bar(s), 1); ptx::mbar_init(empty_bar(s), 1); }
        for (int a = 0; a < 2; ++a) { ptx::mbar_init(tfull_bar(a), 1); ptx::mbar_init(tempty_bar(a), g.y_f32 ? 4 : 16); }
        ptx::mbar_init(wfull_bar, 1);
        ptx::fence_mbar_init();
    }
    if (warp == 1) ptx::tmem_alloc(ptx::smem_u32(tmem_slot), C::TMEM_COLS);
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem = *tmem_slot;

    ptx::griddep_launch_dependents();
    if (warp == 0) {
        // ================= TMA producer =================
        if (ptx::elect_one_sync()) {
            ptx::mbar_arrive_expect_tx(wfull_bar, C::W_BYTES);
            for (int i = 0; i < 27; ++i)
                ptx::tma_load_2d(wsm + g.w_dst[i] * C::ROWB, &map_w, wfull_bar, 0, g.w_src[i] + ch0);
        }
        __syncwarp();
        ptx::griddep_wait();
        int s = 0; uint32_t ph = 0;
        for (int t = tile0; t < g.ntiles; t += tile_step) {
            const long long p0 = (long long)t * 128;
            if (dc_skip(p0, g.P, plane, vol, Dp)) continue;
#pragma unroll 1
            for (int a = 0; a < 4; ++a) {                        // (od, oh) = (a>>1, a&1)
                wait_bar(empty_bar(s), ph ^ 1u);
                if (ptx::elect_one_sync()) {
                    if (g.dbg & 2) {
                        ptx::mbar_arrive(full_bar(s));
                    } else {
                        ptx::mbar_arrive_expect_tx(full_bar(s), C::TX_BYTES);
                        ptx::tma_load_2d(ring + s * C::A_BYTES, &map_a, full_bar(s), 0, (int)(p0 + ((a >> 1) * Hp + (a & 1)) * Wp));
                    }
                }
                __syncwarp();
                if (++s == C::STAGES) { s = 0; ph ^= 1u; }
            }
        }
    } else if (warp == 1) {
        // ================= MMA issuer =================
        constexpr uint32_t idesc0 = ptx::make_idesc_bf16(0);
        wait_bar(wfull_bar, 0);
        ptx::tc_fence_after();
        const uint64_t dsc = ptx::make_kmajor_desc(0u, C::ROWB, 0u);
        const uint32_t desc_hi = (uint32_t)(dsc >> 32);
        const uint32_t ring_lo = (uint32_t)dsc | (ring >> 4);
        const uint32_t w_lo = (uint32_t)dsc | (wsm >> 4);
        int s = 0; uint32_t ph = 0;
        int tcount = 0;
        for (int t = tile0; t < g.ntiles; t += tile_step) {
            const long long p0 = (long long)t * 128;
            if (dc_skip(p0, g.P, plane, vol, Dp)) continue;
            const int acc = tcount & 1;
            const uint32_t acc_ph = (uint32_t)(tcount >> 1) & 1u;
            ++tcount;
            wait_bar(tempty_bar(acc), acc_ph ^ 1u);
            ptx::tc_fence_after();
            const uint32_t d_tmem = tmem + acc * C::ACC_COLS;
#pragma unroll 1
            for (int a = 0; a < 4; ++a) {
                wait_bar(full_bar(s), ph);
                ptx::tc_fence_after();
                if (ptx::elect_one_sync()) {
                    const uint32_t a_lo0 = ring_lo + (uint32_t)s * (C::A_BYTES >> 4);
                    for (int op = g.op_begin[a]; op < g.op_begin[a + 1] && !(g.dbg & 4); ++op) {
                        const DcOp o = g.ops[op];
                        const uint32_t a_lo = a_lo0 + (uint32_t)((o.ow * C::ROWB) >> 4);
                        const uint32_t b_lo = w_lo + (uint32_t)((o.brow * C::ROWB) >> 4);
                        const uint32_t idesc = idesc0 | ((uint32_t)(o.n >> 3) << 17);
#pragma unroll
                        for (int k = 0; k < KC / 16; ++k)        // the very first MMA (offset 000 covers all 8 classes) overwrites
                            ptx::umma_bf16_lohi(d_tmem + o.dcol, a_lo + ((k * 32) >> 4), b_lo + ((k * 32) >> 4), desc_hi, idesc,
                                                (op | k) ? 1u : 0u);
                    }
                    ptx::umma_commit(empty_bar(s));
                    if (a == 3) ptx::umma_commit(tfull_bar(acc));
                }
                __syncwarp();
                if (++s == C::STAGES) { s = 0; ph ^= 1u; }
            }
        }
    } else {
        // ================= epilogue: 16 warps, four per TMEM lane quadrant, 2 classes each =================
        // (the MMAs of a tile take ~2.3k cycles: the epilogue is the longer leg, so it gets the threads)
        ptx::griddep_wait();
        const int q = warp & 3;
        const int half = (warp - 2) >> 2;                       // 0..3: owns accumulator blocks 2*half, 2*half+1
        const int r = q * 32 + lane;
        constexpr int NV = NP / 8;
        if (g.y_f32) {
            // Single output channel (GC-Net l37), fp32 [B][Do][Ho][Wo]: a tile carries 8 floats per voxel, so the
            // per-tile coordinate arithmetic is the cost — ONE warp per lane quadrant does it (the other twelve
            // would only repeat it), reads the 8 class columns and writes the (pw = 0, 1) pairs as 8-byte stores
            // that are contiguous across the warp.
            if (half == 0) {
                int tcount = 0;
                const bool pair_ok = (g.Wo & 1) == 0;
                for (int t = tile0; t < g.ntiles; t += tile_step) {
                    const long long p0 = (long long)t * 128;
                    if (dc_skip(p0, g.P, plane, vol, Dp)) continue;
                    const long long p = p0 + r;
                    bool interior = false;
                    int ob = 0, dz = 0, hy = 0, wx = 0;
                    if (p < g.P) {
                        const FlatPos fp = flat_decode(p, plane, vol);
                        const int dp = fp.dp, rem2 = fp.off;
                        ob = fp.b;
                        const int hp = (int)((unsigned)rem2 / (unsigned)Wp), wp = rem2 - hp * Wp;
                        interior = dp >= 1 && dp <= g.D && hp >= 1 && hp <= g.H && wp >= 1 && wp <= g.W && !(g.dbg & 1);
                        dz = 2 * (dp - 1); hy = 2 * (hp - 1); wx = 2 * (wp - 1);
                    }
                    const int acc = tcount & 1;
                    const uint32_t acc_ph = (uint32_t)(tcount >> 1) & 1u;
                    ++tcount;
                    const uint32_t taddr0 = tmem + ((uint32_t)(q * 32) << 16) + acc * C::ACC_COLS;
                    wait_bar(tfull_bar(acc), acc_ph);
                    __syncwarp();
                    ptx::tc_fence_after();
                    uint32_t v[8];
#pragma unroll
                    for (int c = 0; c < 8; ++c) v[c] = ptx::tmem_ld1(taddr0 + kDcClassPos[c] * NP);     // v[c] = class c, channel 0
                    ptx::tc_wait_ld();
                    consume_tmem_load(v[0], scratch_smem);
                    ptx::tc_fence_before();
                    __syncwarp();
                    if (lane == 0) ptx::mbar_arrive(tempty_bar(acc));
                    if (!interior) continue;
                    const float sc0 = s_scale[0], sh0 = s_shift[0];
                    const float* resf = reinterpret_cast<const float*>(residual);
                    float* yf = reinterpret_cast<float*>(y);
#pragma unroll
                    for (int c2 = 0; c2 < 4; ++c2) {             // (pd, ph) = (c2 >> 1, c2 & 1)
                        const int od = dz + (c2 >> 1), oh = hy + (c2 & 1);
                        if (od >= g.Do || oh >= g.Ho || wx >= g.Wo) continue;
                        const size_t o = (((size_t)ob * g.Do + od) * g.Ho + oh) * g.Wo + wx;
                        const bool two = wx + 1 < g.Wo;
                        float r0 = 0.f, r1 = 0.f;
                        if (resf) { r0 = __ldg(resf + o); if (two) r1 = __ldg(resf + o + 1); }
                        const float a0 = fuse_act(fmaf(__uint_as_float(v[2 * c2]), sc0, sh0), r0, g.relu);
                        const float a1 = fuse_act(fmaf(__uint_as_float(v[2 * c2 + 1]), sc0, sh0), r1, g.relu);
                        if (two && pair_ok) *reinterpret_cast<float2*>(yf + o) = make_float2(a0, a1);
                        else { yf[o] = a0; if (two) yf[o + 1] = a1; }
                    }
                }
            }
        } else {
        const float4* sc4 = reinterpret_cast<const float4*>(s_scale);
        const int emode = epi_mode(g.relu, residual != nullptr);
        const float4* sh4 = reinterpret_cast<const float4*>(s_shift);
        int tcount = 0;
        // where this thread's two output voxels (its two classes) of tile t live; false when neither exists
        auto locate = [&](int t, bool* valid, size_t* off) -> bool {
            const long long p = (long long)t * 128 + r;
            bool interior = false;
            int ob = 0, dz = 0, hy = 0, wx = 0;
            if (p < g.P) {
                const FlatPos fp = flat_decode(p, plane, vol);
                const int dp = fp.dp, rem2 = fp.off;
                const int hp = (int)((unsigned)rem2 / (unsigned)Wp), wp = rem2 - hp * Wp;
                interior = dp >= 1 && dp <= g.D && hp >= 1 && hp <= g.H && wp >= 1 && wp <= g.W;
                ob = fp.b; dz = 2 * (dp - 1); hy = 2 * (hp - 1); wx = 2 * (wp - 1);
            }
#pragma unroll
            for (int jj = 0; jj < 2; ++jj) {
                const int c = kDcPosClass[2 * half + jj];
                const int od = dz + ((c >> 2) & 1), oh = hy + ((c >> 1) & 1), ow = wx + (c & 1);
                valid[jj] = interior && od < g.Do && oh < g.Ho && ow < g.Wo && !(g.dbg & 1);
                off[jj] = g.y_f32 ? (((size_t)ob * g.Do + od) * g.Ho + oh) * g.Wo + ow
                                  : ((((size_t)ob * (g.Do + 2) + od + 1) * (g.Ho + 2) + oh + 1) * (g.Wo + 2) + ow + 1) * (size_t)g.Cout + ch0;
            }
            return interior;
        };
        for (int t = tile0; t < g.ntiles; t += tile_step) {
            const long long p0 = (long long)t * 128;
            if (dc_skip(p0, g.P, plane, vol, Dp)) continue;
            const int acc = tcount & 1;
            const uint32_t acc_ph = (uint32_t)(tcount >> 1) & 1u;
            ++tcount;
            const uint32_t taddr0 = tmem + ((uint32_t)(q * 32) << 16) + acc * C::ACC_COLS;
            bool valid[2]; size_t off[2];
            locate(t, valid, off);
            auto release = [&]() {
                ptx::tc_fence_before();
                __syncwarp();
                if (lane == 0) ptx::mbar_arrive(tempty_bar(acc));
            };
            {
                const __nv_bfloat16* resb = reinterpret_cast<const __nv_bfloat16*>(residual);
                uint4 rv[2][NV];
#pragma unroll
                for (int jj = 0; jj < 2; ++jj) {
                    const bool ld = residual && valid[jj];
#pragma unroll
                    for (int c = 0; c < NV; c += 2) {
                        rv[jj][c] = rv[jj][c + 1] = make_uint4(0u, 0u, 0u, 0u);
                        if (ld) ld_nc_v8(resb + off[jj] + 8 * c, rv[jj][c], rv[jj][c + 1]);
                    }
                }
                wait_bar(tfull_bar(acc), acc_ph);
                __syncwarp();
                ptx::tc_fence_after();
#pragma unroll
                for (int jj = 0; jj < 2; ++jj) {
                    uint32_t v[NP];
                    if (NP == 32) ptx::tmem_ld32(taddr0 + (2 * half + jj) * NP, v); else ptx::tmem_ld16(taddr0 + (2 * half + jj) * NP, v);
                    ptx::tc_wait_ld();
                    consume_tmem_load(v[0], scratch_smem);
                    if (jj == 1) release();
                    if (valid[jj]) {
                        uint4* out = reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(y) + off[jj]);
                        DSM_EPI_DISPATCH(emode, epi_store256, NV, v, sc4, sh4, rv[jj], out)   // one 256-bit store per 16 channels
                    }
                }
            }
        }
        }   // bf16 output
    }

    ptx::tc_fence_before();
    __syncthreads();
    if (warp == 1) ptx::tmem_dealloc(tmem, C::TMEM_COLS);
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
bool encode_map(CUtensorMap* m, const void* base, int rank, const cuuint64_t* dims, const cuuint64_t* strides_bytes,
                const cuuint32_t* box, int row_bytes) {
    return tma_host::encode(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, base, rank, dims, strides_bytes, box,
                            (row_bytes == 128) ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B);
}

// Kernel launch with the programmatic-stream-serialization attribute when `pdl` is set: the kernel's prologue
// (barrier init, TMEM allocation, resident weights) then overlaps the tail of the previous kernel in the stream;
// the kernels order their dependent accesses with griddepcontrol.wait.
thread_local bool g_launch_pdl = false;        // set per call from `variant` bit 7 (host side, per calling thread)

template <typename... KArgs, typename... Args>
void launch_kernel(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr; cfg.numAttrs = g_launch_pdl ? 1 : 0;
    cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}

template <int KC, int NP, int MODE, int WRES = 0>
int launch_cfg(const ConvMaps& maps, const ConvGeom& g, dim3 grid, const float* scale, const float* shift,
               const void* residual, void* y, cudaStream_t st) {
    using C = Cfg<KC, NP, MODE, WRES>;
    auto kern = conv3d_igemm_kernel<KC, NP, MODE, WRES>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM);
    if (e != cudaSuccess) return (int)e;
    int nsm = DSM_NUM_SMS_B200, dev = 0;
    if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, dev);
    int nblocks = nsm * C::CTAS_PER_SM;
    if (nblocks > (int)grid.x) nblocks = (int)grid.x;
    launch_kernel(kern, dim3(nblocks), dim3(C::THREADS), C::SMEM, st, maps, g, scale, shift, residual, y);
    return dsm_launch_status();
}

template <int MODE>
int launch_mode(int KC, int NP, const ConvMaps& maps, const ConvGeom& g, dim3 grid, const float* scale,
                const float* shift, const void* residual, void* y, cudaStream_t st) {
#define DSM_CASE(kc, np) if (KC == kc && NP == np) return launch_cfg<kc, np, MODE>(maps, g, grid, scale, shift, residual, y, st);
    DSM_CASE(32, 16) DSM_CASE(32, 32) DSM_CASE(32, 64) DSM_CASE(32, 128)
    DSM_CASE(64, 16) DSM_CASE(64, 32) DSM_CASE(64, 64) DSM_CASE(64, 128)
#undef DSM_CASE
    return DSM_EUNSUPPORTED;
}

template <int KC, int NP, bool FUSED = false>
int launch_rs(const CUtensorMap& map_a, const CUtensorMap& map_w, const RsGeom& g, const float* scale, const float* shift,
              const void* residual, void* y, cudaStream_t st, const CUtensorMap* map_a2 = nullptr) {
    using C = RsCfg<KC, NP>;
    auto kern = conv3d_rs_kernel<KC, NP, FUSED>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM);
    if (e != cudaSuccess) return (int)e;
    int nsm = DSM_NUM_SMS_B200, dev = 0;
    if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, dev);
    int per_group = nsm / g.ngroups;                            // persistent: one CTA per SM, split evenly over the channel groups
    if (per_group > g.nitems) per_group = g.nitems;
    const int nblocks = per_group * g.ngroups;
    launch_kernel(kern, dim3(nblocks), dim3(C::THREADS + (FUSED ? 128 : 0)), C::SMEM, st, map_a, map_a2 ? *map_a2 : map_a, map_w, g, scale, shift,
                  residual, y);
    return dsm_launch_status();
}

template <int KC, int NP>
int launch_dc(const CUtensorMap& map_a, const CUtensorMap& map_w, const DcGeom& g, const float* scale, const float* shift,
              const void* residual, void* y, cudaStream_t st) {
    using C = DcCfg<KC, NP>;
    auto kern = conv3d_dc_kernel<KC, NP>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM);
    if (e != cudaSuccess) return (int)e;
    int nsm = DSM_NUM_SMS_B200, dev = 0;
    if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, dev);
    int per_group = nsm / g.ngroups;
    if (per_group > g.ntiles) per_group = g.ntiles;
    const int nblocks = per_group * g.ngroups;
    launch_kernel(kern, dim3(nblocks), dim3(C::THREADS), C::SMEM, st, map_a, map_w, g, scale, shift, residual, y);
    return dsm_launch_status();
}

// MMA schedule of the class-sharing transposed conv: for every input offset o in {0,1}^3 (grouped by the
// (od,oh) tile that is loaded, ow is a descriptor shift) the accumulator blocks (Gray-code class order) whose
// class needs that offset, split into contiguous runs; weight blocks are laid out in the same order.
void build_dc_schedule(DcGeom& g, int NP, int CoutTotal) {
    static const int pos_class[8] = {0, 1, 3, 2, 6, 7, 5, 4};
    int nops = 0, nblk = 0;
    for (int a = 0; a < 4; ++a) for (int ow = 0; ow < 2; ++ow) {
        const int od = a >> 1, oh = a & 1;
        if (ow == 0) g.op_begin[a] = (short)nops;
        int pos = 0;
        while (pos < 8) {
            auto uses = [&](int ps) {
                const int c = pos_class[ps];
                return ((c >> 2) & 1) >= od && ((c >> 1) & 1) >= oh && (c & 1) >= ow;
            };
            if (!uses(pos)) { ++pos; continue; }
            int end = pos;
            while (end + 1 < 8 && uses(end + 1)) ++end;
            DcOp& op = g.ops[nops++];
            op.tile = (short)a; op.ow = (short)ow; op.brow = (short)(nblk * NP); op.n = (short)((end - pos + 1) * NP);
            op.dcol = (short)(pos * NP); op.pad = 0;
            for (int ps = pos; ps <= end; ++ps) {
                const int c = pos_class[ps];
                // per axis: parity 0 -> tap 1 (offset 0); parity 1 -> tap 0 at offset 1, tap 2 at offset 0
                auto tap = [](int par, int off) { return par == 0 ? 1 : (off == 1 ? 0 : 2); };
                const int kd = tap((c >> 2) & 1, od), kh = tap((c >> 1) & 1, oh), kw = tap(c & 1, ow);
                g.w_dst[nblk] = (short)(nblk * NP);
                g.w_src[nblk] = (short)(((kd * 3 + kh) * 3 + kw) * CoutTotal);
                ++nblk;
            }
            pos = end + 1;
        }
    }
    g.nops = nops;      // 10 MMAs, 27 weight blocks
    g.op_begin[4] = (short)nops;
}

int conv3d_dispatch(const void* x, const void* w, const float* scale, const float* shift, const void* residual, void* y,
                    int B, int Cin, int Cout, int D, int H, int W, int stride, int transposed, int relu, int y_dtype,
                    int Do, int Ho, int Wo, int variant, void* stream) {
    if (!x || !w || !y || B <= 0 || Cin <= 0 || Cout <= 0 || D <= 0 || H <= 0 || W <= 0) return DSM_EINVAL;
    if (stride != 1 && stride != 2) return DSM_EINVAL;
    if (relu < 0 || relu > 2) return DSM_EINVAL;
    g_launch_pdl = (variant & 128) != 0;
    if (transposed && stride != 2) return DSM_EUNSUPPORTED;
    if (y_dtype != DSM_BF16 && y_dtype != DSM_F32) return DSM_EINVAL;
    if (Cin != 32 && Cin != 64 && Cin != 128) return DSM_EUNSUPPORTED;
    const int KC = (Cin == 32) ? 32 : 64;
    int NP;
    if (y_dtype == DSM_F32) { if (Cout != 1) return DSM_EUNSUPPORTED; NP = 16; }
    else { if (Cout != 16 && Cout != 32 && Cout != 64 && Cout != 128) return DSM_EUNSUPPORTED; NP = Cout; }
    if (!dsm_aligned16(x) || !dsm_aligned16(w) || !dsm_aligned16(y) || (residual && !dsm_aligned16(residual))) return DSM_EALIGN;
    const bool al32 = dsm_aligned32(y) && (!residual || dsm_aligned32(residual));     // the RS / DC epilogues use 256-bit accesses
    // natural output extent
    int nDo, nHo, nWo;
    if (transposed) { nDo = 2 * D; nHo = 2 * H; nWo = 2 * W; }
    else if (stride == 2) { nDo = (D - 1) / 2 + 1; nHo = (H - 1) / 2 + 1; nWo = (W - 1) / 2 + 1; }
    else { nDo = D; nHo = H; nWo = W; }
    if (Do <= 0) { Do = nDo; Ho = nHo; Wo = nWo; }
    if (Do > nDo || Ho > nHo || Wo > nWo) return DSM_EINVAL;

    const int Dp = D + 2, Hp = H + 2, Wp = W + 2;
    const long long P = (long long)B * Dp * Hp * Wp;
    if (P > 0x7fffff00LL) return DSM_EUNSUPPORTED;

    ConvMaps maps;
    ConvGeom g;
    memset(&g, 0, sizeof(g));
    g.B = B; g.Di = D; g.Hi = H; g.Wi = W; g.Do = Do; g.Ho = Ho; g.Wo = Wo;
    g.Cout = Cout; g.transposed = transposed; g.relu = relu; g.y_f32 = (y_dtype == DSM_F32);
    g.nchunks = Cin / KC; g.P = P; g.desc_variant = variant & 1;
    g.dpad = 1; g.ri = 1; g.ro = 1; g.dil = 1; g.ldy = Cout; g.ldr = Cout; g.y_mode = g.y_f32 ? 1 : 0;
    const int row_bytes = KC * 2;
    // bit1 of `variant` SET selects the per-tap kernel (MODE_FLAT); the default for stride-1 convolutions is the
    // row-shifted-descriptor kernel (validated bit-identical on B200), except N=128 whose stage would not fit twice
    const int mode = (!transposed && stride == 2) ? MODE_BOX : ((!(variant & 2) && !transposed && NP <= 64 && !(KC == 64 && NP == 64 && !(variant & 4))) ? MODE_SHIFT : MODE_FLAT);

    {   // weights: [27*NP][Cin]
        cuuint64_t dims[2] = {(cuuint64_t)Cin, (cuuint64_t)27 * NP};
        cuuint64_t strides[1] = {(cuuint64_t)Cin * 2};
        cuuint32_t box[2] = {(cuuint32_t)KC, (cuuint32_t)NP};
        if (!encode_map(&maps.w, w, 2, dims, strides, box, row_bytes)) return DSM_EDRIVER;
    }
    // transposed, Cout <= 32: the class-sharing kernel (variant bit3 set = keep the per-class kernel, for A/B runs)
    // Cout = 64 (hourglass conv5) runs as two 32-channel groups on the same kernel (variant bit 2 keeps the per-class kernel)
    const bool dc_grouped = transposed && y_dtype == DSM_BF16 && Cout == 64 && !(variant & 4);
    if (transposed && (NP <= 32 || dc_grouped) && Cin <= 64 && al32 && !(variant & 8)) {
        const int NPd = dc_grouped ? 32 : NP;
        CUtensorMap map_a;
        cuuint64_t dims[2] = {(cuuint64_t)Cin, (cuuint64_t)P};
        cuuint64_t strides[1] = {(cuuint64_t)Cin * 2};
        cuuint32_t box[2] = {(cuuint32_t)KC, 130u};
        if (!encode_map(&map_a, x, 2, dims, strides, box, row_bytes)) return DSM_EDRIVER;
        DcGeom dg;
        memset(&dg, 0, sizeof(dg));
        dg.B = B; dg.D = D; dg.H = H; dg.W = W; dg.Do = Do; dg.Ho = Ho; dg.Wo = Wo;
        dg.Cout = Cout; dg.relu = relu; dg.y_f32 = (y_dtype == DSM_F32);
        dg.P = P; dg.ntiles = (int)dsm_ceil_div_ll(P, 128);
        dg.dbg = (variant >> 4) & 7;
        dg.ngroups = dc_grouped ? Cout / 32 : 1;
        build_dc_schedule(dg, NPd, NP);                          // packed weights are [27][NP = CoutP][Cin]
        CUtensorMap map_wd = maps.w;
        if (dc_grouped) {                                        // weight boxes of 32 rows instead of CoutP
            cuuint64_t wdims[2] = {(cuuint64_t)Cin, (cuuint64_t)27 * NP};
            cuuint64_t wstrides[1] = {(cuuint64_t)Cin * 2};
            cuuint32_t wbox[2] = {(cuuint32_t)KC, 32u};
            if (!encode_map(&map_wd, w, 2, wdims, wstrides, wbox, row_bytes)) return DSM_EDRIVER;
        }
        cudaStream_t st = (cudaStream_t)stream;
        if (KC == 32 && NPd == 16) return launch_dc<32, 16>(map_a, map_wd, dg, scale, shift, residual, y, st);
        if (KC == 32 && NPd == 32) return launch_dc<32, 32>(map_a, map_wd, dg, scale, shift, residual, y, st);
        if (KC == 64 && NPd == 16) return launch_dc<64, 16>(map_a, map_wd, dg, scale, shift, residual, y, st);
        if (KC == 64 && NPd == 32) return launch_dc<64, 32>(map_a, map_wd, dg, scale, shift, residual, y, st);
        return DSM_EUNSUPPORTED;
    }
    // stride-1, Cout <= 32: the plane-sharing kernel (variant bit3 set = keep the per-tile kernels, for A/B runs)
    // 64->64: two groups of 32 channels — only when there is enough work to fill the machine twice over (each CTA pays a
    // 110 KB weight prologue; the 12x24x78 hourglass bottom is faster on the per-tile kernel)
    const bool rs_grouped = (y_dtype == DSM_BF16 && Cout > 32 && Cout % 32 == 0 && Cout <= 128) &&
                            (long long)B * dsm_ceil_div(Do > 0 ? Do : D, 8) * dsm_ceil_div(Hp * Wp, 128) * (Cout / 32) >= 2LL * DSM_NUM_SMS_B200;
    if (!transposed && stride == 1 && (NP <= 32 || rs_grouped) && Cin <= 64 && al32 && !(variant & 8)) {
        const int NPk = rs_grouped ? 32 : NP;
        CUtensorMap map_a;
        cuuint64_t dims[2] = {(cuuint64_t)Cin, (cuuint64_t)P};
        cuuint64_t strides[1] = {(cuuint64_t)Cin * 2};
        cuuint32_t box[2] = {(cuuint32_t)KC, 130u};
        if (!encode_map(&map_a, x, 2, dims, strides, box, row_bytes)) return DSM_EDRIVER;
        RsGeom rg;
        memset(&rg, 0, sizeof(rg));
        rg.B = B; rg.D = D; rg.H = H; rg.W = W; rg.Do = Do; rg.Ho = Ho; rg.Wo = Wo;
        rg.Cout = Cout; rg.relu = relu; rg.y_f32 = (y_dtype == DSM_F32);
        // variant bit 8: the weights of this single-output-channel layer are packed with kw folded into the output columns
        // (dsm_pack_weight mode 3): one MMA per (kh, K step), tiles of 120 positions (see conv3d_rs_kernel)
        rg.kwfold = ((variant & 256) && KC == 32 && y_dtype == DSM_F32 && Cout == 1) ? 1 : 0;
        if ((variant & 256) && !rg.kwfold) return DSM_EINVAL;
        rg.plane_tiles = dsm_ceil_div(Hp * Wp, rg.kwfold ? 120 : 128);
        rg.nbands = dsm_ceil_div(Do, 8);
        const long long ni = (long long)B * rg.nbands * rg.plane_tiles;
        if (ni > 0x7fffffffLL) return DSM_EUNSUPPORTED;
        rg.nitems = (int)ni;
        rg.dbg = (variant >> 4) & 7;
        rg.ngroups = rs_grouped ? Cout / 32 : 1;
        for (int t = 0; t < 27; ++t) rg.w_row[t] = t * NP;          // packed weights are [27][NP = CoutP][Cin]
        CUtensorMap map_w = maps.w;
        if (rs_grouped) {                                            // weight boxes of 32 rows instead of CoutP
            cuuint64_t wdims[2] = {(cuuint64_t)Cin, (cuuint64_t)27 * NP};
            cuuint64_t wstrides[1] = {(cuuint64_t)Cin * 2};
            cuuint32_t wbox[2] = {(cuuint32_t)KC, 32u};
            if (!encode_map(&map_w, w, 2, wdims, wstrides, wbox, row_bytes)) return DSM_EDRIVER;
        }
        cudaStream_t st = (cudaStream_t)stream;
        if (rg.kwfold) {
            CUtensorMap map_q;                                       // the same activation tensor, boxes of 32 rows
            cuuint32_t qbox[2] = {(cuuint32_t)KC, 32u};
            if (!encode_map(&map_q, x, 2, dims, strides, qbox, row_bytes)) return DSM_EDRIVER;
            return launch_rs<32, 16>(map_a, map_w, rg, scale, shift, residual, y, st, &map_q);
        }
        if (KC == 32 && NPk == 16) return launch_rs<32, 16>(map_a, map_w, rg, scale, shift, residual, y, st);
        if (KC == 32 && NPk == 32) return launch_rs<32, 32>(map_a, map_w, rg, scale, shift, residual, y, st);
        if (KC == 64 && NPk == 16) return launch_rs<64, 16>(map_a, map_w, rg, scale, shift, residual, y, st);
        if (KC == 64 && NPk == 32) return launch_rs<64, 32>(map_a, map_w, rg, scale, shift, residual, y, st);
        return DSM_EUNSUPPORTED;
    }
    dim3 grid;
    if (mode == MODE_BOX) {
        for (int q = 0; q < 8; ++q) {
            const int pd = (q >> 2) & 1, ph = (q >> 1) & 1, pw = q & 1;
            const char* basep = reinterpret_cast<const char*>(x) + (((size_t)pd * Hp + ph) * Wp + pw) * (size_t)Cin * 2;
            cuuint64_t dims[5] = {(cuuint64_t)Cin, (cuuint64_t)((Wp - pw + 1) / 2), (cuuint64_t)((Hp - ph + 1) / 2),
                                  (cuuint64_t)((Dp - pd + 1) / 2), (cuuint64_t)B};
            cuuint64_t strides[4] = {(cuuint64_t)2 * Cin * 2, (cuuint64_t)2 * Wp * Cin * 2,
                                     (cuuint64_t)2 * Hp * Wp * Cin * 2, (cuuint64_t)Dp * Hp * Wp * Cin * 2};
            cuuint32_t box[5] = {(cuuint32_t)KC, 16, 8, 1, 1};
            if (!encode_map(&maps.a[q], basep, 5, dims, strides, box, row_bytes)) return DSM_EDRIVER;
        }
        g.cls_begin[0] = 0; g.cls_begin[1] = 27;
        for (int kd = 0; kd < 3; ++kd) for (int kh = 0; kh < 3; ++kh) for (int kw = 0; kw < 3; ++kw) {
            const int t = (kd * 3 + kh) * 3 + kw;
            const int q = ((kd & 1) << 2) | ((kh & 1) << 1) | (kw & 1);
            g.a_off[t] = q | ((kw >> 1) << 4) | ((kh >> 1) << 5) | ((kd >> 1) << 6);
            g.w_row[t] = t * NP;
        }
        g.tiles_w = dsm_ceil_div(Wo, 16); g.tiles_h = dsm_ceil_div(Ho, 8);
        const long long nt = (long long)B * Do * g.tiles_h * g.tiles_w;
        if (nt > 0x7fffffffLL) return DSM_EUNSUPPORTED;
        g.ntiles = (int)nt; g.mtiles = (int)nt;
        grid = dim3((unsigned)nt, 1, 1);
    } else {
        cuuint64_t dims[2] = {(cuuint64_t)Cin, (cuuint64_t)P};
        cuuint64_t strides[1] = {(cuuint64_t)Cin * 2};
        cuuint32_t box[2] = {(cuuint32_t)KC, (cuuint32_t)(mode == MODE_SHIFT ? 130 : 128)};
        if (!encode_map(&maps.a[0], x, 2, dims, strides, box, row_bytes)) return DSM_EDRIVER;
        for (int q = 1; q < 8; ++q) maps.a[q] = maps.a[0];
        int ncls = 1;
        if (!transposed) {
            g.cls_begin[0] = 0; g.cls_begin[1] = 27;
            for (int kd = 0; kd < 3; ++kd) for (int kh = 0; kh < 3; ++kh) for (int kw = 0; kw < 3; ++kw) {
                const int t = (kd * 3 + kh) * 3 + kw;
                g.a_off[t] = ((kd - 1) * Hp + (kh - 1)) * Wp + (kw - 1);
                g.w_row[t] = t * NP;
            }
        } else {
            // out[2j+par] : par 0 <- (k=1, in j) ; par 1 <- (k=0, in j+1), (k=2, in j)
            ncls = 8;
            int n = 0;
            const int kk[2][2] = {{1, -1}, {0, 2}}, off[2][2] = {{0, 0}, {1, 0}}, cnt[2] = {1, 2};
            for (int q = 0; q < 8; ++q) {
                const int pd = (q >> 2) & 1, ph = (q >> 1) & 1, pw = q & 1;
                g.cls_begin[q] = n;
                for (int a = 0; a < cnt[pd]; ++a) for (int b2 = 0; b2 < cnt[ph]; ++b2) for (int c = 0; c < cnt[pw]; ++c) {
                    g.a_off[n] = (off[pd][a] * Hp + off[ph][b2]) * Wp + off[pw][c];
                    g.w_row[n] = ((kk[pd][a] * 3 + kk[ph][b2]) * 3 + kk[pw][c]) * NP;
                    ++n;
                }
            }
            g.cls_begin[8] = n;   // 27
        }
        g.mtiles = (int)dsm_ceil_div_ll(P, 128);
        g.ntiles = g.mtiles * ncls;
        grid = dim3((unsigned)g.ntiles, 1, 1);
    }
    cudaStream_t st = (cudaStream_t)stream;
    g.tx_bytes = (mode == MODE_SHIFT ? 130 : 128) * row_bytes + (mode == MODE_SHIFT ? 3 : 1) * NP * row_bytes;
    if (mode == MODE_BOX && KC == 32 && NP == 64 && g.nchunks == 1 && !(variant & 512)) {
        // hourglass conv1 (32 -> 64, stride 2): the 27 weight tiles (110 KB) stay resident, only activation boxes stream
        g.tx_bytes = 128 * row_bytes;
        return launch_cfg<32, 64, MODE_BOX, 1>(maps, g, grid, scale, shift, residual, y, st);
    }
    if (mode == MODE_BOX)   return launch_mode<MODE_BOX>(KC, NP, maps, g, grid, scale, shift, residual, y, st);
    if (mode == MODE_SHIFT) return launch_mode<MODE_SHIFT>(KC, NP, maps, g, grid, scale, shift, residual, y, st);
    return launch_mode<MODE_FLAT>(KC, NP, maps, g, grid, scale, shift, residual, y, st);
}

// ---------------------------------------------------------------------------------------------
// 2-D convolutions of the feature-extraction trunks (models/psmnet/submodule.py:65-140, models/gcnet.py:14-29,
// models/util_conv.py:119-132,180-208) on the same implicit-GEMM kernel: an image is a volume without rim planes
// (dpad = 0, Di = 1).  Activations are padded NHWC bf16 [B][H+2r][W+2r][ld] with a zero rim of r >= dilation pixels;
// ld >= C lets a layer read / write a channel slice of a wider tensor (the 320-channel SPP concatenation is never
// copied: its producers store straight into their slices).
//   k = 3, stride 1 : MODE_SHIFT — per kh one (128 + 2*dil)-row tile, the three kw taps are descriptors shifted by dil rows
//   k = 1, stride 1 : MODE_FLAT with a single tap
//   stride 2 (k 3|1): MODE_BOX over the four (h, w) parity sub-lattices (input rim must be 1)
// ---------------------------------------------------------------------------------------------
int conv2d_dispatch(const void* x, const void* w, const float* scale, const float* shift, const void* residual, void* y,
                    int B, int Cin, int Cout, int H, int W, int ksize, int stride, int dil, int relu,
                    int ri, int ro, int ldx, int ldy, int ldr, int y_mode, int variant, void* stream) {
    if (!x || !w || !y || B <= 0 || Cin <= 0 || Cout <= 0 || H <= 0 || W <= 0) return DSM_EINVAL;
    if ((ksize != 1 && ksize != 3) || (stride != 1 && stride != 2) || dil < 1 || dil > 2 || relu < 0 || relu > 2) return DSM_EINVAL;
    if (y_mode != 0 && y_mode != 2) return DSM_EINVAL;
    if (ri < 1 || ri > 2 || ro < 0 || ro > 2 || (ksize == 3 && ri < dil)) return DSM_EINVAL;
    if (stride == 2 && (ri != 1 || dil != 1)) return DSM_EUNSUPPORTED;
    if (y_mode == 2 && residual) return DSM_EUNSUPPORTED;
    if (Cin % 32 != 0 || (Cin > 32 && Cin % 64 != 0)) return DSM_EUNSUPPORTED;
    if (Cout != 16 && Cout != 32 && Cout != 64 && Cout != 128) return DSM_EUNSUPPORTED;
    if (ldx < Cin || (ldx & 7) || (y_mode == 0 && (ldy < Cout || (ldy & 7))) || (residual && (ldr < Cout || (ldr & 7)))) return DSM_EINVAL;
    if (!dsm_aligned16(x) || !dsm_aligned16(w) || !dsm_aligned16(y) || (residual && !dsm_aligned16(residual))) return DSM_EALIGN;
    g_launch_pdl = (variant & 128) != 0;
    const int KC = (Cin == 32) ? 32 : 64, NP = Cout;
    const int row_bytes = KC * 2;
    const int Hp = H + 2 * ri, Wp = W + 2 * ri;
    const long long P = (long long)B * Hp * Wp;
    if (P > 0x7fffff00LL) return DSM_EUNSUPPORTED;
    const int Ho = stride == 2 ? (H - 1) / 2 + 1 : H, Wo = stride == 2 ? (W - 1) / 2 + 1 : W;
    const int ntaps = ksize * ksize;

    ConvMaps maps;
    ConvGeom g;
    memset(&g, 0, sizeof(g));
    g.B = B; g.Di = 1; g.Hi = H; g.Wi = W; g.Do = 1; g.Ho = Ho; g.Wo = Wo;
    g.Cout = Cout; g.transposed = 0; g.relu = relu; g.y_f32 = 0;
    g.nchunks = Cin / KC; g.P = P;
    g.dpad = 0; g.ri = ri; g.ro = ro; g.dil = dil; g.ldy = ldy; g.ldr = ldr; g.y_mode = y_mode;
    {   // weights: [ntaps*NP][Cin]
        cuuint64_t dims[2] = {(cuuint64_t)Cin, (cuuint64_t)ntaps * NP};
        cuuint64_t strides[1] = {(cuuint64_t)Cin * 2};
        cuuint32_t box[2] = {(cuuint32_t)KC, (cuuint32_t)NP};
        if (!encode_map(&maps.w, w, 2, dims, strides, box, row_bytes)) return DSM_EDRIVER;
    }
    g.cls_begin[0] = 0; g.cls_begin[1] = ntaps;
    cudaStream_t st = (cudaStream_t)stream;
    if (stride == 2) {
        for (int q = 0; q < 4; ++q) {
            const int ph = (q >> 1) & 1, pw = q & 1;
            const char* basep = reinterpret_cast<const char*>(x) + ((size_t)ph * Wp + pw) * (size_t)ldx * 2;
            cuuint64_t dims[5] = {(cuuint64_t)Cin, (cuuint64_t)((Wp - pw + 1) / 2), (cuuint64_t)((Hp - ph + 1) / 2), 1, (cuuint64_t)B};
            cuuint64_t strides[4] = {(cuuint64_t)2 * ldx * 2, (cuuint64_t)2 * Wp * ldx * 2,
                                     (cuuint64_t)Hp * Wp * ldx * 2, (cuuint64_t)Hp * Wp * ldx * 2};
            cuuint32_t box[5] = {(cuuint32_t)KC, 16, 8, 1, 1};
            if (!encode_map(&maps.a[q], basep, 5, dims, strides, box, row_bytes)) return DSM_EDRIVER;
        }
        for (int q = 4; q < 8; ++q) maps.a[q] = maps.a[q - 4];
        if (ksize == 3) {
            for (int kh = 0; kh < 3; ++kh) for (int kw = 0; kw < 3; ++kw) {
                const int t = kh * 3 + kw;
                g.a_off[t] = (((kh & 1) << 1) | (kw & 1)) | ((kw >> 1) << 4) | ((kh >> 1) << 5);
                g.w_row[t] = t * NP;
            }
        } else {                                   // 1x1, stride 2: unpadded (2oh, 2ow) = padded (2oh+1, 2ow+1)
            g.a_off[0] = 3; g.w_row[0] = 0;
        }
        g.tiles_w = dsm_ceil_div(Wo, 16); g.tiles_h = dsm_ceil_div(Ho, 8);
        const long long nt = (long long)B * g.tiles_h * g.tiles_w;
        if (nt > 0x7fffffffLL) return DSM_EUNSUPPORTED;
        g.ntiles = (int)nt; g.mtiles = (int)nt;
        g.tx_bytes = 128 * row_bytes + NP * row_bytes;
        return launch_mode<MODE_BOX>(KC, NP, maps, g, dim3((unsigned)nt), scale, shift, residual, y, st);
    }
    const bool shift_mode = (ksize == 3);
    const int a_rows = shift_mode ? 128 + 2 * dil : 128;
    {
        cuuint64_t dims[2] = {(cuuint64_t)Cin, (cuuint64_t)P};
        cuuint64_t strides[1] = {(cuuint64_t)ldx * 2};
        cuuint32_t box[2] = {(cuuint32_t)KC, (cuuint32_t)a_rows};
        if (!encode_map(&maps.a[0], x, 2, dims, strides, box, row_bytes)) return DSM_EDRIVER;
        for (int q = 1; q < 8; ++q) maps.a[q] = maps.a[0];
    }
    if (ksize == 3) {
        for (int kh = 0; kh < 3; ++kh) for (int kw = 0; kw < 3; ++kw) {
            const int t = kh * 3 + kw;
            g.a_off[t] = ((kh - 1) * Wp + (kw - 1)) * dil;
            g.w_row[t] = t * NP;
        }
    } else {
        g.a_off[0] = 0; g.w_row[0] = 0;
    }
    g.mtiles = (int)dsm_ceil_div_ll(P, 128);
    g.ntiles = g.mtiles;
    g.tx_bytes = a_rows * row_bytes + (shift_mode ? 3 : 1) * NP * row_bytes;
    if (shift_mode) return launch_mode<MODE_SHIFT>(KC, NP, maps, g, dim3((unsigned)g.ntiles), scale, shift, residual, y, st);
    return launch_mode<MODE_FLAT>(KC, NP, maps, g, dim3((unsigned)g.ntiles), scale, shift, residual, y, st);
}

}  // namespace

extern "C" int dsm_conv3d_fwd(const void* x, const void* w_packed, const float* scale, const float* shift,
                              const void* residual, void* y,
                              int B, int Cin, int Cout, int D, int H, int W,
                              int stride, int transposed, int relu, int y_dtype,
                              void* ws, size_t ws_bytes, void* stream) {
    DsmDeviceGuard dsm_guard_(x);
    (void)ws; (void)ws_bytes;
    return conv3d_dispatch(x, w_packed, scale, shift, residual, y, B, Cin, Cout, D, H, W, stride, transposed, relu,
                           y_dtype, 0, 0, 0, /*variant=*/0, stream);
}

// Extended form: (Do,Ho,Wo) is the extent of the y / residual buffers when it is a crop of the
// natural output size (the reference's myadd_3d / myAdd3d crop-to-min semantics,
// stackhourglass.py:10-20, util_fun.py:41-51); `variant` bit0 = descriptor base-offset mode (experiment),
// bit1 = force the per-tap kernel instead of the row-shifted-descriptor kernel for stride-1 convolutions.
extern "C" int dsm_conv3d_fwd_ex(const void* x, const void* w_packed, const float* scale, const float* shift,
                                 const void* residual, void* y,
                                 int B, int Cin, int Cout, int D, int H, int W,
                                 int stride, int transposed, int relu, int y_dtype,
                                 int Do, int Ho, int Wo, int variant, void* stream) {
    DsmDeviceGuard dsm_guard_(x);
    return conv3d_dispatch(x, w_packed, scale, shift, residual, y, B, Cin, Cout, D, H, W, stride, transposed, relu,
                           y_dtype, Do, Ho, Wo, variant, stream);
}

// The first 3-D convolution of PSMNet / GC-Net (dres0.0, l19: 64 -> 32, stride 1) fused with the concatenation cost volume
// it reads: see conv3d_rs_kernel<64, 32, FUSED>.  featL / featR: bf16 NHWC with a zero rim, [B][H+2][W+2][32].
extern "C" int dsm_conv3d_volume_fwd(const void* featL, const void* featR, const void* w_packed, const float* scale, const float* shift,
                                     void* y, int B, int C, int Cout, int D, int H, int W, int mode, int relu, int variant, void* stream) {
    DsmDeviceGuard dsm_guard_(featL);
    if (!featL || !featR || !w_packed || !y || B <= 0 || D <= 0 || H <= 0 || W <= 0) return DSM_EINVAL;
    if (mode < DSM_VOL_PSM || mode > DSM_VOL_GC_RIGHT || relu < 0 || relu > 2) return DSM_EINVAL;
    if (C != 32 || Cout != 32) return DSM_EUNSUPPORTED;                  // 2C = 64 input channels, one 128-byte row per voxel
    if (!dsm_aligned16(featL) || !dsm_aligned16(featR) || !dsm_aligned16(w_packed) || !dsm_aligned32(y)) return DSM_EALIGN;
    const int Hp = H + 2, Wp = W + 2;
    if ((long long)B * (D + 2) * Hp * Wp > 0x7fffff00LL) return DSM_EUNSUPPORTED;
    g_launch_pdl = (variant & 128) != 0;
    CUtensorMap map_w;
    {
        cuuint64_t dims[2] = {64, (cuuint64_t)27 * 32};
        cuuint64_t strides[1] = {64 * 2};
        cuuint32_t box[2] = {64, 32};
        if (!encode_map(&map_w, w_packed, 2, dims, strides, box, 128)) return DSM_EDRIVER;
    }
    RsGeom rg;
    memset(&rg, 0, sizeof(rg));
    rg.B = B; rg.D = D; rg.H = H; rg.W = W; rg.Do = D; rg.Ho = H; rg.Wo = W;
    rg.Cout = 32; rg.relu = relu; rg.y_f32 = 0;
    rg.plane_tiles = dsm_ceil_div(Hp * Wp, 128);
    rg.nbands = dsm_ceil_div(D, 8);
    const long long ni = (long long)B * rg.nbands * rg.plane_tiles;
    if (ni > 0x7fffffffLL) return DSM_EUNSUPPORTED;
    rg.nitems = (int)ni;
    rg.dbg = (variant >> 4) & 7;
    rg.ngroups = 1;
    for (int t = 0; t < 27; ++t) rg.w_row[t] = t * 32;
    rg.featL = static_cast<const __nv_bfloat16*>(featL); rg.featR = static_cast<const __nv_bfloat16*>(featR); rg.vol_mode = mode;
    // first half of a voxel: fL (fR for the right-reference volume); second half: the other map, read d pixels to the left (right)
    const void* first = (mode == DSM_VOL_GC_RIGHT) ? featR : featL;
    const void* second = (mode == DSM_VOL_GC_RIGHT) ? featL : featR;
    CUtensorMap map1, map2;
    {
        cuuint64_t dims[2] = {32, (cuuint64_t)B * Hp * Wp};
        cuuint64_t strides[1] = {64};
        cuuint32_t box[2] = {32, 130};
        if (!encode_map(&map1, first, 2, dims, strides, box, 64) || !encode_map(&map2, second, 2, dims, strides, box, 64)) return DSM_EDRIVER;
    }
    return launch_rs<64, 32, true>(map1, map_w, rg, scale, shift, nullptr, y, (cudaStream_t)stream, &map2);
}

// 2-D convolution block of the feature-extraction trunks; see conv2d_dispatch and include/dsmnet_b200.h
extern "C" int dsm_conv2d_rs_fwd(const void* x, const void* w_packed, const float* scale, const float* shift,
                                 const void* residual, void* y, int B, int Cin, int Cout, int H, int W, int dilation, int relu,
                                 int rim_in, int rim_out, int ldx, int ldy, int ldr, int variant, void* stream);
extern "C" int dsm_conv2d_fwd(const void* x, const void* w_packed, const float* scale, const float* shift,
                              const void* residual, void* y,
                              int B, int Cin, int Cout, int H, int W, int ksize, int stride, int dilation, int relu,
                              int rim_in, int rim_out, int ldx, int ldy, int ldr, int y_mode, int variant, void* stream) {
    DsmDeviceGuard dsm_guard_(x);
    // stride-1 3x3 layers with 32 / 64 / 128 channels: the row-sharing kernel (conv2d_rs.cu) unless variant bit 2 asks for the per-tile one
    if (!(variant & 4) && ksize == 3 && stride == 1 && y_mode == 0 && relu <= 1 && (Cin == 32 || Cin == 64 || Cin == 128) && (Cout == 32 || Cout == 64 || Cout == 128) && !(Cin == 128 && Cout == 32) &&
        !(ldy & 15) && !(ldr & 15) && dsm_aligned32(y) && (!residual || dsm_aligned32(residual)))
        return dsm_conv2d_rs_fwd(x, w_packed, scale, shift, residual, y, B, Cin, Cout, H, W, dilation, relu, rim_in, rim_out,
                                 ldx, ldy, ldr, variant, stream);
    return conv2d_dispatch(x, w_packed, scale, shift, residual, y, B, Cin, Cout, H, W, ksize, stride, dilation, relu,
                           rim_in, rim_out, ldx, ldy, ldr, y_mode, variant, stream);
}

// bring-up aid: `host_mapped` = device-visible int[4] the kernel writes progress codes into (NULL = off)
extern "C" int dsm_debug_conv_set_progress(int* host_mapped) {
    return (int)cudaMemcpyToSymbol(g_conv_progress, &host_mapped, sizeof(int*));
}

// 1 (default): a timed-out pipeline wait traps (fatal, visible); 0: bring-up mode, counted and survivable.  Returns the
// previous setting.  Applies to the forward/dgrad kernels and to the tcgen05 weight-gradient kernel.
extern "C" int dsm_debug_wgrad_set_trap(int on);
extern "C" int dsm_debug_conv_set_trap(int on) {
    int prev = 1;
    cudaMemcpyFromSymbol(&prev, g_conv_trap, sizeof(int));
    on = on ? 1 : 0;
    cudaMemcpyToSymbol(g_conv_trap, &on, sizeof(int));
    dsm_debug_wgrad_set_trap(on);
    return prev;
}

// number of pipeline waits that timed out since the library was loaded (0 in a healthy run)
extern "C" int dsm_debug_conv_timeouts(void) {
    int v = -1;
    if (cudaMemcpyFromSymbol(&v, g_conv_timeouts, sizeof(int)) != cudaSuccess) return -1;
    return v;
}
