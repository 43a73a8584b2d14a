// Row-sharing kernel for the stride-1 3x3 convolutions of the 2-D trunks with 32 or 64 output channels
// (models/psmnet/submodule.py:10-13,21-42: firstconv.2/.4, layer1, layer2 — 40 of feature_extraction's 56 convolutions;
// models/gcnet.py:14-29 + util_conv.py:180-208: all of feature2d's blocks).
//
// The per-tile kernel (conv3d_igemm_kernel, MODE_SHIFT) re-loads the 9 weight tiles for every 128-pixel tile and reads every
// input row three times (once per kh); at the trunk's sizes (494 tiles over 148 CTAs for a 64 -> 64 layer of a stereo pair) a
// launch spends 3 us on MMAs inside a 15 us kernel.  This is the plane-sharing idea of conv3d_rs_kernel one dimension down:
//   * a CTA owns a band of R consecutive output ROWS (of equal parity when the dilation is 2) of one 128-column tile, each row
//     with its own NP accumulator columns (R * NP = 256 columns per buffer, two buffers = all of TMEM);
//   * the A tile of input row i feeds output rows i-1, i, i+1 through taps kh = 2, 1, 0, whose weight rows sit back to back in
//     shared memory: ONE MMA with N = 3 * NP per (kw, K step) updates three adjacent accumulator blocks; the kw taps are
//     descriptors shifted by `dil` rows into the same (128 + 2*dil)-row tile — every input row is loaded once per band;
//   * all 9 weight tiles stay resident for the whole kernel; only activation tiles stream through the TMA ring;
//   * three warps issue MMAs, taking the input rows in rotation; the epilogue warps hand every accumulator block back
//     zero-filled (tcgen05.st), so every MMA accumulates and the issue order between warps does not matter.
// Activations: bf16 [B][H+2r][W+2r][ld] with a zero rim of r >= dil pixels (trunk2d.PaddedImage); ld >= C (channel slices).
#include "common.cuh"
#include "ptx.cuh"
#include "epilogue.cuh"
#include "tma_host.cuh"
#include <string.h>

namespace {

__device__ __forceinline__ unsigned long long gtimer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
// bounded wait: a wedged pipeline traps after 2 s (see conv3d.cu)
__device__ __forceinline__ void wait_bar(uint32_t bar, uint32_t parity) {
    if (ptx::mbar_try_wait(bar, parity)) return;
    unsigned long long t0 = 0;
    for (uint32_t spins = 1; ; ++spins) {
        if (ptx::mbar_try_wait(bar, parity)) return;
        if ((spins & 4095u) == 0u) {
            if (t0 == 0) t0 = gtimer_ns();
            if (gtimer_ns() - t0 > 2000000000ULL) __trap();
        }
    }
}
__device__ __forceinline__ void consume_tmem_load(uint32_t v0, uint32_t scratch_smem) {
    asm volatile(
        "{\n\t.reg .pred q;\n\t"
        "setp.eq.u32 q, %0, 0xFFF0DEAD;\n\t"
        "@q st.shared.u32 [%1], %0;\n\t}"
        :: "r"(v0), "r"(scratch_smem) : "memory");
}

struct R2Geom {
    int B, H, W;             // image extent (stride 1: output == input extent)
    int ri, ro, dil;         // rim of the input / of the output and residual buffers; dilation
    int ldy, ldr;            // channels per pixel of the y / residual buffers
    int relu;                // 0 none, 1 after the residual add
    int row_tiles;           // ceil((W + 2*ri) / 128)
    int nbands[2];           // bands of R rows per row parity class (dil = 1: one class)
    int items_per_image;     // row_tiles * (nbands[0] + nbands[1])
    int nitems;              // B * items_per_image
    int Cout;
    int ngroups;             // output-channel groups of NP channels (Cout = 128: two groups of 64); a CTA serves ONE group
    int nissue;              // 1: one warp issues every MMA (bit-reproducible accumulation order); 3: the three warps rotate
};

template <int KC, int NP, int NCH>
struct R2Cfg {
    static constexpr int R = 256 / NP;                           // output rows per band: 8 (NP = 32) or 4 (NP = 64)
    static constexpr int ROWB = KC * 2;
    static constexpr int A_ROWS = 132;                           // 128 + 2 * max dilation
    static constexpr int A_BYTES = ((A_ROWS * ROWB + 1023) / 1024) * 1024;
    static constexpr int W_TILE = 3 * NP * ROWB;                  // the three kh taps of one (kw, K chunk), kh = 2, 1, 0
    static constexpr int W_BYTES = 3 * NCH * W_TILE;              // [kw][chunk] tiles, resident for the whole kernel
    static constexpr int BAR_BYTES = 512;
    static constexpr int BUDGET = 225 * 1024 - W_BYTES - 1024 - BAR_BYTES - 2 * NP * 4;
    static constexpr int S_RAW = BUDGET / A_BYTES;
    // The ring depth is a multiple of the number of MMA issuers NI, so that a stage always belongs to the SAME issuer: that
    // warp then observes every phase of the stage's barrier.  (With rotating ownership a warp looks at a barrier only every
    // few uses, and a parity wait cannot tell "two phases behind" from "done"; having every issuer wait on every stage is not
    // safe either — measured: a warp stalled at the MMA queue gets lapped by the ring.)
    static constexpr int STAGES = S_RAW >= 6 ? 6 : (S_RAW >= 4 ? 4 : S_RAW);
    static constexpr int NI = (STAGES % 3 == 0) ? 3 : ((STAGES % 2 == 0) ? 2 : 1);
    static constexpr int NACC = 2;
    static constexpr int ACC_COLS = R * NP;                       // 256
    static constexpr int TMEM_COLS = NACC * ACC_COLS;             // 512
    static constexpr int CW = 32;                                 // accumulator columns per epilogue step
    static constexpr int SMEM = W_BYTES + STAGES * A_BYTES + 1024 + BAR_BYTES + 2 * NP * 4;
    static constexpr int THREADS = 384;                           // warp 0 TMA, warps 1-3 MMA, warps 4-11 epilogue
    static_assert(STAGES >= 4, "activation ring too shallow");
    static_assert((W_TILE % 1024) == 0 && ((NP * ROWB) % 1024) == 0, "weight sub-tiles must keep the swizzle phase");
};

struct R2Item { int b, sub, o0, nb, tile; };       // output rows o0 + dil * j, j < nb

__device__ __forceinline__ R2Item r2_decode(const R2Geom& g, int t, int R) {
    R2Item it;
    it.b = t / g.items_per_image; t -= it.b * g.items_per_image;
    it.tile = t % g.row_tiles; int band = t / g.row_tiles;
    it.sub = 0;
    if (band >= g.nbands[0]) { band -= g.nbands[0]; it.sub = 1; }
    const int rows = (g.H - it.sub + g.dil - 1) / g.dil;           // rows of this parity class
    it.o0 = it.sub + g.dil * band * R;
    it.nb = min(R, rows - band * R);
    return it;
}

template <int KC, int NP, int NCH>
__global__ void __launch_bounds__(384, 1)
conv2d_rs_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_w,
                 const __grid_constant__ R2Geom g, const float* __restrict__ scale, const float* __restrict__ shift,
                 const void* __restrict__ residual, void* __restrict__ y) {
    using C = R2Cfg<KC, NP, NCH>;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int Hp = g.H + 2 * g.ri, Wp = g.W + 2 * g.ri;
    const int grp = blockIdx.x % g.ngroups;                      // this CTA's output-channel group: its weights never change
    const int t0 = blockIdx.x / g.ngroups, tstep = gridDim.x / g.ngroups;

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
        s_scale[tid] = scale ? __ldg(scale + grp * NP + tid) : 1.f;
        s_shift[tid] = shift ? __ldg(shift + grp * NP + tid) : 0.f;
    }
    if (warp == 0 && lane == 0) {
        ptx::prefetch_tensormap(&map_w);
        ptx::prefetch_tensormap(&map_a);
        for (int s = 0; s < C::STAGES; ++s) { ptx::mbar_init(full_bar(s), 1); ptx::mbar_init(empty_bar(s), 1); }
        for (int a = 0; a < C::NACC; ++a) { ptx::mbar_init(tfull_bar(a), 3); ptx::mbar_init(tempty_bar(a), 8); }
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
        if (ptx::elect_one_sync()) {                            // this group's weight tiles, once: [kw][chunk][kh = 2, 1, 0][NP rows]
            ptx::mbar_arrive_expect_tx(wfull_bar, 9 * NCH * NP * C::ROWB);
            for (int kh = 0; kh < 3; ++kh)
                for (int kw = 0; kw < 3; ++kw)
                    for (int ch = 0; ch < NCH; ++ch)
                        ptx::tma_load_2d(wsm + (kw * NCH + ch) * C::W_TILE + (2 - kh) * NP * C::ROWB, &map_w, wfull_bar,
                                         ch * KC, (kh * 3 + kw) * g.Cout + grp * NP);
        }
        __syncwarp();
        ptx::griddep_wait();                                    // the activations are the previous kernel's output
        int s = 0; uint32_t ph = 0;
        const uint32_t tx = (uint32_t)(128 + 2 * g.dil) * C::ROWB;
        for (int t = t0; t < g.nitems; t += tstep) {
            const R2Item item = r2_decode(g, t, C::R);
            for (int i = -1; i <= item.nb; ++i) {
                const int o = item.o0 + g.dil * i;               // input row (image coordinates)
                if (o < 0 || o >= g.H) continue;                 // a rim row: contributes nothing
#pragma unroll
                for (int ch = 0; ch < NCH; ++ch) {
                    wait_bar(empty_bar(s), ph ^ 1u);
                    if (ptx::elect_one_sync()) {
                        ptx::mbar_arrive_expect_tx(full_bar(s), tx);
                        ptx::tma_load_2d(ring + s * C::A_BYTES, &map_a, full_bar(s), ch * KC,
                                         (item.b * Hp + g.ri + o) * Wp + item.tile * 128 - g.dil);
                    }
                    __syncwarp();
                    if (++s == C::STAGES) { s = 0; ph ^= 1u; }
                }
            }
        }
    } else if (warp <= 3) {
        // ================= MMA issuers: NI of the three warps take the stages in rotation (stage s <-> warp s % NI) =================
        const int my = warp - 1;
        constexpr uint32_t idesc0 = ptx::make_idesc_bf16(0);
        wait_bar(wfull_bar, 0);
        ptx::tc_fence_after();
        const uint64_t dsc = ptx::make_kmajor_desc(0u, C::ROWB, 0u);
        const uint32_t desc_hi = (uint32_t)(dsc >> 32);
        const uint32_t ring_lo = (uint32_t)dsc | (ring >> 4);
        const uint32_t w_lo = (uint32_t)dsc | (wsm >> 4);
        const uint32_t kw_step = (uint32_t)(g.dil * C::ROWB) >> 4;
        int s = 0; uint32_t ph = 0;
        int tcount = 0, rot = 0;
        for (int t = t0; t < g.nitems; t += tstep) {
            const R2Item item = r2_decode(g, t, C::R);
            const int acc = tcount % C::NACC;
            const uint32_t use_ph = (uint32_t)(tcount / C::NACC) & 1u;
            ++tcount;
            wait_bar(tempty_bar(acc), use_ph);                   // drained AND zero-filled by the epilogue warps
            ptx::tc_fence_after();
            const uint32_t d_tmem = tmem + acc * C::ACC_COLS;
            for (int i = -1; i <= item.nb; ++i) {
                const int o = item.o0 + g.dil * i;
                if (o < 0 || o >= g.H) continue;
                const int jlo = max(i - 1, 0), jhi = min(i + 1, item.nb - 1);
                const int brow = (2 - (i - jlo + 1)) * NP;        // weight row of block jlo's tap (kh = i - jlo + 1)
                const uint32_t d_lo = d_tmem + jlo * NP;
                const uint32_t idesc = idesc0 | ((uint32_t)((jhi - jlo + 1) * NP >> 3) << 17);
#pragma unroll
                for (int ch = 0; ch < NCH; ++ch) {
                    if ((g.nissue == 1 ? 0 : rot) == my) {
                        wait_bar(full_bar(s), ph);
                        ptx::tc_fence_after();
                        if (ptx::elect_one_sync()) {
                            const uint32_t a_lo0 = ring_lo + (uint32_t)s * (C::A_BYTES >> 4);
                            const uint32_t b_lo0 = w_lo + ((uint32_t)(brow * C::ROWB + ch * C::W_TILE) >> 4);
#pragma unroll
                            for (int kw = 0; kw < 3; ++kw) {
#pragma unroll
                                for (int k = 0; k < KC / 16; ++k)
                                    ptx::umma_bf16_lohi(d_lo, a_lo0 + kw * kw_step + ((k * 32) >> 4),
                                                        b_lo0 + ((kw * NCH * C::W_TILE + k * 32) >> 4), desc_hi, idesc, 1u);
                            }
                            ptx::umma_commit(empty_bar(s));
                        }
                        __syncwarp();
                    }
                    if (++rot == C::NI) rot = 0;
                    if (++s == C::STAGES) { s = 0; ph ^= 1u; }
                }
            }
            if (ptx::elect_one_sync()) ptx::umma_commit(tfull_bar(acc));   // this warp's share of "accumulators complete"
            __syncwarp();
        }
    } else {
        // ================= epilogue: 8 warps, two per TMEM lane quadrant (even / odd rows of the band) =================
        ptx::griddep_wait();
        const int q = warp & 3;
        const int half = (warp - 4) >> 2;
        const int r = q * 32 + lane;
        constexpr int JJ = C::R / 2;
        {   // both accumulator buffers start zero-filled; arriving completes phase 0 of their "free" barriers
#pragma unroll
            for (int a2 = 0; a2 < C::NACC; ++a2)
#pragma unroll
                for (int jj = 0; jj < JJ; ++jj)
#pragma unroll
                    for (int c0 = 0; c0 < NP; c0 += 32)
                        ptx::tmem_zero32(tmem + ((uint32_t)(q * 32) << 16) + a2 * C::ACC_COLS + (2 * jj + half) * NP + c0);
            ptx::tc_wait_st();
            ptx::tc_fence_before();
            __syncwarp();
            if (lane == 0) {
#pragma unroll
                for (int a2 = 0; a2 < C::NACC; ++a2) ptx::mbar_arrive(tempty_bar(a2));
            }
        }
        const float4* sc4 = reinterpret_cast<const float4*>(s_scale);
        const int emode = epi_mode(g.relu, residual != nullptr);
        const float4* sh4 = reinterpret_cast<const float4*>(s_shift);
        const int Hop = g.H + 2 * g.ro, Wop = g.W + 2 * g.ro;
        int tcount = 0;
        for (int t = t0; t < g.nitems; t += tstep) {
            const R2Item item = r2_decode(g, t, C::R);
            const int wp = item.tile * 128 + r;                  // padded column of this lane
            const int x = wp - g.ri;
            const bool valid = x >= 0 && x < g.W;
            const int acc = tcount % C::NACC;
            const uint32_t acc_ph = (uint32_t)(tcount / C::NACC) & 1u;
            ++tcount;
            const uint32_t taddr0 = tmem + ((uint32_t)(q * 32) << 16) + acc * C::ACC_COLS;
            int my_last = -1;
            for (int jj = 0; jj < JJ; ++jj) if (2 * jj + half < item.nb) my_last = 2 * jj + half;
            auto release = [&]() {
                ptx::tc_wait_st();
                ptx::tc_fence_before();
                __syncwarp();
                if (lane == 0) ptx::mbar_arrive(tempty_bar(acc));
            };
            // pixel offset of (row o0, column x) in the output / residual buffers; consecutive band rows are dil rows apart
            const size_t pix0 = ((size_t)item.b * Hop + g.ro + item.o0) * Wop + g.ro + x;
            const size_t pstep = (size_t)g.dil * Wop;
            const __nv_bfloat16* resb = residual ? reinterpret_cast<const __nv_bfloat16*>(residual) + grp * NP : nullptr;
            wait_bar(tfull_bar(acc), acc_ph);
            __syncwarp();
            ptx::tc_fence_after();
            if (my_last < 0) release();
#pragma unroll
            for (int jj = 0; jj < JJ; ++jj) {
                const int j = 2 * jj + half;
                if (j < item.nb) {
                    const size_t pix = pix0 + (size_t)j * pstep;
                    uint4* out = reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(y) + pix * (size_t)g.ldy + grp * NP);
                    const __nv_bfloat16* rrow = resb ? resb + pix * (size_t)g.ldr : nullptr;
#pragma unroll
                    for (int c0 = 0; c0 < NP; c0 += C::CW) {
                        constexpr int NV = C::CW / 8;                // 8-channel groups per step
                        uint4 rv[NV];
#pragma unroll
                        for (int c = 0; c < NV; c += 2) {
                            rv[c] = rv[c + 1] = make_uint4(0u, 0u, 0u, 0u);
                            if (rrow && valid) ld_nc_v8(rrow + c0 + 8 * c, rv[c], rv[c + 1]);
                        }
                        uint32_t v[C::CW];
                        if constexpr (C::CW == 32) ptx::tmem_ld32(taddr0 + j * NP + c0, v); else ptx::tmem_ld16(taddr0 + j * NP + c0, v);
                        ptx::tc_wait_ld();
                        consume_tmem_load(v[0], scratch_smem);
                        if constexpr (C::CW == 32) ptx::tmem_zero32(taddr0 + j * NP + c0); else ptx::tmem_zero16(taddr0 + j * NP + c0);
                        if (j == my_last && c0 + C::CW >= NP) release();
                        if (valid) {
                            DSM_EPI_DISPATCH(emode, epi_store256, 4, v, sc4 + (c0 >> 2), sh4 + (c0 >> 2), rv, out + (c0 >> 3))   // 256-bit stores
                        }
                    }
                }
            }
        }
    }

    ptx::tc_fence_before();
    __syncthreads();
    if (warp == 1) ptx::tmem_dealloc(tmem, C::TMEM_COLS);
}

template <int KC, int NP, int NCH>
int launch_r2(const CUtensorMap& map_a, const CUtensorMap& map_w, const R2Geom& g, const float* scale, const float* shift,
              const void* residual, void* y, bool pdl, cudaStream_t st) {
    using C = R2Cfg<KC, NP, NCH>;
    auto kern = conv2d_rs_kernel<KC, NP, NCH>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM);
    if (e != cudaSuccess) return (int)e;
    int nsm = DSM_NUM_SMS_B200, dev = 0;
    if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, dev);
    const long long want = (long long)g.nitems * g.ngroups;        // nitems: per group; a multiple of ngroups CTAs
    const int nblocks = want < nsm ? (int)want : nsm / g.ngroups * g.ngroups;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(nblocks); cfg.blockDim = dim3(C::THREADS); cfg.dynamicSmemBytes = C::SMEM; cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr; cfg.numAttrs = pdl ? 1 : 0;
    cudaLaunchKernelEx(&cfg, kern, map_a, map_w, g, scale, shift, residual, y);
    return dsm_launch_status();
}

}  // namespace

// Row-sharing form of dsm_conv2d_fwd for k = 3, stride 1, Cin and Cout in {32, 64}, bf16 padded-NHWC output
// (y = relu?(conv(x) * scale + shift + residual); relu 0 | 1).  Same arguments as dsm_conv2d_fwd; returns
// DSM_EUNSUPPORTED for everything else (the caller then uses dsm_conv2d_fwd).
extern "C" int dsm_conv2d_rs_fwd(const void* x, const void* w_packed, const float* scale, const float* shift,
                                 const void* residual, void* y,
                                 int B, int Cin, int Cout, int H, int W, int dilation, int relu,
                                 int rim_in, int rim_out, int ldx, int ldy, int ldr, int variant, void* stream) {
    DsmDeviceGuard dsm_guard_(x);
    if (!x || !w_packed || !y || B <= 0 || H <= 0 || W <= 0) return DSM_EINVAL;
    if (dilation < 1 || dilation > 2 || relu < 0 || relu > 1) return DSM_EINVAL;
    if ((Cin != 32 && Cin != 64 && Cin != 128) || (Cout != 32 && Cout != 64 && Cout != 128)) return DSM_EUNSUPPORTED;
    if (Cin == 128 && Cout == 32) return DSM_EUNSUPPORTED;
    if (rim_in < dilation || rim_in > 2 || rim_out < 0 || rim_out > 2) return DSM_EINVAL;
    if (ldx < Cin || (ldx & 7) || ldy < Cout || (ldy & 15) || (residual && (ldr < Cout || (ldr & 15)))) return DSM_EINVAL;
    if (!dsm_aligned16(x) || !dsm_aligned16(w_packed) || !dsm_aligned32(y) || (residual && !dsm_aligned32(residual))) return DSM_EALIGN;
    const int Hp = H + 2 * rim_in, Wp = W + 2 * rim_in;
    const long long P = (long long)B * Hp * Wp;
    if (P > 0x7fffff00LL) return DSM_EUNSUPPORTED;
    const int KC = Cin > 64 ? 64 : Cin, NP = Cout > 64 ? 64 : Cout, row_bytes = KC * 2;
    const int R = 256 / NP;
    R2Geom g;
    memset(&g, 0, sizeof(g));
    g.B = B; g.H = H; g.W = W; g.ri = rim_in; g.ro = rim_out; g.dil = dilation; g.ldy = ldy; g.ldr = ldr; g.relu = relu; g.Cout = Cout; g.ngroups = Cout / NP;
    g.row_tiles = dsm_ceil_div(Wp, 128);
    for (int s = 0; s < 2; ++s) {
        const int rows = s < dilation ? (H - s + dilation - 1) / dilation : 0;
        g.nbands[s] = dsm_ceil_div(rows, R);
    }
    g.items_per_image = g.row_tiles * (g.nbands[0] + g.nbands[1]);
    const long long ni = (long long)B * g.items_per_image;
    if (ni > 0x7fffffffLL || ni < 1) return DSM_EUNSUPPORTED;
    g.nitems = (int)ni;
    CUtensorMap map_a, map_w;
    const CUtensorMapSwizzle sw = (row_bytes == 128) ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B;
    {
        cuuint64_t dims[2] = {(cuuint64_t)Cin, (cuuint64_t)P};
        cuuint64_t strides[1] = {(cuuint64_t)ldx * 2};
        cuuint32_t box[2] = {(cuuint32_t)KC, (cuuint32_t)(128 + 2 * dilation)};
        if (!tma_host::encode(&map_a, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, x, 2, dims, strides, box, sw)) return DSM_EDRIVER;
        cuuint64_t wdims[2] = {(cuuint64_t)Cin, (cuuint64_t)9 * Cout};
        cuuint64_t wstrides[1] = {(cuuint64_t)Cin * 2};
        cuuint32_t wbox[2] = {(cuuint32_t)KC, (cuuint32_t)NP};
        if (!tma_host::encode(&map_w, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, w_packed, 2, wdims, wstrides, wbox, sw)) return DSM_EDRIVER;
    }
    const bool pdl = (variant & 128) != 0;
    g.nissue = (variant & 8) ? 1 : 3;          // bit 3: one MMA issuer (bit-reproducible), ~10-20 % slower
    cudaStream_t st = (cudaStream_t)stream;
    if (Cin == 128) return launch_r2<64, 64, 2>(map_a, map_w, g, scale, shift, residual, y, pdl, st);
    if (KC == 32 && NP == 32) return launch_r2<32, 32, 1>(map_a, map_w, g, scale, shift, residual, y, pdl, st);
    if (KC == 32 && NP == 64) return launch_r2<32, 64, 1>(map_a, map_w, g, scale, shift, residual, y, pdl, st);
    if (KC == 64 && NP == 32) return launch_r2<64, 32, 1>(map_a, map_w, g, scale, shift, residual, y, pdl, st);
    return launch_r2<64, 64, 1>(map_a, map_w, g, scale, shift, residual, y, pdl, st);
}
