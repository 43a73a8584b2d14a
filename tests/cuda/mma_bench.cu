// Micro-benchmark (answers SURVEY.md hard part 1): issue/retire rate of tcgen05.mma
// (cta_group::1, kind::f16, M=128, K=16, bf16) as a function of N, for
//   SS : A and B from shared memory (K-major, swizzled rows of 64/128 bytes)
//   TS : A from tensor memory, B from shared memory
// and of tcgen05.cp 128x256b (shared -> tensor memory, one 128x16 bf16 A slab); `mma_bench 8` measures the CTA-pair form
// (cta_group::2, M = 256 per pair: each SM reads its own A slab and half of B).
// The issue loop is unrolled x8 with pre-built descriptors so that the single issuing thread is not
// the limit.  Operand contents are garbage: only rates are measured.
//   build: make tests/cuda/mma_bench ; run: tests/cuda/mma_bench
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include "../../dsmnet_b200/csrc/ptx.cuh"

namespace {
__device__ __forceinline__ void umma_bf16_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
        :: "r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void tmem_cp_128x256b(uint32_t taddr, uint64_t sdesc) {
    asm volatile("tcgen05.cp.cta_group::1.128x256b [%0], %1;" :: "r"(taddr), "l"(sdesc) : "memory");
}
}

enum { SS = 0, TS = 1, CP = 2, TS_CP = 3 };

template <int N, int ROWB, int MODE, int ASHIFT = 0>
__global__ void __launch_bounds__(64) rate_kernel(long long* cycles, int n_iter) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw = ptx::smem_u32(smem_raw);
    const uint32_t base = (raw + 1023u) & ~1023u;
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t slot;
    const int warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) { ptx::mbar_init(ptx::smem_u32(&bar), 1); ptx::fence_mbar_init(); }
    if (warp == 1) ptx::tmem_alloc(ptx::smem_u32(&slot), 512);
    ptx::tc_fence_before(); __syncthreads(); ptx::tc_fence_after();
    const uint32_t tmem = slot;
    if (warp == 0) {
        constexpr uint32_t idesc = ptx::make_idesc_bf16(N);
        constexpr uint32_t a_bytes = 128 * ROWB, b_bytes = N * ROWB;
        const uint32_t sa = base, sb = base + 2 * a_bytes;
        // 8 operand variants: 2 tiles x (ROWB/32) K-slabs, cycled
        uint64_t ad[8], bd[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const uint32_t tile = (u >> 1) & 1, slab = u % (ROWB / 32);
            ad[u] = ptx::make_kmajor_desc(sa + tile * a_bytes + slab * 32 + (ASHIFT % 100) * ROWB, ROWB, 0u);   // ASHIFT rows into the swizzle atom
            bd[u] = ptx::make_kmajor_desc(sb + tile * b_bytes + slab * 32, ROWB, 0u);
        }
        long long t0 = 0, t1 = 0;
        __syncwarp();
        t0 = clock64();
        if (ptx::elect_one_sync()) {
            for (int i = 0; i < n_iter; ++i) {
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    const uint32_t d = tmem + ((ASHIFT >= 100) ? 0 : (u & 1) * 256);   // two accumulators, alternating (ASHIFT>=100: one)
                    if (MODE == SS) ptx::umma_bf16(d, ad[u], bd[u], idesc, 1u);
                    if (MODE == TS) umma_bf16_ts(d, tmem + 480 + (u & 1) * 8, bd[u], idesc, 1u);
                    if (MODE == CP) tmem_cp_128x256b(tmem + 448 + (u & 3) * 8, ad[u]);
                    if (MODE == TS_CP) {                                // one cp feeds three MMAs (the intended reuse)
                        if (u % 3 == 0) tmem_cp_128x256b(tmem + 448 + ((u / 3) & 3) * 8, ad[u]);
                        umma_bf16_ts(d, tmem + 448 + ((u / 3) & 3) * 8, bd[u], idesc, 1u);
                    }
                }
            }
            ptx::umma_commit(ptx::smem_u32(&bar));
        }
        __syncwarp();
        while (!ptx::mbar_try_wait(ptx::smem_u32(&bar), 0)) {}
        t1 = clock64();
        if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
    }
    ptx::tc_fence_before(); __syncthreads();
    if (warp == 1) ptx::tmem_dealloc(tmem, 512);
}

// The plane-sharing conv's exact MMA stream (32->32: rows of 64 B, N=96): per "plane" 3 A tiles (kh) x 3 row shifts (kw)
// x 2 K slabs against 9 distinct 96-row weight tiles, all accumulating into the same 96 columns; NISS warps issue
// (warp w takes kh = w when NISS == 3), one commit per plane per issuer onto a barrier nobody waits for.
template <int NISS>
__global__ void __launch_bounds__(32 * (NISS + 1)) rs_stream_kernel(long long* cycles, int n_planes) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw = ptx::smem_u32(smem_raw);
    const uint32_t base = (raw + 1023u) & ~1023u;
    __shared__ __align__(8) uint64_t bar[4];
    __shared__ uint32_t slot;
    const int warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) { for (int i = 0; i < 4; ++i) ptx::mbar_init(ptx::smem_u32(&bar[i]), 1); ptx::fence_mbar_init(); }
    if (warp == NISS) ptx::tmem_alloc(ptx::smem_u32(&slot), 512);
    ptx::tc_fence_before(); __syncthreads(); ptx::tc_fence_after();
    const uint32_t tmem = slot;
    constexpr uint32_t A_BYTES = 9216, W_TILE = 6144, STAGE = 3 * A_BYTES;
    const uint32_t wsm = base, ring = base + 9 * W_TILE;
    if (warp < NISS) {
        const uint64_t dsc = ptx::make_kmajor_desc(0u, 64, 0u);
        const uint32_t desc_hi = (uint32_t)(dsc >> 32);
        const uint32_t ring_lo = (uint32_t)dsc | (ring >> 4), w_lo = (uint32_t)dsc | (wsm >> 4);
        constexpr uint32_t idesc = ptx::make_idesc_bf16(96);
        __syncwarp();
        const long long t0 = clock64();
        if (ptx::elect_one_sync()) {
            int s = 0;
            for (int pl = 0; pl < n_planes; ++pl) {
                const uint32_t d = tmem + (uint32_t)((pl % 6) * 32);
                for (int kh = (NISS == 3 ? warp : 0); kh < 3; kh += (NISS == 3 ? 3 : 1)) {
                    const uint32_t a0 = ring_lo + (uint32_t)s * (STAGE >> 4) + (uint32_t)((kh * A_BYTES) >> 4);
                    const uint32_t b0 = w_lo + (uint32_t)((kh * 3 * W_TILE) >> 4);
#pragma unroll
                    for (int kw = 0; kw < 3; ++kw)
#pragma unroll
                        for (int k = 0; k < 2; ++k)
                            ptx::umma_bf16_lohi(d, a0 + ((kw * 64 + k * 32) >> 4), b0 + ((kw * W_TILE + k * 32) >> 4), desc_hi, idesc, 1u);
                }
                ptx::umma_commit(ptx::smem_u32(&bar[1]));
                if (++s == 6) s = 0;
            }
            ptx::umma_commit(ptx::smem_u32(&bar[0]) + 8u * 2 + 8u * (warp ? 1 : 0) * 0);
        }
        __syncwarp();
        if (warp == 0) { while (!ptx::mbar_try_wait(ptx::smem_u32(&bar[2]), 0)) {} }
        const long long t1 = clock64();
        if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
    }
    ptx::tc_fence_before(); __syncthreads();
    if (warp == NISS) ptx::tmem_dealloc(tmem, 512);
}

template <int NISS>
void run_rs(const char* what) {
    const int n_planes = 400;
    int nsm = 148; cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, 0);
    const size_t smem = 9 * 6144 + 6 * 3 * 9216 + 2048;
    auto k = rs_stream_kernel<NISS>;
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    long long* d; cudaMalloc(&d, nsm * sizeof(long long));
    k<<<nsm, 32 * (NISS + 1), smem>>>(d, n_planes);
    k<<<nsm, 32 * (NISS + 1), smem>>>(d, n_planes);
    cudaError_t e = cudaDeviceSynchronize();
    long long h[1024]; cudaMemcpy(h, d, nsm * sizeof(long long), cudaMemcpyDeviceToHost);
    double mx = 0; for (int i = 0; i < nsm; ++i) mx = h[i] > mx ? h[i] : mx;
    printf("%-22s : %7.1f cyc/plane (18 MMAs N=96) = %5.1f cyc/MMA (%s)\n", what, mx / n_planes, mx / n_planes / 18.0, cudaGetErrorString(e));
    cudaFree(d);
}

// ---- cta_group::2 rate probe: a CTA pair issues M=256 (2 x 128) SS MMAs; each SM reads its own 4 KB A slab and HALF of
// the B tile per instruction, so the shared-memory operand traffic per SM drops from 4 KB + N*32 B to 4 KB + N*16 B.
// Operand contents are garbage; only the issue/retire rate is measured.  Every wait is bounded.
__device__ __forceinline__ uint32_t cluster_rank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
template <int N, int ROWB>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(64) rate2_kernel(long long* cycles, int n_iter) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw = ptx::smem_u32(smem_raw);
    const uint32_t base = (raw + 1023u) & ~1023u;
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t slot;
    const int warp = threadIdx.x >> 5;
    const uint32_t rank = cluster_rank();
    if (threadIdx.x == 0) { ptx::mbar_init(ptx::smem_u32(&bar), 1); ptx::fence_mbar_init(); }
    __syncthreads();
    cluster_sync_all();
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(ptx::smem_u32(&slot)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
    ptx::tc_fence_before(); __syncthreads(); ptx::tc_fence_after();
    cluster_sync_all();
    const uint32_t tmem = slot;
    long long t0 = 0, t1 = 0;
    if (warp == 0) {
        // instruction descriptor: as make_idesc_bf16 but M = 256
        constexpr uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);
        constexpr uint32_t a_bytes = 128 * ROWB, b_bytes = (N / 2) * ROWB;       // per CTA
        const uint32_t sa = base, sb = base + 2 * a_bytes;
        uint64_t ad[8], bd[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const uint32_t tile = (u >> 1) & 1, slab = u % (ROWB / 32);
            ad[u] = ptx::make_kmajor_desc(sa + tile * a_bytes + slab * 32, ROWB, 0u);
            bd[u] = ptx::make_kmajor_desc(sb + tile * b_bytes + slab * 32, ROWB, 0u);
        }
        __syncwarp();
        t0 = clock64();
        if (rank == 0 && ptx::elect_one_sync()) {
            for (int i = 0; i < n_iter; ++i) {
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    const uint32_t d = tmem + (u & 1) * 256;
                    asm volatile(
                        "{\n\t.reg .pred p;\n\t"
                        "setp.ne.b32 p, %4, 0;\n\t"
                        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                        :: "r"(d), "l"(ad[u]), "l"(bd[u]), "r"(idesc), "r"(1u) : "memory");
                }
            }
            asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                         :: "r"(ptx::smem_u32(&bar)), "h"((uint16_t)3) : "memory");
        }
        __syncwarp();
        long long guard = clock64();
        while (!ptx::mbar_try_wait(ptx::smem_u32(&bar), 0)) { if (clock64() - guard > 4000000000LL) break; }
        t1 = clock64();
        if (threadIdx.x == 0 && rank == 0) cycles[blockIdx.x / 2] = t1 - t0;
    }
    ptx::tc_fence_before(); __syncthreads();
    cluster_sync_all();
    if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" :: "r"(tmem), "r"(512u) : "memory");
}

template <int N, int ROWB>
void run2(const char* what) {
    const int n_iter = 512;
    int nsm = 148; cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, 0);
    const int nblk = (nsm / 2) * 2;
    const size_t smem = (size_t)2 * (128 * ROWB + (N / 2) * ROWB) + 2048 + 8 * ROWB;
    auto k = rate2_kernel<N, ROWB>;
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    long long* d; cudaMalloc(&d, nsm * sizeof(long long)); cudaMemset(d, 0, nsm * sizeof(long long));
    k<<<nblk, 64, smem>>>(d, n_iter);
    k<<<nblk, 64, smem>>>(d, n_iter);
    cudaError_t e = cudaDeviceSynchronize();
    long long h[1024]; cudaMemcpy(h, d, (nblk / 2) * sizeof(long long), cudaMemcpyDeviceToHost);
    double mx = 0; for (int i = 0; i < nblk / 2; ++i) mx = h[i] > mx ? h[i] : mx;
    const double cyc = mx / (n_iter * 8.0);
    printf("%-6s N=%3d rowB=%3d : %6.1f cyc/op (M=256 per CTA pair) -> %6.0f flop/clk/SM (%s)\n", what, N, ROWB, cyc,
           2.0 * 256 * N * 16 / cyc / 2.0, cudaGetErrorString(e));
    cudaFree(d);
}

template <int N, int ROWB, int MODE, int ASHIFT = 0>
void run(const char* what) {
    const int n_iter = 512;
    int nsm = 148; cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, 0);
    const size_t smem = (size_t)2 * (128 * ROWB + N * ROWB) + 2048 + 8 * ROWB;
    auto k = rate_kernel<N, ROWB, MODE, ASHIFT>;
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    long long* d; cudaMalloc(&d, nsm * sizeof(long long));
    k<<<nsm, 64, smem>>>(d, n_iter);
    k<<<nsm, 64, smem>>>(d, n_iter);
    cudaError_t e = cudaDeviceSynchronize();
    long long h[1024]; cudaMemcpy(h, d, nsm * sizeof(long long), cudaMemcpyDeviceToHost);
    double mx = 0; for (int i = 0; i < nsm; ++i) mx = h[i] > mx ? h[i] : mx;
    const double cyc = mx / (n_iter * 8.0);
    const double flop = (MODE == CP) ? 0.0 : 2.0 * 128 * N * 16 / cyc;
    printf("%-6s N=%3d rowB=%3d : %6.1f cyc/op -> %6.0f flop/clk/SM (%s)\n", what, N, ROWB, cyc, flop, cudaGetErrorString(e));
    cudaFree(d);
}

int main(int argc, char** argv) {
    setvbuf(stdout, nullptr, _IONBF, 0);
    const int what = argc > 1 ? atoi(argv[1]) : 0;
    if (what == 0 || what == 1) {
        run<32, 64, SS>("SS");   run<64, 64, SS>("SS");   run<96, 64, SS>("SS");  run<128, 64, SS>("SS");
        run<32, 128, SS>("SS");  run<48, 64, SS>("SS");   run<64, 128, SS>("SS");  run<96, 128, SS>("SS"); run<128, 128, SS>("SS");
        run<192, 128, SS>("SS"); run<256, 128, SS>("SS");
    }
    if (what == 0 || what == 2) {
        run<32, 64, TS>("TS");   run<64, 64, TS>("TS");   run<96, 64, TS>("TS");   run<32, 128, TS>("TS"); run<128, 128, TS>("TS");
    }
    if (what == 0 || what == 3) { run<32, 64, CP>("CP");  run<32, 128, CP>("CP"); }
    if (what == 0 || what == 5) {   // A tile read through a descriptor that starts 1 / 2 / 4 rows into a swizzle atom
        run<96, 64, SS, 1>("SS+1r"); run<96, 64, SS, 2>("SS+2r"); run<96, 64, SS, 4>("SS+4r"); run<32, 64, SS, 1>("SS+1r");
        run<96, 128, SS, 1>("SS+1r"); run<96, 128, SS, 2>("SS+2r"); run<96, 128, SS, 4>("SS+4r");
    }
    if (what == 0 || what == 6) {   // every MMA accumulates into the SAME columns (dependent chain), as in a real K loop
        run<32, 64, SS, 100>("SS-dep"); run<64, 64, SS, 100>("SS-dep"); run<96, 64, SS, 100>("SS-dep"); run<128, 128, SS, 100>("SS-dep");
        run<256, 128, SS, 100>("SS-dep"); run<32, 64, TS, 100>("TS-dep"); run<96, 64, TS, 100>("TS-dep");
    }
    if (what == 0 || what == 7) { run_rs<1>("RS stream, 1 issuer"); run_rs<3>("RS stream, 3 issuers"); }
    if (what == 8) {    // cta_group::2 (not part of the default sweep)
        run2<32, 64>("SS2"); run2<64, 64>("SS2"); run2<96, 64>("SS2"); run2<128, 64>("SS2"); run2<96, 128>("SS2"); run2<256, 128>("SS2");
    }
    if (what == 0 || what == 4) { run<32, 64, TS_CP>("TS+CP"); run<32, 128, TS_CP>("TS+CP"); run<64, 128, TS_CP>("TS+CP"); }
    return 0;
}
