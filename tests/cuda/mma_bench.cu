// Micro-benchmark: how many SM cycles does one tcgen05.mma (cta_group::1, kind::f16, M=128, K=16,
// bf16 K-major operands from shared memory) take as a function of N, of the swizzle mode (row
// bytes 64 / 128) and of the number of CTAs issuing concurrently on an SM?  Answers SURVEY.md
// hard part 1 ("micro-benchmark tcgen05 throughput vs N first").  Operand contents are garbage
// (uninitialised smem): only the issue/retire rate is measured.
//   build: make tests/cuda/mma_bench ; run: tests/cuda/mma_bench
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include "../../dsmnet_b200/csrc/ptx.cuh"

template <int N, int ROWB>
__global__ void __launch_bounds__(64) mma_rate_kernel(long long* cycles, int n_mma, int distinct) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw = ptx::smem_u32(smem_raw);
    const uint32_t base = (raw + 1023u) & ~1023u;
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t slot;
    const int warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) { ptx::mbar_init(ptx::smem_u32(&bar), 1); ptx::fence_mbar_init(); }
    if (warp == 1) ptx::tmem_alloc(ptx::smem_u32(&slot), 256);
    ptx::tc_fence_before(); __syncthreads(); ptx::tc_fence_after();
    const uint32_t tmem = slot;
    if (warp == 0) {
        constexpr uint32_t idesc = ptx::make_idesc_bf16(N);
        const uint32_t a_bytes = 128 * ROWB, b_bytes = N * ROWB;
        long long t0 = 0, t1 = 0;
        __syncwarp();
        t0 = clock64();
        if (ptx::elect_one_sync()) {
            for (int i = 0; i < n_mma; ++i) {
                // `distinct` operand tiles are cycled so that reads do not all hit the same smem lines
                const uint32_t sa = base + (i % distinct) * (a_bytes + b_bytes);
                const uint64_t ad = ptx::make_kmajor_desc(sa + (i & ((ROWB / 32) - 1)) * 32, ROWB, 0u);
                const uint64_t bd = ptx::make_kmajor_desc(sa + a_bytes + (i & ((ROWB / 32) - 1)) * 32, ROWB, 0u);
                ptx::umma_bf16(tmem + (i & 1) * N % 256, ad, bd, idesc, 1u);
            }
            ptx::umma_commit(ptx::smem_u32(&bar));
        }
        __syncwarp();
        while (!ptx::mbar_try_wait(ptx::smem_u32(&bar), 0)) {}
        t1 = clock64();
        if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
    }
    ptx::tc_fence_before(); __syncthreads();
    if (warp == 1) ptx::tmem_dealloc(tmem, 256);
}

template <int N, int ROWB>
void run(int ctas_per_sm, int distinct) {
    const int n_mma = 4096;
    int nsm = 148; cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, 0);
    const int grid = nsm * ctas_per_sm;
    const size_t smem = (size_t)distinct * (128 * ROWB + N * ROWB) + 2048;
    auto k = mma_rate_kernel<N, ROWB>;
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    long long* d; cudaMalloc(&d, grid * sizeof(long long));
    k<<<grid, 64, smem>>>(d, n_mma, distinct);
    k<<<grid, 64, smem>>>(d, n_mma, distinct);
    cudaError_t e = cudaDeviceSynchronize();
    long long h[1024]; cudaMemcpy(h, d, grid * sizeof(long long), cudaMemcpyDeviceToHost);
    double mx = 0; for (int i = 0; i < grid; ++i) mx = h[i] > mx ? h[i] : mx;
    const double cyc = mx / n_mma;                         // per MMA per CTA
    const double flop_per_clk_sm = 2.0 * 128 * N * 16 * ctas_per_sm / cyc;
    printf("N=%3d rowB=%3d ctas/SM=%d distinct=%d : %6.1f cyc/MMA/CTA  -> %6.0f flop/clk/SM (%s)\n", N, ROWB, ctas_per_sm, distinct,
           cyc, flop_per_clk_sm, cudaGetErrorString(e));
    cudaFree(d);
}

int main() {
    setvbuf(stdout, nullptr, _IONBF, 0);
    run<16, 64>(1, 2);  run<32, 64>(1, 2);  run<32, 128>(1, 2); run<64, 128>(1, 2); run<96, 128>(1, 2);
    run<128, 128>(1, 2); run<192, 128>(1, 2); run<256, 128>(1, 2);
    run<32, 64>(2, 2);  run<32, 128>(2, 2); run<64, 128>(2, 2); run<96, 128>(2, 2); run<128, 128>(2, 2); run<256, 128>(2, 2);
    run<32, 128>(1, 1); run<32, 128>(1, 4); run<96, 64>(1, 2); run<96, 64>(2, 2);
    return 0;
}
