// Correctness probe for the operand form the tcgen05 weight-gradient kernel relies on:
//   * MN-major A and B (the contraction index = voxel rows of the NDHWC tile, channels contiguous), SWIZZLE_64B,
//   * A with OVERLAPPING swizzle atoms along M: LBO = 64 bytes, i.e. M-atom j is the same tile shifted by j voxel rows
//     (one resident tile serves the kw = 0,1,2 filter taps in a single M=128 MMA),
//   * B with N = 96 = three separately loaded copies (LBO = copy pitch).
// One MMA (M=128, N=96, K=16) is compared with a CPU sum.  mode 0: overlapped A; mode 1: A as four copies (control).
//   build: make tests/cuda/mnmajor_test ; run: tests/cuda/mnmajor_test
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
#include "../../dsmnet_b200/csrc/ptx.cuh"
#include "../../dsmnet_b200/csrc/tma_host.cuh"

__device__ __forceinline__ uint64_t make_mn_desc_sw64(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    return (uint64_t)((saddr >> 4) & 0x3FFFu) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16) |
           ((uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32) | (1ull << 46) | (4ull << 61);
}

__global__ void __launch_bounds__(128) probe_kernel(const __grid_constant__ CUtensorMap map_g, const __grid_constant__ CUtensorMap map_x,
                                                    const __grid_constant__ CUtensorMap map_g16, float* out, int mode) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (ptx::smem_u32(smem_raw) + 1023u) & ~1023u;
    __shared__ __align__(8) uint64_t bar_tma, bar_mma;
    __shared__ uint32_t slot;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t sA = base, sB = base + 8192;
    if (threadIdx.x == 0) {
        ptx::mbar_init(ptx::smem_u32(&bar_tma), 1); ptx::mbar_init(ptx::smem_u32(&bar_mma), 1); ptx::fence_mbar_init();
    }
    if (warp == 1) ptx::tmem_alloc(ptx::smem_u32(&slot), 128);
    ptx::tc_fence_before(); __syncthreads(); ptx::tc_fence_after();
    const uint32_t tmem = slot;
    if (threadIdx.x == 0) {
        uint32_t bytes = 3 * 16 * 64;
        if (mode == 0) bytes += 24 * 64; else bytes += 4 * 16 * 64;
        ptx::mbar_arrive_expect_tx(ptx::smem_u32(&bar_tma), bytes);
        if (mode == 0) ptx::tma_load_2d(sA, &map_g, ptx::smem_u32(&bar_tma), 0, 0);          // gy rows 0..23
        else for (int j = 0; j < 4; ++j) ptx::tma_load_2d(sA + j * 1024, &map_g16, ptx::smem_u32(&bar_tma), 0, j);
        for (int c = 0; c < 3; ++c) ptx::tma_load_2d(sB + c * 1024, &map_x, ptx::smem_u32(&bar_tma), 0, c * 5);
        while (!ptx::mbar_try_wait(ptx::smem_u32(&bar_tma), 0)) {}
        ptx::tc_fence_after();
        const uint32_t idesc = ptx::make_idesc_bf16(96) | (1u << 15) | (1u << 16);
        const uint64_t ad = make_mn_desc_sw64(sA, mode == 0 ? 64u : 1024u, 512u);
        const uint64_t bd = make_mn_desc_sw64(sB, 1024u, 512u);
        ptx::umma_bf16(tmem, ad, bd, idesc, 0u);
        ptx::umma_commit(ptx::smem_u32(&bar_mma));
    }
    __syncwarp();
    while (!ptx::mbar_try_wait(ptx::smem_u32(&bar_mma), 0)) {}
    ptx::tc_fence_after();
    for (int c0 = 0; c0 < 96; c0 += 32) {
        uint32_t v[32];
        ptx::tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16) + c0, v);
        ptx::tc_wait_ld();
        for (int i = 0; i < 32; ++i) out[(warp * 32 + lane) * 96 + c0 + i] = __uint_as_float(v[i]);
    }
    ptx::tc_fence_before(); __syncthreads();
    if (warp == 1) ptx::tmem_dealloc(tmem, 128);
}

int main() {
    const int RG = 32, RX = 32, C = 32;
    std::vector<__nv_bfloat16> hg(RG * C), hx(RX * C);
    srand(3);
    for (auto& v : hg) v = __float2bfloat16((rand() % 17 - 8) / 8.f);
    for (auto& v : hx) v = __float2bfloat16((rand() % 13 - 6) / 4.f);
    __nv_bfloat16 *dg, *dx; float* dout;
    cudaMalloc(&dg, hg.size() * 2); cudaMalloc(&dx, hx.size() * 2); cudaMalloc(&dout, 128 * 96 * 4);
    cudaMemcpy(dg, hg.data(), hg.size() * 2, cudaMemcpyHostToDevice);
    cudaMemcpy(dx, hx.data(), hx.size() * 2, cudaMemcpyHostToDevice);
    CUtensorMap mg, mx, mg16;
    cuuint64_t dims[2] = {(cuuint64_t)C, (cuuint64_t)RG}, strides[1] = {(cuuint64_t)C * 2};
    cuuint32_t box24[2] = {32, 24}, box16[2] = {32, 16};
    bool ok = tma_host::encode(&mg, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, dg, 2, dims, strides, box24, CU_TENSOR_MAP_SWIZZLE_64B) &&
              tma_host::encode(&mg16, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, dg, 2, dims, strides, box16, CU_TENSOR_MAP_SWIZZLE_64B) &&
              tma_host::encode(&mx, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, dx, 2, dims, strides, box16, CU_TENSOR_MAP_SWIZZLE_64B);
    if (!ok) { printf("tensor map encode failed\n"); return 2; }
    int rc = 0;
    for (int mode = 0; mode < 2; ++mode) {
        cudaMemset(dout, 0, 128 * 96 * 4);
        probe_kernel<<<1, 128, 8192 + 4096 + 1024>>>(mg, mx, mg16, dout, mode);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("mode %d: %s\n", mode, cudaGetErrorString(e)); return 3; }
        std::vector<float> ho(128 * 96);
        cudaMemcpy(ho.data(), dout, ho.size() * 4, cudaMemcpyDeviceToHost);
        double maxerr = 0; int bad = 0;
        for (int j = 0; j < 4; ++j) for (int co = 0; co < 32; ++co) for (int c = 0; c < 3; ++c) for (int ci = 0; ci < 32; ++ci) {
            double ref = 0;
            for (int k = 0; k < 16; ++k) ref += (double)__bfloat162float(hg[(k + j) * C + co]) * (double)__bfloat162float(hx[(c * 5 + k) * C + ci]);
            const double got = ho[(j * 32 + co) * 96 + c * 32 + ci];
            const double err = fabs(got - ref);
            if (err > maxerr) maxerr = err;
            if (err > 1e-3) { if (bad < 5) printf("  mode %d mismatch j=%d co=%d c=%d ci=%d got %f ref %f\n", mode, j, co, c, ci, got, ref); ++bad; }
        }
        printf("mode %d (%s): max err %.3g, mismatches %d / %d\n", mode, mode == 0 ? "overlapped A atoms, LBO=64B" : "A as four copies", maxerr, bad, 128 * 96);
        if (bad) rc = 1;
    }
    return rc;
}
