// Shared helpers for the sm_100a kernels behind include/dsmnet_b200.h.
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include "../../include/dsmnet_b200.h"

#define DSM_NUM_SMS_B200 148

// Status of the library's own launch: cudaGetLastError clears a non-sticky error, so a stale error of an earlier,
// unrelated launch is reported once (to the call that finds it) and not attributed to every later dsm call.
static inline int dsm_launch_status() {
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? 0 : (int)e;
}

// Every entry point runs on the device that owns its first device pointer: cudaFuncSetAttribute, launches and
// cudaMemsetAsync act on the CURRENT device, which need not be the tensor's (a model on cuda:1 while cuda:0 is
// current).  The guard switches to the pointer's device for the duration of the call and restores the caller's.
struct DsmDeviceGuard {
    int prev = -1;
    bool switched = false;
    explicit DsmDeviceGuard(const void* p) {
        cudaPointerAttributes a;
        if (p && cudaPointerGetAttributes(&a, p) == cudaSuccess && a.type == cudaMemoryTypeDevice) {
            if (cudaGetDevice(&prev) == cudaSuccess && prev != a.device) switched = (cudaSetDevice(a.device) == cudaSuccess);
        } else {
            cudaGetLastError();                       // not a device pointer: the entry point's own checks report it
        }
    }
    ~DsmDeviceGuard() { if (switched) cudaSetDevice(prev); }
    DsmDeviceGuard(const DsmDeviceGuard&) = delete;
    DsmDeviceGuard& operator=(const DsmDeviceGuard&) = delete;
};

static inline bool dsm_aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

__host__ __device__ static inline int dsm_ceil_div(int a, int b) { return (a + b - 1) / b; }
__host__ __device__ static inline long long dsm_ceil_div_ll(long long a, long long b) { return (a + b - 1) / b; }

// streaming 128-bit global accesses: data touched once, keep it out of L1
__device__ __forceinline__ float4 ld_stream_f4(const float4* p) {
    float4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
    return r;
}
__device__ __forceinline__ void st_stream_f4(float4* p, const float4& v) {
    asm volatile("st.global.L1::no_allocate.v4.f32 [%0], {%1,%2,%3,%4};"
                 :: "l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ void st_stream_u4(uint4* p, const uint4& v) {
    asm volatile("st.global.L1::no_allocate.v4.u32 [%0], {%1,%2,%3,%4};"
                 :: "l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ float ld_stream_f1(const float* p) {
    float r;
    asm volatile("ld.global.nc.L1::no_allocate.f32 %0, [%1];" : "=f"(r) : "l"(p));
    return r;
}

// 256-bit global accesses (sm_100: LDG.256 / STG.256): one instruction per 32-byte sector, 32-byte aligned addresses
__device__ __forceinline__ void ld_nc_v8(const void* p, uint4& a, uint4& b) {
    asm volatile("ld.global.nc.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(a.x), "=r"(a.y), "=r"(a.z), "=r"(a.w), "=r"(b.x), "=r"(b.y), "=r"(b.z), "=r"(b.w) : "l"(p));
}
__device__ __forceinline__ void st_v8(void* p, const uint4& a, const uint4& b) {
    asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};"
                 :: "l"(p), "r"(a.x), "r"(a.y), "r"(a.z), "r"(a.w), "r"(b.x), "r"(b.y), "r"(b.z), "r"(b.w) : "memory");
}
static inline bool dsm_aligned32(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 31u) == 0; }

__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
    __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);   // .x = lo (low 16 bits)
    return *reinterpret_cast<uint32_t*>(&v);
}
// sm_100 packed fp32 arithmetic: two IEEE fp32 FMAs / adds per instruction (bit-identical to the scalar forms)
__device__ __forceinline__ void ffma2(float& d0, float& d1, float a0, float a1, float b0, float b1, float c0, float c1) {
    asm("{\n\t.reg .b64 ra, rb, rc, rd;\n\t"
        "mov.b64 ra, {%2, %3};\n\tmov.b64 rb, {%4, %5};\n\tmov.b64 rc, {%6, %7};\n\t"
        "fma.rn.f32x2 rd, ra, rb, rc;\n\t"
        "mov.b64 {%0, %1}, rd;\n\t}"
        : "=f"(d0), "=f"(d1) : "f"(a0), "f"(a1), "f"(b0), "f"(b1), "f"(c0), "f"(c1));
}
__device__ __forceinline__ void fadd2(float& d0, float& d1, float a0, float a1, float b0, float b1) {
    asm("{\n\t.reg .b64 ra, rb, rd;\n\t"
        "mov.b64 ra, {%2, %3};\n\tmov.b64 rb, {%4, %5};\n\t"
        "add.rn.f32x2 rd, ra, rb;\n\t"
        "mov.b64 {%0, %1}, rd;\n\t}"
        : "=f"(d0), "=f"(d1) : "f"(a0), "f"(a1), "f"(b0), "f"(b1));
}

// the same with ReLU folded into the conversion (F2FP.RELU): bf16(max(x, 0)) == max(bf16(x), 0), rounding is monotonic
__device__ __forceinline__ uint32_t pack_bf16x2_relu(float lo, float hi) {
    uint32_t r;
    asm("cvt.rn.relu.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));   // first source -> upper half
    return r;
}
__device__ __forceinline__ float bf16_lo(uint32_t v) { return __uint_as_float(v << 16); }
__device__ __forceinline__ float bf16_hi(uint32_t v) { return __uint_as_float(v & 0xffff0000u); }
