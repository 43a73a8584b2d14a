// Stand-alone GPU self-test of the tcgen05 conv3d path (no Python): runs dsm_conv3d_fwd_ex on
// random bf16 data and compares with a naive direct-convolution CUDA kernel that restates the
// definition (reference semantics: nn.Conv3d k3 p1 s{1,2}, nn.ConvTranspose3d k3 s2 p1 op1,
// models/psmnet/submodule.py:16-19, stackhourglass.py:37-41).  Also prints CUDA-event timings
// of PSMNet-sized layers.  Build: see Makefile (tests/cuda/conv3d_selftest).
//   usage: conv3d_selftest [quick|shift|time|timeshift|full]
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cstdint>
#include <ctime>
#include <unistd.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cmath>
#include <vector>
#include "../../include/dsmnet_b200.h"

extern "C" int dsm_conv3d_fwd_ex(const void*, const void*, const float*, const float*, const void*, void*,
                                 int, int, int, int, int, int, int, int, int, int, int, int, int, int, void*);
extern "C" int dsm_debug_conv_timeouts(void);
extern "C" int dsm_debug_conv_set_progress(int*);
static volatile int* g_prog = nullptr;

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(2); } } while (0)

static uint32_t rng_state = 12345u;
static float frand() { rng_state = rng_state * 1664525u + 1013904223u; return ((rng_state >> 8) & 0xffff) / 65536.0f - 0.5f; }

// naive reference: one thread per (b, od, oh, ow, co); x padded NDHWC bf16; w packed [27][CoutP][Cin]
__global__ void ref_conv(const __nv_bfloat16* x, const __nv_bfloat16* w, const float* scale, const float* shift,
                         const __nv_bfloat16* res_b, const float* res_f, float* out,
                         int B, int Cin, int Cout, int CoutP, int D, int H, int W, int Do, int Ho, int Wo,
                         int stride, int transposed, int relu) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    long long n = (long long)B * Do * Ho * Wo * Cout;
    if (i >= n) return;
    int co = i % Cout; long long t = i / Cout;
    int ow = t % Wo; t /= Wo; int oh = t % Ho; t /= Ho; int od = t % Do; int b = t / Do;
    const int Hp = H + 2, Wp = W + 2, Dp = D + 2;
    float acc = 0.f;
    for (int kd = 0; kd < 3; ++kd) for (int kh = 0; kh < 3; ++kh) for (int kw = 0; kw < 3; ++kw) {
        int id, ih, iw;
        if (!transposed) { id = od * stride + kd - 1; ih = oh * stride + kh - 1; iw = ow * stride + kw - 1; }
        else {
            int a = od + 1 - kd, bb = oh + 1 - kh, c = ow + 1 - kw;
            if ((a & 1) || (bb & 1) || (c & 1)) continue;
            id = a / 2; ih = bb / 2; iw = c / 2;
        }
        if (id < 0 || id >= D || ih < 0 || ih >= H || iw < 0 || iw >= W) continue;
        const __nv_bfloat16* xr = x + ((((size_t)b * Dp + id + 1) * Hp + ih + 1) * Wp + iw + 1) * Cin;
        const __nv_bfloat16* wr = w + ((size_t)((kd * 3 + kh) * 3 + kw) * CoutP + co) * Cin;
        for (int ci = 0; ci < Cin; ++ci) acc += __bfloat162float(xr[ci]) * __bfloat162float(wr[ci]);
    }
    float v = acc * (scale ? scale[co] : 1.f) + (shift ? shift[co] : 0.f);
    if (relu == 2) v = fmaxf(v, 0.f);
    if (res_b) v += __bfloat162float(res_b[((((size_t)b * (Do + 2) + od + 1) * (Ho + 2) + oh + 1) * (Wo + 2) + ow + 1) * Cout + co]);
    if (res_f) v += res_f[(((size_t)b * Do + od) * Ho + oh) * Wo + ow];
    if (relu == 1) v = fmaxf(v, 0.f);
    out[i] = v;
}

// gather the library's output into the same dense [B][Do][Ho][Wo][Cout] fp32 order
__global__ void gather_out(const __nv_bfloat16* yb, const float* yf, float* out, int B, int Cout, int Do, int Ho, int Wo) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    long long n = (long long)B * Do * Ho * Wo * Cout;
    if (i >= n) return;
    int co = i % Cout; long long t = i / Cout;
    int ow = t % Wo; t /= Wo; int oh = t % Ho; t /= Ho; int od = t % Do; int b = t / Do;
    if (yf) out[i] = yf[(((size_t)b * Do + od) * Ho + oh) * Wo + ow];
    else out[i] = __bfloat162float(yb[((((size_t)b * (Do + 2) + od + 1) * (Ho + 2) + oh + 1) * (Wo + 2) + ow + 1) * Cout + co]);
}

struct Case { int B, Cin, Cout, D, H, W, stride, transposed, relu, use_res, use_affine, f32, variant; const char* name; };

static int run_case(const Case& c, bool timing, int reps) {
    const int CoutP = c.f32 ? 16 : c.Cout;
    int Do, Ho, Wo;
    if (c.transposed) { Do = 2 * c.D; Ho = 2 * c.H; Wo = 2 * c.W; }
    else if (c.stride == 2) { Do = (c.D - 1) / 2 + 1; Ho = (c.H - 1) / 2 + 1; Wo = (c.W - 1) / 2 + 1; }
    else { Do = c.D; Ho = c.H; Wo = c.W; }
    const size_t nx = (size_t)c.B * (c.D + 2) * (c.H + 2) * (c.W + 2) * c.Cin;
    const size_t nw = (size_t)27 * CoutP * c.Cin;
    const size_t nyp = (size_t)c.B * (Do + 2) * (Ho + 2) * (Wo + 2) * c.Cout;
    const size_t nyd = (size_t)c.B * Do * Ho * Wo * c.Cout;

    std::vector<__nv_bfloat16> hx(nx), hw(nw), hres;
    for (size_t i = 0; i < nx; ++i) hx[i] = __float2bfloat16(0.f);
    for (int b = 0; b < c.B; ++b) for (int d = 1; d <= c.D; ++d) for (int h = 1; h <= c.H; ++h) for (int w = 1; w <= c.W; ++w) {
        size_t o = ((((size_t)b * (c.D + 2) + d) * (c.H + 2) + h) * (c.W + 2) + w) * c.Cin;
        for (int ci = 0; ci < c.Cin; ++ci) hx[o + ci] = __float2bfloat16(frand());
    }
    for (size_t i = 0; i < nw; ++i) hw[i] = __float2bfloat16(0.f);
    for (int t = 0; t < 27; ++t) for (int co = 0; co < c.Cout; ++co) for (int ci = 0; ci < c.Cin; ++ci)
        hw[((size_t)t * CoutP + co) * c.Cin + ci] = __float2bfloat16(frand() * 0.25f);
    std::vector<float> hscale(CoutP, 1.f), hshift(CoutP, 0.f);
    for (int i = 0; i < c.Cout; ++i) { hscale[i] = 0.5f + frand(); hshift[i] = frand(); }

    __nv_bfloat16 *dx, *dw, *dyb = nullptr, *dresb = nullptr; float *dscale, *dshift, *dyf = nullptr, *dresf = nullptr, *dref, *dgot;
    CK(cudaMalloc(&dx, nx * 2)); CK(cudaMalloc(&dw, nw * 2));
    CK(cudaMalloc(&dscale, CoutP * 4)); CK(cudaMalloc(&dshift, CoutP * 4));
    CK(cudaMalloc(&dref, nyd * 4)); CK(cudaMalloc(&dgot, nyd * 4));
    CK(cudaMemcpy(dx, hx.data(), nx * 2, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dw, hw.data(), nw * 2, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dscale, hscale.data(), CoutP * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dshift, hshift.data(), CoutP * 4, cudaMemcpyHostToDevice));
    if (c.f32) {
        CK(cudaMalloc(&dyf, nyd * 4)); CK(cudaMemset(dyf, 0, nyd * 4));
        if (c.use_res) { std::vector<float> r(nyd); for (auto& v : r) v = frand(); CK(cudaMalloc(&dresf, nyd * 4)); CK(cudaMemcpy(dresf, r.data(), nyd * 4, cudaMemcpyHostToDevice)); }
    } else {
        CK(cudaMalloc(&dyb, nyp * 2)); CK(cudaMemset(dyb, 0, nyp * 2));
        if (c.use_res) { hres.resize(nyp); for (auto& v : hres) v = __float2bfloat16(frand()); CK(cudaMalloc(&dresb, nyp * 2)); CK(cudaMemcpy(dresb, hres.data(), nyp * 2, cudaMemcpyHostToDevice)); }
    }
    const float* sc = c.use_affine ? dscale : nullptr; const float* sh = c.use_affine ? dshift : nullptr;
    void* y = c.f32 ? (void*)dyf : (void*)dyb;
    const void* res = c.f32 ? (const void*)dresf : (const void*)dresb;

    printf("  launching %s (variant 0x%x)...\n", c.name, c.variant);
    int rc = dsm_conv3d_fwd_ex(dx, dw, sc, sh, res, y, c.B, c.Cin, c.Cout, c.D, c.H, c.W, c.stride, c.transposed, c.relu,
                               c.f32 ? DSM_F32 : DSM_BF16, 0, 0, 0, c.variant, nullptr);
    {   // watchdog: never sit on a hung kernel; report how far CTA 1 got and leave
        double waited = 0;
        while (cudaStreamQuery(nullptr) == cudaErrorNotReady) {
            struct timespec ts = {0, 20000000}; nanosleep(&ts, nullptr); waited += 0.02;
            if (waited > 6.0) {
                printf("  HUNG after %.1fs: progress per warp of CTA1 = %d %d %d %d\n", waited, g_prog ? g_prog[0] : -1, g_prog ? g_prog[1] : -1, g_prog ? g_prog[2] : -1, g_prog ? g_prog[3] : -1);
                _exit(3);
            }
        }
    }
    cudaError_t se = cudaDeviceSynchronize();
    printf("  progress per warp of CTA1 = %d %d %d %d\n", g_prog ? g_prog[0] : -1, g_prog ? g_prog[1] : -1, g_prog ? g_prog[2] : -1, g_prog ? g_prog[3] : -1);
    printf("  returned rc=%d sync=%s timeouts=%d\n", rc, cudaGetErrorString(se), dsm_debug_conv_timeouts());
    if (rc != 0 || se != cudaSuccess) {
        printf("CASE %-34s v%d : LAUNCH FAILED rc=%d (%s) sync=%s\n", c.name, c.variant, rc, dsm_strerror(rc), cudaGetErrorString(se));
        return 1;
    }
    int fail = 0;
    if (!timing) {
        const long long n = (long long)nyd;
        ref_conv<<<(unsigned)((n + 255) / 256), 256>>>(dx, dw, sc, sh, dresb, dresf, dref, c.B, c.Cin, c.Cout, CoutP, c.D, c.H, c.W, Do, Ho, Wo, c.stride, c.transposed, c.relu);
        gather_out<<<(unsigned)((n + 255) / 256), 256>>>(dyb, dyf, dgot, c.B, c.Cout, Do, Ho, Wo);
        CK(cudaDeviceSynchronize());
        std::vector<float> a(nyd), b(nyd);
        CK(cudaMemcpy(a.data(), dref, nyd * 4, cudaMemcpyDeviceToHost));
        CK(cudaMemcpy(b.data(), dgot, nyd * 4, cudaMemcpyDeviceToHost));
        double maxd = 0, maxr = 0; size_t bad = 0, arg = 0;
        for (size_t i = 0; i < nyd; ++i) {
            double d = fabs((double)a[i] - b[i]);
            double tol = (c.f32 ? 2e-3 : 1e-2) * (1.0 + fabs((double)a[i]));   // bf16 output rounding: 2^-8 relative
            if (d > maxd) { maxd = d; arg = i; }
            if (fabs(a[i]) > 1e-3) maxr = fmax(maxr, d / fabs(a[i]));
            if (!(d <= tol)) ++bad;
        }
        fail = bad != 0;
        printf("CASE %-34s v%d : %s  max|d|=%.4g (ref %.4g got %.4g @%zu) bad=%zu/%zu timeouts=%d\n", c.name, c.variant,
               fail ? "FAIL" : "PASS", maxd, a[arg], b[arg], arg, bad, nyd, dsm_debug_conv_timeouts());
    } else {
        cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
        for (int i = 0; i < 3; ++i)
            dsm_conv3d_fwd_ex(dx, dw, sc, sh, res, y, c.B, c.Cin, c.Cout, c.D, c.H, c.W, c.stride, c.transposed, c.relu, c.f32 ? DSM_F32 : DSM_BF16, 0, 0, 0, c.variant, nullptr);
        CK(cudaEventRecord(e0));
        for (int i = 0; i < reps; ++i)
            dsm_conv3d_fwd_ex(dx, dw, sc, sh, res, y, c.B, c.Cin, c.Cout, c.D, c.H, c.W, c.stride, c.transposed, c.relu, c.f32 ? DSM_F32 : DSM_BF16, 0, 0, 0, c.variant, nullptr);
        CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); ms /= reps;
        const double vox = (double)c.B * (c.transposed ? (double)c.D * c.H * c.W : (double)Do * Ho * Wo);
        const double flops = 2.0 * 27 * c.Cin * c.Cout * vox;
        printf("TIME %-34s v%d : %.3f ms  %.1f TFLOP/s (algorithmic) timeouts=%d", c.name, c.variant, ms, flops / ms * 1e-9, dsm_debug_conv_timeouts());
        if (g_prog && g_prog[1] > 0) printf("  [CTA0: %.1f us at %.0f MHz]", g_prog[1] * 16e-3, 1e3 * g_prog[0] / (double)g_prog[1]);
        printf("\n");
        if (g_prog) g_prog[0] = g_prog[1] = 0;
    }
    cudaFree(dx); cudaFree(dw); cudaFree(dscale); cudaFree(dshift); cudaFree(dref); cudaFree(dgot);
    if (dyb) cudaFree(dyb); if (dyf) cudaFree(dyf); if (dresb) cudaFree(dresb); if (dresf) cudaFree(dresf);
    return fail;
}

__global__ void hello_kernel(int* p) { if (threadIdx.x == 0) *p = 42; }

static int g_level = 0, g_only = -1;

int main(int argc, char** argv) {
    setvbuf(stdout, nullptr, _IONBF, 0);
    const char* what = argc > 1 ? argv[1] : "quick";
    g_level = argc > 2 ? atoi(argv[2]) : 0;      // bring-up level (see conv3d.cu debug_level)
    g_only = argc > 3 ? atoi(argv[3]) : -1;      // run only this case index
    int fails = 0;
    printf("selftest mode=%s level=%d only=%d\n", what, g_level, g_only);
    {
        int* d; CK(cudaMalloc(&d, 4)); hello_kernel<<<1, 32>>>(d); CK(cudaDeviceSynchronize());
        int h = 0; CK(cudaMemcpy(&h, d, 4, cudaMemcpyDeviceToHost)); cudaFree(d);
        printf("context up, hello kernel -> %d\n", h);
        int* hp = nullptr; CK(cudaHostAlloc(&hp, 4 * sizeof(int), cudaHostAllocMapped));
        for (int i = 0; i < 4; ++i) hp[i] = 0;
        int* dp = nullptr; CK(cudaHostGetDevicePointer(&dp, hp, 0));
        g_prog = hp; dsm_debug_conv_set_progress(dp);
        if (!strcmp(what, "hello")) return h == 42 ? 0 : 1;
    }
    if (!strcmp(what, "quick") || !strcmp(what, "full")) {
        std::vector<Case> cases = {
            // B Cin Cout D  H  W  s  T relu res aff f32 var
            {1, 32, 32, 4, 6, 20, 1, 0, 0, 0, 0, 0, 0, "s1 32->32 plain"},
            {1, 32, 32, 4, 6, 20, 1, 0, 1, 1, 1, 0, 0, "s1 32->32 affine+res+relu"},
            {2, 64, 32, 5, 7, 19, 1, 0, 1, 0, 1, 0, 0, "s1 64->32 B=2 odd dims"},
            {1, 64, 64, 4, 6, 20, 1, 0, 1, 1, 1, 0, 0, "s1 64->64"},
            {1, 32, 1, 4, 6, 20, 1, 0, 0, 1, 0, 1, 0, "s1 32->1 fp32 +res"},
            {1, 32, 64, 8, 12, 40, 2, 0, 1, 0, 1, 0, 0, "s2 32->64"},
            {1, 64, 64, 7, 11, 37, 2, 0, 1, 0, 1, 0, 0, "s2 64->64 odd dims"},
            {1, 64, 64, 3, 5, 10, 2, 1, 1, 1, 1, 0, 0, "deconv 64->64 +res"},
            {1, 64, 32, 3, 5, 10, 2, 1, 0, 1, 1, 0, 0, "deconv 64->32 +res"},
            {1, 32, 1, 3, 5, 10, 2, 1, 0, 0, 0, 1, 0, "deconv 32->1 fp32"},
            {1, 128, 128, 3, 4, 9, 1, 0, 1, 0, 1, 0, 0, "s1 128->128 (2 K chunks)"},
            {1, 64, 128, 6, 8, 18, 2, 0, 1, 0, 1, 0, 0, "s2 64->128"},
            {1, 128, 64, 3, 4, 9, 2, 1, 1, 1, 1, 0, 0, "deconv 128->64"},
            // plane-sharing kernel: several bands, a partial last band, single / two planes
            {1, 32, 32, 11, 6, 20, 1, 0, 1, 1, 1, 0, 0, "s1 32->32 D=11 (8+3 planes)"},
            {2, 32, 32, 17, 5, 33, 1, 0, 1, 0, 1, 0, 0, "s1 32->32 B=2 D=17"},
            {1, 64, 32, 9, 5, 23, 1, 0, 1, 1, 1, 0, 0, "s1 64->32 D=9"},
            {1, 32, 1, 10, 7, 21, 1, 0, 0, 1, 0, 1, 0, "s1 32->1 fp32 D=10"},
            {1, 32, 32, 1, 5, 9, 1, 0, 0, 0, 0, 0, 0, "s1 32->32 D=1"},
            {1, 32, 32, 2, 5, 9, 1, 0, 0, 0, 1, 0, 0, "s1 32->32 D=2"},
            {1, 32, 16, 8, 9, 40, 1, 0, 1, 0, 1, 0, 0, "s1 32->16 D=8"},
            {1, 32, 32, 24, 20, 150, 1, 0, 1, 1, 1, 0, 0, "s1 32->32 24x20x150 (many tiles)"},
            // ReLU before the residual add (GC-Net skip connections)
            {1, 64, 32, 3, 5, 10, 2, 1, 2, 1, 1, 0, 0, "deconv 64->32 relu-then-add"},
            {1, 128, 64, 3, 4, 9, 2, 1, 2, 1, 1, 0, 0, "deconv 128->64 relu-then-add"},
            {1, 32, 32, 4, 6, 20, 1, 0, 2, 1, 1, 0, 0, "s1 32->32 relu-then-add"},
            // class-sharing transposed-conv kernel
            {2, 64, 32, 5, 6, 21, 2, 1, 1, 1, 1, 0, 0, "deconv 64->32 B=2 5x6x21"},
            {1, 32, 32, 4, 9, 40, 2, 1, 0, 0, 0, 0, 0, "deconv 32->32 plain"},
            {1, 32, 16, 6, 7, 30, 2, 1, 1, 1, 1, 0, 0, "deconv 32->16"},
            {1, 64, 1, 4, 6, 17, 2, 1, 0, 1, 1, 1, 0, "deconv 64->1 fp32 +res"},
            {1, 64, 32, 12, 24, 78, 2, 1, 0, 1, 1, 0, 0, "deconv 64->32 12x24x78 (many tiles)"},
        };
        { int idx = 0; for (auto& c : cases) { if (g_only < 0 || g_only == idx) { Case cc = c; cc.variant |= g_level << 8; fails += run_case(cc, false, 0); } ++idx; } }
    }
    if (!strcmp(what, "shift") || !strcmp(what, "full")) {
        // the per-tap kernel (variant bit1) — the default (variant 0) is the row-shifted-descriptor kernel
        std::vector<Case> cases = {
            {1, 32, 32, 4, 6, 20, 1, 0, 1, 1, 1, 0, 2, "s1 32->32 per-tap"},
            {1, 64, 32, 5, 7, 19, 1, 0, 1, 0, 1, 0, 2, "s1 64->32 per-tap"},
            {1, 64, 64, 4, 6, 20, 1, 0, 1, 1, 1, 0, 2, "s1 64->64 per-tap"},
            // variant bit3: the per-tile row-shifted-descriptor kernel instead of the plane-sharing kernel
            {1, 32, 32, 4, 6, 20, 1, 0, 1, 1, 1, 0, 8, "s1 32->32 per-tile shift"},
            {1, 64, 32, 5, 7, 19, 1, 0, 1, 0, 1, 0, 8, "s1 64->32 per-tile shift"},
            {1, 32, 1, 4, 6, 20, 1, 0, 0, 1, 0, 1, 8, "s1 32->1 per-tile shift"},
            {1, 64, 32, 3, 5, 10, 2, 1, 0, 1, 1, 0, 8, "deconv 64->32 per-class"},
            {1, 32, 1, 3, 5, 10, 2, 1, 0, 0, 0, 1, 8, "deconv 32->1 per-class"},
        };
        { int idx = 0; for (auto& c : cases) { if (g_only < 0 || g_only == idx) { Case cc = c; cc.variant |= g_level << 8; fails += run_case(cc, false, 0); } ++idx; } }
    }
    if (!strcmp(what, "time") || !strcmp(what, "full")) {
        std::vector<Case> cases = {
            {1, 64, 32, 48, 96, 312, 1, 0, 1, 0, 1, 0, 0, "dres0.0 64->32 @48x96x312"},
            {1, 32, 32, 48, 96, 312, 1, 0, 1, 1, 1, 0, 0, "32->32 @48x96x312"},
            {1, 32, 1, 48, 96, 312, 1, 0, 0, 1, 0, 1, 0, "classif 32->1 @48x96x312"},
            {1, 32, 64, 48, 96, 312, 2, 0, 1, 0, 1, 0, 0, "conv1 s2 32->64"},
            {1, 64, 64, 24, 48, 156, 1, 0, 1, 1, 1, 0, 0, "conv2 64->64 @24x48x156"},
            {1, 64, 64, 24, 48, 156, 2, 0, 1, 0, 1, 0, 0, "conv3 s2 64->64"},
            {1, 64, 64, 12, 24, 78, 1, 0, 1, 0, 1, 0, 0, "conv4 64->64 @12x24x78"},
            {1, 64, 64, 12, 24, 78, 2, 1, 1, 1, 1, 0, 0, "conv5 deconv 64->64"},
            {1, 64, 32, 24, 48, 156, 2, 1, 0, 1, 1, 0, 0, "conv6 deconv 64->32"},
        };
        { int idx = 0; for (auto& c : cases) { if (g_only < 0 || g_only == idx) { Case cc = c; cc.variant |= g_level << 8; fails += run_case(cc, true, 10); } ++idx; } }
    }
    if (!strcmp(what, "timeshift") || !strcmp(what, "full")) {
        std::vector<Case> cases = {
            {1, 64, 32, 48, 96, 312, 1, 0, 1, 0, 1, 0, 2, "dres0.0 64->32 per-tap"},
            {1, 32, 32, 48, 96, 312, 1, 0, 1, 1, 1, 0, 2, "32->32 per-tap"},
            {1, 64, 64, 24, 48, 156, 1, 0, 1, 1, 1, 0, 2, "conv2 64->64 per-tap"},
            {1, 64, 32, 48, 96, 312, 1, 0, 1, 0, 1, 0, 8, "dres0.0 64->32 per-tile shift"},
            {1, 32, 32, 48, 96, 312, 1, 0, 1, 1, 1, 0, 8, "32->32 per-tile shift"},
            {1, 32, 1, 48, 96, 312, 1, 0, 0, 1, 0, 1, 8, "classif 32->1 per-tile shift"},
            {1, 64, 32, 24, 48, 156, 2, 1, 0, 1, 1, 0, 8, "conv6 deconv 64->32 per-class"},
            {1, 64, 64, 12, 24, 78, 1, 0, 1, 0, 1, 0, 4, "conv4 64->64 @12x24x78 per-tile shift"},
            {1, 64, 64, 12, 24, 78, 1, 0, 1, 0, 1, 0, 0, "conv4 64->64 @12x24x78 default"},
        };
        { int idx = 0; for (auto& c : cases) { if (g_only < 0 || g_only == idx) { Case cc = c; cc.variant |= g_level << 8; fails += run_case(cc, true, 10); } ++idx; } }
    }
    if (!strcmp(what, "dbg48")) {
        Case c = {1, 32, 32, 48, 96, 312, 1, 0, 1, 1, 1, 0, 48, "32->32 dbg48"};
        run_case(c, true, 10);
        Case c2 = {1, 32, 32, 48, 96, 312, 1, 0, 1, 1, 1, 0, 32, "32->32 dbg32 (no TMA)"};
        run_case(c2, true, 10);
    }
    if (!strcmp(what, "dbgdc2")) {
        // is the class-sharing kernel's epilogue bound by its residual loads or by its stores?
        for (int v : {0, 96}) {
            Case c = {1, 64, 32, 24, 48, 156, 2, 1, 0, 1, 1, 0, v, "conv6 deconv 64->32 +res"};
            run_case(c, true, 10);
            Case c2 = {1, 64, 32, 24, 48, 156, 2, 1, 0, 0, 1, 0, v, "conv6 deconv 64->32 no res"};
            run_case(c2, true, 10);
        }
    }
    if (!strcmp(what, "dbgdc")) {
        // timing experiments on the class-sharing transposed-conv kernel (results are garbage for v != 0)
        for (int v : {0, 16, 32, 64, 48, 80, 96, 112}) {
            Case c = {1, 64, 32, 24, 48, 156, 2, 1, 0, 1, 1, 0, v, "conv6 deconv 64->32 dbg"};
            run_case(c, true, 10);
            Case c2 = {1, 32, 1, 96, 128, 256, 2, 1, 0, 0, 0, 1, v, "l37 deconv 32->1 dbg"};
            run_case(c2, true, 5);
        }
    }
    if (!strcmp(what, "dbg")) {
        // timing experiments on the plane-sharing kernel (results are garbage): which part of the pipeline bounds it?
        for (int v : {0, 16, 32, 64, 48, 80, 96, 112}) {
            Case c = {1, 32, 32, 48, 96, 312, 1, 0, 1, 1, 1, 0, v, "32->32 dbg"};
            run_case(c, true, 10);
            Case c2 = {1, 64, 32, 48, 96, 312, 1, 0, 1, 0, 1, 0, v, "64->32 dbg"};
            run_case(c2, true, 10);
            Case c3 = {1, 32, 1, 48, 96, 312, 1, 0, 0, 1, 0, 1, v, "classif 32->1 dbg"};
            run_case(c3, true, 10);
        }
    }
    printf("SELFTEST %s: %d failing case(s), timeouts=%d\n", what, fails, dsm_debug_conv_timeouts());
    return fails ? 1 : 0;
}
