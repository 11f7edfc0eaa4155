// peaks_ffma_mufu.cu — measured CUDA-core peaks of this B200: fp32 FFMA, packed FFMA2 (fma.rn.f32x2), MUFU ex2 / rcp.
// MEASURED_PEAKS.json (driver-written) holds only HBM and bf16 tensor peaks; the LV ensemble kernels are bound by the FMA
// and MUFU pipes, so their roofline denominators are measured here.  Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a
// Prints one JSON object.  Every kernel keeps 16 independent dependency chains per thread, 8 warps x 8 blocks per SM.
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(e_)); exit(1); } } while (0)

constexpr int CH = 16;

__global__ void __launch_bounds__(256) k_ffma(float* out, int iters, float a, float b) {
    float acc[CH];
#pragma unroll
    for (int c = 0; c < CH; ++c) acc[c] = threadIdx.x * 1e-3f + c;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int c = 0; c < CH; ++c) acc[c] = fmaf(acc[c], a, b);
    }
    float s = 0.f;
#pragma unroll
    for (int c = 0; c < CH; ++c) s += acc[c];
    if (s == 123.456f) out[0] = s;
}

// three distinct register sources per FFMA (the shape of a dot product with register-resident operands)
__global__ void __launch_bounds__(256) k_ffma3(float* out, int iters, float a, float b) {
    float acc[CH], x[CH];
#pragma unroll
    for (int c = 0; c < CH; ++c) { acc[c] = threadIdx.x * 1e-3f + c; x[c] = a + c * b; }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int c = 0; c < CH; ++c) acc[c] = fmaf(x[c], x[(c + 1) % CH], acc[c]);
#pragma unroll
        for (int c = 0; c < CH; ++c) x[c] = fmaf(acc[c], 1e-9f, x[c]);
    }
    float s = 0.f;
#pragma unroll
    for (int c = 0; c < CH; ++c) s += acc[c] + x[c];
    if (s == 123.456f) out[0] = s;
}

__global__ void __launch_bounds__(256) k_ffma2(float* out, int iters, float a, float b) {
    unsigned long long acc[CH], pa, pb;
    asm("mov.b64 %0, {%1, %2};" : "=l"(pa) : "f"(a), "f"(a));
    asm("mov.b64 %0, {%1, %2};" : "=l"(pb) : "f"(b), "f"(b));
#pragma unroll
    for (int c = 0; c < CH; ++c) { float v = threadIdx.x * 1e-3f + c; asm("mov.b64 %0, {%1, %2};" : "=l"(acc[c]) : "f"(v), "f"(v + 0.5f)); }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int c = 0; c < CH; ++c) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(acc[c]) : "l"(pa), "l"(pb));
    }
    float s = 0.f;
#pragma unroll
    for (int c = 0; c < CH; ++c) { float lo, hi; asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(acc[c])); s += lo + hi; }
    if (s == 123.456f) out[0] = s;
}

__global__ void __launch_bounds__(256) k_ex2(float* out, int iters) {
    float acc[CH];
#pragma unroll
    for (int c = 0; c < CH; ++c) acc[c] = -(threadIdx.x * 1e-3f + c * 0.01f);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int c = 0; c < CH; ++c) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(acc[c]));
    }
    float s = 0.f;
#pragma unroll
    for (int c = 0; c < CH; ++c) s += acc[c];
    if (s == 123.456f) out[0] = s;
}

__global__ void __launch_bounds__(256) k_rcp(float* out, int iters) {
    float acc[CH];
#pragma unroll
    for (int c = 0; c < CH; ++c) acc[c] = 1.0f + threadIdx.x * 1e-3f + c * 0.01f;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int c = 0; c < CH; ++c) asm volatile("rcp.approx.ftz.f32 %0, %0;" : "+f"(acc[c]));
    }
    float s = 0.f;
#pragma unroll
    for (int c = 0; c < CH; ++c) s += acc[c];
    if (s == 123.456f) out[0] = s;
}

// the LV mix: 1 MUFU per 8 FMA-pipe instructions issued together (do the pipes overlap?)
__global__ void __launch_bounds__(256) k_mix(float* out, int iters, float a, float b) {
    float acc[CH], m[2];
#pragma unroll
    for (int c = 0; c < CH; ++c) acc[c] = threadIdx.x * 1e-3f + c;
    m[0] = -0.3f; m[1] = -0.7f;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int c = 0; c < CH; ++c) acc[c] = fmaf(acc[c], a, b);
        asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(m[0]));
        asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(m[1]));
    }
    float s = m[0] + m[1];
#pragma unroll
    for (int c = 0; c < CH; ++c) s += acc[c];
    if (s == 123.456f) out[0] = s;
}

template <class F> float time_ms(F&& launch, int reps) {
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    for (int i = 0; i < 3; ++i) launch();
    CK(cudaDeviceSynchronize());
    float best = 1e30f;
    for (int r = 0; r < reps; ++r) {
        CK(cudaEventRecord(e0)); launch(); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); if (ms < best) best = ms;
    }
    return best;
}

int main() {
    int dev = 0; CK(cudaSetDevice(dev));
    cudaDeviceProp pr; CK(cudaGetDeviceProperties(&pr, dev));
    int clk_khz = 0; CK(cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, dev));
    float* out; CK(cudaMalloc(&out, 4));
    const int sms = pr.multiProcessorCount, blocks = sms * 8, thr = 256, iters = 4096;
    const double n = (double)blocks * thr * iters * CH;
    const float t_ffma = time_ms([&] { k_ffma<<<blocks, thr>>>(out, iters, 1.0000001f, 1e-7f); }, 10);
    const float t_ffma3 = time_ms([&] { k_ffma3<<<blocks, thr>>>(out, iters, 1.0000001f, 1e-7f); }, 10);
    const float t_ffma2 = time_ms([&] { k_ffma2<<<blocks, thr>>>(out, iters, 1.0000001f, 1e-7f); }, 10);
    const float t_ex2 = time_ms([&] { k_ex2<<<blocks, thr>>>(out, iters); }, 10);
    const float t_rcp = time_ms([&] { k_rcp<<<blocks, thr>>>(out, iters); }, 10);
    const float t_mix = time_ms([&] { k_mix<<<blocks, thr>>>(out, iters, 1.0000001f, 1e-7f); }, 10);
    CK(cudaGetLastError());
    printf("{\"gpu\": \"%s\", \"sms\": %d, \"clock_rate_mhz\": %.0f, "
           "\"ffma_tflops\": %.2f, \"ffma_3src_tflops\": %.2f, \"ffma2_tflops\": %.2f, \"mufu_ex2_tops\": %.3f, \"mufu_rcp_tops\": %.3f, "
           "\"mix_16ffma_2ex2\": {\"ffma_tflops\": %.2f, \"ex2_tops\": %.3f}, "
           "\"derived\": {\"ffma_tflops_at_clock_rate\": %.2f, \"mufu_tops_at_clock_rate\": %.3f}, "
           "\"how\": \"16 independent chains per thread, 8 blocks x 256 threads per SM, 4096 iterations, best of 10, CUDA events; ffma_3src has three distinct register sources per FFMA (2 fp32 ops each, plus the x update counted too)\"}\n",
           pr.name, sms, clk_khz / 1e3,
           2 * n / (t_ffma * 1e-3) / 1e12, 2 * (2 * n) / (t_ffma3 * 1e-3) / 1e12, 4 * n / (t_ffma2 * 1e-3) / 1e12, n / (t_ex2 * 1e-3) / 1e12,
           n / (t_rcp * 1e-3) / 1e12, 2 * n / (t_mix * 1e-3) / 1e12, (n / 8) / (t_mix * 1e-3) / 1e12,
           sms * 128.0 * 2 * clk_khz * 1e3 / 1e12, sms * 16.0 * clk_khz * 1e3 / 1e12);
    return 0;
}
