// mcl_peaks.cu -- issue-rate microbenchmarks for the roofline denominators of the kinetics kernel.
// The path is bound by SFU (MUFU.LG2/EX2) and FP32/INT32 issue, not by HBM, so the "peak" the
// throughput is compared against has to be measured on the same device: giga lane-ops per second
// for MUFU, FFMA, IMAD.WIDE (the Philox multiply) and LOP3.
#include "mcl_common.cuh"

namespace mcl {
namespace {

constexpr int ITERS = 4096;
constexpr int CHAINS = 8;

__global__ void __launch_bounds__(256) peak_mufu(float *out, float seed)
{
    float a[CHAINS];
#pragma unroll
    for (int c = 0; c < CHAINS; c++) a[c] = seed + 0.001f * (threadIdx.x + c);
    for (int i = 0; i < ITERS; i++) {
#pragma unroll
        for (int c = 0; c < CHAINS; c++) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a[c]));
#pragma unroll
        for (int c = 0; c < CHAINS; c++) asm volatile("lg2.approx.ftz.f32 %0, %0;" : "+f"(a[c]));
    }
    float s = 0.f;
#pragma unroll
    for (int c = 0; c < CHAINS; c++) s += a[c];
    if (s == 12345.678f) out[0] = s;
}

__global__ void __launch_bounds__(256) peak_ffma(float *out, float seed)
{
    float a[CHAINS];
#pragma unroll
    for (int c = 0; c < CHAINS; c++) a[c] = seed + 0.001f * (threadIdx.x + c);
    const float m = 1.0000001f + seed * 1e-9f, b = seed * 1e-7f;
    for (int i = 0; i < ITERS; i++) {
#pragma unroll
        for (int k = 0; k < 2; k++)
#pragma unroll
            for (int c = 0; c < CHAINS; c++) a[c] = fmaf(a[c], m, b);
    }
    float s = 0.f;
#pragma unroll
    for (int c = 0; c < CHAINS; c++) s += a[c];
    if (s == 12345.678f) out[0] = s;
}

__global__ void __launch_bounds__(256) peak_imad(unsigned *out, unsigned seed)
{
    unsigned a[CHAINS];
#pragma unroll
    for (int c = 0; c < CHAINS; c++) a[c] = seed + threadIdx.x * 2654435761u + c;
    for (int i = 0; i < ITERS; i++) {
#pragma unroll
        for (int k = 0; k < 2; k++)
#pragma unroll
            for (int c = 0; c < CHAINS; c++) {
                unsigned long long p = (unsigned long long)0xD2511F53u * a[c];     // IMAD.WIDE.U32
                a[c] = (unsigned)(p >> 32) + (unsigned)p;                          // + IADD (alu pipe)
            }
    }
    unsigned s = 0;
#pragma unroll
    for (int c = 0; c < CHAINS; c++) s ^= a[c];
    if (s == 0x12345678u) out[0] = s;
}

__global__ void __launch_bounds__(256) peak_lop3(unsigned *out, unsigned seed)
{
    unsigned a[CHAINS];
#pragma unroll
    for (int c = 0; c < CHAINS; c++) a[c] = seed + threadIdx.x * 2654435761u + c;
    unsigned k0 = seed * 3u + 1u, k1 = seed * 5u + 7u;
    for (int i = 0; i < ITERS; i++) {
#pragma unroll
        for (int k = 0; k < 2; k++)
#pragma unroll
            for (int c = 0; c < CHAINS; c++)
                asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a[c]) : "r"(k0), "r"(k1));   // 3-input xor
    }
    unsigned s = 0;
#pragma unroll
    for (int c = 0; c < CHAINS; c++) s ^= a[c];
    if (s == 0x12345678u) out[0] = s;
}

template <typename F>
static double time_ms(F launch)
{
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    launch(); launch();                       // warm-up
    cudaDeviceSynchronize();
    float best = 1e30f;
    for (int rep = 0; rep < 5; rep++) {
        cudaEventRecord(e0);
        launch();
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms = 0.f; cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
    }
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    return (double)best;
}

}  // namespace
}  // namespace mcl

extern "C" int mcl_device_peaks(mcl_peaks *out)
{
    using namespace mcl;
    if (!out) { set_error("mcl_device_peaks: null output"); return MCL_ERR_ARG; }
    int dev = 0, n_sm = 0, khz = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) { set_error("no CUDA device"); return MCL_ERR_CUDA; }
    cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev);
    cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, dev);
    float *fbuf = nullptr;
    if (cudaMalloc(&fbuf, 256) != cudaSuccess) { set_error("cudaMalloc failed"); return MCL_ERR_ALLOC; }
    const int blocks = n_sm * 8, threads = 256;
    const double lanes = (double)blocks * threads;
    double ms;
    ms = time_ms([&] { peak_mufu<<<blocks, threads>>>(fbuf, 0.5f); });
    out->mufu_gops = lanes * ITERS * CHAINS * 2.0 / (ms * 1e6);
    ms = time_ms([&] { peak_ffma<<<blocks, threads>>>(fbuf, 0.5f); });
    out->ffma_gops = lanes * ITERS * CHAINS * 2.0 / (ms * 1e6);
    ms = time_ms([&] { peak_imad<<<blocks, threads>>>((unsigned *)fbuf, 7u); });
    out->imad_gops = lanes * ITERS * CHAINS * 2.0 / (ms * 1e6);
    ms = time_ms([&] { peak_lop3<<<blocks, threads>>>((unsigned *)fbuf, 7u); });
    out->lop3_gops = lanes * ITERS * CHAINS * 2.0 / (ms * 1e6);
    out->sm_clock_mhz = khz / 1000.0;
    out->n_sm = n_sm; out->reserved = 0;
    cudaFree(fbuf);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { set_error("peaks: %s", cudaGetErrorString(e)); return MCL_ERR_CUDA; }
    return MCL_OK;
}
