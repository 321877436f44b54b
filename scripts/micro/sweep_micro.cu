// Microbenchmark: the identical-channel sweep of mcl_philox.cu in isolation (no barrier, no events), to see how many
// cycles one sweep of a 10^4-slot box costs a 256-thread CTA with 1, 2 or 3 CTAs resident per SM.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o sweep_micro sweep_micro.cu && ./sweep_micro
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
struct RoundKeys { uint32_t k[20]; };
__device__ __forceinline__ void philox(uint32_t &c0, uint32_t &c1, uint32_t &c2, uint32_t &c3, const RoundKeys &K) {
#pragma unroll
    for (int r = 0; r < 10; r++) {
        unsigned long long p0 = 0xD2511F53ull * c0, p1 = 0xCD9E8D57ull * c2;
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ K.k[2 * r], n2 = (uint32_t)(p0 >> 32) ^ c3 ^ K.k[2 * r + 1];
        c1 = (uint32_t)p1; c3 = (uint32_t)p0; c0 = n0; c2 = n2;
    }
}
__device__ __forceinline__ float lg2f_(float x) { float y; asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float u01(uint32_t r) { return __uint_as_float((r >> 9) | 0x3f800000u) - 0.99999994f; }
template <int NCH, bool BARRIER, int MODE>
__global__ void __launch_bounds__(256, 3) sweep_kernel(const RoundKeys K, int n_slots, int steps, float *out, long long *cyc) {
    extern __shared__ __align__(16) float cr[];
    const int tid = threadIdx.x, NT = 256;
    for (int s = tid; s < 10048; s += NT) cr[s] = 100.0f + (float)((s * 2654435761u) >> 20) * 1e-3f;
    __syncthreads();
    const int n_chunks = n_slots / 4;
    const float4 *cr4 = reinterpret_cast<const float4 *>(cr);
    float acc = 0.f;
    const long long t0 = clock64();
    for (int step = 0; step < steps; step++) {
        float best = __builtin_huge_valf(); int bslot = -1;
        for (int b0 = tid; b0 < n_chunks; b0 += NCH * NT) {
            float cs[NCH][4]; uint32_t w[NCH][4];
#pragma unroll
            for (int q = 0; q < NCH; q++) {
                const int b = min(b0 + q * NT, n_chunks - 1);
                const float4 c = cr4[b]; cs[q][0] = c.x; cs[q][1] = c.y; cs[q][2] = c.z; cs[q][3] = c.w;
                w[q][0] = b; w[q][1] = step; w[q][2] = blockIdx.x; w[q][3] = 0x40000000u;
            }
#pragma unroll
            for (int q = 0; q < NCH; q++) philox(w[q][0], w[q][1], w[q][2], w[q][3], K);
#pragma unroll
            for (int q = 0; q < NCH; q++) {
                float l[4];
#pragma unroll
                for (int k = 0; k < 4; k++) l[k] = lg2f_(-lg2f_(u01(w[q][k]))) + cs[q][k];
                if (MODE == 0) {
#pragma unroll
                    for (int k = 0; k < 4; k++) if (l[k] < best) { best = l[k]; bslot = 4 * (b0 + q * NT) + k; }
                } else {
                    const bool p01 = l[1] < l[0], p23 = l[3] < l[2];
                    const float m01 = p01 ? l[1] : l[0], m23 = p23 ? l[3] : l[2];
                    const bool ph = m23 < m01;
                    const float m = ph ? m23 : m01;
                    const int km = ph ? (p23 ? 3 : 2) : (p01 ? 1 : 0);
                    if (m < best) { best = m; bslot = 4 * (b0 + q * NT) + km; }
                }
            }
        }
        acc += best + (float)bslot;
        if (BARRIER) __syncthreads();
    }
    const long long t1 = clock64();
    if (tid == 0) cyc[blockIdx.x] = t1 - t0;
    out[blockIdx.x * NT + tid] = acc;
}
template <int NCH, bool BARRIER, int MODE> void run(int ctas_per_sm, int n_slots) {
    RoundKeys K; for (int i = 0; i < 20; i++) K.k[i] = 0x9E3779B9u * (i + 1);
    const int grid = 148 * ctas_per_sm, steps = 400;
    float *out; long long *cyc; cudaMalloc(&out, grid * 256 * 4); cudaMalloc(&cyc, grid * 8);
    cudaFuncSetAttribute(sweep_kernel<NCH, BARRIER, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 61440);
    sweep_kernel<NCH, BARRIER, MODE><<<grid, 256, 61440>>>(K, n_slots, steps, out, cyc);
    cudaDeviceSynchronize();
    long long h[148 * 3]; cudaMemcpy(h, cyc, grid * 8, cudaMemcpyDeviceToHost);
    double m = 0; for (int i = 0; i < grid; i++) m += h[i]; m /= grid;
    const double per_sweep = m / steps;
    printf("MODE=%d NCH=%d barrier=%d CTAs/SM=%d slots=%d: %.0f cycles per sweep per CTA -> %.2f cycles per slot per SM, %.3e e-steps/s at 1.965 GHz (%s)\n",
           MODE, NCH, (int)BARRIER, ctas_per_sm, n_slots, per_sweep, per_sweep / n_slots / ctas_per_sm,
           148.0 * 1.965e9 / (per_sweep / n_slots / ctas_per_sm), cudaGetErrorString(cudaGetLastError()));
    cudaFree(out); cudaFree(cyc);
}
int main() {
    for (int c = 1; c <= 3; c++) { run<2, false, 0>(c, 10000); run<2, false, 1>(c, 10000); run<4, false, 0>(c, 10000); run<4, false, 1>(c, 10000); run<3, false, 1>(c, 10000); }
    return 0;
}
