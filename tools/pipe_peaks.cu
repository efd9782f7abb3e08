// pipe_peaks.cu — measures the per-SM issue rates that bound the fused cloth kernel besides HBM:
// binary64 DFMA/DADD/DMUL, MUFU.RSQ64H, 128-bit shared-memory loads+stores.  Prints one JSON line.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o build/pipe_peaks tools/pipe_peaks.cu
#include <cstdio>
#include <cuda_runtime.h>

constexpr int kThreads = 512, kIters = 4096, kChains = 8;

__global__ void __launch_bounds__(kThreads) dfma_kernel(double* out, double a, double b) {
    double v[kChains];
#pragma unroll
    for (int q = 0; q < kChains; ++q) v[q] = threadIdx.x + q;
    for (int i = 0; i < kIters; ++i) {
#pragma unroll
        for (int q = 0; q < kChains; ++q) v[q] = fma(v[q], a, b);
    }
    double s = 0;
#pragma unroll
    for (int q = 0; q < kChains; ++q) s += v[q];
    if (s == 123.456) out[0] = s;
}

__global__ void __launch_bounds__(kThreads) dadd_dmul_kernel(double* out, double a, double b) {
    double v[kChains];
#pragma unroll
    for (int q = 0; q < kChains; ++q) v[q] = threadIdx.x + q;
    for (int i = 0; i < kIters / 2; ++i) {
#pragma unroll
        for (int q = 0; q < kChains; ++q) v[q] = __dadd_rn(__dmul_rn(v[q], a), b);
    }
    double s = 0;
#pragma unroll
    for (int q = 0; q < kChains; ++q) s += v[q];
    if (s == 123.456) out[0] = s;
}

__global__ void __launch_bounds__(kThreads) rsq_kernel(double* out, double a) {
    double v[kChains];
#pragma unroll
    for (int q = 0; q < kChains; ++q) v[q] = threadIdx.x + q + a;
    for (int i = 0; i < kIters; ++i) {
#pragma unroll
        for (int q = 0; q < kChains; ++q) asm volatile("rsqrt.approx.ftz.f64 %0, %0;" : "+d"(v[q]));
    }
    double s = 0;
#pragma unroll
    for (int q = 0; q < kChains; ++q) s += v[q];
    if (s == 123.456) out[0] = s;
}

// each thread: load 16 B, store 16 B, unit stride across lanes
__global__ void __launch_bounds__(kThreads) smem_kernel(double* out) {
    extern __shared__ double2 sm[];
    for (int i = threadIdx.x; i < 8192; i += kThreads) sm[i] = make_double2(i, i);
    __syncthreads();
    double2 acc = make_double2(0, 0);
    int at = threadIdx.x;
    for (int i = 0; i < kIters; ++i) {
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            double2 v = sm[(at + q * kThreads) & 8191];
            acc.x += v.y;
            v.x = acc.x;
            sm[(at + q * kThreads + 4096) & 8191] = v;
        }
        at = (at + 32) & 8191;
    }
    if (acc.x == 123.456) out[0] = acc.x;
}

// dependent-issue latency: one warp, one chain, clock64 around it
__global__ void latency_kernel(double* out, long long* cyc, double a, double b) {
    double v = threadIdx.x;
    long long t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < 1024; ++i) v = fma(v, a, b);
    long long t1 = clock64();
    double w = v + 2.0;
#pragma unroll 16
    for (int i = 0; i < 1024; ++i) asm volatile("rsqrt.approx.ftz.f64 %0, %0;" : "+d"(w));
    long long t2 = clock64();
    double z = w;
#pragma unroll 16
    for (int i = 0; i < 1024; ++i) z = __dadd_rn(z, a);
    long long t3 = clock64();
    if (threadIdx.x == 0) {
        cyc[0] = t1 - t0;
        cyc[1] = t2 - t1;
        cyc[2] = t3 - t2;
    }
    if (z == 123.456) out[0] = z;
}

template <typename F>
float time_ms(F f) {
    cudaEvent_t a, b;
    cudaEventCreate(&a);
    cudaEventCreate(&b);
    f();
    cudaDeviceSynchronize();
    float best = 1e30f;
    for (int r = 0; r < 5; ++r) {
        cudaEventRecord(a);
        f();
        cudaEventRecord(b);
        cudaEventSynchronize(b);
        float ms;
        cudaEventElapsedTime(&ms, a, b);
        if (ms < best) best = ms;
    }
    return best;
}

int main() {
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    const int sms = p.multiProcessorCount;
    const int blocks = sms * 2;  // 2 x 512 threads per SM = 8 warps per scheduler
    double* out;
    cudaMalloc(&out, 64);
    int clk_khz = 0;
    cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
    const double n_thread_ops = (double)blocks * kThreads * kIters * kChains;
    const float t_fma = time_ms([&] { dfma_kernel<<<blocks, kThreads>>>(out, 1.0000001, 1e-9); });
    const float t_am = time_ms([&] { dadd_dmul_kernel<<<blocks, kThreads>>>(out, 1.0000001, 1e-9); });
    const float t_rsq = time_ms([&] { rsq_kernel<<<blocks, kThreads>>>(out, 1.5); });
    cudaFuncSetAttribute(smem_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 8192 * 16);
    const float t_sm = time_ms([&] { smem_kernel<<<sms, kThreads, 8192 * 16>>>(out); });
    long long* cyc;
    cudaMallocManaged(&cyc, 64);
    latency_kernel<<<1, 32>>>(out, cyc, 1.0000001, 1e-9);
    cudaDeviceSynchronize();
    printf("{\"dfma_dep_latency_cycles\": %.2f, \"rsq64h_dep_latency_cycles\": %.2f, \"dadd_dep_latency_cycles\": %.2f}\n",
           cyc[0] / 1024.0, cyc[1] / 1024.0, cyc[2] / 1024.0);
    const double smem_bytes = (double)sms * kThreads * kIters * 4 * 32.0;
    printf("{\"device\": \"%s\", \"sms\": %d, \"clock_mhz_attr\": %d, "
           "\"dfma_thread_inst_per_s\": %.4g, \"dadd_dmul_thread_inst_per_s\": %.4g, \"rsq64h_thread_inst_per_s\": %.4g, "
           "\"dfma_per_sm_per_clk_at_1965\": %.2f, \"dadd_dmul_per_sm_per_clk_at_1965\": %.2f, "
           "\"rsq_per_sm_per_clk_at_1965\": %.2f, \"smem_ldst128_bytes_per_s\": %.4g, "
           "\"smem_bytes_per_sm_per_clk_at_1965\": %.1f}\n",
           p.name, sms, clk_khz / 1000, n_thread_ops / (t_fma * 1e-3), n_thread_ops / (t_am * 1e-3),
           n_thread_ops / (t_rsq * 1e-3), n_thread_ops / (t_fma * 1e-3) / sms / 1.965e9,
           n_thread_ops / (t_am * 1e-3) / sms / 1.965e9, n_thread_ops / (t_rsq * 1e-3) / sms / 1.965e9,
           smem_bytes / (t_sm * 1e-3), smem_bytes / (t_sm * 1e-3) / sms / 1.965e9);
    return 0;
}
