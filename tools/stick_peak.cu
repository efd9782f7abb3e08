// stick_peak.cu — how fast can the B200 relax sticks when everything stays in registers?
//
// Runs the product's own relax_n<double, false, N> (cloth_fused.cuh) on N sticks per thread in a loop, no memory
// traffic at all, for N = 1, 2, 4, 8 chains per thread and 4..24 warps per SM.  The best figure is the arithmetic
// "speed of light" of the relaxation (every stick takes the full stretched path), the table says how many independent
// chains a scheduler needs to get there.  Build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -fmad=false
//   -I cloth-physics_b200/csrc tools/stick_peak.cu -o build/stick_peak
#include <cstdio>

#include "cloth_fused.cuh"

using namespace clothgpu_impl;

constexpr int kIters = 2048;

template <int N>
__global__ void __launch_bounds__(256) chain_kernel(double* out, FrameU<double> u, double stretch) {
    double2 h[N], t[N];
    double2* hp[N];
    double2* tp[N];
#pragma unroll
    for (int q = 0; q < N; ++q) {
        h[q].x = 100.0 + threadIdx.x * 0.001 + q, h[q].y = 200.0 + blockIdx.x * 0.01;
        t[q].x = h[q].x + stretch, t[q].y = h[q].y + 0.25 * q;
        hp[q] = &h[q], tp[q] = &t[q];
    }
    bool touched = false;
#pragma unroll 1
    for (int i = 0; i < kIters; ++i) {
        relax_n<double, false, N>(hp, tp, (1u << N) - 1u, 0u, u, false, touched);
#pragma unroll
        for (int q = 0; q < N; ++q) t[q].x += 1.0;  // pull the stick apart again: it stays on the stretched path
    }
    double acc = 0;
#pragma unroll
    for (int q = 0; q < N; ++q) acc += h[q].x + t[q].y;
    if (acc == 123.456 || !touched) out[0] = acc;
}

template <int N>
double run(int sms, int warps_per_sm, double* out, const FrameU<double>& u) {
    const int threads = 256, blocks = sms * warps_per_sm * 32 / threads;
    cudaEvent_t a, b;
    cudaEventCreate(&a);
    cudaEventCreate(&b);
    chain_kernel<N><<<blocks, threads>>>(out, u, 9.0);
    cudaDeviceSynchronize();
    float best = 1e30f;
    for (int r = 0; r < 3; ++r) {
        cudaEventRecord(a);
        chain_kernel<N><<<blocks, threads>>>(out, u, 9.0);
        cudaEventRecord(b);
        cudaEventSynchronize(b);
        float ms;
        cudaEventElapsedTime(&ms, a, b);
        if (ms < best) best = ms;
    }
    return (double)blocks * threads * N * kIters / (best * 1e-3);
}

int main() {
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    const int sms = p.multiProcessorCount;
    double* out;
    cudaMalloc(&out, 64);
    FrameU<double> u = {};
    u.L = 6.0, u.s_rest = 36.0, u.gain = 0.35, u.tear_len = 150.0;
    printf("{\"device\": \"%s\", \"sms\": %d, \"unit\": \"G stick relaxations/s, registers only\", \"rows\": [\n", p.name, sms);
    const int warps[] = {4, 8, 16, 24, 32};
    for (int wi = 0; wi < 5; ++wi) {
        const int w = warps[wi];
        printf("  {\"warps_per_sm\": %d, \"chains1\": %.1f, \"chains2\": %.1f, \"chains4\": %.1f, \"chains8\": %.1f}%s\n", w,
               run<1>(sms, w, out, u) * 1e-9, run<2>(sms, w, out, u) * 1e-9, run<4>(sms, w, out, u) * 1e-9,
               run<8>(sms, w, out, u) * 1e-9, wi == 4 ? "" : ",");
    }
    printf("]}\n");
    return 0;
}
