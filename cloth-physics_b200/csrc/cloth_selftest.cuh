// cloth_selftest.cuh — device self-test of the binary64 stick arithmetic.
//
// Every fast kernel relaxes a stretched stick through StickMath<double>::mul_dist (cloth_fused.cuh): sqrt(s), (L - d)/d
// and L/d from ONE reciprocal-square-root seed.  The bit-exactness contract needs those three results to be the
// correctly rounded IEEE values Go computes with SQRTSD / DIVSD (physics/constraint.go:26-42).  This kernel checks
// exactly that on the device: the expansion against CUDA's own sqrt.rn.f64 / div.rn.f64, operand by operand,
//   * on pseudo-random squared lengths (random mantissa, exponent uniform over [L^2, 2^61], rest length 1..64), and
//   * on an explicit operand list — the hard cases of the reciprocal step (tools/stickmath_hard_cases.py): stick
//     lengths whose reciprocal lies within a few 2^-106 of a rounding midpoint, for every rest length.
// It is a test hook of the product library, not a compute path: nothing else calls it.
#pragma once
#include "cloth_fused.cuh"

namespace clothgpu_impl {

struct StickMathReport {
    unsigned long long tested;      // (s, L) pairs compared
    unsigned long long mismatches;  // pairs whose dist or mul differ from the IEEE operators
    double first_s, first_L;        // the first mismatching pair (by atomic order)
    double got_dist, want_dist, got_mul, want_mul;
};

__device__ __forceinline__ unsigned long long splitmix64(unsigned long long x) {
    x += 0x9e3779b97f4a7c15ull;
    x = (x ^ (x >> 30)) * 0xbf58476d1ce4e5b9ull;
    x = (x ^ (x >> 27)) * 0x94d049bb133111ebull;
    return x ^ (x >> 31);
}

__device__ __forceinline__ bool stickmath_check_one(double s, double L, double gain, StickMathReport* rep) {
    FrameU<double> u{};
    u.L = L;
    u.gain = gain;
    double mul, dist;
    StickMath<double>::mul_dist(s, u, mul, dist);
    // the reference's operators, one rounding each (constraint.go:28, 41-42)
    const double d = __dsqrt_rn(s);
    const double diff = __ddiv_rn(__dsub_rn(L, d), d);
    const double want = __dmul_rn(__dmul_rn(diff, gain), __dsub_rn(1.0, __ddiv_rn(L, d)));
    const bool ok = __double_as_longlong(d) == __double_as_longlong(dist) && __double_as_longlong(want) == __double_as_longlong(mul);
    if (!ok) {
        if (atomicAdd(&rep->mismatches, 1ull) == 0ull) {
            rep->first_s = s, rep->first_L = L;
            rep->got_dist = dist, rep->want_dist = d, rep->got_mul = mul, rep->want_mul = want;
        }
    }
    return ok;
}

// `samples` pseudo-random operands, grid-stride.  Sample i: rest length L = 1 + (h & 63); s = 2^e * (1 + m/2^52) with the
// exponent uniform over [2*ceil(log2 L), 60] — a stretched stick (s >= L^2, constraint.go:30) up to the longest the int32
// window allows — and all 52 mantissa bits random.
__global__ void __launch_bounds__(256) stickmath_random_kernel(unsigned long long seed, unsigned long long samples, double gain,
                                                               StickMathReport* rep) {
    unsigned long long done = 0;
    for (unsigned long long i = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x; i < samples;
         i += (unsigned long long)gridDim.x * blockDim.x) {
        const unsigned long long h = splitmix64(seed ^ (i * 0xd1342543de82ef95ull));
        const unsigned long long h2 = splitmix64(h);
        const int Li = 1 + (int)(h & 63ull);
        const int emin = 2 * (32 - __clz(Li - 1));  // 2^emin >= L^2
        const int e = emin + (int)((h >> 6) % (unsigned long long)(61 - emin));
        const double s = __longlong_as_double(((long long)(1023 + e) << 52) | (long long)(h2 >> 12));
        stickmath_check_one(s, (double)Li, gain, rep);
        ++done;
    }
    for (int o = 16; o > 0; o >>= 1) done += __shfl_down_sync(0xffffffffu, done, o);
    if ((threadIdx.x & 31) == 0 && done) atomicAdd(&rep->tested, done);
}

// Every listed s against every rest length 1..64 that leaves the stick stretched (s >= L*L).
__global__ void __launch_bounds__(256) stickmath_list_kernel(const double* __restrict__ s_list, unsigned long long n, double gain,
                                                             StickMathReport* rep) {
    const unsigned long long total = n * 64ull;
    unsigned long long done = 0;
    for (unsigned long long i = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x; i < total;
         i += (unsigned long long)gridDim.x * blockDim.x) {
        const double s = s_list[i >> 6];
        const double L = (double)(1 + (int)(i & 63ull));
        if (!(s >= L * L) || !(s < 0x1p900)) continue;
        stickmath_check_one(s, L, gain, rep);
        ++done;
    }
    for (int o = 16; o > 0; o >>= 1) done += __shfl_down_sync(0xffffffffu, done, o);
    if ((threadIdx.x & 31) == 0 && done) atomicAdd(&rep->tested, done);
}

// binary32, exhaustive: every s = the float with bit pattern b, b in [bits(1.0f), bits(2^63)), with every rest length
// 1..64 that leaves the stick stretched.  Grid-stride over b; the 64 rest lengths in the inner loop.
__global__ void __launch_bounds__(256) stickmath_f32_exhaustive_kernel(float gain, StickMathReport* rep) {
    constexpr unsigned int b0 = 0x3f800000u, b1 = 0x5f000000u;  // 1.0f, 2^63
    unsigned long long done = 0;
    for (unsigned long long i = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x; i < (unsigned long long)(b1 - b0);
         i += (unsigned long long)gridDim.x * blockDim.x) {
        const float s = __uint_as_float(b0 + (unsigned int)i);
        const float d = __fsqrt_rn(s);
        for (int Li = 1; Li <= 64; ++Li) {
            const float L = (float)Li;
            if (!(s >= L * L)) break;
            FrameU<float> u{};
            u.L = L;
            u.gain = gain;
            float mul, dist;
            StickMath<float>::mul_dist(s, u, mul, dist);
            const float diff = __fdiv_rn(__fsub_rn(L, d), d);
            const float want = __fmul_rn(__fmul_rn(diff, gain), __fsub_rn(1.0f, __fdiv_rn(L, d)));
            if (__float_as_uint(d) != __float_as_uint(dist) || __float_as_uint(want) != __float_as_uint(mul)) {
                if (atomicAdd(&rep->mismatches, 1ull) == 0ull) {
                    rep->first_s = s, rep->first_L = L;
                    rep->got_dist = dist, rep->want_dist = d, rep->got_mul = mul, rep->want_mul = want;
                }
            }
            ++done;
        }
    }
    for (int o = 16; o > 0; o >>= 1) done += __shfl_down_sync(0xffffffffu, done, o);
    if ((threadIdx.x & 31) == 0 && done) atomicAdd(&rep->tested, done);
}

}  // namespace clothgpu_impl
