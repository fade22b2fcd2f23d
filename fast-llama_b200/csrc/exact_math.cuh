// exact_math.cuh — device restatements of the scalar arithmetic on the reference's hot path.
//
// Everything here must produce the SAME BITS as the reference's x86 build (oracle/_ref, AVX2+FMA):
// explicit IEEE round-to-nearest intrinsics only (the file is compiled with -fmad=false as a second
// line of defence), denormals kept, true division / sqrt.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace fl {

// --------------------------------------------------------------------------------------------
// glibc 2.39 expf (sysdeps/ieee754/flt-32/e_expf.c, the FMA ifunc variant) in fp64.
// The reference calls libm expf in softmax_sisd (src/blas/tf_operators.cpp:180) and swiglu
// (src/platforms/arch/x86_simd.cpp:1768).  oracle/ref_port.c:port_expf_emul is the same algorithm and
// is checked against libm on all 2^32 inputs; tests/test_ops_gpu.py checks this copy against it.
// --------------------------------------------------------------------------------------------
__constant__ uint64_t c_exp2f_tab[32] = {
    0x3ff0000000000000ull, 0x3fefd9b0d3158574ull, 0x3fefb5586cf9890full, 0x3fef9301d0125b51ull,
    0x3fef72b83c7d517bull, 0x3fef54873168b9aaull, 0x3fef387a6e756238ull, 0x3fef1e9df51fdee1ull,
    0x3fef06fe0a31b715ull, 0x3feef1a7373aa9cbull, 0x3feedea64c123422ull, 0x3feece086061892dull,
    0x3feebfdad5362a27ull, 0x3feeb42b569d4f82ull, 0x3feeab07dd485429ull, 0x3feea47eb03a5585ull,
    0x3feea09e667f3bcdull, 0x3fee9f75e8ec5f74ull, 0x3feea11473eb0187ull, 0x3feea589994cce13ull,
    0x3feeace5422aa0dbull, 0x3feeb737b0cdc5e5ull, 0x3feec49182a3f090ull, 0x3feed503b23e255dull,
    0x3feee89f995ad3adull, 0x3feeff76f2fb5e47ull, 0x3fef199bdd85529cull, 0x3fef3720dcef9069ull,
    0x3fef5818dcfba487ull, 0x3fef7c97337b9b5full, 0x3fefa4afa2a490daull, 0x3fefd0765b6e4540ull,
};

__device__ __forceinline__ float expf_exact(float x) {
    const double InvLn2N = 0x1.71547652b82fep+0 * 32.0;
    const double SHIFT = 0x1.8p+52;
    const double C0 = 0x1.c6af84b912394p-5 / 32.0 / 32.0 / 32.0;
    const double C1 = 0x1.ebfce50fac4f3p-3 / 32.0 / 32.0;
    const double C2 = 0x1.62e42ff0c52d6p-1 / 32.0;
    const uint32_t ix = __float_as_uint(x);
    const uint32_t abstop = (ix >> 20) & 0x7ffu;
    if (abstop >= 0x42bu) {                       // |x| >= 88 or NaN
        if (ix == 0xff800000u) return 0.0f;
        if (abstop >= 0x7f8u) return __fadd_rn(x, x);
        if (x > 0x1.62e42ep6f) return __int_as_float(0x7f800000);
        if (x < -0x1.9fe368p6f) return 0.0f;
        if (x < -0x1.9d1d9ep6f) return __uint_as_float(1u);   // 2^-149
    }
    const double xd = (double)x;
    const double z = __dmul_rn(InvLn2N, xd);
    double kd = __dadd_rn(z, SHIFT);
    const uint64_t ki = (uint64_t)__double_as_longlong(kd);
    kd = __dsub_rn(kd, SHIFT);
    const double r = __fma_rn(InvLn2N, xd, -kd);
    uint64_t t = c_exp2f_tab[ki & 31];
    t += ki << 47;
    const double s = __longlong_as_double((long long)t);
    const double p = __fma_rn(C0, r, C1);
    const double r2 = __dmul_rn(r, r);
    double y = __fma_rn(C2, r, 1.0);
    y = __fma_rn(p, r2, y);
    y = __dmul_rn(y, s);
    return __double2float_rn(y);
}

// simd::swiglu (x86_simd.cpp:1766-1770): float( double(a) / (1.0 + double(expf(-a))) * double(b) )
__device__ __forceinline__ float swiglu_exact(float a, float b) {
    const double e = (double)expf_exact(-a);
    const double q = __ddiv_rn((double)a, __dadd_rn(e, 1.0));
    return __double2float_rn(__dmul_rn(q, (double)b));
}

// (T)(x / r) as g++ compiled it: vcvttps2dq, low bits kept; NaN / out of range -> 0x80000000 -> low bits 0
__device__ __forceinline__ int cvtt_x86(float v) {
    if (v >= -2147483648.0f && v < 2147483648.0f) return __float2int_rz(v);
    return (int)0x80000000;
}

// rope_v2 (tf_operators.cpp:397-400 as compiled): o0 = fma(c,x0,-(s*x1)), o1 = fma(s,x0,c*x1)
__device__ __forceinline__ void rope_pair(float c, float s, float x0, float x1, float& o0, float& o1) {
    o0 = __fmaf_rn(c, x0, -__fmul_rn(s, x1));
    o1 = __fmaf_rn(s, x0, __fmul_rn(c, x1));
}

}  // namespace fl
