// sbx_pk.cuh -- packed FP32 pairs for sm_100a: FADD2 / FMUL2 / FFMA2 (PTX add/mul/fma.rn.f32x2, new with Blackwell).
//
// One instruction works on two independent f32 lanes held in an aligned 64-bit register pair, with the same
// IEEE round-to-nearest results as the scalar instruction on each lane (no contraction, no ftz).  The fused frame
// kernel is bound by instruction issue, not by the FP32 pipe, so running the SAME arithmetic of two independent
// transforms in the two lanes halves the issue slots it costs.  `pk` carries the pair as a b64 value so that ptxas
// keeps it in a register pair from the 128-bit shared-memory load to the 128-bit store; lane 0 = .lo, lane 1 = .hi.
//
// cpk = a complex number in each lane: re = (re0, re1), im = (im0, im1) -- the structure-of-arrays form in which
// multiplication by -i is a register renaming and every butterfly operation is lane-uniform.
#pragma once
#ifdef SBX_EMU
#include "emu_shim.h"
#else
#include <cuda_runtime.h>
#endif
#include <stdint.h>

namespace sbx {

#ifdef SBX_EMU
struct pk { float lo, hi; };
__device__ __forceinline__ pk pk_make(float lo, float hi) { pk r; r.lo = lo; r.hi = hi; return r; }
__device__ __forceinline__ float pk_lo(pk a) { return a.lo; }
__device__ __forceinline__ float pk_hi(pk a) { return a.hi; }
__device__ __forceinline__ pk pk_add(pk a, pk b) { return pk_make(a.lo + b.lo, a.hi + b.hi); }
__device__ __forceinline__ pk pk_sub(pk a, pk b) { return pk_make(a.lo - b.lo, a.hi - b.hi); }
__device__ __forceinline__ pk pk_mul(pk a, pk b) { return pk_make(a.lo * b.lo, a.hi * b.hi); }
__device__ __forceinline__ pk pk_fma(pk a, pk b, pk c) { return pk_make(fmaf(a.lo, b.lo, c.lo), fmaf(a.hi, b.hi, c.hi)); }
#else
typedef unsigned long long pk;
__device__ __forceinline__ pk pk_make(float lo, float hi)
{
    pk r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ float pk_lo(pk a)
{
    float lo, hi;
    asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(a));
    return lo;
}
__device__ __forceinline__ float pk_hi(pk a)
{
    float lo, hi;
    asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(a));
    return hi;
}
__device__ __forceinline__ pk pk_add(pk a, pk b) { pk r; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ pk pk_sub(pk a, pk b) { pk r; asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ pk pk_mul(pk a, pk b) { pk r; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ pk pk_fma(pk a, pk b, pk c)
{
    pk r;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
    return r;
}
#endif
__device__ __forceinline__ pk pk_dup(float c) { return pk_make(c, c); }
// -a: ptxas folds the two scalar negations into the operand modifier of the consuming FADD2 / FMUL2 / FFMA2
__device__ __forceinline__ pk pk_neg(pk a) { return pk_make(-pk_lo(a), -pk_hi(a)); }

struct cpk { pk re, im; };

__device__ __forceinline__ cpk cpk_make(pk re, pk im) { cpk r; r.re = re; r.im = im; return r; }
__device__ __forceinline__ cpk cpk_dup(float2 w) { return cpk_make(pk_dup(w.x), pk_dup(w.y)); }
__device__ __forceinline__ cpk cpk_add(cpk a, cpk b) { return cpk_make(pk_add(a.re, b.re), pk_add(a.im, b.im)); }
__device__ __forceinline__ cpk cpk_sub(cpk a, cpk b) { return cpk_make(pk_sub(a.re, b.re), pk_sub(a.im, b.im)); }
// a * b = (a.re b.re - a.im b.im, a.re b.im + a.im b.re): 2 FMUL2 + 2 FFMA2
__device__ __forceinline__ cpk cpk_mul(cpk a, cpk b)
{
    return cpk_make(pk_fma(a.re, b.re, pk_neg(pk_mul(a.im, b.im))), pk_fma(a.re, b.im, pk_mul(a.im, b.re)));
}
__device__ __forceinline__ cpk cpk_sqr(cpk a)
{
    const pk t = pk_mul(a.re, a.im);
    return cpk_make(pk_fma(a.re, a.re, pk_neg(pk_mul(a.im, a.im))), pk_add(t, t));
}

// 128-bit shared-memory element (re0, re1, im0, im1) <-> cpk
__device__ __forceinline__ cpk cpk_lds(const float4 *p)
{
#ifdef SBX_EMU
    const float4 v = *p;
    return cpk_make(pk_make(v.x, v.y), pk_make(v.z, v.w));
#else
    const ulonglong2 v = *reinterpret_cast<const ulonglong2 *>(p);
    return cpk_make(v.x, v.y);
#endif
}
__device__ __forceinline__ void cpk_sts(float4 *p, cpk v)
{
#ifdef SBX_EMU
    *p = make_float4(v.re.lo, v.re.hi, v.im.lo, v.im.hi);
#else
    *p = make_float4(pk_lo(v.re), pk_hi(v.re), pk_lo(v.im), pk_hi(v.im));
#endif
}

// ---- forward DFT butterflies on cpk (both lanes run the same butterfly) ---------------------------------------
__device__ __forceinline__ void pk_bf2(cpk &a, cpk &b)
{
    const cpk t = cpk_sub(a, b);
    a = cpk_add(a, b);
    b = t;
}
// 4-point DFT, outputs in natural order: 16 packed instructions
__device__ __forceinline__ void pk_bf4(cpk &a0, cpk &a1, cpk &a2, cpk &a3)
{
    const cpk t0 = cpk_add(a0, a2), t1 = cpk_sub(a0, a2), t2 = cpk_add(a1, a3), t3 = cpk_sub(a1, a3);
    a0 = cpk_add(t0, t2);
    a2 = cpk_sub(t0, t2);
    a1 = cpk_make(pk_add(t1.re, t3.im), pk_sub(t1.im, t3.re));     // t1 - i t3
    a3 = cpk_make(pk_sub(t1.re, t3.im), pk_add(t1.im, t3.re));     // t1 + i t3
}
#define SBX_PK_RSQRT2 0.70710678118654752440f
#define SBX_PK_COS_PI_8 0.92387953251128675613f
#define SBX_PK_SIN_PI_8 0.38268343236508977173f
// v * (c - i s) for lane-uniform constants given as dup'ed pairs
__device__ __forceinline__ cpk pk_rot(cpk v, pk c, pk s)
{
    return cpk_make(pk_fma(v.re, c, pk_mul(v.im, s)), pk_fma(v.im, c, pk_neg(pk_mul(v.re, s))));
}
__device__ __forceinline__ cpk pk_mul_w8_1(cpk v, pk r)             // * W8^1 = (1 - i) / sqrt 2
{
    return cpk_make(pk_mul(pk_add(v.re, v.im), r), pk_mul(pk_sub(v.im, v.re), r));
}
__device__ __forceinline__ cpk pk_mul_w8_3(cpk v, pk r)             // * W8^3 = (-1 - i) / sqrt 2
{
    return cpk_make(pk_mul(pk_sub(v.im, v.re), r), pk_mul(pk_add(v.re, v.im), pk_neg(r)));
}
__device__ __forceinline__ cpk pk_mul_neg_i(cpk v) { return cpk_make(v.im, pk_neg(v.re)); }

template <int R> struct PkButterfly;
template <> struct PkButterfly<2> {
    static __device__ __forceinline__ void run(cpk (&v)[2]) { pk_bf2(v[0], v[1]); }
};
template <> struct PkButterfly<4> {
    static __device__ __forceinline__ void run(cpk (&v)[4]) { pk_bf4(v[0], v[1], v[2], v[3]); }
};
template <> struct PkButterfly<8> {
    static __device__ __forceinline__ void run(cpk (&v)[8])
    {
        const pk r = pk_dup(SBX_PK_RSQRT2);
        pk_bf4(v[0], v[2], v[4], v[6]);
        pk_bf4(v[1], v[3], v[5], v[7]);
        v[3] = pk_mul_w8_1(v[3], r);
        v[5] = pk_mul_neg_i(v[5]);
        v[7] = pk_mul_w8_3(v[7], r);
        const cpk x0 = cpk_add(v[0], v[1]), x4 = cpk_sub(v[0], v[1]);
        const cpk x1 = cpk_add(v[2], v[3]), x5 = cpk_sub(v[2], v[3]);
        const cpk x2 = cpk_add(v[4], v[5]), x6 = cpk_sub(v[4], v[5]);
        const cpk x3 = cpk_add(v[6], v[7]), x7 = cpk_sub(v[6], v[7]);
        v[0] = x0; v[1] = x1; v[2] = x2; v[3] = x3; v[4] = x4; v[5] = x5; v[6] = x6; v[7] = x7;
    }
};
template <> struct PkButterfly<16> {
    // 16 = 4 x 4 exactly as Butterfly<16> of stft_kernels.cuh: 2 x 64 add/sub + 8 rotations x 4 = 160 packed instructions
    static __device__ __forceinline__ void run(cpk (&v)[16])
    {
        pk_bf4(v[0], v[4], v[8], v[12]);
        pk_bf4(v[1], v[5], v[9], v[13]);
        pk_bf4(v[2], v[6], v[10], v[14]);
        pk_bf4(v[3], v[7], v[11], v[15]);
        const pk c = pk_dup(SBX_PK_COS_PI_8), s = pk_dup(SBX_PK_SIN_PI_8), r = pk_dup(SBX_PK_RSQRT2);
        v[5] = pk_rot(v[5], c, s);                       // W16^1 = c - i s
        v[6] = pk_mul_w8_1(v[6], r);                     // W16^2
        v[7] = pk_rot(v[7], s, c);                       // W16^3 = s - i c
        v[9] = pk_mul_w8_1(v[9], r);                     // W16^2
        v[10] = pk_mul_neg_i(v[10]);                     // W16^4
        v[11] = pk_mul_w8_3(v[11], r);                   // W16^6
        v[13] = pk_rot(v[13], s, c);                     // W16^3
        v[14] = pk_mul_w8_3(v[14], r);                   // W16^6
        v[15] = pk_rot(v[15], pk_neg(c), pk_neg(s));     // W16^9 = -c + i s
        pk_bf4(v[0], v[1], v[2], v[3]);
        pk_bf4(v[4], v[5], v[6], v[7]);
        pk_bf4(v[8], v[9], v[10], v[11]);
        pk_bf4(v[12], v[13], v[14], v[15]);
        cpk t;
        t = v[1]; v[1] = v[4]; v[4] = t;
        t = v[2]; v[2] = v[8]; v[8] = t;
        t = v[3]; v[3] = v[12]; v[12] = t;
        t = v[6]; v[6] = v[9]; v[9] = t;
        t = v[7]; v[7] = v[13]; v[13] = t;
        t = v[11]; v[11] = v[14]; v[14] = t;
    }
};

// v[q] *= w^q, q = 1..R-1, from w^1 (and w^4 for R >= 8) given as dup'ed pairs; the other powers by products, as
// apply_twiddles of stft_kernels.cuh with SBX_TW_LOADS = 2 (both lanes share the twiddle: same butterfly index)
template <int R>
__device__ __forceinline__ void pk_apply_twiddles(cpk (&v)[R], cpk w1, cpk w4)
{
    v[1] = cpk_mul(v[1], w1);
    if constexpr (R >= 4) {
        const cpk w2 = cpk_sqr(w1), w3 = cpk_mul(w2, w1);
        v[2] = cpk_mul(v[2], w2);
        v[3] = cpk_mul(v[3], w3);
        if constexpr (R >= 8) {
            const cpk w5 = cpk_mul(w4, w1), w6 = cpk_mul(w4, w2), w7 = cpk_mul(w4, w3);
            v[4] = cpk_mul(v[4], w4);
            v[5] = cpk_mul(v[5], w5);
            v[6] = cpk_mul(v[6], w6);
            v[7] = cpk_mul(v[7], w7);
            if constexpr (R >= 16) {
                const cpk w8 = cpk_sqr(w4);
                v[8] = cpk_mul(v[8], w8);
                v[9] = cpk_mul(v[9], cpk_mul(w8, w1));
                v[10] = cpk_mul(v[10], cpk_mul(w8, w2));
                v[11] = cpk_mul(v[11], cpk_mul(w8, w3));
                v[12] = cpk_mul(v[12], cpk_mul(w8, w4));
                v[13] = cpk_mul(v[13], cpk_mul(w8, w5));
                v[14] = cpk_mul(v[14], cpk_mul(w8, w6));
                v[15] = cpk_mul(v[15], cpk_mul(w8, w7));
            }
        }
    }
}
__device__ __forceinline__ void pk_bf16(cpk (&v)[16]) { PkButterfly<16>::run(v); }
__device__ __forceinline__ void pk_tw16(cpk (&v)[16], cpk w1, cpk w4) { pk_apply_twiddles<16>(v, w1, w4); }

}  // namespace sbx
