// strict_math.h -- float32 arithmetic with a fixed evaluation order and NO fused multiply-add, usable from
// both nvcc device code and a plain host compiler.
//
// Why: parity with the reference is defined against IEEE binary32 evaluation of its OpenCL expressions in
// source order.  +, -, *, / and sqrt are correctly rounded on sm_100 and on x86-64, so a
// kernel written with these helpers produces bit-identical hit records on the GPU and in the CPU oracle.
// The conventions are those of oracle/clc_shim.inc: dot() sums left to right, normalize(v) = v / sqrt(dot).
#ifndef YUNE_STRICT_MATH_H
#define YUNE_STRICT_MATH_H

#include <math.h>
#include <stdint.h>

#if defined(__CUDACC__)
  #define YUNE_HD __host__ __device__ __forceinline__
  #define YUNE_HD_CALL __host__ __device__ __forceinline__   /* big leaf functions; real calls were measured 30 % slower (stack traffic) */
  /* Leaf functions whose arguments and results are plain values (registers only across the call).  bdpt.cu compiles them as
     real calls: its shade kernel inlined them ~10 times each (250 KB of code) and ncu showed warps waiting for instruction
     fetch 7 cycles per issued instruction. */
  #ifdef YUNE_LEAF_NOINLINE
    #define YUNE_HD_LEAF __host__ __device__ __noinline__ inline
  #else
    #define YUNE_HD_LEAF __host__ __device__ __forceinline__
  #endif
#else
  #define YUNE_HD inline
  #define YUNE_HD_CALL inline
  #define YUNE_HD_LEAF inline
#endif

#if defined(__CUDA_ARCH__)
  #define YUNE_NO_UNROLL _Pragma("unroll 1")
#else
  #define YUNE_NO_UNROLL
#endif

#if defined(__CUDA_ARCH__) && defined(YUNE_MEASURE_FMA)
  /* MEASUREMENT ONLY (tools/build_variant.py fma -DYUNE_MEASURE_FMA -fmad=true): plain operators, which nvcc contracts into
     FFMA.  Hit records are then no longer bit-identical to the oracle; the product is never built this way.  It exists to
     price the no-contraction convention. */
  #define YF_MUL(a, b) ((a) * (b))
  #define YF_ADD(a, b) ((a) + (b))
  #define YF_SUB(a, b) ((a) - (b))
  #define YF_DIV(a, b) __fdiv_rn((a), (b))
  #define YF_SQRT(a)   __fsqrt_rn((a))
  #define YF_ASINT(f)  __float_as_int(f)
  #define YF_ASFLOAT(i) __int_as_float(i)
#elif defined(__CUDA_ARCH__)
  #define YF_MUL(a, b) __fmul_rn((a), (b))
  #define YF_ADD(a, b) __fadd_rn((a), (b))
  #define YF_SUB(a, b) __fsub_rn((a), (b))
  #define YF_DIV(a, b) __fdiv_rn((a), (b))
  #define YF_SQRT(a)   __fsqrt_rn((a))
  #define YF_ASINT(f)  __float_as_int(f)
  #define YF_ASFLOAT(i) __int_as_float(i)
#else
  #define YF_MUL(a, b) ((a) * (b))
  #define YF_ADD(a, b) ((a) + (b))
  #define YF_SUB(a, b) ((a) - (b))
  #define YF_DIV(a, b) ((a) / (b))
  #define YF_SQRT(a)   sqrtf((a))
  static inline int   yf_asint(float f) { union { float f; int i; } u; u.f = f; return u.i; }
  static inline float yf_asfloat(int i) { union { float f; int i; } u; u.i = i; return u.f; }
  #define YF_ASINT(f)  yf_asint(f)
  #define YF_ASFLOAT(i) yf_asfloat(i)
#endif

namespace yune {

struct V3 { float x, y, z; };

YUNE_HD V3 v3(float x, float y, float z) { V3 r; r.x = x; r.y = y; r.z = z; return r; }
YUNE_HD V3 vadd(V3 a, V3 b) { return v3(YF_ADD(a.x, b.x), YF_ADD(a.y, b.y), YF_ADD(a.z, b.z)); }
YUNE_HD V3 vsub(V3 a, V3 b) { return v3(YF_SUB(a.x, b.x), YF_SUB(a.y, b.y), YF_SUB(a.z, b.z)); }
YUNE_HD V3 vmul(V3 a, V3 b) { return v3(YF_MUL(a.x, b.x), YF_MUL(a.y, b.y), YF_MUL(a.z, b.z)); }
YUNE_HD V3 vscale(V3 a, float s) { return v3(YF_MUL(a.x, s), YF_MUL(a.y, s), YF_MUL(a.z, s)); }

YUNE_HD V3 vneg(V3 a) { return v3(-a.x, -a.y, -a.z); }
// OpenCL dot() of two float4 whose w product is +0: ((x*x' + y*y') + z*z') [+ 0]
YUNE_HD float vdot(V3 a, V3 b) { return YF_ADD(YF_ADD(YF_MUL(a.x, b.x), YF_MUL(a.y, b.y)), YF_MUL(a.z, b.z)); }
YUNE_HD V3 vcross(V3 a, V3 b)
{
    return v3(YF_SUB(YF_MUL(a.y, b.z), YF_MUL(a.z, b.y)),
              YF_SUB(YF_MUL(a.z, b.x), YF_MUL(a.x, b.z)),
              YF_SUB(YF_MUL(a.x, b.y), YF_MUL(a.y, b.x)));
}
YUNE_HD float vlength(V3 a) { return YF_SQRT(vdot(a, a)); }


// Exact x / s for s > 0 that never enters the division's slow path because of a ZERO NUMERATOR.  IEEE division on the GPU is a
// reciprocal + refinement guarded by a range check (FCHK) that sends zero / denormal / huge operands to a ~30-instruction
// subroutine; axis-aligned normals and coloured albedos are full of zeros, and ncu attributed 18 % of the shade kernel's
// instructions to that subroutine.  0 / s is just the (signed) zero, so the division is fed a harmless 1 instead and the zero is
// put back.  Same bits as x / s in every case (s <= 0 or NaN falls through to the true division).
YUNE_HD float div_pos(float x, float s)
{
    const bool z = (x == 0.0f) && (s > 0.0f);
    const float q = YF_DIV(z ? 1.0f : x, s);
    return z ? x : q;
}

YUNE_HD_LEAF V3 vdivs(V3 a, float s) { return v3(div_pos(a.x, s), div_pos(a.y, s), div_pos(a.z, s)); }
YUNE_HD V3 vnormalize(V3 a) { float l = YF_SQRT(vdot(a, a)); return vdivs(a, l); }

// OpenCL min/max per the specification text (clc_shim.inc): only used where a NaN may reach them.
YUNE_HD float cl_min(float x, float y) { return (y < x) ? y : x; }
YUNE_HD float cl_max(float x, float y) { return (x < y) ? y : x; }

} // namespace yune
#endif
