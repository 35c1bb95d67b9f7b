// Register-resident small DFTs (radix 2/4/8/16) used by the fused STFT kernel.
//
// Every function works on a fully unrolled array of complex values so the data
// stays in registers; inputs and outputs are in natural order, forward transform
// (e^{-2*pi*i*n*k/R}).  The code is generic over the complex value type V:
//   * float2           -- one frame per thread (scalar FADD/FMUL/FFMA);
//   * c2 = {pk x, y}   -- TWO frames per thread, the same element of frame A and
//                         frame B packed in one 64-bit register pair and processed
//                         by Blackwell's packed FP32 instructions (FADD2 / FMUL2 /
//                         FFMA2: one issue slot, both lanes), which halves the
//                         issue-slot cost of the transform.
// The functions are __host__ __device__ (the packed type has a plain two-float
// host fallback) so the host emulator under tests/emu/ exercises the exact same
// arithmetic and index algebra as the kernel.
#pragma once
#include <cuda_runtime.h>

#define MMF_HD __host__ __device__ __forceinline__

namespace mmf {

// ---------------------------------------------------------------------------
// scalar layer: float, and pk = (frame A, frame B) packed
// ---------------------------------------------------------------------------
struct alignas(8) pk {
  unsigned long long v;  // low 32 bits: frame A, high 32 bits: frame B
};

MMF_HD pk pmake(float a, float b) {
  pk r;
#ifdef __CUDA_ARCH__
  asm("mov.b64 %0, {%1, %2};" : "=l"(r.v) : "f"(a), "f"(b));
#else
  unsigned ua, ub;
  __builtin_memcpy(&ua, &a, 4);
  __builtin_memcpy(&ub, &b, 4);
  r.v = ((unsigned long long)ub << 32) | ua;
#endif
  return r;
}
MMF_HD float plo(pk p) {
#ifdef __CUDA_ARCH__
  float a, b;
  asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(p.v));
  return a;
#else
  unsigned u = (unsigned)(p.v & 0xffffffffu);
  float a;
  __builtin_memcpy(&a, &u, 4);
  return a;
#endif
}
MMF_HD float phi(pk p) {
#ifdef __CUDA_ARCH__
  float a, b;
  asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(p.v));
  return b;
#else
  unsigned u = (unsigned)(p.v >> 32);
  float b;
  __builtin_memcpy(&b, &u, 4);
  return b;
#endif
}

MMF_HD float sadd(float a, float b) { return a + b; }
MMF_HD float ssub(float a, float b) { return a - b; }
MMF_HD float smul(float a, float b) { return a * b; }
MMF_HD float sfma(float a, float b, float c) { return fmaf(a, b, c); }
MMF_HD float sfnma(float a, float b, float c) { return fmaf(-a, b, c); }  // c - a*b
MMF_HD float sneg(float a) { return -a; }
MMF_HD pk sadd(pk a, pk b) {
#ifdef __CUDA_ARCH__
  pk r;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r.v) : "l"(a.v), "l"(b.v));
  return r;
#else
  return pmake(plo(a) + plo(b), phi(a) + phi(b));
#endif
}
MMF_HD pk ssub(pk a, pk b) {
#ifdef __CUDA_ARCH__
  pk r;
  asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r.v) : "l"(a.v), "l"(b.v));
  return r;
#else
  return pmake(plo(a) - plo(b), phi(a) - phi(b));
#endif
}
MMF_HD pk smul(pk a, pk b) {
#ifdef __CUDA_ARCH__
  pk r;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r.v) : "l"(a.v), "l"(b.v));
  return r;
#else
  return pmake(plo(a) * plo(b), phi(a) * phi(b));
#endif
}
MMF_HD pk sfma(pk a, pk b, pk c) {
#ifdef __CUDA_ARCH__
  pk r;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r.v) : "l"(a.v), "l"(b.v), "l"(c.v));
  return r;
#else
  return pmake(fmaf(plo(a), plo(b), plo(c)), fmaf(phi(a), phi(b), phi(c)));
#endif
}

// c - a*b and -a: the negations are written on the halves; ptxas folds them into the
// operand-negate modifier of the consuming FFMA2 / FADD2
MMF_HD pk sneg(pk a) {
#ifdef __CUDA_ARCH__
  pk r;
  asm("{\n .reg .f32 lo, hi;\n mov.b64 {lo, hi}, %1;\n neg.f32 lo, lo;\n neg.f32 hi, hi;\n mov.b64 %0, {lo, hi};\n}"
      : "=l"(r.v)
      : "l"(a.v));
  return r;
#else
  return pmake(-plo(a), -phi(a));
#endif
}
MMF_HD pk sfnma(pk a, pk b, pk c) { return sfma(sneg(a), b, c); }

template <typename S>
struct ScalarOps;
template <>
struct ScalarOps<float> {
  static MMF_HD float dup(float c) { return c; }
};
template <>
struct ScalarOps<pk> {
  static MMF_HD pk dup(float c) { return pmake(c, c); }
};

// ---------------------------------------------------------------------------
// complex layer
// ---------------------------------------------------------------------------
struct alignas(16) c2 {
  pk x, y;  // real parts of (frame A, frame B); imaginary parts of (frame A, frame B)
};

template <typename V>
struct CxTraits;
template <>
struct CxTraits<float2> {
  using S = float;      // scalar of one component
  using Tw = float2;    // twiddle as stored in shared memory
  using Xe = float2;    // exchange-buffer element (8 bytes): the whole complex value
  static constexpr int kFrames = 1;
  static constexpr int kXParts = 1;  // exchange rounds
  static MMF_HD float2 make(float x, float y) { return make_float2(x, y); }
};
template <>
struct CxTraits<c2> {
  using S = pk;
  using Tw = c2;        // (wx, wx), (wy, wy)
  using Xe = pk;        // one component of both frames (8 bytes); exchanges run in two rounds
  static constexpr int kFrames = 2;
  static constexpr int kXParts = 2;
  static MMF_HD c2 make(pk x, pk y) {
    c2 r;
    r.x = x;
    r.y = y;
    return r;
  }
};

template <typename V>
MMF_HD V cx(typename CxTraits<V>::S x, typename CxTraits<V>::S y) {
  return CxTraits<V>::make(x, y);
}
template <typename V>
MMF_HD typename CxTraits<V>::S sconst(float c) {
  return ScalarOps<typename CxTraits<V>::S>::dup(c);
}

// twiddle of the value type from a scalar complex constant
MMF_HD float2 tw_from(float2 w, const float2*) { return w; }
MMF_HD c2 tw_from(float2 w, const c2*) { return CxTraits<c2>::make(pmake(w.x, w.x), pmake(w.y, w.y)); }
template <typename V>
MMF_HD typename CxTraits<V>::Tw make_tw(float2 w) {
  return tw_from(w, (const V*)nullptr);
}

template <typename V>
MMF_HD V cadd(V a, V b) {
  return cx<V>(sadd(a.x, b.x), sadd(a.y, b.y));
}
template <typename V>
MMF_HD V csub(V a, V b) {
  return cx<V>(ssub(a.x, b.x), ssub(a.y, b.y));
}
// a * w (w a twiddle of the matching type): (ax*wx - ay*wy, ax*wy + ay*wx)
template <typename V, typename W>
MMF_HD V cmul(V a, W w) {
  return cx<V>(sfnma(a.y, w.y, smul(a.x, w.x)), sfma(a.x, w.y, smul(a.y, w.x)));
}
// multiply by -i
template <typename V>
MMF_HD V cmul_mi(V a) {
  return cx<V>(a.y, sneg(a.x));
}

// e^{-2*pi*i*J/16} for the J that occur inside the 16-point butterfly.
template <int J, typename V>
MMF_HD V mul_w16(V a) {
  constexpr float C1 = 0.92387953251128674f;  // cos(pi/8)
  constexpr float S1 = 0.38268343236508977f;  // sin(pi/8)
  constexpr float R = 0.70710678118654752f;   // sqrt(1/2)
  if constexpr (J == 0) {
    return a;
  } else if constexpr (J == 1) {
    return cx<V>(sfma(a.y, sconst<V>(S1), smul(a.x, sconst<V>(C1))),
                 sfnma(a.x, sconst<V>(S1), smul(a.y, sconst<V>(C1))));
  } else if constexpr (J == 2) {
    return cx<V>(smul(sadd(a.x, a.y), sconst<V>(R)), smul(ssub(a.y, a.x), sconst<V>(R)));
  } else if constexpr (J == 3) {
    return cx<V>(sfma(a.y, sconst<V>(C1), smul(a.x, sconst<V>(S1))),
                 sfnma(a.x, sconst<V>(C1), smul(a.y, sconst<V>(S1))));
  } else if constexpr (J == 4) {
    return cmul_mi(a);
  } else if constexpr (J == 6) {
    return cx<V>(smul(ssub(a.y, a.x), sconst<V>(R)), smul(sadd(a.x, a.y), sconst<V>(-R)));
  } else {
    static_assert(J == 9, "unexpected twiddle");
    return cx<V>(sfnma(a.y, sconst<V>(S1), smul(a.x, sconst<V>(-C1))),
                 sfnma(a.y, sconst<V>(C1), smul(a.x, sconst<V>(S1))));
  }
}

template <typename V>
MMF_HD void dft2(V& a, V& b) {
  V t = a;
  a = cadd(t, b);
  b = csub(t, b);
}

template <typename V>
MMF_HD void dft4(V& a0, V& a1, V& a2, V& a3) {
  V s02 = cadd(a0, a2), d02 = csub(a0, a2);
  V s13 = cadd(a1, a3), d13 = csub(a1, a3);
  a0 = cadd(s02, s13);
  a2 = csub(s02, s13);
  a1 = cx<V>(sadd(d02.x, d13.y), ssub(d02.y, d13.x));  // d02 - i*d13
  a3 = cx<V>(ssub(d02.x, d13.y), sadd(d02.y, d13.x));  // d02 + i*d13
}

// 8-point DFT as 2 (n1) x 4 (n2):  n = n1 + 2*n2,  k = 4*k1 + k2.
template <typename V>
MMF_HD void dft8(V (&v)[8]) {
  dft4(v[0], v[2], v[4], v[6]);  // n1 = 0 -> a[0][k2] at v[2*k2]
  dft4(v[1], v[3], v[5], v[7]);  // n1 = 1 -> a[1][k2] at v[1 + 2*k2]
  v[3] = mul_w16<2>(v[3]);       // W8^1
  v[5] = mul_w16<4>(v[5]);       // W8^2
  v[7] = mul_w16<6>(v[7]);       // W8^3
  dft2(v[0], v[1]);
  dft2(v[2], v[3]);
  dft2(v[4], v[5]);
  dft2(v[6], v[7]);
  // result for k = 4*k1 + k2 sits at v[k1 + 2*k2]
  V o[8];
#pragma unroll
  for (int k2 = 0; k2 < 4; ++k2) {
    o[k2] = v[2 * k2];
    o[4 + k2] = v[2 * k2 + 1];
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) v[i] = o[i];
}

// 16-point DFT as 4 x 4:  n = n1 + 4*n2,  k = 4*k1 + k2.
template <typename V>
MMF_HD void dft16(V (&v)[16]) {
  dft4(v[0], v[4], v[8], v[12]);   // a[n1][k2] at v[n1 + 4*k2]
  dft4(v[1], v[5], v[9], v[13]);
  dft4(v[2], v[6], v[10], v[14]);
  dft4(v[3], v[7], v[11], v[15]);
  // twiddle W16^{n1*k2}
  v[5] = mul_w16<1>(v[5]);
  v[6] = mul_w16<2>(v[6]);
  v[7] = mul_w16<3>(v[7]);
  v[9] = mul_w16<2>(v[9]);
  v[10] = mul_w16<4>(v[10]);
  v[11] = mul_w16<6>(v[11]);
  v[13] = mul_w16<3>(v[13]);
  v[14] = mul_w16<6>(v[14]);
  v[15] = mul_w16<9>(v[15]);
  dft4(v[0], v[1], v[2], v[3]);    // over n1 -> k1, result at v[k1 + 4*k2]
  dft4(v[4], v[5], v[6], v[7]);
  dft4(v[8], v[9], v[10], v[11]);
  dft4(v[12], v[13], v[14], v[15]);
  V o[16];
#pragma unroll
  for (int k1 = 0; k1 < 4; ++k1)
#pragma unroll
    for (int k2 = 0; k2 < 4; ++k2) o[4 * k1 + k2] = v[k1 + 4 * k2];
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = o[i];
}

// R-point DFT on v[BASE .. BASE+R) of a 16-register file.
template <int R, int BASE, typename V>
MMF_HD void dft_sub(V (&v)[16]) {
  if constexpr (R == 16) {
    dft16(v);
  } else if constexpr (R == 8) {
    V t[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) t[i] = v[BASE + i];
    dft8(t);
#pragma unroll
    for (int i = 0; i < 8; ++i) v[BASE + i] = t[i];
  } else if constexpr (R == 4) {
    dft4(v[BASE], v[BASE + 1], v[BASE + 2], v[BASE + 3]);
  } else if constexpr (R == 2) {
    dft2(v[BASE], v[BASE + 1]);
  }
}

template <int R, int NB, int B = 0, typename V>
MMF_HD void dft_groups(V (&v)[16]) {
  if constexpr (B < NB) {
    dft_sub<R, B * R>(v);
    dft_groups<R, NB, B + 1>(v);
  }
}

}  // namespace mmf
