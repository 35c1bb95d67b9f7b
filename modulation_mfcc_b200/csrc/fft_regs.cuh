// Register-resident small DFTs (radix 2/4/8/16) used by the fused STFT kernel.
//
// Every function works on a fully unrolled float2 array so the values stay in
// registers; inputs and outputs are in natural order, forward transform
// (e^{-2*pi*i*n*k/R}).  The functions are __host__ __device__ so the host
// emulator under tests/emu/ exercises the exact same arithmetic as the kernel.
#pragma once
#include <cuda_runtime.h>

#define MMF_HD __host__ __device__ __forceinline__

namespace mmf {

MMF_HD float2 cadd(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
MMF_HD float2 csub(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }
MMF_HD float2 cmul(float2 a, float2 w) {
  return make_float2(a.x * w.x - a.y * w.y, a.x * w.y + a.y * w.x);
}
// multiply by -i
MMF_HD float2 cmul_mi(float2 a) { return make_float2(a.y, -a.x); }

// e^{-2*pi*i*J/16} for the J that occur inside the 16-point butterfly.
template <int J>
MMF_HD float2 mul_w16(float2 a) {
  constexpr float C1 = 0.92387953251128674f;  // cos(pi/8)
  constexpr float S1 = 0.38268343236508977f;  // sin(pi/8)
  constexpr float R = 0.70710678118654752f;   // sqrt(1/2)
  if constexpr (J == 0) {
    return a;
  } else if constexpr (J == 1) {
    return make_float2(a.x * C1 + a.y * S1, a.y * C1 - a.x * S1);
  } else if constexpr (J == 2) {
    return make_float2((a.x + a.y) * R, (a.y - a.x) * R);
  } else if constexpr (J == 3) {
    return make_float2(a.x * S1 + a.y * C1, a.y * S1 - a.x * C1);
  } else if constexpr (J == 4) {
    return make_float2(a.y, -a.x);
  } else if constexpr (J == 6) {
    return make_float2((a.y - a.x) * R, -(a.x + a.y) * R);
  } else {
    static_assert(J == 9, "unexpected twiddle");
    return make_float2(-a.x * C1 - a.y * S1, a.x * S1 - a.y * C1);
  }
}

MMF_HD void dft2(float2& a, float2& b) {
  float2 t = a;
  a = cadd(t, b);
  b = csub(t, b);
}

MMF_HD void dft4(float2& a0, float2& a1, float2& a2, float2& a3) {
  float2 s02 = cadd(a0, a2), d02 = csub(a0, a2);
  float2 s13 = cadd(a1, a3), d13 = csub(a1, a3);
  a0 = cadd(s02, s13);
  a2 = csub(s02, s13);
  a1 = make_float2(d02.x + d13.y, d02.y - d13.x);  // d02 - i*d13
  a3 = make_float2(d02.x - d13.y, d02.y + d13.x);  // d02 + i*d13
}

// 8-point DFT as 2 (n1) x 4 (n2):  n = n1 + 2*n2,  k = 4*k1 + k2.
MMF_HD void dft8(float2 (&v)[8]) {
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
  float2 o[8];
#pragma unroll
  for (int k2 = 0; k2 < 4; ++k2) {
    o[k2] = v[2 * k2];
    o[4 + k2] = v[2 * k2 + 1];
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) v[i] = o[i];
}

// 16-point DFT as 4 x 4:  n = n1 + 4*n2,  k = 4*k1 + k2.
MMF_HD void dft16(float2 (&v)[16]) {
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
  float2 o[16];
#pragma unroll
  for (int k1 = 0; k1 < 4; ++k1)
#pragma unroll
    for (int k2 = 0; k2 < 4; ++k2) o[4 * k1 + k2] = v[k1 + 4 * k2];
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = o[i];
}

// R-point DFT on v[BASE .. BASE+R) of a 16-register file.
template <int R, int BASE>
MMF_HD void dft_sub(float2 (&v)[16]) {
  if constexpr (R == 16) {
    dft16(v);
  } else if constexpr (R == 8) {
    float2 t[8];
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

template <int R, int NB, int B = 0>
MMF_HD void dft_groups(float2 (&v)[16]) {
  if constexpr (B < NB) {
    dft_sub<R, B * R>(v);
    dft_groups<R, NB, B + 1>(v);
  }
}

}  // namespace mmf
