// Hilbert amplitude envelope |scipy.signal.hilbert(x)| in O(n log n) for any length n
// (script/calc.py:284-286, method 'Hilb': FFT of length n, negative frequencies zeroed, inverse FFT).
//
// scipy transforms at the signal's own length (160 000 for a 10 s clip, 57.6 M for an hour): not a power of
// two, often with large prime factors.  Bluestein's identity turns a length-n DFT into a circular convolution
// of length M = 2^k >= 2n - 1 with the chirp c[j] = exp(-i pi j^2 / n):
//     X[k] = c[k] * sum_j (x[j] c[j]) conj(c[k - j]).
// The convolutions run through a radix-2 Stockham FFT of length M (natural order in and out, log2 M streaming
// passes, twiddles from a table built once per call); chirp phases are reduced exactly (j^2 mod 2n in integer
// arithmetic).  Everything is float64 -- the flops are negligible next to the streaming passes, and the result
// is then exact to ~1e-12 (scipy itself computes float32 input in single precision).
// Envelope = | IDFT_n( H . DFT_n(x) ) | with H = 1 (k = 0, and n/2 for even n), 2 (0 < k < n/2), 0 otherwise;
// the inverse transform reuses the forward machinery on conjugated data.
#include <cmath>

#include "mmf_internal.h"

namespace mmf {

namespace {

__device__ __forceinline__ double2 cmul_d(double2 a, double2 b) {
  return make_double2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}

// c[j] = exp(-i pi j^2 / n), j < n, with j^2 reduced mod 2n exactly
__global__ void chirp_kernel(long n, double2* __restrict__ c) {
  const long j = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n) return;
  const unsigned long long r = ((unsigned long long)j * (unsigned long long)j) % (unsigned long long)(2 * n);  // j < 2^26
  double s, co;
  sincospi(-(double)r / (double)n, &s, &co);
  c[j] = make_double2(co, s);
}

// W[k] = exp(-2 pi i k / M), k < M/2
__global__ void twiddle_kernel(long half, double2* __restrict__ w) {
  const long k = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= half) return;
  double s, co;
  sincospi(-(double)k / (double)half, &s, &co);
  w[k] = make_double2(co, s);
}

// b[j] = conj(c[|j|]) wrapped to length M: b[j] for j < n, b[M - j] for 0 < j < n, zero between
__global__ void chirp_filter_kernel(long n, long M, const double2* __restrict__ c, double2* __restrict__ b) {
  const long j = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= M) return;
  double2 v = make_double2(0.0, 0.0);
  if (j < n) v = make_double2(c[j].x, -c[j].y);
  else if (M - j < n) v = make_double2(c[M - j].x, -c[M - j].y);
  b[j] = v;
}

// one radix-2 Stockham pass: sub-transforms of length Ns -> 2 Ns
__global__ void stockham_pass_kernel(const double2* __restrict__ in, double2* __restrict__ out, long half, long Ns,
                                     const double2* __restrict__ w, long wstride) {
  const long j = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= half) return;
  const long k = j & (Ns - 1);
  const double2 a = in[j];
  const double2 t = cmul_d(in[j + half], w[k * wstride]);
  const long j0 = ((j - k) << 1) + k;
  out[j0] = make_double2(a.x + t.x, a.y + t.y);
  out[j0 + Ns] = make_double2(a.x - t.x, a.y - t.y);
}

// a[j] = x[j] c[j] (real input), zero padded to M
__global__ void load_real_kernel(const float* __restrict__ x, long n, long M, const double2* __restrict__ c,
                                 double2* __restrict__ a) {
  const long j = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= M) return;
  double2 v = make_double2(0.0, 0.0);
  if (j < n) {
    const double xv = (double)x[j];
    v = make_double2(xv * c[j].x, xv * c[j].y);
  }
  a[j] = v;
}

// pointwise product with the transformed chirp filter, conjugated for the inverse-by-conjugation trick:
// out = conj(a * bf)
__global__ void mul_conj_kernel(double2* __restrict__ a, const double2* __restrict__ bf, long M) {
  const long j = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= M) return;
  const double2 p = cmul_d(a[j], bf[j]);
  a[j] = make_double2(p.x, -p.y);
}

// after the second forward FFT of conj(A.Bf): conv[k] = conj(buf[k]) / M; X[k] = c[k] conv[k].
// Apply the analytic-signal mask and prepare the inverse DFT_n by conjugation:
// next[j] = conj(H[j] X[j]) c[j] for j < n, zero padded to M
__global__ void mask_kernel(const double2* __restrict__ buf, long n, long M, const double2* __restrict__ c,
                            double2* __restrict__ next) {
  const long j = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= M) return;
  double2 v = make_double2(0.0, 0.0);
  if (j < n) {
    const double inv = 1.0 / (double)M;
    const double2 conv = make_double2(buf[j].x * inv, -buf[j].y * inv);
    const double2 X = cmul_d(c[j], conv);
    double h;
    if (n % 2 == 0)
      h = (j == 0 || j == n / 2) ? 1.0 : (j < n / 2 ? 2.0 : 0.0);
    else
      h = j == 0 ? 1.0 : (j < (n + 1) / 2 ? 2.0 : 0.0);
    const double2 Yc = make_double2(h * X.x, -h * X.y);  // conj(H X)
    v = cmul_d(Yc, c[j]);
  }
  next[j] = v;
}

// z[k] = conj(c[k] conv[k]) / n with conv[k] = conj(buf[k]) / M; envelope = |z|
__global__ void envelope_kernel(const double2* __restrict__ buf, long n, long M, const double2* __restrict__ c,
                                float* __restrict__ amp) {
  const long j = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n) return;
  const double inv = 1.0 / ((double)M * (double)n);
  const double2 conv = make_double2(buf[j].x * inv, -buf[j].y * inv);
  const double2 z = cmul_d(c[j], conv);
  amp[j] = (float)sqrt(z.x * z.x + z.y * z.y);
}

unsigned blocks_for(long n) { return (unsigned)((n + 255) / 256); }

// forward FFT of length M (power of two) from `a`; returns the buffer holding the result (a or tmp)
double2* fft_pow2(double2* a, double2* tmp, long M, const double2* w, cudaStream_t st) {
  const long half = M / 2;
  double2 *in = a, *out = tmp;
  for (long Ns = 1; Ns < M; Ns <<= 1) {
    stockham_pass_kernel<<<blocks_for(half), 256, 0, st>>>(in, out, half, Ns, w, half / Ns);
    count_launch();
    double2* t = in;
    in = out;
    out = t;
  }
  return in;
}

}  // namespace

bool hilbert_fft_supported(long n) { return n >= 2 && n <= (1L << 26); }

cudaError_t hilbert_fft_launch(const float* x, long n_clips, long n, long stride, float* amp, long amp_stride,
                               cudaStream_t st) {
  long M = 1;
  while (M < 2 * n - 1) M <<= 1;
  double2 *c = nullptr, *w = nullptr, *bf = nullptr, *a = nullptr, *t = nullptr;
  cudaError_t e;
  if ((e = cudaMallocAsync((void**)&c, (size_t)n * 16, st)) != cudaSuccess) return e;
  if ((e = cudaMallocAsync((void**)&w, (size_t)(M / 2) * 16, st)) != cudaSuccess) return e;
  if ((e = cudaMallocAsync((void**)&bf, (size_t)M * 16, st)) != cudaSuccess) return e;
  if ((e = cudaMallocAsync((void**)&a, (size_t)M * 16, st)) != cudaSuccess) return e;
  if ((e = cudaMallocAsync((void**)&t, (size_t)M * 16, st)) != cudaSuccess) return e;
  chirp_kernel<<<blocks_for(n), 256, 0, st>>>(n, c);
  twiddle_kernel<<<blocks_for(M / 2), 256, 0, st>>>(M / 2, w);
  // transformed chirp filter, once per call (left in whichever buffer the FFT ends in; copy to bf if needed)
  chirp_filter_kernel<<<blocks_for(M), 256, 0, st>>>(n, M, c, a);
  count_launch(3);
  double2* r = fft_pow2(a, t, M, w, st);
  if ((e = cudaMemcpyAsync(bf, r, (size_t)M * 16, cudaMemcpyDeviceToDevice, st)) != cudaSuccess) return e;
  for (long clip = 0; clip < n_clips; ++clip) {
    load_real_kernel<<<blocks_for(M), 256, 0, st>>>(x + clip * stride, n, M, c, a);
    r = fft_pow2(a, t, M, w, st);
    double2* o = r == a ? t : a;
    mul_conj_kernel<<<blocks_for(M), 256, 0, st>>>(r, bf, M);
    r = fft_pow2(r, o, M, w, st);
    o = r == a ? t : a;
    mask_kernel<<<blocks_for(M), 256, 0, st>>>(r, n, M, c, o);
    r = fft_pow2(o, r, M, w, st);
    o = r == a ? t : a;
    mul_conj_kernel<<<blocks_for(M), 256, 0, st>>>(r, bf, M);
    r = fft_pow2(r, o, M, w, st);
    envelope_kernel<<<blocks_for(n), 256, 0, st>>>(r, n, M, c, amp + clip * amp_stride);
    count_launch(5);
  }
  cudaFreeAsync(c, st);
  cudaFreeAsync(w, st);
  cudaFreeAsync(bf, st);
  cudaFreeAsync(a, st);
  cudaFreeAsync(t, st);
  return cudaGetLastError();
}

}  // namespace mmf
