// Generic frame transform for the n_fft the register FFT does not cover (anything that is not a power of two in
// [256, 4096]: librosa.stft under script/mfcc.py:387 takes any n_fft, and the GUI's n_fft field is free text).
//
// |rfft(w * frame)|^2 as a plain FP32 matrix product: C[frames x 2F] = A[frames x n_fft] . B[n_fft x 2F] with
// A[t][n] = y_pad[t*hop + n] (the centre padding is the bounds check), B[n][2k] = w[n] cos(2 pi n k / n_fft),
// B[n][2k+1] = -w[n] sin(2 pi n k / n_fft) (host table, angles reduced exactly, float64 -> float32), then
// re^2 + im^2 in the epilogue.  64 x 64 output tiles, 4 x 4 per thread, operands staged in shared memory.
// A fallback by design: O(n_fft^2) per frame, the power spectrum goes through HBM, and the mel projection is a
// second small kernel (the same sparse walk, mel_column, one frame per thread, reading the spectrum from HBM).
#include <cfloat>
#include <cmath>
#include <vector>

#include "mmf_internal.h"
#include "stft_core.cuh"

namespace mmf {

namespace {

constexpr int kBM = 64, kBN = 64, kBK = 16, kGThreads = 256;

struct DftArgs {
  const float* pcm;
  long n_samples, clip_stride;
  int T, hop, n_fft, F;
  int ld_b;  // columns of the table (2 * F rounded up to kBN)
  float preemph;
  const float* btab;
  float* power;  // [n_clips, F, T]
};

__global__ void __launch_bounds__(kGThreads) dft_power_kernel(const DftArgs p) {
  __shared__ __align__(16) float As[kBK][kBM + 4];
  __shared__ __align__(16) float Bs[kBK][kBN];
  const int tid = threadIdx.x;
  const int t0 = blockIdx.x * kBM, c0 = blockIdx.y * kBN;
  const long clip = blockIdx.z;
  const float* y = p.pcm + (size_t)clip * p.clip_stride;
  const int tx = tid & 15, ty = tid >> 4;  // thread -> columns 4*tx.., frames 4*ty..
  float acc[4][4] = {};
  const int pad = p.n_fft / 2;
  for (int k0 = 0; k0 < p.n_fft; k0 += kBK) {
    // A tile: kBK samples of kBM frames (frames along the fast index: conflict-free reads below)
    for (int e = tid; e < kBK * kBM; e += kGThreads) {
      const int kk = e % kBK, m = e / kBK;
      const int n = k0 + kk, t = t0 + m;
      float v = 0.0f;
      if (n < p.n_fft && t < p.T) {
        const long i = (long)t * p.hop + n - pad;
        if (i >= 0 && i < p.n_samples) {
          v = __ldg(y + i);
          if (p.preemph != 0.0f && i > 0) v -= p.preemph * __ldg(y + i - 1);
        }
      }
      As[kk][m] = v;
    }
    for (int e = tid; e < kBK * kBN / 4; e += kGThreads) {
      const int kk = e / (kBN / 4), c4 = e % (kBN / 4);
      const int n = k0 + kk;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (n < p.n_fft) v = __ldg(reinterpret_cast<const float4*>(p.btab + (size_t)n * p.ld_b + c0) + c4);
      *reinterpret_cast<float4*>(&Bs[kk][4 * c4]) = v;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < kBK; ++kk) {
      const float4 b = *reinterpret_cast<const float4*>(&Bs[kk][4 * tx]);
      const float4 a = *reinterpret_cast<const float4*>(&As[kk][4 * ty]);
      const float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    __syncthreads();
  }
  // columns 4*tx .. 4*tx+3 = (re, im) of bins (c0 + 4*tx)/2 and +1
  float* out = p.power + (size_t)clip * p.F * p.T;
#pragma unroll
  for (int jb = 0; jb < 2; ++jb) {
    const int k = (c0 + 4 * tx) / 2 + jb;
    if (k >= p.F) continue;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int t = t0 + 4 * ty + i;
      if (t < p.T) out[(size_t)k * p.T + t] = acc[i][2 * jb] * acc[i][2 * jb] + acc[i][2 * jb + 1] * acc[i][2 * jb + 1];
    }
  }
}

struct MelPowArgs {
  const float* power;  // [n_clips, F, T]
  int F, T, n_mels;
  float amin;
  const int* seg_start;
  const float2* w2;
  float* logmel;  // [n_clips, n_mels, T]
  int* clipmax;
};

__device__ __forceinline__ int g_float_key(float f) {
  const int b = __float_as_int(f);
  return b >= 0 ? b : b ^ 0x7FFFFFFF;
}

__global__ void __launch_bounds__(128) mel_from_power_kernel(const MelPowArgs p) {
  const long clip = blockIdx.y;
  const int t = blockIdx.x * 128 + threadIdx.x;
  float mx = -FLT_MAX;
  if (t < p.T) {
    const float* col = p.power + (size_t)clip * p.F * p.T + t;
    float* dst = p.logmel + (size_t)clip * p.n_mels * p.T + t;
    mel_column<0>(col, p.T, p.seg_start, p.w2, 0, p.n_mels, [&](int m, float val) {
      float lg;
      asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(lg) : "f"(fmaxf(p.amin, val)));
      const float db = 3.01029995663981195f * lg;
      dst[(size_t)m * p.T] = db;
      mx = fmaxf(mx, db);
    });
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  if ((threadIdx.x & 31) == 0 && mx > -FLT_MAX) atomicMax(p.clipmax + clip, g_float_key(mx));
}

}  // namespace

// [n_fft][ld] table, ld = 2*F rounded up to the column tile; window folded in
void dft_generic_table(int n_fft, const std::vector<float>& window, std::vector<float>& tab, int* ld_out) {
  const int F = n_fft / 2 + 1;
  const int ld = (2 * F + kBN - 1) / kBN * kBN;
  tab.assign((size_t)n_fft * ld, 0.0f);
  const double kPi = 3.14159265358979323846;
  for (int n = 0; n < n_fft; ++n) {
    const double w = window[n];
    if (w == 0.0) continue;
    for (int k = 0; k < F; ++k) {
      const long q = ((long)n * k) % n_fft;  // exact angle reduction
      const double ang = 2.0 * kPi * (double)q / (double)n_fft;
      tab[(size_t)n * ld + 2 * k] = (float)(w * std::cos(ang));
      tab[(size_t)n * ld + 2 * k + 1] = (float)(-w * std::sin(ang));
    }
  }
  *ld_out = ld;
}

cudaError_t dft_generic_power_launch(const float* pcm, long n_clips, long n_samples, long clip_stride, int T, int hop,
                                     int n_fft, float preemph, const float* btab, int ld_b, float* power,
                                     cudaStream_t st) {
  DftArgs a{pcm, n_samples, clip_stride, T, hop, n_fft, n_fft / 2 + 1, ld_b, preemph, btab, power};
  for (long c0 = 0; c0 < n_clips; c0 += 65535) {
    const long nc = std::min<long>(65535, n_clips - c0);
    DftArgs b = a;
    b.pcm = pcm + (size_t)c0 * clip_stride;
    b.power = power + (size_t)c0 * a.F * T;
    dim3 grid((T + kBM - 1) / kBM, ld_b / kBN, (unsigned)nc);
    dft_power_kernel<<<grid, kGThreads, 0, st>>>(b);
    count_launch();
  }
  return cudaGetLastError();
}

cudaError_t mel_from_power_launch(const float* power, long n_clips, int F, int T, int n_mels, float amin,
                                  const int* seg_start, const float2* w2, float* logmel, int* clipmax,
                                  cudaStream_t st) {
  for (long c0 = 0; c0 < n_clips; c0 += 65535) {
    const long nc = std::min<long>(65535, n_clips - c0);
    MelPowArgs a{power + (size_t)c0 * F * T, F, T, n_mels, amin, seg_start, w2, logmel + (size_t)c0 * n_mels * T,
                 clipmax + c0};
    dim3 grid((T + 127) / 128, (unsigned)nc);
    mel_from_power_kernel<<<grid, 128, 0, st>>>(a);
    count_launch();
  }
  return cudaGetLastError();
}

}  // namespace mmf
