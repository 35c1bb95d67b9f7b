// 512-point frames on the 5th-generation tensor cores (opt-in: MMF_FLAG_TC_FFT, power spectrum only so far).
//
// librosa.stft under script/mfcc.py:387 for n_fft = 512: the real frame is packed into 256 complex points
// z[n] = x[2n] + i x[2n+1], transformed as two radix-16 stages (n = n1 + 16 n2, k = 16 k1 + k2), each stage a
// real GEMM  [128 rows x 32] . [32 x 32]  (rows = 8 frames x 16 sub-vectors, columns = re/im interleaved)
// against the real representation R of the 16-point DFT matrix, issued as tcgen05.mma kind::f16 with the
// accumulators in tensor memory:
//
//   operands are fp16 pairs   x * s = hi + lo / 2048          (s: power of two per frame, |x s| < 64)
//                             R     = Rhi + Rlo / 2048
//   D[:, 0:32]  = hi . Rhi                                      (one MMA per K = 16 slab, N = 64 with B = [Rhi | Rlo])
//   D[:, 32:64] = hi . Rlo + lo . Rhi                            (second MMA, N = 32, accumulating)
//   result      = (D[:, 0:32] + D[:, 32:64] / 2048) / s          -- 22-bit operands, fp32 accumulation
//
// tests/studies/tf32_dft_study.py: mel-power error of this split 1.6e-6 (plain fp32 FFT: 1.6e-6); tools/ubench/
// tcgen05_f16.cu: descriptors validated against the host, 48 + 47 cycles per K = 16 slab.
//
// Stage 1 A operand: K-major (thread (f, n1) owns a row and writes 16-byte chunks).  Stage 2 needs rows
// (f, k2) with K = n1, i.e. the transpose inside every frame: written MN-major (the thread owns one K pair
// and eight consecutive rows per 16-byte store), with the core-matrix stride padded to 144 bytes against bank
// aliasing -- the tensor core does the transpose through the descriptor.
#include <cuda_fp16.h>

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>

#include "mmf_internal.h"

namespace mmf {

namespace {

constexpr int kTcThreads = 128;  // one thread per GEMM row = per TMEM lane
constexpr int kFB = 8;           // frames per row block
constexpr int kZP = 257;         // pitch (complex) of the transformed frames in shared memory
constexpr int kA2Lbo = 144;      // bytes between K-groups of the MN-major stage-2 operand (128 + 16 pad)
constexpr int kA2Sbo = 4 * kA2Lbo;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
  uint64_t d = (uint64_t)((saddr & 0x3FFFFu) >> 4);
  d |= (uint64_t)((lbo >> 4) & 0x3FFFu) << 16;
  d |= (uint64_t)((sbo >> 4) & 0x3FFFu) << 32;
  d |= 1ull << 46;  // descriptor version; no swizzle, base offset 0
  return d;
}

// fp16 operands, fp32 accumulate, B K-major; a_mn: A is MN-major
__device__ __forceinline__ uint32_t make_idesc(int N, bool a_mn) {
  return (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24) | (a_mn ? (1u << 15) : 0u);
}

__device__ __forceinline__ void mma_f16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, {%5, %5, %5, %5}, p;\n\t"
      "}\n" ::"r"(d_tmem),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc), "r"(0u)
      : "memory");
}

__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok = 0;
  while (!ok) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t"
        "}\n"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
  }
}

// 32 consecutive accumulator columns of this thread's TMEM lane
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float* v) {
  uint32_t* u = reinterpret_cast<uint32_t*>(v);
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : "=r"(u[0]), "=r"(u[1]), "=r"(u[2]), "=r"(u[3]), "=r"(u[4]), "=r"(u[5]), "=r"(u[6]), "=r"(u[7]), "=r"(u[8]),
        "=r"(u[9]), "=r"(u[10]), "=r"(u[11]), "=r"(u[12]), "=r"(u[13]), "=r"(u[14]), "=r"(u[15]), "=r"(u[16]),
        "=r"(u[17]), "=r"(u[18]), "=r"(u[19]), "=r"(u[20]), "=r"(u[21]), "=r"(u[22]), "=r"(u[23]), "=r"(u[24]),
        "=r"(u[25]), "=r"(u[26]), "=r"(u[27]), "=r"(u[28]), "=r"(u[29]), "=r"(u[30]), "=r"(u[31])
      : "r"(taddr)
      : "memory");
}

// power-of-two scale that brings `m` (>= 0) into [32, 64); returns the scale, *inv its inverse
__device__ __forceinline__ float pow2_scale(float m, float* inv) {
  int e = (int)((__float_as_uint(m) >> 23) & 0xffu) - 127;  // floor(log2 m) for normal m
  if (m < 1e-30f) e = 5;                                    // silent frame: scale 1
  e = max(-100, min(100, e));
  *inv = __uint_as_float((uint32_t)(e - 5 + 127) << 23);
  return __uint_as_float((uint32_t)(5 - e + 127) << 23);
}

// v * s -> fp16 hi, fp16 lo with v s = hi + lo / 2048 (to 22 bits)
__device__ __forceinline__ void split_f16(float v, __half* hi, __half* lo) {
  const __half h = __float2half_rn(v);
  *hi = h;
  *lo = __float2half_rn((v - __half2float(h)) * 2048.0f);
}

struct alignas(16) Half8 {
  __half2 a, b, c, d;
};

}  // namespace

struct TcFftArgs {
  const float* pcm;
  long n_samples, clip_stride;
  int T, hop;
  int blocks_per_clip;
  long n_blocks;
  const float* window;   // [512] padded Hann
  const __half* btab;    // canonical K-major [Rhi | Rlo] (64 x 32) followed by Rhi (32 x 32)
  const float2* tw;      // [16][16] W256^(n1 k2)
  float* power;          // [clips][257][T]
};

__global__ void __launch_bounds__(kTcThreads) tc_fft512_kernel(const TcFftArgs p) {
  extern __shared__ __align__(1024) unsigned char sm_tc[];
  // carve (bytes): B tables 6144 | A1 hi 8192 | A1 lo 8192 | A2 hi 9216 | A2 lo 9216 | window 2048 | tw 2048 | W512 2064 | span
  __half* sB64 = reinterpret_cast<__half*>(sm_tc);
  __half* sB32 = sB64 + 64 * 32;
  unsigned char* sA1h = sm_tc + 6144;
  unsigned char* sA1l = sA1h + 8192;
  unsigned char* sA2h = sA1l + 8192;
  unsigned char* sA2l = sA2h + 16 * kA2Sbo;
  float* s_win = reinterpret_cast<float*>(sA2l + 16 * kA2Sbo);
  float2* s_tw = reinterpret_cast<float2*>(s_win + 512);
  float2* s_w512 = s_tw + 256;  // [257] (cos, sin)(pi k / 256), 8-byte padded to 258
  float* s_span = reinterpret_cast<float*>(s_w512 + 258);
  // the transformed frames reuse the stage-1 operand area: 8 x 257 complex = 16448 bytes > 16384, so they
  // start at A1 and run 64 bytes into A2 hi, which is dead by then (stage 2 has completed)
  float2* sZ = reinterpret_cast<float2*>(sA1h);
  __shared__ __align__(8) unsigned long long bar;
  __shared__ uint32_t tmem_base;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  for (int i = tid; i < (64 + 32) * 32 / 8; i += kTcThreads)
    reinterpret_cast<uint4*>(sB64)[i] = reinterpret_cast<const uint4*>(p.btab)[i];
  for (int i = tid; i < 512; i += kTcThreads) s_win[i] = p.window[i];
  for (int i = tid; i < 256; i += kTcThreads) s_tw[i] = p.tw[i];
  for (int i = tid; i < 257; i += kTcThreads) {
    float sn, cs;
    sincospif((float)i * (1.0f / 256.0f), &sn, &cs);
    s_w512[i] = make_float2(cs, sn);
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base)), "r"(128)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t d1 = tmem_base, d2 = tmem_base + 64;
  const uint32_t lane_base = (uint32_t)(warp * 32) << 16;
  const uint32_t bar_a = smem_u32(&bar);
  uint32_t parity = 0;

  // descriptors (uniform per CTA)
  const uint64_t b64_desc = make_desc(smem_u32(sB64), 128, 512);
  const uint64_t b32_desc = make_desc(smem_u32(sB32), 128, 512);
  const uint64_t a1h_desc = make_desc(smem_u32(sA1h), 128, 512);
  const uint64_t a1l_desc = make_desc(smem_u32(sA1l), 128, 512);
  const uint64_t a2h_desc = make_desc(smem_u32(sA2h), kA2Lbo, kA2Sbo);
  const uint64_t a2l_desc = make_desc(smem_u32(sA2l), kA2Lbo, kA2Sbo);
  const uint32_t i64k = make_idesc(64, false), i32k = make_idesc(32, false);
  const uint32_t i64m = make_idesc(64, true), i32m = make_idesc(32, true);

  const bool vec_ok = (p.hop & 1) == 0;
  const int f = tid >> 4, n1 = tid & 15;  // stage 1: row = (frame, n1); stage 2: row = (frame, k2 = n1)
  const int span_len = (kFB - 1) * p.hop + 512;
  float2 wreg[16];  // window at the thread's own positions 2 (n1 + 16 n2), +1
#pragma unroll
  for (int n2 = 0; n2 < 16; ++n2) wreg[n2] = *reinterpret_cast<const float2*>(s_win + 2 * (n1 + 16 * n2));

  for (long blk = blockIdx.x; blk < p.n_blocks; blk += gridDim.x) {
    const long clip = blk / p.blocks_per_clip;
    const int t0 = (int)(blk - clip * p.blocks_per_clip) * kFB;
    // ---- PCM span of the 8 frames, zero outside the clip (librosa center padding)
    {
      const long s0 = (long)t0 * p.hop - 256;
      const float* src = p.pcm + clip * p.clip_stride;
      for (int i = tid; i < span_len; i += kTcThreads) {
        const long s = s0 + i;
        s_span[i] = (s >= 0 && s < p.n_samples) ? __ldg(src + s) : 0.0f;
      }
    }
    __syncthreads();
    // ---- stage-1 operand: window, per-frame scale, fp16 split, K-major rows
    float inv1;
    {
      float v[32];
      const float* x = s_span + f * p.hop;
      float m = 0.0f;
#pragma unroll
      for (int n2 = 0; n2 < 16; ++n2) {
        const int n = 2 * (n1 + 16 * n2);
        float2 xv;
        if (vec_ok) xv = *reinterpret_cast<const float2*>(x + n);  // (3) hop even: 8-byte aligned
        else xv = make_float2(x[n], x[n + 1]);
        v[2 * n2] = xv.x * wreg[n2].x;
        v[2 * n2 + 1] = xv.y * wreg[n2].y;
        m = fmaxf(m, fmaxf(fabsf(v[2 * n2]), fabsf(v[2 * n2 + 1])));
      }
#pragma unroll
      for (int o = 8; o >= 1; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
      const float s = pow2_scale(m, &inv1);
#pragma unroll
      for (int kc = 0; kc < 4; ++kc) {
        __half h[8], l[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) split_f16(v[8 * kc + j] * s, &h[j], &l[j]);
        const int off = ((tid >> 3) * 4 + kc) * 128 + (tid & 7) * 16;
        *reinterpret_cast<Half8*>(sA1h + off) = Half8{__halves2half2(h[0], h[1]), __halves2half2(h[2], h[3]),
                                                      __halves2half2(h[4], h[5]), __halves2half2(h[6], h[7])};
        *reinterpret_cast<Half8*>(sA1l + off) = Half8{__halves2half2(l[0], l[1]), __halves2half2(l[2], l[3]),
                                                      __halves2half2(l[4], l[5]), __halves2half2(l[6], l[7])};
      }
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (tid == 0) {
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      mma_f16(d1, a1h_desc, b64_desc, i64k, 0u);
      mma_f16(d1, a1h_desc + 16, b64_desc + 16, i64k, 1u);
      mma_f16(d1 + 32, a1l_desc, b32_desc, i32k, 1u);
      mma_f16(d1 + 32, a1l_desc + 16, b32_desc + 16, i32k, 1u);
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar_a) : "memory");
    }
    mbar_wait(bar_a, parity);
    parity ^= 1u;
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    // ---- stage-1 result of row (f, n1): combine, twiddle W256^(n1 k2), rescale, split, MN-major store
    float inv2;
    {
      float c0[32], c1[32];
      tmem_ld32(d1 + lane_base, c0);
      tmem_ld32(d1 + lane_base + 32, c1);
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      float m = 0.0f;
#pragma unroll
      for (int k2 = 0; k2 < 16; ++k2) {
        const float re = fmaf(c1[2 * k2], 1.0f / 2048.0f, c0[2 * k2]) * inv1;
        const float im = fmaf(c1[2 * k2 + 1], 1.0f / 2048.0f, c0[2 * k2 + 1]) * inv1;
        const float2 w = s_tw[n1 * 16 + k2];
        c0[2 * k2] = re * w.x - im * w.y;
        c0[2 * k2 + 1] = re * w.y + im * w.x;
        m = fmaxf(m, fmaxf(fabsf(c0[2 * k2]), fabsf(c0[2 * k2 + 1])));
      }
#pragma unroll
      for (int o = 8; o >= 1; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
      const float s = pow2_scale(m, &inv2);
      // element (row m = 16 f + k2, k = 2 n1 + c) at (m >> 3) * SBO + (k >> 3) * LBO + (k & 7) * 16 + (m & 7) * 2
#pragma unroll
      for (int c = 0; c < 2; ++c) {
#pragma unroll
        for (int g = 0; g < 2; ++g) {
          __half h[8], l[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) split_f16(c0[2 * (8 * g + j) + c] * s, &h[j], &l[j]);
          const int k = 2 * n1 + c;
          const int off = (2 * f + g) * kA2Sbo + (k >> 3) * kA2Lbo + (k & 7) * 16;
          *reinterpret_cast<Half8*>(sA2h + off) = Half8{__halves2half2(h[0], h[1]), __halves2half2(h[2], h[3]),
                                                        __halves2half2(h[4], h[5]), __halves2half2(h[6], h[7])};
          *reinterpret_cast<Half8*>(sA2l + off) = Half8{__halves2half2(l[0], l[1]), __halves2half2(l[2], l[3]),
                                                        __halves2half2(l[4], l[5]), __halves2half2(l[6], l[7])};
        }
      }
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (tid == 0) {
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      constexpr uint64_t kStep = 2 * kA2Lbo / 16;  // descriptor units per K = 16 slab
      mma_f16(d2, a2h_desc, b64_desc, i64m, 0u);
      mma_f16(d2, a2h_desc + kStep, b64_desc + 16, i64m, 1u);
      mma_f16(d2 + 32, a2l_desc, b32_desc, i32m, 1u);
      mma_f16(d2 + 32, a2l_desc + kStep, b32_desc + 16, i32m, 1u);
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar_a) : "memory");
    }
    mbar_wait(bar_a, parity);
    parity ^= 1u;
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    // ---- stage-2 result of row (f, k2): Z[16 k1 + k2]
    {
      float c0[32], c1[32];
      tmem_ld32(d2 + lane_base, c0);
      tmem_ld32(d2 + lane_base + 32, c1);
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      // the scale of stage 2 was taken per frame over the rows (f, n1); rows (f, k2) of the same frame share it
      float2* z = sZ + f * kZP + n1;
#pragma unroll
      for (int k1 = 0; k1 < 16; ++k1) {
        const float re = fmaf(c1[2 * k1], 1.0f / 2048.0f, c0[2 * k1]) * inv2;
        const float im = fmaf(c1[2 * k1 + 1], 1.0f / 2048.0f, c0[2 * k1 + 1]) * inv2;
        z[16 * k1] = make_float2(re, im);
      }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    // ---- real-FFT split step and |X|^2, bins 0 .. 256 of the 8 frames
    {
      float* dst = p.power + (size_t)clip * 257 * p.T + t0;
      const int t_valid = min(kFB, p.T - t0);
      for (int e = tid; e < 257 * kFB; e += kTcThreads) {
        const int k = e >> 3, t = e & 7;
        const float2 a = sZ[t * kZP + (k & 255)];
        const float2 b = sZ[t * kZP + ((256 - k) & 255)];
        // X[k] = (Z[k] + conj Z[256-k]) / 2 - i W512^k (Z[k] - conj Z[256-k]) / 2
        const float er = 0.5f * (a.x + b.x), ei = 0.5f * (a.y - b.y);
        const float orr = 0.5f * (a.x - b.x), oi = 0.5f * (a.y + b.y);
        const float cs = s_w512[k].x, sn = s_w512[k].y;  // W512^k = cs - i sn
        // -i W (o) with o = orr + i oi:  W o = (cs orr + sn oi) + i (cs oi - sn orr);  -i (x + i y) = y - i x
        const float wr = cs * orr + sn * oi, wi = cs * oi - sn * orr;
        const float xr = er + wi, xi = ei - wr;
        if (t < t_valid) dst[(size_t)k * p.T + t] = xr * xr + xi * xi;
      }
    }
    __syncthreads();  // sZ (aliasing the operands) and the span are free again
  }

  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(128) : "memory");
}

// ---------------------------------------------------------------------------------------------------
// host
// ---------------------------------------------------------------------------------------------------
void tc_fft_tables(std::vector<uint16_t>& btab, std::vector<float>& tw) {
  // R: real representation of the 16-point DFT matrix acting on interleaved (re, im) row vectors
  const double kPi = 3.14159265358979323846;
  std::vector<float> R(32 * 32, 0.0f);
  for (int n = 0; n < 16; ++n)
    for (int j = 0; j < 16; ++j) {
      const int q = (n * j) % 16;
      double c = std::cos(2.0 * kPi * q / 16.0), s = -std::sin(2.0 * kPi * q / 16.0);
      if (q % 4 == 0) {  // exact zeros and ones
        c = (q == 0) ? 1.0 : (q == 8) ? -1.0 : 0.0;
        s = (q == 4) ? -1.0 : (q == 12) ? 1.0 : 0.0;
      }
      R[(2 * n) * 32 + 2 * j] = (float)c;
      R[(2 * n) * 32 + 2 * j + 1] = (float)s;
      R[(2 * n + 1) * 32 + 2 * j] = (float)-s;
      R[(2 * n + 1) * 32 + 2 * j + 1] = (float)c;
    }
  auto to_bits = [](__half h) {
    uint16_t u;
    std::memcpy(&u, &h, 2);
    return u;
  };
  // B[n][k] = R[k][n]; canonical K-major: (n / 8) * 4 * 64 + (k / 8) * 64 + (n % 8) * 8 + (k % 8) halves
  btab.assign((64 + 32) * 32, 0);
  for (int n = 0; n < 64; ++n)
    for (int k = 0; k < 32; ++k) {
      const float r = R[k * 32 + (n & 31)];
      const __half hi = __float2half_rn(r);
      const __half lo = __float2half_rn((r - __half2float(hi)) * 2048.0f);
      const int idx = (n / 8) * 4 * 64 + (k / 8) * 64 + (n % 8) * 8 + (k % 8);
      btab[idx] = to_bits(n < 32 ? hi : lo);
      if (n < 32) btab[64 * 32 + idx] = to_bits(hi);
    }
  tw.assign(2 * 256, 0.0f);
  for (int a = 0; a < 16; ++a)
    for (int b = 0; b < 16; ++b) {
      tw[2 * (a * 16 + b)] = (float)std::cos(2.0 * kPi * (a * b) / 256.0);
      tw[2 * (a * 16 + b) + 1] = (float)-std::sin(2.0 * kPi * (a * b) / 256.0);
    }
}

bool tc_fft_supported(int n_fft, int hop, float preemph) { return n_fft == 512 && hop >= 1 && hop <= 512 && preemph == 0.0f; }

cudaError_t tc_fft_power_launch(const float* pcm, long n_clips, long n_samples, long clip_stride, int T, int hop,
                                const float* window, const void* btab, const float2* tw, float* power, int sm_count,
                                cudaStream_t st) {
  TcFftArgs a{};
  a.pcm = pcm;
  a.n_samples = n_samples;
  a.clip_stride = clip_stride;
  a.T = T;
  a.hop = hop;
  a.blocks_per_clip = (T + kFB - 1) / kFB;
  a.n_blocks = (long)a.blocks_per_clip * n_clips;
  a.window = window;
  a.btab = reinterpret_cast<const __half*>(btab);
  a.tw = tw;
  a.power = power;
  const size_t smem = 6144 + 2 * 8192 + 2 * 16 * kA2Sbo + 2048 + 2048 + 258 * 8 + (size_t)((kFB - 1) * hop + 512) * 4;
  MMF_SMEM_ONCE(tc_fft512_kernel, 100 * 1024);
  // four CTAs per SM: 128 of the 512 TMEM columns and ~52 KB of shared memory each
  const long grid = std::min<long>(a.n_blocks, (long)sm_count * 4);
  tc_fft512_kernel<<<(unsigned)grid, kTcThreads, smem, st>>>(a);
  count_launch();
  return cudaGetLastError();
}

}  // namespace mmf
