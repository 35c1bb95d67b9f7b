// K3 on the 5th-generation tensor cores: clamp + DCT-II (+ delta) as a tcgen05 GEMM.
//
// power_to_db's top_db clamp and scipy.fftpack.dct(type 2, norm='ortho') under script/mfcc.py:387 for 128
// consecutive frames of a clip at a time:  D[128 frames x n_mfcc] = A[128 x n_mels] . DCT^T, issued as
// tcgen05.mma kind::f16 with fp16 operand pairs (x s = hi + lo / 2048 with a power-of-two s per frame,
// DCT = Bhi + Blo / 2048): D0 = hi.Bhi and D1 = hi.Blo + lo.Bhi in adjacent TMEM columns (one N = 32 MMA per K
// slab for hi against [Bhi ; Blo], one N = 16 MMA for lo against Bhi), result (D0 + D1 / 2048) / s.  A goes from
// registers into tensor memory (one TMEM lane per frame), the 3 KB DCT operand sits in shared memory in the
// canonical K-major layout.  Same structure and descriptors as modspec_tc.cu / tools/ubench/tcgen05_f16.cu.
//
// The FP32 kernel (mfcc_kernel, post_kernels.cu) spends 640 FFMAs per frame; here a frame costs its 40 loads,
// the clamp, the fp16 split and 13 outputs.  delta = np.gradient along time (calc.py:642-645) needs t-1 / t+1:
// a tile owns 126 of its 128 frames (one-frame halo each side, recomputed by the neighbouring tile).
#include <cuda_fp16.h>

#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>

#include "mmf_internal.h"

namespace mmf {

namespace {

constexpr int kDtThreads = 128;
constexpr int kDtN = 16;  // n_mfcc <= 16 output columns per accumulator

__device__ __forceinline__ uint32_t dt_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ uint64_t dt_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
  uint64_t d = (uint64_t)((saddr & 0x3FFFFu) >> 4);
  d |= (uint64_t)((lbo >> 4) & 0x3FFFu) << 16;
  d |= (uint64_t)((sbo >> 4) & 0x3FFFu) << 32;
  d |= 1ull << 46;
  return d;
}

__device__ __forceinline__ void dt_mma_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, {%5, %5, %5, %5}, p;\n\t"
      "}\n" ::"r"(d_tmem),
      "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(acc), "r"(0u)
      : "memory");
}

__device__ __forceinline__ void dt_mbar_wait(uint32_t bar, uint32_t parity) {
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

__device__ __forceinline__ void dt_tmem_ld16(uint32_t taddr, float* v) {
  uint32_t* u = reinterpret_cast<uint32_t*>(v);
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(u[0]), "=r"(u[1]), "=r"(u[2]), "=r"(u[3]), "=r"(u[4]), "=r"(u[5]), "=r"(u[6]), "=r"(u[7]), "=r"(u[8]),
        "=r"(u[9]), "=r"(u[10]), "=r"(u[11]), "=r"(u[12]), "=r"(u[13]), "=r"(u[14]), "=r"(u[15])
      : "r"(taddr)
      : "memory");
}

__device__ __forceinline__ void dt_tmem_st8(uint32_t taddr, const uint32_t* v) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"r"(taddr), "r"(v[0]),
               "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7])
               : "memory");
}

__device__ __forceinline__ float dt_key_to_float(int k) { return __int_as_float(k >= 0 ? k : k ^ 0x7FFFFFFF); }

}  // namespace

struct MfccTcArgs {
  float* logmel;
  const int* clipmax;
  long T;
  int n_mels, n_mfcc;
  float top_db;
  int tiles_per_clip;
  long n_tiles;
  int halo;            // 1 when delta is wanted
  const __half* btab;  // canonical K-major: 32 rows ([Bhi (16) ; Blo (16)]) x (16 KS) halves
  float* mfcc;
  float* delta;
  int clamp_in_place;
};

template <int KS>
__global__ void __launch_bounds__(kDtThreads) mfcc_tc_kernel(const MfccTcArgs p) {
  constexpr int KP = 16 * KS;
  __shared__ __align__(128) __half sB[32 * KP];
  __shared__ float s_col[kDtN][kDtThreads + 1];
  __shared__ __align__(8) unsigned long long bar;
  __shared__ uint32_t tmem_base;
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < 32 * KP / 8; i += kDtThreads)
    reinterpret_cast<uint4*>(sB)[i] = reinterpret_cast<const uint4*>(p.btab)[i];
  constexpr int kCols = 128;  // D0 | D1 (32 columns) + A hi (KP / 2) + A lo (KP / 2), KP <= 96
  static_assert(32 + KP <= kCols, "TMEM budget");
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dt_smem_u32(&tmem_base)), "r"(kCols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(dt_smem_u32(&bar)) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t d_tm = tmem_base, ah_tm = tmem_base + 32, al_tm = ah_tm + KP / 2;
  const uint32_t lane_base = (uint32_t)(warp * 32) << 16;
  const uint32_t bar_a = dt_smem_u32(&bar);
  uint32_t parity = 0;
  const uint64_t b_desc = dt_desc(dt_smem_u32(sB), 128, (uint32_t)(KP / 8) * 128);
  const uint32_t idesc32 = (1u << 4) | ((uint32_t)(32 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
  const uint32_t idesc16 = (1u << 4) | ((uint32_t)(16 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
  const int per_tile = kDtThreads - 2 * p.halo;

  for (long tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x) {
    const long clip = tile / p.tiles_per_clip;
    const int tin = (int)(tile - clip * p.tiles_per_clip);
    const long t = (long)tin * per_tile - p.halo + tid;
    const bool valid = t >= 0 && t < p.T;
    const bool own = valid && tid >= p.halo && tid < kDtThreads - p.halo;
    const float thr = p.top_db >= 0.0f ? dt_key_to_float(p.clipmax[clip]) - p.top_db : -FLT_MAX;
    // ---- this thread's frame: all mel rows in flight at once (clamped index, no predicates), top_db clamp
    float inv_s;
    {
      const long tc = min(max(t, 0L), p.T - 1);
      float* col = p.logmel + (size_t)clip * p.n_mels * p.T + tc;
      float v[KP];
#pragma unroll
      for (int m = 0; m < KP; ++m) v[m] = (m < KP - 16 || m < p.n_mels) ? __ldcs(col + (size_t)min(m, p.n_mels - 1) * p.T) : 0.0f;
      float mx = 0.0f;
#pragma unroll
      for (int m = 0; m < KP; ++m) {
        if (m < KP - 16 || m < p.n_mels) {
          v[m] = fmaxf(v[m], thr);
          if (p.clamp_in_place && own) col[(size_t)m * p.T] = v[m];
          mx = fmaxf(mx, fabsf(v[m]));
        }
      }
      int e = (int)((__float_as_uint(mx) >> 23) & 0xffu) - 127;
      if (mx < 1e-30f) e = 5;
      e = max(-100, min(100, e));
      const float s = __uint_as_float((uint32_t)(5 - e + 127) << 23);  // mx s in [32, 64)
      inv_s = __uint_as_float((uint32_t)(e - 5 + 127) << 23);
#pragma unroll
      for (int kc = 0; kc < KS; ++kc) {
        uint32_t hi[8], lo[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float a = v[16 * kc + 2 * j] * s, b = v[16 * kc + 2 * j + 1] * s;
          const __half2 h2 = __floats2half2_rn(a, b);
          const float2 f2 = __half22float2(h2);
          const __half2 l2 = __floats2half2_rn((a - f2.x) * 2048.0f, (b - f2.y) * 2048.0f);
          hi[j] = *reinterpret_cast<const uint32_t*>(&h2);
          lo[j] = *reinterpret_cast<const uint32_t*>(&l2);
        }
        dt_tmem_st8(ah_tm + lane_base + 8 * kc, hi);
        dt_tmem_st8(al_tm + lane_base + 8 * kc, lo);
      }
      asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (tid == 0) {
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll
      for (int s = 0; s < KS; ++s) dt_mma_ts(d_tm, ah_tm + 8 * s, b_desc + 16 * s, idesc32, s > 0 ? 1u : 0u);
#pragma unroll
      for (int s = 0; s < KS; ++s) dt_mma_ts(d_tm + 16, al_tm + 8 * s, b_desc + 16 * s, idesc16, 1u);
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar_a) : "memory");
    }
    dt_mbar_wait(bar_a, parity);
    parity ^= 1u;
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    // ---- cepstral coefficients of this frame
    float c[kDtN];
    {
      float d0[16], d1[16];
      dt_tmem_ld16(d_tm + lane_base, d0);
      dt_tmem_ld16(d_tm + lane_base + 16, d1);
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
      for (int j = 0; j < kDtN; ++j) c[j] = fmaf(d1[j], 1.0f / 2048.0f, d0[j]) * inv_s;
    }
    if (own) {
      float* dst = p.mfcc + (size_t)clip * p.n_mfcc * p.T + t;
#pragma unroll
      for (int j = 0; j < kDtN; ++j)
        if (j < p.n_mfcc) dst[(size_t)j * p.T] = c[j];
    }
    if (p.delta != nullptr) {
#pragma unroll
      for (int j = 0; j < kDtN; ++j) s_col[j][tid] = c[j];
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();  // the accumulators have been read; neighbours' coefficients are visible
    if (p.delta != nullptr && own) {
      float* dst = p.delta + (size_t)clip * p.n_mfcc * p.T + t;
#pragma unroll
      for (int j = 0; j < kDtN; ++j) {
        if (j < p.n_mfcc) {
          const float* cc = &s_col[j][tid];
          float d;
          if (p.T == 1) d = 0.0f;
          else if (t == 0) d = cc[1] - cc[0];
          else if (t == p.T - 1) d = cc[0] - cc[-1];
          else d = (cc[1] - cc[-1]) / 2.0f;
          dst[(size_t)j * p.T] = d;
        }
      }
    }
    if (p.delta != nullptr) __syncthreads();  // s_col is rewritten by the next tile
  }

  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0)
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kCols) : "memory");
}

// ---------------------------------------------------------------------------------------------------
// host
// ---------------------------------------------------------------------------------------------------
bool mfcc_tc_supported(int n_mfcc, int n_mels) { return n_mfcc >= 1 && n_mfcc <= kDtN && n_mels >= 1 && n_mels <= 96; }

// dct: [n_mfcc][n_mels] row-major.  Table: 32 rows x kp halves canonical K-major, rows 0..15 = Bhi (zero past
// n_mfcc), rows 16..31 = Blo * 2048
void mfcc_tc_table(const float* dct, int n_mfcc, int n_mels, std::vector<uint16_t>& tab, int* kp_out) {
  const int kp = (n_mels + 15) / 16 * 16;
  tab.assign((size_t)32 * kp, 0);
  auto bits = [](__half h) {
    uint16_t u;
    std::memcpy(&u, &h, 2);
    return u;
  };
  auto idx = [&](int n, int k) { return (size_t)(n / 8) * (kp / 8) * 64 + (size_t)(k / 8) * 64 + (n % 8) * 8 + (k % 8); };
  for (int n = 0; n < n_mfcc; ++n)
    for (int k = 0; k < n_mels; ++k) {
      const float v = dct[(size_t)n * n_mels + k];
      const __half hi = __float2half_rn(v);
      const __half lo = __float2half_rn((v - __half2float(hi)) * 2048.0f);
      tab[idx(n, k)] = bits(hi);
      tab[idx(16 + n, k)] = bits(lo);
    }
  *kp_out = kp;
}

cudaError_t mfcc_tc_launch(const void* btab, int kp, float* logmel, const int* clipmax, long n_clips, long T, int n_mels,
                           int n_mfcc, float top_db, float* mfcc, float* delta, int clamp_in_place, int sm_count,
                           cudaStream_t st) {
  MfccTcArgs a{};
  a.logmel = logmel;
  a.clipmax = clipmax;
  a.T = T;
  a.n_mels = n_mels;
  a.n_mfcc = n_mfcc;
  a.top_db = top_db;
  a.halo = delta != nullptr ? 1 : 0;
  const int per_tile = kDtThreads - 2 * a.halo;
  a.tiles_per_clip = (int)((T + per_tile - 1) / per_tile);
  a.n_tiles = (long)a.tiles_per_clip * n_clips;
  a.btab = reinterpret_cast<const __half*>(btab);
  a.mfcc = mfcc;
  a.delta = delta;
  a.clamp_in_place = clamp_in_place;
  // four persistent CTAs per SM: 128 of the 512 TMEM columns each
  const unsigned grid = (unsigned)std::min<long>(a.n_tiles, 4L * sm_count);
  switch (kp / 16) {
#define MMF_DT_CASE(KS) \
  case KS: mfcc_tc_kernel<KS><<<grid, kDtThreads, 0, st>>>(a); break;
    MMF_DT_CASE(1)
    MMF_DT_CASE(2)
    MMF_DT_CASE(3)
    MMF_DT_CASE(4)
    MMF_DT_CASE(5)
    MMF_DT_CASE(6)
#undef MMF_DT_CASE
    default: return cudaErrorNotSupported;
  }
  count_launch();
  return cudaGetLastError();
}

}  // namespace mmf
