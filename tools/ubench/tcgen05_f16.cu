// Micro-benchmark + known-answer test: hand-written tcgen05.mma kind::f16 (fp16 operands, fp32
// accumulate) on sm_100a -- the fp16 x3 alternative to the TF32 x3 split (tests/studies/tf32_dft_study.py).
// Generated from tcgen05_tf32.cu's structure; same descriptors, 8 halves per 16-byte chunk, K = 16 per MMA.
//
//   D[128 x N] (TMEM, fp32) = A[128 x K] . B[N x K]^T,  A and B K-major in shared memory in the
//   no-swizzle ("interleave") canonical layout: 8-row x 16-byte core matrices, 128 contiguous bytes
//   each; LBO = distance between the two 16-byte K-chunks of one MMA (K = 8 tf32), SBO = distance
//   between 8-row groups.
//
// What it answers for the round-2 transform kernel:
//   * are the shared-memory / instruction descriptors right (result checked against the host),
//   * cycles per MMA for M = 128, K = 8 and N = 32 ... 256, one issuing thread per SM,
//   * TMEM -> register read-back rate (tcgen05.ld 32x32b).
//
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tcgen05_tf32 tcgen05_tf32.cu
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include <cuda_fp16.h>
#include <cuda_runtime.h>

constexpr int kM = 128;
constexpr int kK = 32;  // K per tile: two MMAs of K = 16

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);   // start address, 16-byte units
  d |= (uint64_t)((lbo >> 4) & 0x3FFFu) << 16;  // leading byte offset
  d |= (uint64_t)((sbo >> 4) & 0x3FFFu) << 32;  // stride byte offset
  d |= 1ull << 46;                              // descriptor version (Blackwell)
  return d;                                     // base_offset 0, layout_type 0 = no swizzle
}

__device__ __forceinline__ uint32_t make_idesc(int M, int N) {
  // c_format F32 (1) bits [4,6); a_format / b_format TF32 (2) bits [7,10) / [10,13); both K-major;
  // n_dim = N >> 3 bits [17,23); m_dim = M >> 4 bits [24,29)
  return (1u << 4) | (0u << 7) | (0u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

__device__ __forceinline__ void mma_tf32(uint32_t taddr, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, {%5, %5, %5, %5}, p;\n\t"
      "}\n" ::"r"(taddr),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc), "r"(0u)
      : "memory");
}

// A operand from tensor memory (lane = row, one 32-bit column per tf32 element), B from shared memory
__device__ __forceinline__ void mma_tf32_ts(uint32_t taddr, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, {%5, %5, %5, %5}, p;\n\t"
      "}\n" ::"r"(taddr),
      "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(acc), "r"(0u)
      : "memory");
}

__device__ __forceinline__ bool mbar_wait(uint32_t bar, uint32_t parity) {
  for (int spin = 0; spin < (1 << 24); ++spin) {
    uint32_t ok;
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t"
        "}\n"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    if (ok) return true;
  }
  return false;  // bounded: a wrong descriptor must not hang the box
}

// tile element (r, k) -> float index inside the canonical no-swizzle K-major tile
__host__ __device__ inline int tile_index(int r, int k, int K) {
  return (r / 8) * (K / 8) * 64 + (k / 8) * 64 + (r % 8) * 8 + (k % 8);
}

// a_in_tmem is a template parameter: the issuing thread's loop must stay free of branches and of
// descriptor arithmetic (UTCHMMA takes uniform registers; a runtime switch here doubled the issue time)
// two_issuers: lane 0 of warp 1 issues the same stream into a second accumulator -- tells whether the
// ~45-cycle floor per small MMA is the issuing thread or the tensor pipe
// A_MN_LBO > 0: A stored MN-major (M contiguous in 8-element chunks; the 8 K-rows of a core matrix 16 bytes
// apart; K-groups A_MN_LBO bytes apart, 144 = padded against bank aliasing; M-groups SBO = 4 * A_MN_LBO apart)
template <bool a_in_tmem, bool two_issuers = false, int A_MN_LBO = 0>
__global__ void __launch_bounds__(128)
    f16_kernel(const __half* __restrict__ A, const __half* __restrict__ B, float* __restrict__ D, int N, int iters,
                long long* __restrict__ cycles, int* __restrict__ err) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __half* sA = reinterpret_cast<__half*>(smem);  // 128 x 32 halves = 8 KB
  __half* sB = sA + (A_MN_LBO > 128 ? (kM / 8) * (kK / 8) * A_MN_LBO / 2 : kM * kK);  // N x 32 halves
  __shared__ __align__(8) unsigned long long bar;
  __shared__ __align__(8) unsigned long long bar2;
  __shared__ uint32_t tmem_base;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  for (int i = tid; i < kM * kK; i += blockDim.x) {
    const int r = i / kK, k = i % kK;
    if (A_MN_LBO > 0) sA[((r / 8) * (kK / 8) * A_MN_LBO + (k / 8) * A_MN_LBO + (k % 8) * 16 + (r % 8) * 2) / 2] = A[i];
    else sA[tile_index(r, k, kK)] = A[i];
  }
  for (int i = tid; i < N * kK; i += blockDim.x) sB[tile_index(i / kK, i % kK, kK)] = B[i];
  // generic-proxy writes -> visible to the tensor core (async proxy)
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");

  int ncols = 32;
  while (ncols < N + (a_in_tmem ? kK : 0) + (two_issuers ? N : 0)) ncols <<= 1;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base)), "r"(ncols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)) : "memory");
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar2)) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t taddr = tmem_base;
  const uint32_t a_tmem = taddr + (uint32_t)N;  // columns N .. N+K-1 hold A when a_in_tmem
  static_assert(!a_in_tmem, "fp16 variant: A from shared memory only");

  if (tid == 0) {
    const uint32_t lbo = 128, sbo = (kK / 8) * 128;
    const uint64_t adesc0 = A_MN_LBO > 0 ? make_desc(smem_u32(sA), A_MN_LBO, (kK / 8) * A_MN_LBO) : make_desc(smem_u32(sA), lbo, sbo);
    const uint64_t bdesc0 = make_desc(smem_u32(sB), lbo, sbo);
    const uint32_t idesc = make_idesc(kM, N) | (A_MN_LBO > 0 ? (1u << 15) : 0u);  // bit 15: A is MN-major
    constexpr int a_step = A_MN_LBO > 0 ? 2 * A_MN_LBO / 16 : 16;  // descriptor units per K = 16 slab
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int s = 0; s < kK / 16; ++s) {
        // next K = 8 slab: two 16-byte chunks further = 2 * LBO = 256 bytes = 16 descriptor units
        if (a_in_tmem) mma_tf32_ts(taddr, a_tmem + (uint32_t)(s * 8), bdesc0 + (uint64_t)(s * 16), idesc, (it | s) ? 1u : 0u);
        else mma_tf32(taddr, adesc0 + (uint64_t)(s * a_step), bdesc0 + (uint64_t)(s * 16), idesc, (it | s) ? 1u : 0u);
      }
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar))
                 : "memory");
    const bool ok = mbar_wait(smem_u32(&bar), 0);
    const long long t1 = clock64();
    if (!ok) atomicExch(err, 1);
    cycles[blockIdx.x] = t1 - t0;
  }
  if (two_issuers && tid == 32) {
    const uint32_t lbo = 128, sbo = (kK / 8) * 128;
    const uint64_t adesc0 = make_desc(smem_u32(sA), lbo, sbo);
    const uint64_t bdesc0 = make_desc(smem_u32(sB), lbo, sbo);
    const uint32_t idesc = make_idesc(kM, N);
    const uint32_t d2 = taddr + (uint32_t)(N + (a_in_tmem ? kK : 0));
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int s = 0; s < kK / 16; ++s) {
        if (a_in_tmem) mma_tf32_ts(d2, a_tmem + (uint32_t)(s * 8), bdesc0 + (uint64_t)(s * 16), idesc, (it | s) ? 1u : 0u);
        else mma_tf32(d2, adesc0 + (uint64_t)(s * 16), bdesc0 + (uint64_t)(s * 16), idesc, (it | s) ? 1u : 0u);
      }
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar2))
                 : "memory");
    if (!mbar_wait(smem_u32(&bar2), 0)) atomicExch(err, 1);
  }
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");

  // read back: warp w owns TMEM lanes 32w .. 32w+31 = rows of D
  // (a) timing only: all loads of the tile in flight, one wait
  const long long q0 = clock64();
  {
    uint32_t acc = 0;
    for (int c0 = 0; c0 < N; c0 += 32) {
      uint32_t v[32];
      asm volatile(
          "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
          "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
          : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
            "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
            "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
            "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
          : "r"(taddr + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0)
          : "memory");
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
      for (int j = 0; j < 32; ++j) acc ^= v[j];
    }
    if (acc == 0x12345678u) atomicExch(err, 2);  // keep the loads alive
  }
  const long long q1 = clock64();
  if (tid == 0) cycles[2 * gridDim.x + blockIdx.x] = q1 - q0;
  // (b) known-answer read-back, eight columns at a time
  const long long r0 = clock64();
  for (int c0 = 0; c0 < N; c0 += 8) {
    uint32_t v[8];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
                 : "r"(taddr + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0)
                 : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    if (blockIdx.x == 0) {
      const int row = warp * 32 + lane;
#pragma unroll
      for (int j = 0; j < 8; ++j) D[(size_t)row * N + c0 + j] = __uint_as_float(v[j]);
    }
  }
  const long long r1 = clock64();
  if (tid == 0) cycles[gridDim.x + blockIdx.x] = r1 - r0;

  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}

static float to_f16(float x) { return __half2float(__float2half(x)); }

int main() {
  int dev = 0, nsm = 0, khz = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, dev);
  cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, dev);
  std::printf("SMs %d, max clock %.0f MHz\n", nsm, khz / 1000.0);
  std::vector<float> hA(kM * kK), hB(256 * kK);
  srand(7);
  for (auto& v : hA) v = to_f16((float)rand() / RAND_MAX - 0.5f);
  for (auto& v : hB) v = to_f16((float)rand() / RAND_MAX - 0.5f);
  std::vector<__half> hAh(hA.size()), hBh(hB.size());
  for (size_t i = 0; i < hA.size(); ++i) hAh[i] = __float2half(hA[i]);
  for (size_t i = 0; i < hB.size(); ++i) hBh[i] = __float2half(hB[i]);
  __half *dA, *dB;
  float* dD;
  long long* dC;
  int* dErr;
  cudaMalloc(&dA, hA.size() * 2);
  cudaMalloc(&dB, hB.size() * 2);
  cudaMalloc(&dD, kM * 256 * 4);
  cudaMalloc(&dC, 3 * nsm * sizeof(long long));
  cudaMalloc(&dErr, 4);
  cudaMemcpy(dA, hAh.data(), hA.size() * 2, cudaMemcpyHostToDevice);
  cudaMemcpy(dB, hBh.data(), hB.size() * 2, cudaMemcpyHostToDevice);
  cudaMemset(dErr, 0, 4);
  cudaFuncSetAttribute(f16_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  cudaFuncSetAttribute(f16_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  cudaFuncSetAttribute(f16_kernel<false, false, 128>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  cudaFuncSetAttribute(f16_kernel<false, false, 144>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  for (int mode : {0, 2, 3, 4})
  for (int N : {32, 64, 128, 256}) {
    if (mode >= 2 && N > 128) continue;
    const char* tag = mode == 4 ? "A MN-major LBO 144" : mode == 3 ? "A MN-major LBO 128" : mode == 2 ? "A in smem, 2 issuers" : "A in smem";
    const size_t smem = (size_t)(kM + N) * kK * 2 + (mode == 4 ? 2048 : 0);
    // known answer: one pass
    cudaMemset(dD, 0, kM * 256 * 4);
    if (mode == 4) f16_kernel<false, false, 144><<<1, 128, smem>>>(dA, dB, dD, N, 1, dC, dErr);
    else if (mode == 3) f16_kernel<false, false, 128><<<1, 128, smem>>>(dA, dB, dD, N, 1, dC, dErr);
    else if (mode == 2) f16_kernel<false, true><<<1, 128, smem>>>(dA, dB, dD, N, 1, dC, dErr);
    else f16_kernel<false><<<1, 128, smem>>>(dA, dB, dD, N, 1, dC, dErr);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) {
      std::printf("N %d: CUDA error %s\n", N, cudaGetErrorString(e));
      return 1;
    }
    std::vector<float> hD((size_t)kM * N);
    cudaMemcpy(hD.data(), dD, hD.size() * 4, cudaMemcpyDeviceToHost);
    double worst = 0.0;
    for (int i = 0; i < kM; ++i)
      for (int j = 0; j < N; ++j) {
        double ref = 0.0;
        for (int k = 0; k < kK; ++k) ref += (double)hA[i * kK + k] * (double)hB[j * kK + k];
        worst = std::fmax(worst, std::fabs(ref - (double)hD[(size_t)i * N + j]));
      }
    int herr = 0;
    cudaMemcpy(&herr, dErr, 4, cudaMemcpyDeviceToHost);
    std::printf("%s N %3d: known answer max |err| %.3g %s%s\n", tag, N, worst, worst < 1e-5 ? "OK" : "MISMATCH",
                herr == 1 ? " (barrier wait timed out)" : "");
    if (herr) cudaMemset(dErr, 0, 4);
    if (herr || !(worst < 1e-5)) continue;
    // rate: every SM issuing
    const int iters = 4000;
    if (mode == 4) f16_kernel<false, false, 144><<<nsm, 128, smem>>>(dA, dB, dD, N, iters, dC, dErr);
    else if (mode == 3) f16_kernel<false, false, 128><<<nsm, 128, smem>>>(dA, dB, dD, N, iters, dC, dErr);
    else if (mode == 2) f16_kernel<false, true><<<nsm, 128, smem>>>(dA, dB, dD, N, iters, dC, dErr);
    else f16_kernel<false><<<nsm, 128, smem>>>(dA, dB, dD, N, iters, dC, dErr);
    e = cudaDeviceSynchronize();
    if (e != cudaSuccess) {
      std::printf("N %d: CUDA error %s\n", N, cudaGetErrorString(e));
      return 1;
    }
    std::vector<long long> hC(3 * nsm);
    cudaMemcpy(hC.data(), dC, hC.size() * sizeof(long long), cudaMemcpyDeviceToHost);
    double cyc = 0.0, rd = 0.0, rd32 = 0.0;
    for (int i = 0; i < nsm; ++i) {
      cyc += (double)hC[i];
      rd += (double)hC[nsm + i];
      rd32 += (double)hC[2 * nsm + i];
    }
    cyc /= nsm;
    rd /= nsm;
    rd32 /= nsm;
    const double per_mma = cyc / (iters * (kK / 16)) / (mode == 2 ? 2 : 1);
    const double flop_per_mma = 2.0 * kM * N * 16;
    std::printf("%s N %3d: %.1f cycles per MMA (m128 n%d k16, fp16) = %.0f flop/clk/SM = %.0f TFLOP/s at %.0f MHz on %d SMs; "
                "read-back of 128 x %d fp32 by 4 warps: %.0f cycles (x32 loads), %.0f (x8 loads + stores)\n",
                tag, N, per_mma, N, flop_per_mma / per_mma, flop_per_mma / per_mma * nsm * (khz * 1e3) / 1e12, khz / 1000.0,
                nsm, N, rd32, rd);
  }
  return 0;
}
