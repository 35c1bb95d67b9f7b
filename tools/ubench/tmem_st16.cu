// Known-answer micro-test for the tcgen05 mel projection of K1 (csrc/stft_mel_tc.cu):
//   (1) tcgen05.st.16x32bx2 -- a warp writes 16 TMEM lanes (threads 0-15 -> column c, threads 16-31 -> column
//       c + imm) -- at lane offsets 0 AND 16 inside the warp's 32-lane quarter, read back with 32x32b;
//   (2) tcgen05.mma kind::f16 with BF16 operands, A from tensor memory (two bf16 per 32-bit column, low half =
//       even k), B from shared memory (canonical no-swizzle K-major), D fp32 in TMEM, checked against the host.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tmem_st16 tmem_st16.cu
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include <cuda_bf16.h>
#include <cuda_runtime.h>

constexpr int kN = 32, kK = 32;  // D[128 x 32] = A[128 x 32] . B[32 x 32]^T

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
  uint64_t d = (uint64_t)((saddr & 0x3FFFFu) >> 4);
  d |= (uint64_t)((lbo >> 4) & 0x3FFFu) << 16;
  d |= (uint64_t)((sbo >> 4) & 0x3FFFu) << 32;
  d |= 1ull << 46;
  return d;
}

__global__ void __launch_bounds__(128) k(const uint32_t* a_packed /* [128][kK/2] */, const uint16_t* b_canon,
                                         uint32_t* rb /* [128][2] read-back of test 1 */, float* d_out /* [128][kN] */,
                                         int* flags) {
  __shared__ __align__(1024) uint16_t sB[kN * kK];
  __shared__ __align__(8) unsigned long long bar;
  __shared__ uint32_t tmem_base;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  for (int i = tid; i < kN * kK; i += 128) sB[i] = b_canon[i];
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base)), "r"(128) : "memory");
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
  const uint32_t tb = tmem_base;
  const uint32_t q_base = tb + ((uint32_t)(warp * 32) << 16);

  // ---- test 1: fill columns 64, 65 with 1000 + lane (32x32b), then overwrite lanes 16..31 and lanes 0..15
  // separately with the 16-lane shape
  {
    const uint32_t v0 = 1000u + tid, v1 = 5000u + tid;
    asm volatile("tcgen05.st.sync.aligned.32x32b.x2.b32 [%0], {%1, %2};" ::"r"(q_base + 64), "r"(v0), "r"(v1) : "memory");
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    // upper 16 lanes of the quarter: thread t < 16 -> lane 16 + t, column 64; t >= 16 -> lane t, column 65
    const uint32_t w = 20000u + tid;
    asm volatile("tcgen05.st.sync.aligned.16x32bx2.x1.b32 [%0], 1, {%1};" ::"r"(q_base + (16u << 16) + 64), "r"(w) : "memory");
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    uint32_t r0, r1;
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x2.b32 {%0, %1}, [%2];" : "=r"(r0), "=r"(r1) : "r"(q_base + 64) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    rb[tid * 2] = r0;
    rb[tid * 2 + 1] = r1;
  }
  __syncthreads();

  // ---- test 2: A (bf16 pairs) into columns 32 .. 32 + kK/2 through the 16-lane shape, both halves of the quarter
  {
    // half-warp h2 = lane / 16 handles columns [8 h2, 8 h2 + 8) of the 16 A columns; lanes in two rounds
#pragma unroll
    for (int half = 0; half < 2; ++half) {
      const int row = warp * 32 + half * 16 + (lane & 15);
      const int c0 = (lane >> 4) * 8;
      uint32_t v[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] = a_packed[row * (kK / 2) + c0 + j];
      asm volatile("tcgen05.st.sync.aligned.16x32bx2.x8.b32 [%0], 8, {%1,%2,%3,%4,%5,%6,%7,%8};" ::"r"(
                       q_base + ((uint32_t)(half * 16) << 16) + 32),
                   "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7])
                   : "memory");
    }
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (tid == 0) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    // bf16 x bf16 -> fp32: c_format F32 (1) bits [4,6), a_format / b_format BF16 (1) bits [7,10) / [10,13)
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(kN >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    const uint32_t sbo = (uint32_t)(kK / 8) * 128;
    const uint64_t bd = make_desc(smem_u32(sB), 128, sbo);
#pragma unroll
    for (int s = 0; s < kK / 16; ++s) {
      const uint32_t acc = s > 0;
      asm volatile(
          "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
          "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, {%5, %5, %5, %5}, p;\n\t}\n" ::"r"(tb),
          "r"(tb + 32 + 8 * s), "l"(bd + 16 * s), "r"(idesc), "r"(acc), "r"(0u)
          : "memory");
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
  }
  {
    uint32_t ok = 0;
    int spin = 0;
    while (!ok && spin < (1 << 22)) {
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
                   : "=r"(ok)
                   : "r"(smem_u32(&bar)), "r"(0u)
                   : "memory");
      ++spin;
    }
    if (!ok && tid == 0) flags[0] = 1;
  }
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  {
    uint32_t u[32];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
        : "=r"(u[0]), "=r"(u[1]), "=r"(u[2]), "=r"(u[3]), "=r"(u[4]), "=r"(u[5]), "=r"(u[6]), "=r"(u[7]), "=r"(u[8]), "=r"(u[9]),
          "=r"(u[10]), "=r"(u[11]), "=r"(u[12]), "=r"(u[13]), "=r"(u[14]), "=r"(u[15]), "=r"(u[16]), "=r"(u[17]), "=r"(u[18]),
          "=r"(u[19]), "=r"(u[20]), "=r"(u[21]), "=r"(u[22]), "=r"(u[23]), "=r"(u[24]), "=r"(u[25]), "=r"(u[26]), "=r"(u[27]),
          "=r"(u[28]), "=r"(u[29]), "=r"(u[30]), "=r"(u[31])
        : "r"(q_base)
        : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    for (int j = 0; j < kN; ++j) d_out[tid * kN + j] = __uint_as_float(u[j]);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tb), "r"(128) : "memory");
}

static uint16_t bf16_bits(float f) {
  uint32_t u;
  std::memcpy(&u, &f, 4);
  u += 0x7fffu + ((u >> 16) & 1u);
  return (uint16_t)(u >> 16);
}
static float bf16_val(uint16_t b) {
  uint32_t u = (uint32_t)b << 16;
  float f;
  std::memcpy(&f, &u, 4);
  return f;
}

int main() {
  std::vector<uint16_t> A(128 * kK), B(kN * kK), Bc(kN * kK);
  srand(1);
  for (auto& x : A) x = bf16_bits((float)(rand() % 2001 - 1000) / 256.0f);
  for (auto& x : B) x = bf16_bits((float)(rand() % 2001 - 1000) / 512.0f);
  // canonical K-major: ((n / 8) * (K / 8) + k / 8) * 64 + (n % 8) * 8 + k % 8
  for (int n = 0; n < kN; ++n)
    for (int kk = 0; kk < kK; ++kk) Bc[((n / 8) * (kK / 8) + kk / 8) * 64 + (n % 8) * 8 + kk % 8] = B[n * kK + kk];
  std::vector<uint32_t> Ap(128 * kK / 2);
  for (int r = 0; r < 128; ++r)
    for (int j = 0; j < kK / 2; ++j) Ap[r * (kK / 2) + j] = (uint32_t)A[r * kK + 2 * j] | ((uint32_t)A[r * kK + 2 * j + 1] << 16);
  uint32_t *dA, *dRb;
  uint16_t* dB;
  float* dD;
  int* dF;
  cudaMalloc(&dA, Ap.size() * 4);
  cudaMalloc(&dB, Bc.size() * 2);
  cudaMalloc(&dRb, 128 * 2 * 4);
  cudaMalloc(&dD, 128 * kN * 4);
  cudaMalloc(&dF, 4);
  cudaMemset(dF, 0, 4);
  cudaMemcpy(dA, Ap.data(), Ap.size() * 4, cudaMemcpyHostToDevice);
  cudaMemcpy(dB, Bc.data(), Bc.size() * 2, cudaMemcpyHostToDevice);
  k<<<1, 128>>>(dA, dB, dRb, dD, dF);
  cudaError_t e = cudaDeviceSynchronize();
  printf("kernel: %s\n", cudaGetErrorString(e));
  if (e != cudaSuccess) return 1;
  std::vector<uint32_t> rb(256);
  std::vector<float> D(128 * kN);
  int fl = 0;
  cudaMemcpy(rb.data(), dRb, 1024, cudaMemcpyDeviceToHost);
  cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost);
  cudaMemcpy(&fl, dF, 4, cudaMemcpyDeviceToHost);
  // test 1 expectation: lanes 0..15 of every quarter untouched (1000 + tid, 5000 + tid); lane 16 + t: column 64 =
  // 20000 + (32 warp + t), column 65 = 20000 + (32 warp + 16 + t)
  int bad1 = 0;
  for (int t = 0; t < 128; ++t) {
    const int w = t / 32, l = t % 32;
    uint32_t e0 = l < 16 ? 1000u + t : 20000u + (32 * w + (l - 16));
    uint32_t e1 = l < 16 ? 5000u + t : 20000u + (32 * w + 16 + (l - 16));
    if (rb[2 * t] != e0 || rb[2 * t + 1] != e1) {
      if (bad1 < 8) printf("  test1 lane %d: got (%u, %u) expected (%u, %u)\n", t, rb[2 * t], rb[2 * t + 1], e0, e1);
      ++bad1;
    }
  }
  printf("test1 (16x32bx2 at lane offset 16): %s (%d mismatches)\n", bad1 ? "FAIL" : "ok", bad1);
  double maxerr = 0;
  for (int r = 0; r < 128; ++r)
    for (int n = 0; n < kN; ++n) {
      double acc = 0;
      for (int kk = 0; kk < kK; ++kk) acc += (double)bf16_val(A[r * kK + kk]) * (double)bf16_val(B[n * kK + kk]);
      maxerr = std::fmax(maxerr, std::fabs(acc - (double)D[r * kN + n]));
    }
  printf("test2 (bf16 MMA, A via 16-lane stores): max abs err %.3e %s, mbarrier timeout flag %d\n", maxerr,
         maxerr < 1e-3 ? "ok" : "FAIL", fl);
  return (bad1 || maxerr >= 1e-3 || fl) ? 2 : 0;
}
