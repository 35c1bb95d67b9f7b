// Which TMEM lanes does a cta_group::1 M = 64 tcgen05.mma (kind::f16, A from tensor memory) read A from and write D to?
// A is filled so that lane L holds the value (L + 1) in k = 0 and zeros elsewhere; B = one row (n = 0) with B[0][0] = 1:
// D[row][0] = A[row][0].  Every D lane is pre-set to -7 so untouched lanes show.  Run for the two candidate base lanes
// (0 and 16).  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tmem_m64 tmem_m64.cu
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <vector>

#include <cuda_runtime.h>

constexpr int kN = 16, kK = 16;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
  uint64_t d = (uint64_t)((saddr & 0x3FFFFu) >> 4);
  d |= (uint64_t)((lbo >> 4) & 0x3FFFu) << 16;
  d |= (uint64_t)((sbo >> 4) & 0x3FFFu) << 32;
  d |= 1ull << 46;
  return d;
}

__global__ void __launch_bounds__(128) k(int base_lane, float* d_out /* [128] */, int* flags) {
  __shared__ __align__(1024) uint16_t sB[kN * kK];
  __shared__ __align__(8) unsigned long long bar;
  __shared__ uint32_t tmem_base;
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < kN * kK; i += 128) sB[i] = 0;
  __syncthreads();
  if (tid == 0) sB[0] = 0x3F80;  // bf16 1.0 at (n = 0, k = 0) of the canonical layout
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base)), "r"(64) : "memory");
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
  {
    // A: columns 32 .. 39 (K = 16 bf16 = 8 columns): column 32 low half = bf16(lane + 1), rest 0
    uint32_t v[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    const float f = (float)(tid + 1);
    v[0] = __float_as_uint(f) >> 16;  // exact for values <= 256
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"r"(q_base + 32), "r"(v[0]), "r"(v[1]),
                 "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7])
                 : "memory");
    const uint32_t m7 = __float_as_uint(-7.0f);
    asm volatile("tcgen05.st.sync.aligned.32x32b.x1.b32 [%0], {%1};" ::"r"(q_base + 0), "r"(m7) : "memory");
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (tid == 0) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    // bf16 x bf16 -> fp32, M = 64, N = 16
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(kN >> 3) << 17) | ((uint32_t)(64 >> 4) << 24);
    const uint64_t bd = make_desc(smem_u32(sB), 128, (uint32_t)(kK / 8) * 128);
    const uint32_t lane_off = (uint32_t)base_lane << 16;
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, {%5, %5, %5, %5}, p;\n\t}\n" ::"r"(tb + lane_off),
        "r"(tb + lane_off + 32), "l"(bd), "r"(idesc), "r"(0u), "r"(0u)
        : "memory");
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
    uint32_t u;
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1];" : "=r"(u) : "r"(q_base) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    d_out[tid] = __uint_as_float(u);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tb), "r"(64) : "memory");
}

int main() {
  float* dD;
  int* dF;
  cudaMalloc(&dD, 128 * 4);
  cudaMalloc(&dF, 4);
  for (int base : {0, 16}) {
    cudaMemset(dF, 0, 4);
    k<<<1, 128>>>(base, dD, dF);
    cudaError_t e = cudaDeviceSynchronize();
    printf("base lane %d: %s\n", base, cudaGetErrorString(e));
    if (e != cudaSuccess) return 1;
    std::vector<float> D(128);
    int fl = 0;
    cudaMemcpy(D.data(), dD, 512, cudaMemcpyDeviceToHost);
    cudaMemcpy(&fl, dF, 4, cudaMemcpyDeviceToHost);
    printf("  timeout flag %d; D[lane][0] (A lane value + 1 expected where the MMA wrote, -7 elsewhere):\n", fl);
    for (int l = 0; l < 128; ++l) printf("%s%4.0f", (l % 32 == 0) ? "\n   " : " ", D[l]);
    printf("\n");
  }
  return 0;
}
