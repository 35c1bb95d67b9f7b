// Micro-benchmark: issue rate of legacy mma.sync (TF32 m16n8k8, BF16 m16n8k16) on sm_100a.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mma_sync mma_sync.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

template <int MODE, int CH>
__global__ void k(float* out, int iters) {
  float d[CH][4];
  uint32_t a[4] = {0x3f800000u + threadIdx.x, 0x3f900000u, 0x3fa00000u, 0x3fb00000u};
  uint32_t b0 = 0x3f800000u, b1 = 0x3f000000u;
  for (int c = 0; c < CH; ++c)
    for (int e = 0; e < 4; ++e) d[c][e] = (float)(c + e);
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int c = 0; c < CH; ++c) {
      if (MODE == 0)
        asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                     : "+f"(d[c][0]), "+f"(d[c][1]), "+f"(d[c][2]), "+f"(d[c][3])
                     : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
      else
        asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                     : "+f"(d[c][0]), "+f"(d[c][1]), "+f"(d[c][2]), "+f"(d[c][3])
                     : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
    }
  }
  float r = 0;
  for (int c = 0; c < CH; ++c)
    for (int e = 0; e < 4; ++e) r += d[c][e];
  out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}

template <int MODE, int CH>
void run(const char* name, double flop_per_mma) {
  float* out;
  cudaMalloc(&out, 148 * 1024 * 4);
  const int iters = 2048;
  for (int warps : {4, 8, 16, 32}) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    k<MODE, CH><<<148, warps * 32>>>(out, 8);
    cudaEventRecord(e0);
    k<MODE, CH><<<148, warps * 32>>>(out, iters);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    const double mmas_per_sm = (double)iters * CH * warps;
    const double clk = ms * 1e-3 * 1.965e9;
    printf("%-22s chains %d warps/SM %2d: %.3f ms  %.3f mma/clk/SM  %.1f TFLOP/s chip\n", name, CH, warps, ms,
           mmas_per_sm / clk, mmas_per_sm * 148 * flop_per_mma / (ms * 1e-3) / 1e12);
  }
  cudaFree(out);
}

int main() {
  run<0, 1>("tf32 m16n8k8 (dep.)", 2.0 * 16 * 8 * 8);
  run<0, 8>("tf32 m16n8k8", 2.0 * 16 * 8 * 8);
  run<1, 8>("bf16 m16n8k16", 2.0 * 16 * 8 * 16);
  return 0;
}
