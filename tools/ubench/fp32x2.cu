// Micro-benchmark: issue rate of scalar vs packed (f32x2) FP32 instructions on sm_100a.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp32x2 fp32x2.cu
#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned long long u64;
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c){ u64 d; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
__device__ __forceinline__ u64 add2(u64 a, u64 b){ u64 d; asm volatile("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ u64 mul2(u64 a, u64 b){ u64 d; asm volatile("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ float fma1(float a, float b, float c){ float d; asm volatile("fma.rn.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c)); return d; }
__device__ __forceinline__ float add1(float a, float b){ float d; asm volatile("add.rn.f32 %0, %1, %2;" : "=f"(d) : "f"(a), "f"(b)); return d; }

constexpr int CH = 8;   // independent chains per thread
template <int MODE>
__global__ void k(float* out, int iters, float s){
  float a[CH]; u64 p[CH];
  for (int i=0;i<CH;++i){ a[i] = s + i + threadIdx.x; p[i] = ((u64)__float_as_uint(a[i])<<32) | __float_as_uint(a[i]*0.5f); }
  const float b = s*0.999f, c = s*0.001f;
  const u64 pb = ((u64)__float_as_uint(b)<<32)|__float_as_uint(b), pc = ((u64)__float_as_uint(c)<<32)|__float_as_uint(c);
  for (int it=0; it<iters; ++it){
#pragma unroll
    for (int u=0;u<4;++u){
#pragma unroll
      for (int i=0;i<CH;++i){
        if (MODE==0) a[i] = fma1(a[i], b, c);
        if (MODE==1) a[i] = add1(a[i], c);
        if (MODE==2) p[i] = fma2(p[i], pb, pc);
        if (MODE==3) p[i] = add2(p[i], pc);
        if (MODE==4) p[i] = mul2(p[i], pb);
        if (MODE==5) { a[i] = fma1(a[i], b, c); p[i] = add2(p[i], pc); }   // mix scalar FFMA + packed FADD2
        if (MODE==6) { a[i] = add1(a[i], c); p[i] = fma2(p[i], pb, pc); }
      }
    }
  }
  float r = 0; for (int i=0;i<CH;++i) r += a[i] + __uint_as_float((unsigned)p[i]) + __uint_as_float((unsigned)(p[i]>>32));
  out[blockIdx.x*blockDim.x+threadIdx.x] = r;
}
template <int MODE> void run(const char* name, int per_iter_instr, int flops_per_instr_thread){
  float* out; cudaMalloc(&out, 148*8*1024*4);
  const int iters = 4096;
  for (int warps : {4, 8, 16, 32}) {
    const int threads = warps*32;  // one CTA per SM
    cudaEvent_t e0,e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<MODE><<<148, threads>>>(out, 16, 1.0f);
    cudaEventRecord(e0);
    k<MODE><<<148, threads>>>(out, iters, 1.0f);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double winstr = (double)iters*4*CH*per_iter_instr*warps;   // warp instrs per SM
    double ghz = 1.965;
    printf("%-22s warps/SM %2d  %.3f ms  warp-instr/clk/SM (at 1.965GHz) %.2f  flop/clk/SM %.1f\n", name, warps, ms,
           winstr/(ms*1e-3*ghz*1e9), winstr*32*flops_per_instr_thread/per_iter_instr/(ms*1e-3*ghz*1e9));
  }
  cudaFree(out);
}
int main(){
  run<0>("FFMA", 1, 2);
  run<1>("FADD", 1, 1);
  run<2>("FFMA2", 1, 4);
  run<3>("FADD2", 1, 2);
  run<4>("FMUL2", 1, 2);
  run<5>("FFMA+FADD2", 2, 4);
  run<6>("FADD+FFMA2", 2, 5);
  return 0;
}
