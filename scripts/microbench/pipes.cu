// Issue-rate microbenchmarks that decide the STFT / YIN kernel design on sm_100a:
// scalar FFMA vs packed FFMA2 / FADD2 / FMUL2 (fma.rn.f32x2 ...), DFMA, and the same mixed with shared-memory loads.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o pipes pipes.cu && ./pipes
#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned long long u64;
__device__ __forceinline__ u64 pk(float a, float b) { u64 r; asm("mov.b64 %0, {%1,%2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ u64 ffma2(u64 a, u64 b, u64 c) { u64 r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }
__device__ __forceinline__ u64 fadd2(u64 a, u64 b) { u64 r; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ u64 fmul2(u64 a, u64 b) { u64 r; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }

template <int MODE>
__global__ void kern(float* out, int iters, float s, const float* tab) {
  extern __shared__ float sm[];
  for (int i = threadIdx.x; i < 4096; i += blockDim.x) sm[i] = tab ? tab[i] : 1.0f;
  __syncthreads();
  float r = 0.f;
  if (MODE == 0) {  // 16 independent FFMA (3 distinct registers each)
    float a[16]; for (int j = 0; j < 16; ++j) a[j] = threadIdx.x + j;
    float c = 0.5f + s;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
      for (int j = 0; j < 16; ++j) a[j] = fmaf(a[j], s, c);
    }
    for (int j = 0; j < 16; ++j) r += a[j];
  } else if (MODE == 1) {  // 8 independent FFMA2 = the same flops
    u64 a[8]; for (int j = 0; j < 8; ++j) a[j] = pk(threadIdx.x + j, j);
    u64 m = pk(s, s), c = pk(0.5f + s, 0.25f);
    for (int i = 0; i < iters; ++i) {
#pragma unroll
      for (int j = 0; j < 8; ++j) a[j] = ffma2(a[j], m, c);
    }
    u64 x = 0; for (int j = 0; j < 8; ++j) x ^= a[j];
    r = __uint_as_float((unsigned)(x ^ (x >> 32)));
  } else if (MODE == 2) {  // 16 independent FFMA2 = twice the flops of MODE 0 in the same instruction count
    u64 a[16]; for (int j = 0; j < 16; ++j) a[j] = pk(threadIdx.x + j, j);
    u64 m = pk(s, s), c = pk(0.5f + s, 0.25f);
    for (int i = 0; i < iters; ++i) {
#pragma unroll
      for (int j = 0; j < 16; ++j) a[j] = ffma2(a[j], m, c);
    }
    u64 x = 0; for (int j = 0; j < 16; ++j) x ^= a[j];
    r = __uint_as_float((unsigned)(x ^ (x >> 32)));
  } else if (MODE == 3) {  // FADD2 + FMUL2 mix, 16 per iteration
    u64 a[16]; for (int j = 0; j < 16; ++j) a[j] = pk(threadIdx.x + j, j);
    u64 m = pk(s, s), c = pk(0.5f + s, 0.25f);
    for (int i = 0; i < iters; ++i) {
#pragma unroll
      for (int j = 0; j < 16; j += 2) { a[j] = fadd2(a[j], c); a[j + 1] = fmul2(a[j + 1], m); }
    }
    u64 x = 0; for (int j = 0; j < 16; ++j) x ^= a[j];
    r = __uint_as_float((unsigned)(x ^ (x >> 32)));
  } else if (MODE == 4) {  // 16 FADD/FMUL scalar mix
    float a[16]; for (int j = 0; j < 16; ++j) a[j] = threadIdx.x + j;
    float c = 0.5f + s;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
      for (int j = 0; j < 16; j += 2) { a[j] = a[j] + c; a[j + 1] = a[j + 1] * s; }
    }
    for (int j = 0; j < 16; ++j) r += a[j];
  } else if (MODE == 5) {  // 16 DFMA
    double a[16]; for (int j = 0; j < 16; ++j) a[j] = threadIdx.x + j;
    double c = 0.5 + s, m = s;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
      for (int j = 0; j < 16; ++j) a[j] = fma(a[j], m, c);
    }
    double x = 0; for (int j = 0; j < 16; ++j) x += a[j];
    r = (float)x;
  } else if (MODE == 6) {  // 8 FFMA + 8 LDS.32 (conflict-free) per iteration
    float a[8]; for (int j = 0; j < 8; ++j) a[j] = threadIdx.x + j;
    int idx = threadIdx.x & 31;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
      for (int j = 0; j < 8; ++j) a[j] = fmaf(a[j], s, sm[(idx + 32 * j + i) & 4095]);
    }
    for (int j = 0; j < 8; ++j) r += a[j];
  } else if (MODE == 7) {  // 8 FFMA2 + 8 LDS.64 per iteration (twice the flops and bytes of MODE 6 per instruction)
    u64 a[8]; for (int j = 0; j < 8; ++j) a[j] = pk(threadIdx.x + j, j);
    u64 m = pk(s, s);
    const u64* sm2 = reinterpret_cast<const u64*>(sm);
    int idx = threadIdx.x & 31;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
      for (int j = 0; j < 8; ++j) a[j] = ffma2(a[j], m, sm2[(idx + 32 * j + i) & 2047]);
    }
    u64 x = 0; for (int j = 0; j < 8; ++j) x ^= a[j];
    r = __uint_as_float((unsigned)(x ^ (x >> 32)));
  } else if (MODE == 8) {  // 8 FFMA2 x2 + 8 LDS.128 per iteration
    u64 a[16]; for (int j = 0; j < 16; ++j) a[j] = pk(threadIdx.x + j, j);
    u64 m = pk(s, s);
    const ulonglong2* sm4 = reinterpret_cast<const ulonglong2*>(sm);
    int idx = threadIdx.x & 31;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
      for (int j = 0; j < 8; ++j) { ulonglong2 t = sm4[(idx + 32 * j + i) & 1023]; a[2 * j] = ffma2(a[2 * j], m, t.x); a[2 * j + 1] = ffma2(a[2 * j + 1], m, t.y); }
    }
    u64 x = 0; for (int j = 0; j < 16; ++j) x ^= a[j];
    r = __uint_as_float((unsigned)(x ^ (x >> 32)));
  } else if (MODE == 9) {  // 8 FFMA + 8 SHFL
    float a[8]; for (int j = 0; j < 8; ++j) a[j] = threadIdx.x + j;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
      for (int j = 0; j < 8; ++j) a[j] = fmaf(__shfl_xor_sync(0xffffffffu, a[j], 1 + (j & 3)), s, a[(j + 1) & 7]);
    }
    for (int j = 0; j < 8; ++j) r += a[j];
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}

template <int MODE>
void run(const char* name, double flop_per_thread_iter, double lsu_per_thread_iter, float* d) {
  int iters = 4000;
  for (int bs : {128, 384, 1024}) {
    int per_sm = 2048 / bs; if (bs == 384) per_sm = 1;  // 384 = the STFT kernel's 12 warps per SM
    int grid = 148 * per_sm;
    cudaFuncSetAttribute(kern<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 16384);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    kern<MODE><<<grid, bs, 16384>>>(d, iters, 0.999f, nullptr);
    cudaEventRecord(e0);
    kern<MODE><<<grid, bs, 16384>>>(d, iters, 0.999f, nullptr);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double thr = (double)grid * bs * iters;
    printf("%-28s warps/SM=%2d  %.3f ms  %7.2f TFLOP/s  %6.3f inst/clk/SMSP(@1.9GHz, fp+lsu)\n", name, per_sm * bs / 32, ms,
           thr * flop_per_thread_iter / ms * 1e-9,
           (thr / 32) * (flop_per_thread_iter > 0 ? 1 : 0) / (ms * 1e-3) / (148.0 * 4 * 1.9e9));
    (void)lsu_per_thread_iter;
  }
}
int main() {
  float* d; cudaMalloc(&d, 148 * 16 * 1024 * 4);
  run<0>("16 FFMA", 32, 0, d);
  run<1>("8 FFMA2", 32, 0, d);
  run<2>("16 FFMA2", 64, 0, d);
  run<3>("8 FADD2 + 8 FMUL2", 32, 0, d);
  run<4>("8 FADD + 8 FMUL", 16, 0, d);
  run<5>("16 DFMA", 32, 0, d);
  run<6>("8 FFMA + 8 LDS.32", 16, 8, d);
  run<7>("8 FFMA2 + 8 LDS.64", 32, 8, d);
  run<8>("16 FFMA2 + 8 LDS.128", 64, 8, d);
  run<9>("8 FFMA + 8 SHFL", 16, 8, d);
  cudaError_t e = cudaDeviceSynchronize();
  printf("status: %s\n", cudaGetErrorString(e));
  return e != cudaSuccess;
}
