// Instruction-delivery microbenchmark: a loop whose body is straight-line code of NI instructions (8 independent FFMA
// chains per thread, no memory traffic), W warps per SM started at different phases.  Reports issue slots per
// scheduler-cycle against the size of the loop body: the fused STFT / pitch kernels have 43-86 KB loop bodies.
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#define R8(x) x x x x x x x x
#define BODY8 "fma.rn.f32 %0, %0, %8, %9;\n\tfma.rn.f32 %1, %1, %8, %9;\n\tfma.rn.f32 %2, %2, %8, %9;\n\tfma.rn.f32 %3, %3, %8, %9;\n\t" \
              "fma.rn.f32 %4, %4, %8, %9;\n\tfma.rn.f32 %5, %5, %8, %9;\n\tfma.rn.f32 %6, %6, %8, %9;\n\tfma.rn.f32 %7, %7, %8, %9;\n\t"
#define ASM8() asm volatile(BODY8 : "+f"(a0), "+f"(a1), "+f"(a2), "+f"(a3), "+f"(a4), "+f"(a5), "+f"(a6), "+f"(a7) : "f"(m), "f"(c));
#define ASM64() R8(ASM8())
#define ASM256() ASM64() ASM64() ASM64() ASM64()  // 4 KB

template <int KB>  // loop body of KB kilobytes = KB * 64 instructions
__global__ void __launch_bounds__(512, 1) k(float* out, int iters, int stagger, unsigned long long* cyc) {
  float a0 = threadIdx.x, a1 = 1, a2 = 2, a3 = 3, a4 = 4, a5 = 5, a6 = 6, a7 = 7;
  const float m = 0.999f, c = 0.001f;
  const int warp = threadIdx.x >> 5;
  // different phases: warp w first runs w * stagger cycles of nothing
  const long long t0 = clock64();
  while (clock64() - t0 < (long long)warp * stagger) {}
  const long long t1 = clock64();
#pragma unroll 1  // the loop body stays KB kilobytes (unrolled, it would be 4 x that)
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int r = 0; r < KB / 4; ++r) { ASM256() }
  }
  const long long t2 = clock64();
  out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
  if ((threadIdx.x & 31) == 0) atomicMax(&cyc[blockIdx.x], (unsigned long long)(t2 - t1));
}

template <int KB>
void run(int warps, int stagger) {
  float* out;
  unsigned long long* cyc;
  cudaMalloc(&out, 148 * 512 * 4);
  cudaMalloc(&cyc, 148 * 8);
  const long long total_inst = 1LL << 22;  // per warp
  const int iters = (int)(total_inst / (KB * 64));
  for (int rep = 0; rep < 2; ++rep) {
    cudaMemset(cyc, 0, 148 * 8);
    k<KB><<<148, warps * 32>>>(out, iters, stagger, cyc);
    cudaDeviceSynchronize();
  }
  unsigned long long h[148];
  cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
  double mx = 0;
  for (int i = 0; i < 148; ++i) mx = h[i] > mx ? h[i] : mx;
  const double ipc = (double)iters * KB * 64 * warps / 4.0 / mx;
  printf("body %3d KB  warps/SM %2d  stagger %5d  -> %.3f issue slots / scheduler-cycle (%s)\n", KB, warps, stagger, ipc,
         cudaGetErrorString(cudaGetLastError()));
  cudaFree(out);
  cudaFree(cyc);
}

int main() {
  for (int warps : {4, 12, 16}) {
    for (int stagger : {0, 3001}) {
      run<8>(warps, stagger);
      run<16>(warps, stagger);
      run<24>(warps, stagger);
      run<28>(warps, stagger);
      run<32>(warps, stagger);
      run<36>(warps, stagger);
      run<40>(warps, stagger);
      run<48>(warps, stagger);
      run<64>(warps, stagger);
      run<96>(warps, stagger);
      run<160>(warps, stagger);
    }
  }
  return 0;
}
