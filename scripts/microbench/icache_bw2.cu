// Finer instruction-delivery sweep (see icache_bw.cu): loop bodies of 4..20 KB, and FOUR DISTINCT bodies selected by
// warp % 4 (one role per scheduler) or by warp / 4 (all roles on every scheduler).
#include <cstdio>
#include <cuda_runtime.h>
#define R2(x) x x
#define R4(x) x x x x
#define R8(x) x x x x x x x x
#define BODY8 "fma.rn.f32 %0, %0, %8, %9;\n\tfma.rn.f32 %1, %1, %8, %9;\n\tfma.rn.f32 %2, %2, %8, %9;\n\tfma.rn.f32 %3, %3, %8, %9;\n\t" \
              "fma.rn.f32 %4, %4, %8, %9;\n\tfma.rn.f32 %5, %5, %8, %9;\n\tfma.rn.f32 %6, %6, %8, %9;\n\tfma.rn.f32 %7, %7, %8, %9;\n\t"
#define ASM8() asm volatile(BODY8 : "+f"(a0), "+f"(a1), "+f"(a2), "+f"(a3), "+f"(a4), "+f"(a5), "+f"(a6), "+f"(a7) : "f"(m), "f"(c));
#define ASM64() R8(ASM8())   // 1 KB

// MODE 0: one body for every warp; 1: four copies, selected by warp % 4; 2: four copies, selected by warp / (warps/4)
template <int KB, int MODE>
__global__ void __launch_bounds__(512, 1) k(float* out, int iters, int stagger, unsigned long long* cyc) {
  float a0 = threadIdx.x, a1 = 1, a2 = 2, a3 = 3, a4 = 4, a5 = 5, a6 = 6, a7 = 7;
  const float m = 0.999f, c = 0.001f;
  const int warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  const int role = MODE == 0 ? 0 : (MODE == 1 ? (warp & 3) : warp / (nw / 4));
  const long long t0 = clock64();
  while (clock64() - t0 < (long long)warp * stagger) {}
  const long long t1 = clock64();
#define LOOP() for (int it = 0; it < iters; ++it) { _Pragma("unroll") for (int r = 0; r < KB; ++r) { ASM64() } }
  if (role == 0) { LOOP() }
  else if (role == 1) { LOOP() a0 += 1.f; }
  else if (role == 2) { LOOP() a1 += 1.f; }
  else { LOOP() a2 += 1.f; }
  const long long t2 = clock64();
  out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
  if ((threadIdx.x & 31) == 0) atomicMax(&cyc[blockIdx.x], (unsigned long long)(t2 - t1));
}
template <int KB, int MODE>
void run(int warps, int stagger) {
  float* out; unsigned long long* cyc;
  cudaMalloc(&out, 148 * 512 * 4); cudaMalloc(&cyc, 148 * 8);
  const int iters = (int)((1LL << 21) / (KB * 64));
  for (int rep = 0; rep < 2; ++rep) { cudaMemset(cyc, 0, 148 * 8); k<KB, MODE><<<148, warps * 32>>>(out, iters, stagger, cyc); cudaDeviceSynchronize(); }
  unsigned long long h[148]; cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
  double mx = 0; for (int i = 0; i < 148; ++i) mx = h[i] > mx ? h[i] : mx;
  printf("body %2d KB mode %d warps/SM %2d stagger %5d -> %.3f issue slots / scheduler-cycle (%s)\n", KB, MODE, warps, stagger,
         (double)iters * KB * 64 * warps / 4.0 / mx, cudaGetErrorString(cudaGetLastError()));
  cudaFree(out); cudaFree(cyc);
}
template <int MODE> void sweep(int warps, int stagger) {
  run<4, MODE>(warps, stagger); run<6, MODE>(warps, stagger); run<8, MODE>(warps, stagger); run<10, MODE>(warps, stagger);
  run<12, MODE>(warps, stagger); run<14, MODE>(warps, stagger); run<16, MODE>(warps, stagger); run<20, MODE>(warps, stagger);
}
int main() {
  for (int warps : {12, 16}) for (int stagger : {0, 3001}) { sweep<0>(warps, stagger); sweep<1>(warps, stagger); sweep<2>(warps, stagger); }
  return 0;
}
