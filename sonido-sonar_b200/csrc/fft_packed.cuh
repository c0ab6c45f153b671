// Packed complex FP32 arithmetic for sm_100a: a complex number (re, im) lives in one 64-bit register pair and is
// processed by the Blackwell packed-FP32 instructions FADD2 / FMUL2 / FFMA2 (PTX add/mul/fma.rn.f32x2, CUDA
// __fadd2_rn / __fmul2_rn / __ffma2_rn).  Their operand modifiers (swap the halves, negate one half, broadcast a
// scalar register) make
//     complex add / subtract      1 instruction   (2 scalar)
//     multiply by -i              0 instructions  (folded into the consumer's operand modifier)
//     complex multiply            2 instructions  (4 scalar: FMUL2 (im, re) x (-wi, wi), FFMA2 x (wr, wr))
// so a radix-32 transform in registers is 160 FADD2 + 34 FMUL2 + 34 FFMA2 = 228 issue slots instead of ~510 scalar
// ones (scripts/microbench: the packed instructions run at the same FLOP rate as the scalar ones, they halve the
// ISSUE slots, which is what bounds the STFT and YIN kernels).
#pragma once
#include <cuda_runtime.h>

#include "fft_regs.cuh"

namespace sonar {

// XOR-swizzled shared-memory rows read 16 bytes at a time by lanes that each own a run of consecutive elements, and
// written 4 / 8 bytes at a time by lanes that own consecutive elements (both conflict free, no padding):
// table rows (one float per bin): 16-byte chunk c = k >> 2 lives at c ^ ((c >> 3) & 7)
__host__ __device__ __forceinline__ int spos(int k) { return ((((k >> 2) ^ ((k >> 5) & 7))) << 2) | (k & 3); }
// magnitude pair rows (one float2 per bin): 16-byte chunk c = k >> 1 lives at c ^ ((c >> 3) & 7).  Conflict free both
// for pass 2's 8-byte stores (32 consecutive bins per instruction) and for the scan's 16-byte loads (BPL contiguous bins
// per lane).
__host__ __device__ __forceinline__ int ppos(int k) { return ((((k >> 1) ^ ((k >> 4) & 7))) << 1) | (k & 1); }

namespace pk {

__device__ __forceinline__ float2 add(float2 a, float2 b) { return __fadd2_rn(a, b); }
__device__ __forceinline__ float2 sub(float2 a, float2 b) { return __fadd2_rn(a, make_float2(-b.x, -b.y)); }
__device__ __forceinline__ float2 mul(float2 a, float2 w) {  // complex product a w
  return __ffma2_rn(a, make_float2(w.x, w.x), __fmul2_rn(make_float2(a.y, a.x), make_float2(-w.y, w.y)));
}
__device__ __forceinline__ float2 mul_conj(float2 a, float2 w) {  // a conj(w)
  return __ffma2_rn(a, make_float2(w.x, w.x), __fmul2_rn(make_float2(a.y, a.x), make_float2(w.y, -w.y)));
}
__device__ __forceinline__ float2 mul_mi(float2 a) { return make_float2(a.y, -a.x); }  // a (-i)
__device__ __forceinline__ float2 scale(float2 a, float s) { return __fmul2_rn(a, make_float2(s, s)); }
__device__ __forceinline__ float2 fma(float2 a, float s, float2 c) { return __ffma2_rn(a, make_float2(s, s), c); }

template <int R, int K>
__device__ __forceinline__ float2 mul_tw(float2 a) {  // a exp(-2 pi i K / R), K in [0, R/2)
  if constexpr (K == 0) {
    return a;
  } else if constexpr (4 * K == R) {
    return mul_mi(a);
  } else {
    static_assert(64 % R == 0, "radix must divide 64");
    return mul(a, w64(K * (64 / R)));
  }
}

// Forward DFT of R points held in registers (radix-2 decimation in time, compile-time recursion), natural order in
// and out.
template <int R>
struct Fft {
  template <int K>
  static __device__ __forceinline__ void combine(float2 (&v)[R], const float2 (&e)[R / 2], const float2 (&o)[R / 2]) {
    if constexpr (K < R / 2) {
      const float2 t = mul_tw<R, K>(o[K]);
      v[K] = add(e[K], t);
      v[K + R / 2] = sub(e[K], t);
      combine<K + 1>(v, e, o);
    }
  }
  static __device__ __forceinline__ void run(float2 (&v)[R]) {
    float2 e[R / 2], o[R / 2];
#pragma unroll
    for (int i = 0; i < R / 2; ++i) {
      e[i] = v[2 * i];
      o[i] = v[2 * i + 1];
    }
    Fft<R / 2>::run(e);
    Fft<R / 2>::run(o);
    combine<0>(v, e, o);
  }
};
template <>
struct Fft<2> {
  static __device__ __forceinline__ void run(float2 (&v)[2]) {
    const float2 a = v[0], b = v[1];
    v[0] = add(a, b);
    v[1] = sub(a, b);
  }
};
template <>
struct Fft<1> {
  static __device__ __forceinline__ void run(float2 (&)[1]) {}
};

}  // namespace pk
}  // namespace sonar
