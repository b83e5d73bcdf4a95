// Microbenchmark: FP32 issue / pipe rates on sm_100a (scalar FFMA vs packed FFMA2/FADD2, with and without LDS.128
// broadcasts).  Prints lane-FMAs per clock per SM.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ uint64_t pk(float a, float b){ uint64_t r; asm("mov.b64 %0, {%1,%2};":"=l"(r):"f"(a),"f"(b)); return r;}
__device__ __forceinline__ uint64_t sub2(uint64_t a, uint64_t b){ uint64_t r; asm volatile("sub.rn.f32x2 %0, %1, %2;":"=l"(r):"l"(a),"l"(b)); return r;}
__device__ __forceinline__ uint64_t fma2(uint64_t a, uint64_t b, uint64_t c){ uint64_t r; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;":"=l"(r):"l"(a),"l"(b),"l"(c)); return r;}
__constant__ ulonglong2 cm[4096];
template<int MODE> __global__ void k(float* out, long long* cyc, int iters, float seed) {
  __shared__ __align__(16) float sm[1024];
  for (int i = threadIdx.x; i < 1024; i += blockDim.x) sm[i] = seed * i;
  __syncthreads();
  float a[16]; uint64_t p[8], x[8];
  for (int i = 0; i < 16; i++) a[i] = seed + i;
  for (int i = 0; i < 8; i++) { p[i] = pk(seed + i, seed - i); x[i] = pk(seed * i, 1.0f + seed); }
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    if (MODE == 0) {  // scalar FFMA, 16 independent chains, 64 per iter
#pragma unroll
      for (int r = 0; r < 4; r++)
#pragma unroll
        for (int i = 0; i < 16; i++) a[i] = __fmaf_rn(a[i], 0.999f, seed);
    } else if (MODE == 1) {  // FFMA2, 8 independent chains, 32 per iter (=64 lane-fma per thread)
#pragma unroll
      for (int r = 0; r < 4; r++)
#pragma unroll
        for (int i = 0; i < 8; i++) p[i] = fma2(p[i], x[i], x[(i + 1) & 7]);
    } else if (MODE == 2) {  // FADD2 -> FFMA2 dependent pairs, 8 chains
#pragma unroll
      for (int r = 0; r < 2; r++)
#pragma unroll
        for (int i = 0; i < 8; i++) { uint64_t d = sub2(x[i], p[(i + 3) & 7]); p[i] = fma2(d, d, p[i]); }
    } else if (MODE == 3) {  // same + LDS.128 broadcast feeding the subs (like the GMM kernel: 1 LDS per 8 packed ops)
#pragma unroll
      for (int r = 0; r < 4; r++) {
        const ulonglong2 m = *reinterpret_cast<const ulonglong2*>(&sm[((it * 4 + r) & 63) * 4]);
        uint64_t d;
        d = sub2(m.x, x[0]); p[0] = fma2(d, d, p[0]);
        d = sub2(m.y, x[1]); p[1] = fma2(d, d, p[1]);
        d = sub2(m.x, x[2]); p[2] = fma2(d, d, p[2]);
        d = sub2(m.y, x[3]); p[3] = fma2(d, d, p[3]);
      }
    } else if (MODE == 5 || MODE == 6) {  // constant-bank fed (LDCU -> uniform register operand), streaming 64 KB
      // MODE 5: 11 x 16 B per 'row' of 40 packed ops for 2 frames (GMM C2 shape); MODE 6: same, all warps in lockstep rows
      const int row = (MODE == 5 ? (it + (threadIdx.x >> 5) * 23) : it) % 372;
#pragma unroll
      for (int r = 0; r < 4; r++) {
        const ulonglong2 m = cm[row * 11 + r];
        uint64_t d;
        d = sub2(m.x, x[0]); p[0] = fma2(d, d, p[0]);
        d = sub2(m.y, x[1]); p[1] = fma2(d, d, p[1]);
        d = sub2(m.x, x[2]); p[2] = fma2(d, d, p[2]);
        d = sub2(m.y, x[3]); p[3] = fma2(d, d, p[3]);
      }
    } else if (MODE == 4) {  // mix: FFMA2 + scalar FFMA interleaved 1:1 (do they share the pipe?)
#pragma unroll
      for (int r = 0; r < 2; r++)
#pragma unroll
        for (int i = 0; i < 8; i++) { p[i] = fma2(p[i], x[i], x[(i + 1) & 7]); a[i] = __fmaf_rn(a[i], 0.999f, seed); a[i+8] = __fmaf_rn(a[i+8], 0.999f, seed);}
    }
  }
  long long t1 = clock64();
  float s = 0; for (int i = 0; i < 16; i++) s += a[i];
  for (int i = 0; i < 8; i++) s += __uint_as_float((uint32_t)p[i]) + __uint_as_float((uint32_t)(p[i] >> 32));
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
template<int MODE> void run(const char* name, int threads, double lane_fma_per_thread_iter) {
  float* out; long long* cyc; cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&cyc, 148 * 8);
  const int iters = 20000;
  k<MODE><<<148, threads>>>(out, cyc, iters, 0.5f); cudaDeviceSynchronize();
  k<MODE><<<148, threads>>>(out, cyc, iters, 0.5f); cudaDeviceSynchronize();
  long long h[148]; cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
  double c = 0; for (int i = 0; i < 148; i++) c += h[i]; c /= 148;
  printf("%-28s threads/SM %4d  lane-fma/clk/SM %.1f  (warp-inst/clk/SMSP %.3f)\n", name, threads,
         lane_fma_per_thread_iter * iters * threads / c, lane_fma_per_thread_iter * iters * threads / c / 128.0);
  cudaFree(out); cudaFree(cyc);
}
int main() {
  for (int th : {256, 512}) {
    run<0>("scalar FFMA", th, 64);
    run<1>("FFMA2", th, 64);
    run<2>("FADD2->FFMA2 pairs", th, 64);   // counts both ops as lane-ops: 16 sub2+16 fma2 = 64 lane ops
    run<3>("LDS.128 + FADD2->FFMA2", th, 64);
    run<5>("LDCU + FADD2->FFMA2 (skewed)", th, 64);
    run<6>("LDCU + FADD2->FFMA2 (lockstep)", th, 64);
  }
  return 0;
}
