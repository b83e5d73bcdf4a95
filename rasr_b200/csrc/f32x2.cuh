// f32x2.cuh -- packed f32x2 arithmetic (sm_100: FADD2 / FMUL2 / FFMA2 on 64-bit register pairs) and the
// reference-order evaluation of ONE density row of Mm::BatchFloatFeatureScorer (src/Mm/BatchFeatureScorer.cc:207-253),
// shared by the direct kernel, the refinement kernel (gmm.cu) and the fused screening + refinement epilogue
// (gmm_tensor.cu).
#pragma once

#include <stdint.h>

namespace rbf32x2 {

// packed pairs of f32 in one 64-bit register pair (sm_100 FADD2 / FMUL2 / FFMA2)
__device__ __forceinline__ uint64_t pack2(float lo, float hi) {
    uint64_t r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ float lo2(uint64_t v) {
    return __uint_as_float((uint32_t)v);
}
__device__ __forceinline__ float hi2(uint64_t v) {
    return __uint_as_float((uint32_t)(v >> 32));
}
__device__ __forceinline__ uint64_t sub2(uint64_t a, uint64_t b) {
    uint64_t r;
    asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ uint64_t add2(uint64_t a, uint64_t b) {
    uint64_t r;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ uint64_t mul2(uint64_t a, uint64_t b) {
    uint64_t r;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ uint64_t fma2(uint64_t a, uint64_t b, uint64_t c) {
    uint64_t r;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
    return r;
}

// strict (non-contracted) a + d*d per lane.  ptxas 12.9 contracts mul.rn.f32x2 + add.rn.f32x2 into FFMA2 despite the
// explicit rounding modifiers, so the product is formed with scalar FMULs whose .rn is honoured.
__device__ __forceinline__ uint64_t sqadd2(uint64_t d, uint64_t a) {
    const float dl = lo2(d), dh = hi2(d);
    return pack2(__fadd_rn(lo2(a), __fmul_rn(dl, dl)), __fadd_rn(hi2(a), __fmul_rn(dh, dh)));
}

// refinement rows: [ mu' (NB*8) | c | 0 ] = NB*8 + 2 floats = an ODD number (4 NB + 1) of 8-byte words, so that the 16
// lanes of an LDS.64 phase that pick 16 different rows of a mixture hit 16 different bank pairs and the per-lane gather
// runs at the full shared-memory rate (with the 176-byte rows of the direct kernel, rows j and j + 8 collide and the
// gather cost 2.2 wavefronts per phase: 291 us per 100k frames).
__host__ __device__ constexpr int refine_pitch(int nb) {
    return nb * 8 + 2;
}

template<int NB, bool FUSE>
__device__ __forceinline__ float batch_row_score(const float* row, const uint64_t (&x)[NB * 4]) {
    const uint64_t* r    = reinterpret_cast<const uint64_t*>(row);
    uint64_t        a[4] = {r[NB * 4], 0ull, 0ull, 0ull};  // (c, 0): constant first, lane 0 of the first accumulator
#pragma unroll
    for (int b = 0; b < NB; ++b) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const uint64_t d = sub2(r[4 * b + j], x[4 * b + j]);
            a[j]             = FUSE ? fma2(d, d, a[j]) : sqadd2(d, a[j]);
        }
    }
    const uint64_t q = add2(add2(a[0], a[2]), add2(a[1], a[3]));
    return __fadd_rn(lo2(q), hi2(q));
}

}  // namespace rbf32x2
