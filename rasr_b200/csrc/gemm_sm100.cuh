// gemm_sm100.cuh -- hand-written tcgen05 GEMM for sm_100a:  D[M x N] = epilogue(A[M x K] * B[N x K]^T)
//
// A and B are 16-bit (bf16 or fp16) row-major with K contiguous ("K-major"), K a multiple of 64.
//   * operands are staged by TMA (cp.async.bulk.tensor.2d, SWIZZLE_128B) into a 4-stage ring,
//   * one elected thread issues tcgen05.mma.cta_group::1.kind::f16 (UMMA 128x256x16), accumulators
//     live in TMEM (2 x 256 columns, double buffered so the epilogue of tile i overlaps the MMAs of i+1),
//   * eight epilogue warps (two per TMEM lane quarter, 128 accumulator columns each) read the accumulator with
//     tcgen05.ld (32 lanes x 32 columns per call) and hand 32 consecutive columns of one row to the epilogue functor
//     (bias / activation ...).  With four warps -- one per scheduler, nothing to hide its latencies behind -- the
//     epilogue of a 128 x 256 tile took ~7.7 us: longer than the MMAs of a K = 448 tile (2.7 us), so the Nn input
//     layer ran at 561 TFLOP/s, and the f32 score layer's epilogue stuck out from under its MMAs too,
//   * persistent grid (one CTA per SM), tiles ordered n-fastest so the A tile of a frame block is
//     re-read from L2, never from HBM.
// Warp roles: 0 = TMA producer, 1 = MMA issuer, 2 = TMEM allocator, 3 = idle, 4..11 = epilogue.
#pragma once

#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include "common.cuh"

namespace rbgemm {

using namespace rbdev;

constexpr int BM = 128, BN = 256, BK = 64, STAGES = 4;
constexpr int A_BYTES     = BM * BK * 2;
constexpr int B_BYTES     = BN * BK * 2;
constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
constexpr int THREADS     = 384;  // 4 control warps + 8 epilogue warps
constexpr int TMEM_COLS   = 512;
constexpr int SMEM_BYTES  = STAGES * STAGE_BYTES + 256 + 1024;  // ring + barriers + alignment slack
constexpr int OUT_STAGE_BYTES = 4 * 2 * 32 * 32 * 4;             // TMA-store epilogue: 4 warps x 2 boxes of 32 x 32 f32
constexpr int SMEM_BYTES_TMA  = STAGES * STAGE_BYTES + 1024 + OUT_STAGE_BYTES + 1024;

enum { FMT_F16 = 0, FMT_BF16 = 1 };

// instruction descriptor (cute::UMMA::InstrDescriptor): f32 accumulate, K-major A and B
__host__ __device__ constexpr uint32_t instr_desc(int fmt, int tileN = BN) {
    return (1u << 4) | ((uint32_t)fmt << 7) | ((uint32_t)fmt << 10) | ((uint32_t)(tileN >> 3) << 17) |
           ((uint32_t)(BM >> 4) << 24);
}
// shared memory of the register-epilogue kernel with TN-column tiles (TN = BN, or 128 for batches whose 128 x 256
// tiles would leave most SMs idle)
__host__ __device__ constexpr int smem_bytes(int tileN) {
    return STAGES * (A_BYTES + tileN * BK * 2) + 256 + 1024;
}

// shared-memory matrix descriptor: K-major, SWIZZLE_128B, 8-row groups 1024 bytes apart
__device__ __forceinline__ uint64_t smem_desc(uint32_t saddr) {
    return (uint64_t)((saddr & 0x3FFFFu) >> 4) | (1ull << 16) | (64ull << 32) | (1ull << 46) | (2ull << 61);
}

__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, int c0, int c1, uint64_t* bar) {
    asm volatile(
            "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::
                    "r"(dst),
            "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
            : "memory");
}
// shared -> global tile store through the TMA unit (SASS: UTMASTG); completion is tracked by bulk groups
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* map, uint32_t src, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(map), "r"(src),
                 "r"(c0), "r"(c1)
                 : "memory");
}
__device__ __forceinline__ void bulk_commit() {
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
template<int N>
__device__ __forceinline__ void bulk_wait_read() {  // at most N bulk groups of this thread still read shared memory
    asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* map) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(map) : "memory");
}
__device__ __forceinline__ void tc_fence_before() {
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_after() {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void tc_mma(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                       uint32_t accumulate) {
    asm volatile(
            "{\n\t"
            ".reg .pred p;\n\t"
            "setp.ne.b32 p, %4, 0;\n\t"
            "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
            "}\n" ::"r"(tmem_d),
            "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
            : "memory");
}
__device__ __forceinline__ void tmem_ld32_issue(uint32_t taddr, float (&v)[32]) {
    uint32_t* r = reinterpret_cast<uint32_t*>(v);
    asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, "
            "%15, %16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
            : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
              "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
              "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
              "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
            : "r"(taddr)
            : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() {
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float (&v)[32]) {
    uint32_t* r = reinterpret_cast<uint32_t*>(v);
    asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, "
            "%15, %16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
            : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
              "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
              "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
              "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
            : "r"(taddr)
            : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// Epi must provide
//   struct State;                                   per-thread state that lives across the chunks of one tile row
//   __device__ void begin(State&, int row) const;   once per (tile, row)
//   __device__ void chunk(State&, int row, int col0, const float (&v)[32]) const;   columns col0..col0+31
// row < M is guaranteed by the caller; columns may exceed N -- the functor masks them.  The two column halves of a
// tile row are handled by different warps: a functor must not carry state from chunk to chunk.
// Epilogues with  static constexpr bool kTmaStore = true  provide  transform(col0, v, o)  instead of chunk():
// the kernel stages each 32 x 32 f32 block in (128-byte swizzled) shared memory and hands it to the TMA unit, which
// writes whole lines and clips at the matrix edges; the LSU never sees the row-strided stores that otherwise make a
// wide f32 epilogue 2x longer than the MMAs of a tile.
template<class Epi, bool TMA_STORE = false, int TN = BN>
__global__ void __launch_bounds__(THREADS, 1)
        gemm16_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                      const __grid_constant__ CUtensorMap tmOut, int M, int N, int K, uint32_t idesc, const Epi epi) {
    static_assert(TN == BN || (TN == 128 && !TMA_STORE), "tile widths: 256, or 128 with the register epilogue");
    constexpr int STAGE_BYTES = A_BYTES + TN * BK * 2;  // (shadows the namespace constant: this kernel's ring stage)
    // epilogue warps: eight (two per TMEM lane quarter) for the register epilogues; the TMA-store variant keeps four --
    // its staging boxes would cost a ring stage, and with three stages the score layer lost more (721 -> 790 us per
    // 18944 frames) than the second set of warps gave
    constexpr int EPI_WARPS = TMA_STORE ? 4 : 8;
    extern __shared__ unsigned char smem_dyn[];
    const uint32_t raw   = smem_u32(smem_dyn);
    const uint32_t pad   = (1024u - (raw & 1023u)) & 1023u;
    unsigned char* base  = smem_dyn + pad;
    const uint32_t sbase = raw + pad;
    uint64_t* full       = reinterpret_cast<uint64_t*>(base + STAGES * STAGE_BYTES);
    uint64_t* empty      = full + STAGES;
    uint64_t* tfull      = empty + STAGES;
    uint64_t* tempty     = tfull + 2;
    uint32_t* tmemPtr    = reinterpret_cast<uint32_t*>(tempty + 2);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nMB = (M + BM - 1) / BM, nNB = (N + TN - 1) / TN, nTiles = nMB * nNB, nKB = K / BK;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmA);
        tma_prefetch_desc(&tmB);
        if (TMA_STORE)
            tma_prefetch_desc(&tmOut);
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], 1);
        }
        for (int a = 0; a < 2; ++a) {
            mbar_init(&tfull[a], 1);
            mbar_init(&tempty[a], EPI_WARPS);
        }
        mbar_fence_init();
    }
    if (warp == 2) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmemPtr)),
                     "r"((uint32_t)TMEM_COLS)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmemBase = *tmemPtr;

    if (warp == 0) {
        if (lane == 0) {
            uint32_t it = 0;
            for (int tile = blockIdx.x; tile < nTiles; tile += gridDim.x) {
                const int mb = tile / nNB, nb = tile - mb * nNB;
                for (int kb = 0; kb < nKB; ++kb, ++it) {
                    const uint32_t s = it % STAGES, ph = (it / STAGES) & 1u;
                    mbar_wait(&empty[s], ph ^ 1u);
                    mbar_expect_tx(&full[s], STAGE_BYTES);
                    tma_load_2d(sbase + s * STAGE_BYTES, &tmA, kb * BK, mb * BM, &full[s]);
                    tma_load_2d(sbase + s * STAGE_BYTES + A_BYTES, &tmB, kb * BK, nb * TN, &full[s]);
                }
            }
        }
    }
    else if (warp == 1) {
        if (lane == 0) {
            uint32_t it = 0, tc = 0;
            for (int tile = blockIdx.x; tile < nTiles; tile += gridDim.x, ++tc) {
                const uint32_t a = tc & 1u, aph = (tc >> 1) & 1u;
                mbar_wait(&tempty[a], aph ^ 1u);
                tc_fence_after();
                const uint32_t dTmem = tmemBase + a * TN;
                for (int kb = 0; kb < nKB; ++kb, ++it) {
                    const uint32_t s = it % STAGES, ph = (it / STAGES) & 1u;
                    mbar_wait(&full[s], ph);
                    tc_fence_after();
                    const uint64_t ad = smem_desc(sbase + s * STAGE_BYTES);
                    const uint64_t bd = smem_desc(sbase + s * STAGE_BYTES + A_BYTES);
#pragma unroll
                    for (int k = 0; k < BK / 16; ++k)  // 32 bytes of K per UMMA, inside the 128-byte swizzle row
                        tc_mma(dTmem, ad + 2 * k, bd + 2 * k, idesc, (uint32_t)((kb | k) != 0));
                    tc_commit(&empty[s]);  // frees the stage once the MMAs above have read it
                }
                tc_commit(&tfull[a]);  // accumulator complete
            }
        }
    }
    else if (warp >= 4 && warp < 4 + EPI_WARPS) {
        const int q = warp & 3;          // TMEM lane quarter this warp may read
        const int h = (warp - 4) >> 2;   // eight warps: which half of the tile's columns
        constexpr int CPW = TN / 32 / (EPI_WARPS / 4);  // 32-column chunks per warp
        uint32_t  tc = 0;
        // TMA-store staging: two 4 KB boxes per epilogue warp behind the operand ring (1024-byte aligned)
        const uint32_t outStage = sbase + STAGES * STAGE_BYTES + 1024 + (warp - 4) * 2 * 4096;
        for (int tile = blockIdx.x; tile < nTiles; tile += gridDim.x, ++tc) {
            const int      mb = tile / nNB, nb = tile - mb * nNB;
            const uint32_t a = tc & 1u, aph = (tc >> 1) & 1u;
            mbar_wait(&tfull[a], aph);
            tc_fence_after();
            const int row = mb * BM + q * 32 + lane;
            typename Epi::State st;
            if (row < M)
                epi.begin(st, row);
#pragma unroll 1
            for (int c = h * CPW; c < (h + 1) * CPW; c += 2) {  // two TMEM loads in flight per wait
                float          v0[32], v1[32];
                const uint32_t ta = tmemBase + ((uint32_t)(q * 32) << 16) + a * TN + c * 32;
                tmem_ld32_issue(ta, v0);
                tmem_ld32_issue(ta + 32, v1);
                tmem_ld_wait();
                if constexpr (TMA_STORE) {
#pragma unroll
                    for (int hlf = 0; hlf < 2; ++hlf) {
                        const int col0 = nb * TN + (c + hlf) * 32;
                        if (col0 >= N)  // warp-uniform: the whole box lies outside the matrix
                            continue;
                        float o[32];
                        epi.transform(col0, hlf ? v1 : v0, o);
                        if (lane == 0)
                            bulk_wait_read<1>();  // the store that used this box two steps ago has read it
                        __syncwarp();
                        // row = lane, 16-byte chunk j of the row lives at chunk j ^ (row & 7)  (SWIZZLE_128B)
                        const uint32_t dst = outStage + hlf * 4096 + lane * 128;
#pragma unroll
                        for (int j = 0; j < 8; ++j)
                            asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(dst + ((j ^ (lane & 7)) << 4)),
                                         "f"(o[4 * j]), "f"(o[4 * j + 1]), "f"(o[4 * j + 2]), "f"(o[4 * j + 3])
                                         : "memory");
                        fence_proxy_async();
                        __syncwarp();
                        if (lane == 0) {
                            tma_store_2d(&tmOut, outStage + hlf * 4096, col0, mb * BM + q * 32);
                            bulk_commit();
                        }
                    }
                }
                else if (row < M) {
                    epi.chunk(st, row, nb * TN + c * 32, v0);
                    epi.chunk(st, row, nb * TN + c * 32 + 32, v1);
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0)
                mbar_arrive(&tempty[a]);
        }
        if (TMA_STORE && lane == 0)
            asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");  // all stores complete before the CTA exits
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmemBase), "r"((uint32_t)TMEM_COLS)
                     : "memory");
    }
}

// ---------------------------------------------------------------- host: tensor maps
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline EncodeTiledFn encode_fn() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void*                           p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = (EncodeTiledFn)p;
    }
    return fn;
}

// 2-D map over a row-major [rows x cols] 16-bit matrix with row pitch ld elements; box = boxRows x 64
inline int make_map(CUtensorMap* map, const void* ptr, uint64_t rows, uint64_t cols, uint64_t ld, uint32_t boxRows,
                    bool bf16) {
    EncodeTiledFn fn = encode_fn();
    if (!fn) {
        rb::set_error("cuTensorMapEncodeTiled is not available from the driver");
        return RB_ERR_CUDA;
    }
    cuuint64_t dims[2]    = {cols, rows};
    cuuint64_t strides[1] = {ld * 2};
    cuuint32_t box[2]     = {64, boxRows};
    cuuint32_t estr[2]    = {1, 1};
    CUresult   r = fn(map, bf16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2,
                    const_cast<void*>(ptr), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        rb::set_error("cuTensorMapEncodeTiled failed with CUresult %d (rows %llu cols %llu ld %llu)", (int)r,
                      (unsigned long long)rows, (unsigned long long)cols, (unsigned long long)ld);
        return RB_ERR_CUDA;
    }
    return RB_OK;
}

// tmB must have been made with TN-row boxes (make_map(..., TN, ...))
template<class Epi, int TN = BN>
int launch(const CUtensorMap& tmA, const CUtensorMap& tmB, int M, int N, int K, int fmt, const Epi& epi, int smCount,
           cudaStream_t s) {
    RB_CUDA(cudaFuncSetAttribute(gemm16_kernel<Epi, false, TN>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes(TN)));
    RB_REQUIRE(K % BK == 0 && K > 0, "GEMM K=%d must be a positive multiple of %d", K, BK);
    const int nTiles = ((M + BM - 1) / BM) * ((N + TN - 1) / TN);
    const int grid   = std::min(nTiles, smCount);
    gemm16_kernel<Epi, false, TN><<<grid, THREADS, smem_bytes(TN), s>>>(tmA, tmB, tmA, M, N, K, instr_desc(fmt, TN), epi);
    RB_LAUNCH_CHECK();
    return RB_OK;
}

// f32 output through TMA stores; tmOut: map over the [M x N] f32 output with 32 x 32 boxes (make_map_out_f32)
template<class Epi>
int launch_tma_store(const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmOut, int M, int N, int K,
                     int fmt, const Epi& epi, int smCount, cudaStream_t s) {
    RB_CUDA(cudaFuncSetAttribute(gemm16_kernel<Epi, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 SMEM_BYTES_TMA));
    RB_REQUIRE(K % BK == 0 && K > 0, "GEMM K=%d must be a positive multiple of %d", K, BK);
    const int nTiles = ((M + BM - 1) / BM) * ((N + BN - 1) / BN);
    const int grid   = std::min(nTiles, smCount);
    gemm16_kernel<Epi, true><<<grid, THREADS, SMEM_BYTES_TMA, s>>>(tmA, tmB, tmOut, M, N, K, instr_desc(fmt), epi);
    RB_LAUNCH_CHECK();
    return RB_OK;
}

// 2-D map over a row-major [rows x cols] f32 matrix with row pitch ld elements (ld * 4 a multiple of 16); box 32 x 32
inline int make_map_out_f32(CUtensorMap* map, const void* ptr, uint64_t rows, uint64_t cols, uint64_t ld) {
    EncodeTiledFn fn = encode_fn();
    if (!fn) {
        rb::set_error("cuTensorMapEncodeTiled is not available from the driver");
        return RB_ERR_CUDA;
    }
    cuuint64_t dims[2]    = {cols, rows};
    cuuint64_t strides[1] = {ld * 4};
    cuuint32_t box[2]     = {32, 32};
    cuuint32_t estr[2]    = {1, 1};
    CUresult   r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<void*>(ptr), dims, strides, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        rb::set_error("cuTensorMapEncodeTiled (f32 output) failed with CUresult %d (rows %llu cols %llu ld %llu)", (int)r,
                      (unsigned long long)rows, (unsigned long long)cols, (unsigned long long)ld);
        return RB_ERR_CUDA;
    }
    return RB_OK;
}

}  // namespace rbgemm
