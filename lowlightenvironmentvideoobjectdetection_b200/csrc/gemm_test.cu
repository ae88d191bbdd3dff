// gemm_test.cu -- a plain tcgen05 GEMM (D = A * B^T) used by tests/test_gpu_tc.py to validate the descriptor / pipeline
// building blocks of tc.cuh (and the tensor maps of tmap.cu) in isolation before they are trusted inside the fused SELSA and
// most-similar-location kernels.  NOT part of libvodagg.so: it is linked only into libvodagg_selftest.so
// (include/vodagg_selftest.h, built by the same build.py), so the product library carries no test kernels.
//
// Kernel shape: one CTA per 128x128 output tile, 4-stage TMA->smem ring (128-byte K slices, SWIZZLE_128B),
// warp 0 = TMA producer, warp 1 = MMA issuer (one elected thread, accumulator in TMEM),
// warps 2..5 = epilogue (tcgen05.ld 32x32b -> global).
#include <cuda_bf16.h>

#include "common.cuh"
#include "tc.cuh"
#include "../../include/vodagg_selftest.h"

namespace vod {

constexpr int kGtStages = 4;
constexpr int kGtTileBytes = 128 * 128;  // 128 rows x 128 B
constexpr int kGtThreads = 192;

template <bool BF16>
__global__ void __launch_bounds__(kGtThreads, 1)
gemm_nt_kernel(const __grid_constant__ CUtensorMap tm_a, const __grid_constant__ CUtensorMap tm_b, float *__restrict__ D,
               int M, int N, int K) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t *sa = smem;                                   // [stages][16 KB]
    uint8_t *sb = smem + kGtStages * kGtTileBytes;        // [stages][16 KB]
    __shared__ uint64_t full_bar[kGtStages], empty_bar[kGtStages], acc_bar;
    __shared__ uint32_t tmem_slot;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int m0 = blockIdx.y * 128, n0 = blockIdx.x * 128;
    constexpr int kElemsPerSlice = BF16 ? 64 : 32;  // 128 bytes of K
    const int nkb = (K + kElemsPerSlice - 1) / kElemsPerSlice;

    if (threadIdx.x == 0) {
        for (int s = 0; s < kGtStages; ++s) { tc::mbar_init(&full_bar[s], 1); tc::mbar_init(&empty_bar[s], 1); }
        tc::mbar_init(&acc_bar, 1);
        tc::fence_barrier_init();
    }
    if (warp == 1) tc::tmem_alloc(&tmem_slot, 128);
    tc::tcgen05_fence_before();
    __syncthreads();
    tc::tcgen05_fence_after();
    const uint32_t tmem = tmem_slot;

    if (warp == 0) {
        if (tc::elect_one()) {
            tc::tma_prefetch_desc(&tm_a);
            tc::tma_prefetch_desc(&tm_b);
            for (int kb = 0; kb < nkb; ++kb) {
                const int s = kb % kGtStages;
                const uint32_t ph = (kb / kGtStages) & 1;
                tc::mbar_wait(&empty_bar[s], ph ^ 1);
                tc::mbar_arrive_expect_tx(&full_bar[s], 2 * kGtTileBytes);
                tc::tma_load_2d(sa + s * kGtTileBytes, &tm_a, &full_bar[s], kb * kElemsPerSlice, m0);
                tc::tma_load_2d(sb + s * kGtTileBytes, &tm_b, &full_bar[s], kb * kElemsPerSlice, n0);
            }
        }
    } else if (warp == 1) {
        constexpr uint32_t idesc = tc::umma_idesc(BF16 ? tc::kFmtBF16 : tc::kFmtTF32, 128, 128);
        for (int kb = 0; kb < nkb; ++kb) {
            const int s = kb % kGtStages;
            const uint32_t ph = (kb / kGtStages) & 1;
            tc::mbar_wait(&full_bar[s], ph);
            tc::tcgen05_fence_after();
            if (tc::elect_one()) {
                const uint32_t a0 = tc::smem_u32(sa + s * kGtTileBytes), b0 = tc::smem_u32(sb + s * kGtTileBytes);
#pragma unroll
                for (int k = 0; k < 4; ++k) {  // 4 x 32-byte K steps per 128-byte slice (UMMA_K = 16 bf16 / 8 tf32)
                    const uint64_t ad = tc::umma_desc_k_sw128(a0 + k * 32), bd = tc::umma_desc_k_sw128(b0 + k * 32);
                    if (BF16) tc::umma_f16(tmem, ad, bd, idesc, (kb | k) != 0);
                    else tc::umma_tf32(tmem, ad, bd, idesc, (kb | k) != 0);
                }
                tc::umma_commit(&empty_bar[s]);
                if (kb == nkb - 1) tc::umma_commit(&acc_bar);
            }
            __syncwarp();
        }
    } else {
        const int quarter = warp & 3;  // TMEM lane quarter this warp may access
        tc::mbar_wait(&acc_bar, 0);
        tc::tcgen05_fence_after();
        const int row = m0 + quarter * 32 + lane;
#pragma unroll 1
        for (int c = 0; c < 4; ++c) {
            uint32_t r[32];
            tc::tmem_ld_32x32(tmem + ((uint32_t)(quarter * 32) << 16) + c * 32, r);
            tc::tmem_ld_wait();
            if (row < M) {
                for (int j = 0; j < 32; ++j) {
                    const int col = n0 + c * 32 + j;
                    if (col < N) D[(size_t)row * N + col] = __uint_as_float(r[j]);
                }
            }
        }
    }
    tc::tcgen05_fence_before();
    __syncthreads();
    if (warp == 1) tc::tmem_dealloc(tmem, 128);
}

// ---------------------------------------------------------------------------------------------------------
// Variant with the A operand resident in TMEM (tcgen05.mma ... [d], [a_tmem], bdesc ...): validates the layout the
// most-similar-location kernel relies on -- row r of A lives in TMEM lane r, its K elements are packed two bf16 per
// 32-bit column (element k in column k/2, low half first), written with tcgen05.st.32x32b by the warp that owns
// the lane quarter.  bf16 only, K <= 512 and K % 64 == 0.  TMEM columns: [0,128) accumulator, [256, 256+K/2) A.
__global__ void __launch_bounds__(kGtThreads, 1)
gemm_nt_ts_kernel(const __nv_bfloat16 *__restrict__ A, const __grid_constant__ CUtensorMap tm_b, float *__restrict__ D,
                  int M, int N, int K) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t *sb = smem;  // [stages][16 KB]
    __shared__ uint64_t full_bar[kGtStages], empty_bar[kGtStages], acc_bar, a_bar;
    __shared__ uint32_t tmem_slot;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int m0 = blockIdx.y * 128, n0 = blockIdx.x * 128;
    const int nkb = K / 64;
    if (threadIdx.x == 0) {
        for (int s = 0; s < kGtStages; ++s) { tc::mbar_init(&full_bar[s], 1); tc::mbar_init(&empty_bar[s], 1); }
        tc::mbar_init(&acc_bar, 1);
        tc::mbar_init(&a_bar, 128);
        tc::fence_barrier_init();
    }
    if (warp == 1) tc::tmem_alloc(&tmem_slot, 512);
    tc::tcgen05_fence_before();
    __syncthreads();
    tc::tcgen05_fence_after();
    const uint32_t tmem = tmem_slot;
    constexpr uint32_t kACol = 256;

    if (warp == 0) {
        if (tc::elect_one()) {
            tc::tma_prefetch_desc(&tm_b);
            for (int kb = 0; kb < nkb; ++kb) {
                const int s = kb % kGtStages;
                tc::mbar_wait(&empty_bar[s], ((kb / kGtStages) & 1) ^ 1);
                tc::mbar_arrive_expect_tx(&full_bar[s], kGtTileBytes);
                tc::tma_load_2d(sb + s * kGtTileBytes, &tm_b, &full_bar[s], kb * 64, n0);
            }
        }
    } else if (warp == 1) {
        constexpr uint32_t idesc = tc::umma_idesc(tc::kFmtBF16, 128, 128);
        tc::mbar_wait(&a_bar, 0);
        tc::tcgen05_fence_after();
        for (int kb = 0; kb < nkb; ++kb) {
            const int s = kb % kGtStages;
            tc::mbar_wait(&full_bar[s], (kb / kGtStages) & 1);
            tc::tcgen05_fence_after();
            if (tc::elect_one()) {
                const uint32_t b0 = tc::smem_u32(sb + s * kGtTileBytes);
#pragma unroll
                for (int k = 0; k < 4; ++k)   // 16 bf16 of K per MMA = 8 TMEM columns of A, 32 bytes of the B slice
                    tc::umma_f16_ts(tmem, tmem + kACol + kb * 32 + k * 8, tc::umma_desc_k_sw128(b0 + k * 32), idesc,
                                    (kb | k) != 0);
                tc::umma_commit(&empty_bar[s]);
                if (kb == nkb - 1) tc::umma_commit(&acc_bar);
            }
            __syncwarp();
        }
    } else {
        const int quarter = warp & 3;
        const int row = m0 + quarter * 32 + lane;
        const uint32_t tl = tmem + ((uint32_t)(quarter * 32) << 16);
        // A row -> TMEM: 16 packed columns (32 bf16) per tcgen05.st
        for (int c = 0; c < K / 2; c += 16) {
            uint32_t v[16];
            if (row < M) {
                const uint4 *src = reinterpret_cast<const uint4 *>(A + (size_t)row * K + 2 * c);
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const uint4 u = __ldg(src + q);
                    v[4 * q] = u.x; v[4 * q + 1] = u.y; v[4 * q + 2] = u.z; v[4 * q + 3] = u.w;
                }
            } else {
#pragma unroll
                for (int q = 0; q < 16; ++q) v[q] = 0u;
            }
            tc::tmem_st_32x16(tl + kACol + c, v);
        }
        tc::tmem_st_wait();
        tc::tcgen05_fence_before();
        tc::mbar_arrive(&a_bar);
        tc::mbar_wait(&acc_bar, 0);
        tc::tcgen05_fence_after();
#pragma unroll 1
        for (int c = 0; c < 4; ++c) {
            uint32_t r[32];
            tc::tmem_ld_32x32(tl + c * 32, r);
            tc::tmem_ld_wait();
            if (row < M)
                for (int j = 0; j < 32; ++j) {
                    const int col = n0 + c * 32 + j;
                    if (col < N) D[(size_t)row * N + col] = __uint_as_float(r[j]);
                }
        }
    }
    tc::tcgen05_fence_before();
    __syncthreads();
    if (warp == 1) tc::tmem_dealloc(tmem, 512);
}

}  // namespace vod

using namespace vod;

extern "C" int vod_test_gemm_nt(const void *a, const void *b, float *d, int M, int N, int K, int dtype,
                                vod_stream_t stream) {
    VOD_REQUIRE(a && b && d && M > 0 && N > 0 && K > 0, "vod_test_gemm_nt: bad args");
    VOD_REQUIRE(dtype == VOD_DTYPE_F32 || dtype == VOD_DTYPE_BF16 || dtype == 2, "vod_test_gemm_nt: dtype");
    if (!vod_device_is_sm100()) return fail(VOD_E_UNSUPPORTED, "vod_test_gemm_nt: device is not sm_100");
    if (dtype == 2) {   // bf16, A operand staged in TMEM
        VOD_REQUIRE(K % 64 == 0 && K <= 512, "vod_test_gemm_nt: TMEM-A variant needs K %% 64 == 0 and K <= 512");
        CUtensorMap tb2;
        int rc2 = make_tmap_2d_sw128(&tb2, b, 2, N, K, (uint64_t)K * 2, 128);
        if (rc2) return rc2;
        const int smem2 = kGtStages * kGtTileBytes + 1024;
        cudaFuncSetAttribute(gemm_nt_ts_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem2);
        dim3 grid2(ceil_div(N, 128), ceil_div(M, 128));
        gemm_nt_ts_kernel<<<grid2, kGtThreads, smem2, as_stream(stream)>>>(reinterpret_cast<const __nv_bfloat16 *>(a), tb2, d, M, N, K); note_launch();
        return check_launch("vod_test_gemm_nt(tmem-A)");
    }
    const int eb = dtype == VOD_DTYPE_BF16 ? 2 : 4;
    CUtensorMap ta, tb;
    int rc = make_tmap_2d_sw128(&ta, a, eb, M, K, (uint64_t)K * eb, 128);
    if (rc) return rc;
    rc = make_tmap_2d_sw128(&tb, b, eb, N, K, (uint64_t)K * eb, 128);
    if (rc) return rc;
    const int smem = 2 * kGtStages * kGtTileBytes + 1024;
    dim3 grid(ceil_div(N, 128), ceil_div(M, 128));
    if (dtype == VOD_DTYPE_BF16) {
        cudaFuncSetAttribute(gemm_nt_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        gemm_nt_kernel<true><<<grid, kGtThreads, smem, as_stream(stream)>>>(ta, tb, d, M, N, K); note_launch();
    } else {
        cudaFuncSetAttribute(gemm_nt_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        gemm_nt_kernel<false><<<grid, kGtThreads, smem, as_stream(stream)>>>(ta, tb, d, M, N, K); note_launch();
    }
    return check_launch("vod_test_gemm_nt");
}
