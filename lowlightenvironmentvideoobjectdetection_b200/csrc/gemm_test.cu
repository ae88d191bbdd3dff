// gemm_test.cu -- host-side TMA tensor-map construction + a plain tcgen05 GEMM (D = A * B^T) used by
// tests/test_gpu_tc.py to validate the descriptor / pipeline building blocks of tc.cuh in isolation
// before they are trusted inside the fused SELSA and most-similar-location kernels.
//
// Kernel shape: one CTA per 128x128 output tile, 4-stage TMA->smem ring (128-byte K slices, SWIZZLE_128B),
// warp 0 = TMA producer, warp 1 = MMA issuer (one elected thread, accumulator in TMEM),
// warps 2..5 = epilogue (tcgen05.ld 32x32b -> global).
#include <cuda_bf16.h>

#include "common.cuh"
#include "tc.cuh"

namespace vod {

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
    static EncodeTiledFn fn = nullptr;  // immutable once resolved
    if (!fn) {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
        else
            cudaGetLastError();
    }
    return fn;
}

int make_tmap_2d_sw128(CUtensorMap *map, const void *base, int elem_bytes, uint64_t rows, uint64_t cols,
                       uint64_t row_stride_bytes, uint32_t box_rows) {
    EncodeTiledFn enc = get_encode_fn();
    if (!enc) return fail(VOD_E_LAUNCH, "cuTensorMapEncodeTiled entry point unavailable");
    if ((reinterpret_cast<uintptr_t>(base) & 15) || (row_stride_bytes & 15))
        return fail(VOD_E_BADARG, "TMA operand needs 16-byte aligned base and row stride");
    // fp32 operands of kind::tf32 MMAs: TFLOAT32 makes the TMA unit round to tf32 while loading (the MMA
    // would otherwise truncate the low 13 mantissa bits, a biased error of ~5e-4 relative).
    CUtensorMapDataType dt = elem_bytes == 2 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_TFLOAT32;
    cuuint64_t dims[2] = {cols, rows};
    cuuint64_t strides[1] = {row_stride_bytes};
    cuuint32_t box[2] = {(cuuint32_t)(128 / elem_bytes), box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(map, dt, 2, const_cast<void *>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(VOD_E_LAUNCH, "cuTensorMapEncodeTiled failed (%d)", (int)r);
    return VOD_OK;
}

// bf16 [rows, C] matrix seen as (64 channels, rows, C/64 K-slices): one box = box_rows x 128 B x box_slices, landing in
// shared memory as box_slices consecutive K-major SWIZZLE_128B tiles (slice stride 128 B < row stride: dimensions may be
// listed in any order, only the innermost one must be contiguous).
int make_tmap_kslices_sw128(CUtensorMap *map, const void *base, uint64_t rows, uint64_t C, uint32_t box_rows, uint32_t box_slices) {
    EncodeTiledFn enc = get_encode_fn();
    if (!enc) return fail(VOD_E_LAUNCH, "cuTensorMapEncodeTiled entry point unavailable");
    if ((reinterpret_cast<uintptr_t>(base) & 15) || C % 64 != 0)
        return fail(VOD_E_BADARG, "TMA operand needs a 16-byte aligned base and C %% 64 == 0");
    cuuint64_t dims[3] = {64, rows, C / 64};
    cuuint64_t strides[2] = {C * 2, 128};
    cuuint32_t box[3] = {64, box_rows, box_slices};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void *>(base), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(VOD_E_LAUNCH, "cuTensorMapEncodeTiled (k-slices) failed (%d)", (int)r);
    return VOD_OK;
}

// B operand of the msra similarity GEMM: bf16 unit rows [T*HW, C], seen as a 5-D tensor
//   (64 channels | A = location / 4 | g = location % 4 | frame t | K slice of 64 channels)
// so that one box (64, 32, 2, 1, slices) lands in shared memory as `slices` K-major SWIZZLE_128B tiles of 64 rows ordered
// r = g_local * 32 + a, i.e. location = 128 * tile + 4 * a + g: neighbouring locations fall into different 32-row groups.
// When HW % 4 != 0 the last A of a frame touches up to 3 rows of the next frame (masked by the consumer); for the last
// frame these lie past T*HW, hence the 3 rows of readable padding the C ABI asks for.
int make_tmap_msra_b(CUtensorMap *map, const void *base, uint64_t T, uint64_t HW, uint64_t C, uint32_t box_slices) {
    EncodeTiledFn enc = get_encode_fn();
    if (!enc) return fail(VOD_E_LAUNCH, "cuTensorMapEncodeTiled entry point unavailable");
    if ((reinterpret_cast<uintptr_t>(base) & 15) || C % 64 != 0)
        return fail(VOD_E_BADARG, "TMA operand needs a 16-byte aligned base and C %% 64 == 0");
    cuuint64_t dims[5] = {64, (HW + 3) / 4, 4, T, C / 64};
    cuuint64_t strides[4] = {4 * C * 2, C * 2, HW * C * 2, 128};
    cuuint32_t box[5] = {64, 32, 2, 1, box_slices};
    cuuint32_t estr[5] = {1, 1, 1, 1, 1};
    CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, const_cast<void *>(base), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(VOD_E_LAUNCH, "cuTensorMapEncodeTiled (msra B) failed (%d)", (int)r);
    return VOD_OK;
}

// fp32 [d2][d1][d0] tensor (d0 contiguous), dense un-swizzled box (b0, b1, b2): lands in shared memory as [b2][b1][b0].
// Out-of-range coordinates are zero-filled.  Used for the frame tiles of the key-projected TAFA logits kernel.
int make_tmap_f32_3d(CUtensorMap *map, const void *base, uint64_t d0, uint64_t d1, uint64_t d2, uint64_t stride1_bytes,
                     uint64_t stride2_bytes, uint32_t b0, uint32_t b1, uint32_t b2) {
    EncodeTiledFn enc = get_encode_fn();
    if (!enc) return fail(VOD_E_LAUNCH, "cuTensorMapEncodeTiled entry point unavailable");
    if ((reinterpret_cast<uintptr_t>(base) & 15) || (stride1_bytes & 15) || (stride2_bytes & 15) || (b0 * 4) % 16 != 0 ||
        b0 > 256 || b1 > 256 || b2 > 256)
        return fail(VOD_E_BADARG, "TMA operand needs 16-byte aligned base/strides and box dims <= 256");
    cuuint64_t dims[3] = {d0, d1, d2};
    cuuint64_t strides[2] = {stride1_bytes, stride2_bytes};
    cuuint32_t box[3] = {b0, b1, b2};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<void *>(base), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(VOD_E_LAUNCH, "cuTensorMapEncodeTiled (f32 3d) failed (%d)", (int)r);
    return VOD_OK;
}

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
