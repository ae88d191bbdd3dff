// tmap.cu -- host-side construction of the TMA tensor maps the tcgen05 / TMA kernels take as __grid_constant__ arguments
// (cuTensorMapEncodeTiled resolved through the runtime's driver entry point: no link-time dependency on libcuda).
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

// fp32 [d2][d1][d0] tensor (d0 contiguous), box (b0, b1, b2): lands in shared memory as [b2][b1][b0].
// Out-of-range coordinates are zero-filled.  Used for the frame tiles of the key-projected TAFA logits kernel.
// tf32_sw128 = false: dense, un-swizzled, bit-exact fp32.  true: the box rows must be 128 bytes (b0 = 32); they are stored with
// the 128-byte swizzle (16-byte chunk index ^= 128-byte line index mod 8, destination 1024-byte aligned) and the TMA unit rounds
// every value to tf32 on the way in -- the operand form of the mma.sync tf32 variant of that kernel.
int make_tmap_f32_3d(CUtensorMap *map, const void *base, uint64_t d0, uint64_t d1, uint64_t d2, uint64_t stride1_bytes,
                     uint64_t stride2_bytes, uint32_t b0, uint32_t b1, uint32_t b2, bool tf32_sw128) {
    EncodeTiledFn enc = get_encode_fn();
    if (!enc) return fail(VOD_E_LAUNCH, "cuTensorMapEncodeTiled entry point unavailable");
    if ((reinterpret_cast<uintptr_t>(base) & 15) || (stride1_bytes & 15) || (stride2_bytes & 15) || (b0 * 4) % 16 != 0 ||
        b0 > 256 || b1 > 256 || b2 > 256)
        return fail(VOD_E_BADARG, "TMA operand needs 16-byte aligned base/strides and box dims <= 256");
    if (tf32_sw128 && b0 != 32) return fail(VOD_E_BADARG, "swizzled fp32 box needs 128-byte rows");
    cuuint64_t dims[3] = {d0, d1, d2};
    cuuint64_t strides[2] = {stride1_bytes, stride2_bytes};
    cuuint32_t box[3] = {b0, b1, b2};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = enc(map, tf32_sw128 ? CU_TENSOR_MAP_DATA_TYPE_TFLOAT32 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<void *>(base),
                     dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     tf32_sw128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(VOD_E_LAUNCH, "cuTensorMapEncodeTiled (f32 3d) failed (%d)", (int)r);
    return VOD_OK;
}

}  // namespace vod
