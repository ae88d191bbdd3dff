// msra.cuh -- shared declarations of the most-similar-RoI-align kernels (tafa.cu, msra_gemm.cu).
#pragma once
#include "common.cuh"

namespace vod {

constexpr int kMsraMaxK = 4;      // num_most_similar_points <= 4 (reference default 2)
constexpr int kMsraCand = 16;     // candidates per (row, frame) kept by the tensor-core pass (4 column groups x top-4)
// A candidate is re-scored in fp32 when its bf16-GEMM similarity is within this margin of the 2nd best one.
// bf16 operand rounding gives an error of ~1e-4 (sigma) on unit vectors, the 20-bit key truncation < 3.5e-4.
constexpr float kMsraMargin = 2.5e-3f;
// Bound used on |bf16-GEMM key value - exact fp32 similarity| when a full candidate list is tested against the EXACT k-th best
// similarity (overflow detection): half the re-score margin (3 key quanta: 2.4e-4 of key rounding + 12 sigma of bf16 noise).
constexpr float kMsraKeyErr = 1.5e-3f;

int msra_launch_scan(const float *roi, const float *ref, const float *roi_norm, const float *ref_norm, float *out,
                     int *idx_out, float *val_out, int NP, int C, int T, int HW, int k, cudaStream_t st);
struct MsraOvf;
int msra_launch_rescore(const float *roi, const float *ref, const float *roi_norm, const float *ref_norm,
                        const uint32_t *cand, int KC, float *out, int *idx_out, float *val_out, int NP, int C, int T,
                        int HW, int k, const MsraOvf *ovf, cudaStream_t st);
// tensor-core candidate pass (msra_gemm.cu): bf16 unit rows -> cand [NP, T, kMsraCand]
bool msra_gemm_supported(int NP, int C, int T, int HW);
int msra_launch_gemm_topk(const void *roi_unit_bf16, const void *ref_unit_bf16, uint32_t *cand, int NP, int NP_pad, int C,
                          int T, int HW, cudaStream_t st);

// ---- overflow bookkeeping of the tensor-core candidate pass (exact-by-construction top-k) --------------------------
// The candidate pass keeps the 4 best bf16 similarities of each of four interleaved location groups.  A location can only be
// missing from the lists if its group's list is full, i.e. if the group's 4th key is at least as large as the location's own
// key; such a location can belong to the true top-k only if that 4th key lies within the re-score margin of the k-th best
// key of the (row, frame).  The re-score kernels detect exactly that condition per group, append the (row, frame) to a work
// list in the caller's workspace, and msra_overflow_fix re-scans the flagged groups of those pairs in exact fp32
// (msra_overflow.cu).  No flagged pair -> the fix-up kernels find empty lists and return.
struct MsraOvf {
    int *ctrl;          // [0] number of flagged pairs, [1 + 4*t + g] entries of bin (frame t, group g); zeroed before the re-score
    int4 *pair_list;    // [NP*T] (row, frame, group mask, -)
    float4 *pair_top;   // [NP*T] the re-scored exact top-2 of the pair: (v0, bits(l0), v1, bits(l1))
    int2 *bin_list;     // [4*T][NP] (row, pair slot) of the pairs whose group g of frame t overflowed
    float4 *ovf_top;    // [NP*T][4] exact top-2 of group g of pair slot `pos`, written by the scan kernel
    int *done;          // [kMsraOvfSplitChunks] arrival counters of the location-split scan (zeroed with ctrl)
    float4 *split_top;  // [kMsraOvfSplitChunks][8 splits][8 rows] partial top-2 of a location slice
};
constexpr int kMsraOvfSplitChunks = 512;   // a short work list is scanned by up to 8 CTAs per item (latency); longer lists are not split
int msra_overflow_fix(const float *roi, const float *ref, const float *roi_norm, const float *ref_norm, const MsraOvf &o,
                      float *out, int *idx_out, float *val_out, int NP, int C, int T, int HW, int k, cudaStream_t st);

// Shared tail: given this warp's k best (value, location) pairs (sorted descending, warp-uniform),
// softmax over the k values (temporal_roi_align.py:155) and weighted gather of the raw reference
// features (temporal_roi_align.py:165-176).
template <int KMAX>
__device__ __forceinline__ void msra_emit(const float *__restrict__ ref_t /*[HW,C]*/, float *__restrict__ out_row,
                                          int *__restrict__ idx_out, float *__restrict__ val_out, const float *val,
                                          const int *loc, int k, int C, int lane) {
    // No comparable candidate (an all-zero RoI row or reference pixel makes every similarity NaN, the
    // reference divides by a zero norm without an epsilon): the reference's output row is NaN too.
    if (!(loc[0] >= 0 && loc[0] != 0x7fffffff) || (k > 1 && !(loc[k - 1] >= 0 && loc[k - 1] != 0x7fffffff))) {
        const float qnan = __int_as_float(0x7fc00000);
        for (int c = lane; c < C; c += 32) out_row[c] = qnan;
        if (lane < k) {
            if (idx_out) idx_out[lane] = 0;
            if (val_out) val_out[lane] = qnan;
        }
        return;
    }
    float w[KMAX];
    float m = val[0], sum = 0.f;
#pragma unroll
    for (int q = 0; q < KMAX; ++q) if (q < k) { w[q] = expf(val[q] - m); sum += w[q]; }
#pragma unroll
    for (int q = 0; q < KMAX; ++q) if (q < k) w[q] = w[q] / sum;
    if (lane < k) {
        // pick element `lane` without dynamic register indexing
        float v = val[0]; int l = loc[0];
#pragma unroll
        for (int q = 1; q < KMAX; ++q) if (lane == q) { v = val[q]; l = loc[q]; }
        if (idx_out) idx_out[lane] = l;
        if (val_out) val_out[lane] = v;
    }
    if ((C & 3) == 0) {
        for (int c = lane * 4; c < C; c += 128) {
            float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
            for (int q = 0; q < KMAX; ++q) if (q < k) {
                float4 v = ldg_f4(ref_t + (size_t)loc[q] * C + c);
                // topk_feats * topk_weights summed over k (temporal_roi_align.py:170-172)
                acc.x += v.x * w[q]; acc.y += v.y * w[q]; acc.z += v.z * w[q]; acc.w += v.w * w[q];
            }
            stg_cs_f4(out_row + c, acc);
        }
    } else {
        for (int c = lane; c < C; c += 32) {
            float acc = 0.f;
#pragma unroll
            for (int q = 0; q < KMAX; ++q) if (q < k) acc += __ldg(ref_t + (size_t)loc[q] * C + c) * w[q];
            out_row[c] = acc;
        }
    }
}

// insert (v, l) into a descending sorted list of length KMAX (ties: smaller location first)
template <int KMAX>
__device__ __forceinline__ void topk_insert(float (&val)[KMAX], int (&loc)[KMAX], float v, int l) {
#pragma unroll
    for (int q = 0; q < KMAX; ++q) {
        bool better = (v > val[q]) || (v == val[q] && l < loc[q]);
        if (better) {
            float tv = val[q]; int tl = loc[q];
            val[q] = v; loc[q] = l; v = tv; l = tl;
        }
    }
}

// warp-wide merge of per-lane sorted lists into the warp's top-k (result warp-uniform)
template <int KMAX>
__device__ __forceinline__ void topk_warp_merge(float (&val)[KMAX], int (&loc)[KMAX], int k, int lane) {
    float rv[KMAX]; int rl[KMAX];
#pragma unroll
    for (int q = 0; q < KMAX; ++q) { rv[q] = -INFINITY; rl[q] = 0x7fffffff; }
#pragma unroll
    for (int r = 0; r < KMAX; ++r) {
        if (r < k) {
            float bv = val[0]; int bl = loc[0];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                float ov = __shfl_xor_sync(0xffffffffu, bv, o);
                int ol = __shfl_xor_sync(0xffffffffu, bl, o);
                if (ov > bv || (ov == bv && ol < bl)) { bv = ov; bl = ol; }
            }
            rv[r] = bv; rl[r] = bl;
            if (val[0] == bv && loc[0] == bl) {  // the owning lane pops its head
#pragma unroll
                for (int q = 0; q + 1 < KMAX; ++q) { val[q] = val[q + 1]; loc[q] = loc[q + 1]; }
                val[KMAX - 1] = -INFINITY; loc[KMAX - 1] = 0x7fffffff;
            }
        }
    }
#pragma unroll
    for (int q = 0; q < KMAX; ++q) { val[q] = rv[q]; loc[q] = rl[q]; }
}

}  // namespace vod
