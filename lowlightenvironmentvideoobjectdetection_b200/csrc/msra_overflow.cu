// msra_overflow.cu -- makes the tensor-core most-similar-location search exact by construction.
//
// TemporalRoIAlign.most_similar_roi_align takes an exact top-k over all H*W locations of every frame
// (mmtracking/mmtrack/models/roi_heads/roi_extractors/temporal_roi_align.py:142-155).  The tensor-core candidate pass
// (msra_gemm.cu) keeps only the 4 best bf16 similarities of each of four interleaved location groups; the re-score kernels
// (tafa.cu) flag every (row, frame, group) whose full list ends within the re-score margin of the k-th best key -- the only
// situation in which a location that fell off a list could belong to the exact top-k (msra.cuh).  The two kernels here
// resolve those flags on the same stream, with fixed grids (CUDA-graph capturable; an empty work list costs two ~2 us launches):
//
//   msra_overflow_scan_kernel      work item = (frame t, location group g, up to 8 flagged RoI rows): exact fp32 similarity of
//                                  those rows against EVERY location of the group (warp = 4 locations per pass, lanes split the
//                                  channels exactly as the re-score kernel does; a reference row is read once for 8 RoI rows),
//                                  exact top-2 per row -> ovf_top[pair][g]
//   msra_overflow_finalize_kernel  warp = flagged pair: merges the re-scored top-2 with the scanned groups' top-2
//                                  (de-duplicated by location) and, if the selection changed, re-emits softmax + gather.
//
// The exact top-2 of a group contains every member of the overall exact top-2 that lies in that group, so the merged result is
// the exact top-k over all locations -- whatever the bf16 pass dropped.
#include "common.cuh"
#include "msra.cuh"

namespace vod {

constexpr int kOvfRows = 8;        // flagged rows of one (frame, group) bin scanned together
constexpr int kOvfThreads = 256;
constexpr int kOvfWarps = kOvfThreads / 32;
constexpr int kOvfMaxC = 512;      // the tensor-core pass supports C <= 512
constexpr int kOvfMaxBins = 1024;  // 4 * T, T <= 256
static_assert(kOvfWarps == kOvfRows, "the block-level merge assigns one warp per row");

// warp = kOvfLB locations at a time; lane owns channels lane*4 + 128*i (the re-score kernel's mapping and arithmetic, so a
// location both kernels evaluate gets the bit-identical similarity); the R x kOvfLB = 32 partial dot products of a pass are
// reduced with ONE transposing butterfly (31 shuffles; lane k ends up with the complete sum k = row*4 + location) instead of
// 32 five-step reductions.  Reading every reference row once for 8 RoI rows and every RoI row once for 4 locations keeps the
// kernel FMA-bound: the first version (thread = location, rows of 2 KB per thread) was bound by L1 wavefronts at 100 us per
// work item; this one takes ~15 us per item and ~60 us for the ~3400 items of the iid-noise benchmark input.
constexpr int kOvfLB = 4;
static_assert(kOvfRows * kOvfLB == 32, "one butterfly reduces rows x locations = 32 values");

template <int NQ>
#ifndef VOD_OVF_MINB
#define VOD_OVF_MINB 2
#endif
__global__ void __launch_bounds__(kOvfThreads, VOD_OVF_MINB)
msra_overflow_scan_kernel(const float *__restrict__ roi, const float *__restrict__ ref, const float *__restrict__ roi_norm,
                          const float *__restrict__ ref_norm, const MsraOvf o, int NP, int C, int T, int HW) {
    constexpr int R = kOvfRows, LB = kOvfLB, CP = 128 * NQ;     // CP: padded channel count
    __shared__ __align__(16) float s_q[R * CP];
    __shared__ int s_prefix[kOvfMaxBins + 1];
    __shared__ int s_warp_sum[kOvfWarps];
    __shared__ float s_mv[kOvfWarps][R][2];
    __shared__ int s_ml[kOvfWarps][R][2];
    __shared__ int s_row[R], s_pos[R];
    __shared__ int s_last;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int nb = 4 * T;

    // chunk prefix over the bins: thread i owns bins 4i .. 4i+3
    int c4[4], local = 0;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const int b = 4 * tid + j;
        c4[j] = b < nb ? ceil_div(o.ctrl[1 + b], R) : 0;
        local += c4[j];
    }
    int incl = local;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const int v = __shfl_up_sync(0xffffffffu, incl, d);
        if (lane >= d) incl += v;
    }
    if (lane == 31) s_warp_sum[warp] = incl;
    __syncthreads();
    int base = 0;
    for (int w = 0; w < warp; ++w) base += s_warp_sum[w];
    int run = base + incl - local;
    if (tid == 0) s_prefix[0] = 0;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        run += c4[j];
        if (4 * tid + j < nb) s_prefix[4 * tid + j + 1] = run;
    }
    __syncthreads();
    const int total = s_prefix[nb];
    // A short work list (real video: a handful of flagged pairs) would leave one CTA scanning 600 locations on its own -- 60 us
    // of latency in every step.  Its locations are then split over up to 8 CTAs per item; the last slice to arrive merges.
    const int S = (total > 0 && total <= kMsraOvfSplitChunks) ? min(8, max(1, (int)gridDim.x / total)) : 1;

    for (int item = blockIdx.x; item < total * S; item += gridDim.x) {
        const int chunk = item / S, split = item - chunk * S;
        // bin of this chunk: largest b with s_prefix[b] <= chunk
        int lo = 0, hi = nb;
        while (hi - lo > 1) {
            const int mid = (lo + hi) >> 1;
            if (s_prefix[mid] <= chunk) lo = mid; else hi = mid;
        }
        const int bin = lo, t = bin >> 2, g = bin & 3;
        const int first = (chunk - s_prefix[bin]) * R;
        const int nrows = min(R, o.ctrl[1 + bin] - first);
        __syncthreads();                                   // the previous item's shared-memory reads are complete
        if (tid < R) {
            int2 e = make_int2(-1, -1);
            if (tid < nrows) e = o.bin_list[(size_t)bin * NP + first + tid];
            s_row[tid] = e.x; s_pos[tid] = e.y;
        }
        __syncthreads();
        // normalised RoI rows, same arithmetic as the re-score kernels: x * (1 / |x|); zero beyond C and for absent rows
        for (int i = tid; i < R * (CP / 4); i += kOvfThreads) {
            const int r = i / (CP / 4), cq = i - r * (CP / 4), row = s_row[r];
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (row >= 0 && 4 * cq < C) {
                const float qinv = 1.0f / __ldg(roi_norm + row);
                v = ldg_f4(roi + (size_t)row * C + 4 * cq);
                v.x *= qinv; v.y *= qinv; v.z *= qinv; v.w *= qinv;
            }
            *reinterpret_cast<float4 *>(s_q + r * CP + 4 * cq) = v;
        }
        __syncthreads();

        // lane k = 4*row + location-in-pass keeps the running top-2 of ITS row over ITS quarter of the locations
        float tv0 = -INFINITY, tv1 = -INFINITY;
        int tl0 = 0x7fffffff, tl1 = 0x7fffffff;
        const int nloc = (HW - g + 3) >> 2;                // locations 4*i + g < HW
        const float *ref_t = ref + (size_t)t * HW * C;
        for (int i0 = (split * kOvfWarps + warp) * LB; i0 < nloc; i0 += S * kOvfWarps * LB) {
            float4 v[LB][NQ];
#pragma unroll
            for (int j = 0; j < LB; ++j) {
                const int l = 4 * (i0 + j) + g;
                const bool ok = i0 + j < nloc;
                const float rinv = ok ? 1.0f / __ldg(ref_norm + (size_t)t * HW + l) : 0.f;
#pragma unroll
                for (int i = 0; i < NQ; ++i) {
                    float4 x = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (ok && lane * 4 + 128 * i < C) x = ldg_f4(ref_t + (size_t)l * C + lane * 4 + 128 * i);
                    x.x *= rinv; x.y *= rinv; x.z *= rinv; x.w *= rinv;
                    v[j][i] = x;
                }
            }
            float acc[R * LB];
#pragma unroll
            for (int r = 0; r < R; ++r) {
                float4 q[NQ];
#pragma unroll
                for (int i = 0; i < NQ; ++i) q[i] = *reinterpret_cast<const float4 *>(s_q + r * CP + lane * 4 + 128 * i);
#pragma unroll
                for (int j = 0; j < LB; ++j) {
                    float s = 0.f;
#pragma unroll
                    for (int i = 0; i < NQ; ++i) {
                        s = fmaf(q[i].x, v[j][i].x, s); s = fmaf(q[i].y, v[j][i].y, s);
                        s = fmaf(q[i].z, v[j][i].z, s); s = fmaf(q[i].w, v[j][i].w, s);
                    }
                    acc[r * LB + j] = s;
                }
            }
            // transposing butterfly: after the step with offset o, a lane keeps the half of its values selected by its bit o;
            // the additions pair the same lanes in the same order as warp_sum (xor 16, 8, 4, 2, 1)
#pragma unroll
            for (int o = 16, n = 16; o >= 1; o >>= 1, n >>= 1) {
                const bool up = (lane & o) != 0;
#pragma unroll
                for (int k = 0; k < n; ++k) {
                    const float keep = up ? acc[k + n] : acc[k];
                    const float send = up ? acc[k] : acc[k + n];
                    acc[k] = keep + __shfl_xor_sync(0xffffffffu, send, o);
                }
            }
            const int j = lane & (LB - 1);
            const float sres = acc[0];
            if (i0 + j < nloc) {
                const int l = 4 * (i0 + j) + g;
                // a lane visits its locations in increasing order: on equal similarity the earlier (smaller) location stays ahead
                if (sres > tv0) { tv1 = tv0; tl1 = tl0; tv0 = sres; tl0 = l; }
                else if (sres > tv1) { tv1 = sres; tl1 = l; }
            }
        }
        // merge the 4 lanes of a row (xor 1, 2), then the 8 warps through shared memory, one warp per row
        {
            float v2[2] = {tv0, tv1};
            int l2[2] = {tl0, tl1};
#pragma unroll
            for (int o = 1; o <= 2; o <<= 1) {
                const float ov0 = __shfl_xor_sync(0xffffffffu, v2[0], o), ov1 = __shfl_xor_sync(0xffffffffu, v2[1], o);
                const int ol0 = __shfl_xor_sync(0xffffffffu, l2[0], o), ol1 = __shfl_xor_sync(0xffffffffu, l2[1], o);
                topk_insert<2>(v2, l2, ov0, ol0);
                topk_insert<2>(v2, l2, ov1, ol1);
            }
            if ((lane & 3) == 0) {
                const int r = lane >> 2;
                s_mv[warp][r][0] = v2[0]; s_mv[warp][r][1] = v2[1]; s_ml[warp][r][0] = l2[0]; s_ml[warp][r][1] = l2[1];
            }
        }
        __syncthreads();
        if (warp < R) {
            const int r = warp;
            float v[2] = {-INFINITY, -INFINITY};
            int l[2] = {0x7fffffff, 0x7fffffff};
            if (lane < 2 * kOvfWarps) { v[0] = s_mv[lane >> 1][r][lane & 1]; l[0] = s_ml[lane >> 1][r][lane & 1]; }
            topk_warp_merge<2>(v, l, 2, lane);
            if (lane == 0) {
                const float4 res = make_float4(v[0], __int_as_float(l[0]), v[1], __int_as_float(l[1]));
                if (S == 1) { if (s_pos[r] >= 0) o.ovf_top[(size_t)s_pos[r] * 4 + g] = res; }
                else o.split_top[((size_t)chunk * 8 + split) * R + r] = res;
            }
        }
        if (S > 1) {
            __threadfence();
            __syncthreads();
            if (tid == 0) s_last = atomicAdd(o.done + chunk, 1) == S - 1;
            __syncthreads();
            if (s_last) {
                __threadfence();
                if (warp < R) {
                    const int r = warp;
                    float v[2] = {-INFINITY, -INFINITY};
                    int l[2] = {0x7fffffff, 0x7fffffff};
                    if (lane < S) {
                        const float4 e = __ldcg(o.split_top + ((size_t)chunk * 8 + lane) * R + r);
                        v[0] = e.x; l[0] = __float_as_int(e.y); v[1] = e.z; l[1] = __float_as_int(e.w);
                    }
                    topk_warp_merge<2>(v, l, 2, lane);
                    if (lane == 0 && s_pos[r] >= 0)
                        o.ovf_top[(size_t)s_pos[r] * 4 + g] = make_float4(v[0], __int_as_float(l[0]), v[1], __int_as_float(l[1]));
                }
            }
        }
    }
}

__global__ void __launch_bounds__(kOvfThreads)
msra_overflow_finalize_kernel(const float *__restrict__ ref, const MsraOvf o, float *__restrict__ out, int *__restrict__ idx_out,
                              float *__restrict__ val_out, int NP, int C, int T, int HW, int k) {
    const int lane = threadIdx.x & 31;
    const int npairs = o.ctrl[0];
    const int nwarps = gridDim.x * kOvfWarps;
    for (int pos = blockIdx.x * kOvfWarps + (threadIdx.x >> 5); pos < npairs; pos += nwarps) {
        const int4 pl = o.pair_list[pos];
        const int row = pl.x, t = pl.y;
        const unsigned mask = (unsigned)pl.z;
        const float4 pt = o.pair_top[pos];
        float val[2] = {pt.x, pt.z};
        int loc[2] = {__float_as_int(pt.y), __float_as_int(pt.w)};
        const int ol0 = loc[0], ol1 = loc[1];
        for (int g = 0; g < 4; ++g) {
            if (!(mask >> g & 1u)) continue;
            const float4 e = o.ovf_top[(size_t)pos * 4 + g];
            const float ev[2] = {e.x, e.z};
            const int el[2] = {__float_as_int(e.y), __float_as_int(e.w)};
#pragma unroll
            for (int j = 0; j < 2; ++j)
                // a location the re-score already holds keeps its re-scored value (the two sums differ in rounding order only)
                if (el[j] != 0x7fffffff && el[j] != loc[0] && el[j] != loc[1]) topk_insert<2>(val, loc, ev[j], el[j]);
        }
        if (loc[0] == ol0 && (k < 2 || loc[1] == ol1)) continue;      // the candidate pass had lost nothing that matters
        msra_emit<2>(ref + (size_t)t * HW * C, out + ((size_t)t * NP + row) * C,
                     idx_out ? idx_out + ((size_t)row * T + t) * k : nullptr, val_out ? val_out + ((size_t)row * T + t) * k : nullptr,
                     val, loc, k, C, lane);
    }
}

int msra_overflow_fix(const float *roi, const float *ref, const float *roi_norm, const float *ref_norm, const MsraOvf &o,
                      float *out, int *idx_out, float *val_out, int NP, int C, int T, int HW, int k, cudaStream_t st) {
    if (C > kOvfMaxC || (C & 3) || 4 * T > kOvfMaxBins || k > 2)
        return fail(VOD_E_UNSUPPORTED, "msra_overflow_fix: C=%d T=%d k=%d outside the tensor-core path's range", C, T, k);
    const int sms = num_sms();
    switch ((C + 127) >> 7) {
        case 1: msra_overflow_scan_kernel<1><<<2 * sms, kOvfThreads, 0, st>>>(roi, ref, roi_norm, ref_norm, o, NP, C, T, HW); break;
        case 2: msra_overflow_scan_kernel<2><<<2 * sms, kOvfThreads, 0, st>>>(roi, ref, roi_norm, ref_norm, o, NP, C, T, HW); break;
        case 3: msra_overflow_scan_kernel<3><<<2 * sms, kOvfThreads, 0, st>>>(roi, ref, roi_norm, ref_norm, o, NP, C, T, HW); break;
        default: msra_overflow_scan_kernel<4><<<2 * sms, kOvfThreads, 0, st>>>(roi, ref, roi_norm, ref_norm, o, NP, C, T, HW); break;
    }
    note_launch();
    int rc = check_launch("msra_overflow_scan");
    if (rc) return rc;
    msra_overflow_finalize_kernel<<<2 * sms, kOvfThreads, 0, st>>>(ref, o, out, idx_out, val_out, NP, C, T, HW, k);
    note_launch();
    return check_launch("msra_overflow_finalize");
}

}  // namespace vod
