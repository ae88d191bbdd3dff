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
//                                  those rows against EVERY location of the group (thread = location, the 8 normalised rows
//                                  broadcast from shared memory, so a reference row is read once for 8 RoI rows), exact top-2
//                                  per row -> ovf_top[pair][g]
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

template <int R>
__global__ void __launch_bounds__(kOvfThreads)
msra_overflow_scan_kernel(const float *__restrict__ roi, const float *__restrict__ ref, const float *__restrict__ roi_norm,
                          const float *__restrict__ ref_norm, const MsraOvf o, int NP, int C, int T, int HW) {
    __shared__ __align__(16) float s_q[R * kOvfMaxC];
    __shared__ int s_prefix[kOvfMaxBins + 1];
    __shared__ int s_warp_sum[kOvfWarps];
    __shared__ float s_mv[kOvfWarps][R][2];
    __shared__ int s_ml[kOvfWarps][R][2];
    __shared__ int s_row[R], s_pos[R];
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

    for (int chunk = blockIdx.x; chunk < total; chunk += gridDim.x) {
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
        // normalised RoI rows, same arithmetic as the re-score kernels: x * (1 / |x|)
        const int c4n = C >> 2;
        for (int i = tid; i < R * c4n; i += kOvfThreads) {
            const int r = i / c4n, cq = i - r * c4n, row = s_row[r];
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (row >= 0) {
                const float qinv = 1.0f / __ldg(roi_norm + row);
                v = ldg_f4(roi + (size_t)row * C + 4 * cq);
                v.x *= qinv; v.y *= qinv; v.z *= qinv; v.w *= qinv;
            }
            *reinterpret_cast<float4 *>(s_q + r * C + 4 * cq) = v;
        }
        __syncthreads();

        float tv0[R], tv1[R]; int tl0[R], tl1[R];
#pragma unroll
        for (int r = 0; r < R; ++r) { tv0[r] = tv1[r] = -INFINITY; tl0[r] = tl1[r] = 0x7fffffff; }
        const int nloc = (HW - g + 3) >> 2;                // locations 4*i + g < HW
        for (int i = tid; i < nloc; i += kOvfThreads) {
            const int l = 4 * i + g;
            const float rinv = 1.0f / __ldg(ref_norm + (size_t)t * HW + l);
            const float *rp = ref + ((size_t)t * HW + l) * C;
            float s[R];
#pragma unroll
            for (int r = 0; r < R; ++r) s[r] = 0.f;
#pragma unroll 4
            for (int c = 0; c < C; c += 4) {
                float4 v = ldg_f4(rp + c);
                v.x *= rinv; v.y *= rinv; v.z *= rinv; v.w *= rinv;
#pragma unroll
                for (int r = 0; r < R; ++r) {
                    const float4 q = *reinterpret_cast<const float4 *>(s_q + r * C + c);
                    s[r] = fmaf(q.x, v.x, s[r]); s[r] = fmaf(q.y, v.y, s[r]);
                    s[r] = fmaf(q.z, v.z, s[r]); s[r] = fmaf(q.w, v.w, s[r]);
                }
            }
#pragma unroll
            for (int r = 0; r < R; ++r) {
                // a thread visits its locations in increasing order: on equal similarity the earlier (smaller) location stays ahead
                if (s[r] > tv0[r]) { tv1[r] = tv0[r]; tl1[r] = tl0[r]; tv0[r] = s[r]; tl0[r] = l; }
                else if (s[r] > tv1[r]) { tv1[r] = s[r]; tl1[r] = l; }
            }
        }
        // warp-level merge of the 32 per-thread top-2 lists of every row, then one warp per row merges the 8 warps' results
#pragma unroll
        for (int r = 0; r < R; ++r) {
            float v[2] = {tv0[r], tv1[r]};
            int l[2] = {tl0[r], tl1[r]};
            topk_warp_merge<2>(v, l, 2, lane);
            if (lane == 0) { s_mv[warp][r][0] = v[0]; s_mv[warp][r][1] = v[1]; s_ml[warp][r][0] = l[0]; s_ml[warp][r][1] = l[1]; }
        }
        __syncthreads();
        if (warp < R) {
            const int r = warp;
            float v[2] = {-INFINITY, -INFINITY};
            int l[2] = {0x7fffffff, 0x7fffffff};
            if (lane < 2 * kOvfWarps) { v[0] = s_mv[lane >> 1][r][lane & 1]; l[0] = s_ml[lane >> 1][r][lane & 1]; }
            topk_warp_merge<2>(v, l, 2, lane);
            if (lane == 0 && s_pos[r] >= 0)
                o.ovf_top[(size_t)s_pos[r] * 4 + g] = make_float4(v[0], __int_as_float(l[0]), v[1], __int_as_float(l[1]));
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
    msra_overflow_scan_kernel<kOvfRows><<<2 * sms, kOvfThreads, 0, st>>>(roi, ref, roi_norm, ref_norm, o, NP, C, T, HW);
    note_launch();
    int rc = check_launch("msra_overflow_scan");
    if (rc) return rc;
    msra_overflow_finalize_kernel<<<2 * sms, kOvfThreads, 0, st>>>(ref, o, out, idx_out, val_out, NP, C, T, HW, k);
    note_launch();
    return check_launch("msra_overflow_finalize");
}

}  // namespace vod
