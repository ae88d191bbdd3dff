// tafa.cu -- TemporalRoIAlign device code that is not the tensor-core GEMM:
//   * tafa_weighted_sum : weighting half of temporal_attentional_feature_aggregation,
//       mmtracking/mmtrack/models/roi_heads/roi_extractors/temporal_roi_align.py:77-97  (HBM-bound)
//   * msra_scan_kernel  : exact fp32 most-similar-location scan (generic shapes / small cases),
//       temporal_roi_align.py:124-181
//   * msra_rescore_kernel: exact fp32 re-score of the tensor-core candidates + top-k + softmax +
//       gather (the tail of the tcgen05 path in msra_gemm.cu)
// All operate on NHWC-style rows ([.., C] contiguous): a warp covers 128 channels per 128-bit load.
#include "common.cuh"
#include "msra.cuh"

namespace vod {

// ------------------------------------------------------------------------------------ TAFA
constexpr int kTafaWarps = 8;
constexpr int kTafaMaxT = 64;

// CTA = (n, head).  VEC=4: lanes own 4 consecutive channels, chunks of 128 channels.
template <int VEC, bool FLAT>
// (3 CTAs per SM = 79 registers without spills: 110 us in logits mode against 105 us for 4 CTAs with 52 bytes of spills)
#ifndef VOD_TAFA_MINB
#define VOD_TAFA_MINB 4
#endif
__global__ void __launch_bounds__(kTafaWarps * 32, VOD_TAFA_MINB)
tafa_kernel(const float *__restrict__ x_all, const float *__restrict__ emb_all, const float *__restrict__ emb_bias,
            float *__restrict__ out, int T1, int N, int P, int C, int hs, float scale, int use_attn, int out_layout,
            const float *__restrict__ logit_parts, int nparts) {
    extern __shared__ __align__(16) float tile[];          // [hs][P] (layout 0)
    __shared__ float s_w[kTafaWarps][kTafaMaxT];
    constexpr int CH = 32 * VEC;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    // layout 0 (NCHW output, transposed through shared memory): CTA = (n, head), its warps stride over the P positions.
    // layout 1 (channels_last output): nothing ties a CTA to one RoI, so the grid is flat over (n, p, head) warp tasks
    // -- consecutive warps read consecutive 512-byte head slices, and 7350 small CTAs leave no wave tail (the
    // (300, 4) grid ran 2.03 waves of 592 resident CTAs: a third wave for 16 CTAs).
    int n, head, p0, p1, pstep;
    if (FLAT) {
        const int groups = ceil_div(C, hs);
        const long task = (long)blockIdx.x * kTafaWarps + warp;
        if (task >= (long)N * P * groups) return;
        head = (int)(task % groups);
        const long np = task / groups;
        n = (int)(np / P); p0 = (int)(np % P); p1 = p0 + 1; pstep = 1;
    } else {
        n = blockIdx.x; head = blockIdx.y; p0 = warp; p1 = P; pstep = kTafaWarps;
    }
    const int c_head = head * hs;
    const int hs_eff = min(hs, C - c_head);
    const size_t frame_stride = (size_t)N * P * C;
    const int nch = ceil_div(hs_eff, CH);

    for (int p = p0; p < p1; p += pstep) {
        const size_t base = ((size_t)n * P + p) * C + c_head + lane * VEC;
        float *w = s_w[warp];
        if (use_attn && logit_parts) {
            // logits come from the key-projected contraction (tafa_keyproj.cu) as per-channel-chunk partial sums
            // [nparts][N][P][groups][T1]; sum the chunks in a fixed order
            const int groups = ceil_div(C, hs);
            const size_t part_stride = (size_t)N * P * groups * T1;
            const float *src = logit_parts + (((size_t)n * P + p) * groups + head) * T1;
            for (int t = lane; t < T1; t += 32) {
                float s = 0.f;
                for (int k = 0; k < nparts; ++k) s += __ldg(src + (size_t)k * part_stride + t);
                w[t] = s * scale;
            }
            __syncwarp();
        } else if (use_attn) {
            // logits_t = <emb[t] + b, emb[0] + b>_head * scale   (b = the embed conv's bias, folded in here so
            // the conv can run bias-free and the [T1,N,P,C] embedding is not rewritten by a bias-add pass).
            // Frames are processed 8 at a time: all loads of a batch are issued before the first reduction.
            constexpr int TB = 8;
            for (int t0 = 0; t0 < T1; t0 += TB) {
                float d[TB];
#pragma unroll
                for (int u = 0; u < TB; ++u) d[u] = 0.f;
                for (int ch = 0; ch < nch; ++ch) {
                    if (ch * CH + lane * VEC < hs_eff) {
                        if (VEC == 4) {
                            float4 a = ldg_f4(emb_all + base + ch * CH);
                            float4 b[TB];
#pragma unroll
                            for (int u = 0; u < TB; ++u)
                                b[u] = (t0 + u < T1) ? ldg_f4(emb_all + (size_t)(t0 + u) * frame_stride + base + ch * CH)
                                                     : make_float4(0.f, 0.f, 0.f, 0.f);
                            float4 bb = make_float4(0.f, 0.f, 0.f, 0.f);
                            if (emb_bias) bb = ldg_f4(emb_bias + c_head + lane * VEC + ch * CH);
                            a.x += bb.x; a.y += bb.y; a.z += bb.z; a.w += bb.w;
#pragma unroll
                            for (int u = 0; u < TB; ++u) {
                                d[u] = fmaf(a.x, b[u].x + bb.x, d[u]); d[u] = fmaf(a.y, b[u].y + bb.y, d[u]);
                                d[u] = fmaf(a.z, b[u].z + bb.z, d[u]); d[u] = fmaf(a.w, b[u].w + bb.w, d[u]);
                            }
                        } else {
                            float a = __ldg(emb_all + base + ch * CH);
                            const float bb = emb_bias ? __ldg(emb_bias + c_head + lane + ch * CH) : 0.f;
                            a += bb;
#pragma unroll
                            for (int u = 0; u < TB; ++u)
                                if (t0 + u < T1)
                                    d[u] = fmaf(a, __ldg(emb_all + (size_t)(t0 + u) * frame_stride + base + ch * CH) + bb, d[u]);
                        }
                    }
                }
#pragma unroll
                for (int u = 0; u < TB; ++u) d[u] = warp_sum(d[u]);
                if (lane == 0) {
#pragma unroll
                    for (int u = 0; u < TB; ++u) if (t0 + u < T1) w[t0 + u] = d[u] * scale;
                }
            }
        }
        if (use_attn) {
            // softmax over the T1 <= 64 frames, lane t owns frames t and t + 32: one expf per lane and two warp reductions
            // (every lane looping over all frames cost ~3 T1 expf per lane: ncu showed the kernel 69 % issue-bound on them)
            __syncwarp();
            const float v0 = lane < T1 ? w[lane] : -INFINITY, v1 = lane + 32 < T1 ? w[lane + 32] : -INFINITY;
            const float m = warp_max(fmaxf(v0, v1));
            const float e0 = lane < T1 ? expf(v0 - m) : 0.f, e1 = lane + 32 < T1 ? expf(v1 - m) : 0.f;
            const float inv = 1.f / warp_sum(e0 + e1);
            __syncwarp();
            if (lane < T1) w[lane] = e0 * inv;
            if (lane + 32 < T1) w[lane + 32] = e1 * inv;
            __syncwarp();
        }
        const float uniform = 1.0f / (float)T1;
        for (int ch = 0; ch < nch; ++ch) {
            const bool on = ch * CH + lane * VEC < hs_eff;
            float acc[VEC];
#pragma unroll
            for (int q = 0; q < VEC; ++q) acc[q] = 0.f;
            if (on) {
                constexpr int TB2 = 8;
                for (int t0 = 0; t0 < T1; t0 += TB2) {
                    if (VEC == 4) {
                        float4 v[TB2];
#pragma unroll
                        for (int u = 0; u < TB2; ++u)
                            v[u] = (t0 + u < T1) ? ldg_f4(x_all + (size_t)(t0 + u) * frame_stride + base + ch * CH)
                                                 : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
                        for (int u = 0; u < TB2; ++u) {
                            const float wt = (t0 + u < T1) ? (use_attn ? w[t0 + u] : uniform) : 0.f;
                            acc[0] = fmaf(wt, v[u].x, acc[0]); acc[1 % VEC] = fmaf(wt, v[u].y, acc[1 % VEC]);
                            acc[2 % VEC] = fmaf(wt, v[u].z, acc[2 % VEC]); acc[3 % VEC] = fmaf(wt, v[u].w, acc[3 % VEC]);
                        }
                    } else {
#pragma unroll
                        for (int u = 0; u < TB2; ++u)
                            if (t0 + u < T1)
                                acc[0] = fmaf(use_attn ? w[t0 + u] : uniform,
                                              __ldg(x_all + (size_t)(t0 + u) * frame_stride + base + ch * CH), acc[0]);
                    }
                }
                if (!FLAT) {
#pragma unroll
                    for (int q = 0; q < VEC; ++q) tile[(ch * CH + lane * VEC + q) * P + p] = acc[q];
                } else {
                    float *dst = out + base + ch * CH;
                    if (VEC == 4) stg_cs_f4(dst, make_float4(acc[0], acc[1 % VEC], acc[2 % VEC], acc[3 % VEC]));
                    else dst[0] = acc[0];
                }
            }
        }
        __syncwarp();
    }
    if (FLAT) return;
    __syncthreads();
    const int total = hs_eff * P;
    float *dst = out + ((size_t)n * C + c_head) * P;
    if (VEC == 4 && (total & 3) == 0 && ((((size_t)n * C + c_head) * P) & 3) == 0) {
        const float4 *s4 = reinterpret_cast<const float4 *>(tile);
        for (int t = threadIdx.x; t < total / 4; t += kTafaWarps * 32) stg_cs_f4(dst + 4 * t, s4[t]);
    } else {
        for (int t = threadIdx.x; t < total; t += kTafaWarps * 32) dst[t] = tile[t];
    }
}

// ------------------------------------------------------------------------------------ MSRA
constexpr int kScanWarps = 8;

// warp = (row, t) task; lane scans locations lane, lane+32, ...; exact fp32:
//   sim = sum_c (roi[c] / |roi|) * (ref[c] / |ref|)      (temporal_roi_align.py:127-142)
__global__ void __launch_bounds__(kScanWarps * 32)
msra_scan_kernel(const float *__restrict__ roi, const float *__restrict__ ref, const float *__restrict__ roi_norm,
                 const float *__restrict__ ref_norm, float *__restrict__ out, int *__restrict__ idx_out,
                 float *__restrict__ val_out, int NP, int C, int T, int HW, int k) {
    extern __shared__ float s_roi[];  // [kScanWarps][C]
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const long task = (long)blockIdx.x * kScanWarps + warp;
    if (task >= (long)NP * T) return;
    const int row = (int)(task / T), t = (int)(task % T);
    float *q = s_roi + (size_t)warp * C;
    const float qn = roi_norm[row];
    for (int c = lane; c < C; c += 32) q[c] = __fdiv_rn(__ldg(roi + (size_t)row * C + c), qn);
    __syncwarp();
    const float *ref_t = ref + (size_t)t * HW * C;
    float val[kMsraMaxK]; int loc[kMsraMaxK];
#pragma unroll
    for (int i = 0; i < kMsraMaxK; ++i) { val[i] = -INFINITY; loc[i] = 0x7fffffff; }
    for (int l = lane; l < HW; l += 32) {
        const float rn = ref_norm[(size_t)t * HW + l];
        const float *r = ref_t + (size_t)l * C;
        float s = 0.f;
        for (int c = 0; c < C; ++c) s = fmaf(q[c], __fdiv_rn(__ldg(r + c), rn), s);
        topk_insert<kMsraMaxK>(val, loc, s, l);
    }
    topk_warp_merge<kMsraMaxK>(val, loc, k, lane);
    msra_emit<kMsraMaxK>(ref_t, out + ((size_t)t * NP + row) * C, idx_out ? idx_out + ((size_t)row * T + t) * k : nullptr,
                         val_out ? val_out + ((size_t)row * T + t) * k : nullptr, val, loc, k, C, lane);
}

// Appends the (row, frame) to the overflow work lists (msra.cuh): `mask` bit g = the candidate list of location group g was
// full AND its 4th key lies within the re-score margin of the k-th best key, so a location that fell off that list could still
// belong to the exact top-k.  Called by one lane; (v0, l0, v1, l1) is the exact top-2 of the re-scored candidates.
__device__ __forceinline__ void msra_flag_overflow(const MsraOvf &o, int NP, int row, int t, unsigned mask, float v0, int l0,
                                                   float v1, int l1) {
    const int pos = atomicAdd(o.ctrl, 1);
    o.pair_list[pos] = make_int4(row, t, (int)mask, 0);
    o.pair_top[pos] = make_float4(v0, __int_as_float(l0), v1, __int_as_float(l1));
    for (int g = 0; g < 4; ++g)
        if (mask >> g & 1u) {
            const int bin = 4 * t + g;
            const int j = atomicAdd(o.ctrl + 1 + bin, 1);
            o.bin_list[(size_t)bin * NP + j] = make_int2(row, pos);
        }
}

// lanes 3, 7, 11, 15 of a ballot (the 4th key of each group's list) -> 4-bit group mask
__device__ __forceinline__ unsigned msra_group_mask(unsigned ballot) {
    return (ballot >> 3 & 1u) | (ballot >> 6 & 2u) | (ballot >> 9 & 4u) | (ballot >> 12 & 8u);
}

// Tail of the tensor-core path.  warp = one RoI row (all T frames): the row is normalised once; for every
// frame the 16 packed candidate keys of msra_gemm_topk_kernel are decoded, candidates whose bf16-GEMM
// similarity lies within kMsraMargin of the 2nd best are re-scored in exact fp32
//   sim = sum_c (roi[c] / |roi|) * (ref[c] / |ref|)      (temporal_roi_align.py:127-142)
// and the exact top-k feeds the softmax + gather.  Typically 2-3 of the 16 candidates are re-scored.
template <int KM>
__global__ void __launch_bounds__(kScanWarps * 32)
msra_rescore_kernel(const float *__restrict__ roi, const float *__restrict__ ref, const float *__restrict__ roi_norm,
                    const float *__restrict__ ref_norm, const uint32_t *__restrict__ cand, int KC,
                    float *__restrict__ out, int *__restrict__ idx_out, float *__restrict__ val_out, int NP, int C,
                    int T, int HW, int k, const MsraOvf ovf) {
    extern __shared__ float s_roi[];  // [kScanWarps][C]
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int row = blockIdx.x * kScanWarps + warp;
    if (row >= NP) return;
    float *q = s_roi + (size_t)warp * C;
    // x * (1/|x|) instead of x / |x|: an IEEE division per element takes the slow path whenever a lane holds an
    // exact zero (half of all post-ReLU features); the two differ by <= 1 ulp per element, far below the
    // spacing of neighbouring similarities.
    const float qinv = 1.0f / roi_norm[row];
    for (int c = lane; c < C; c += 32) q[c] = __ldg(roi + (size_t)row * C + c) * qinv;
    __syncwarp();
    for (int t = 0; t < T; ++t) {
        const float *ref_t = ref + (size_t)t * HW * C;
        // decode this frame's candidate keys (lane j < KC holds candidate j)
        uint32_t key = lane < KC ? cand[((size_t)row * T + t) * KC + lane] : 0u;
        // key = round((1.5 + sim) * 2^11) << 12 | location (msra_gemm.cu); a NaN similarity has the value field 0xFFFFF
        const float approx = !key ? -INFINITY : key >= 0xFFFFF000u ? __int_as_float(0x7fc00000)
                                              : (float)(key >> 12) * (1.0f / 2048.0f) - 1.5f;
        // k-th largest approximate similarity (k <= 4): repeatedly remove the maximum
        float kth = approx;
        {
            float mine = approx;
            bool removed = false;
            for (int r = 0; r < k; ++r) {
                float cur = removed ? -INFINITY : mine;
                float m = warp_max(cur);
                kth = m;
                // remove exactly one holder of the maximum (lowest lane)
                unsigned holders = __ballot_sync(0xffffffffu, !removed && cur == m);
                if (holders && lane == __ffs(holders) - 1) removed = true;
            }
        }
        // NaN approximations (zero-norm vectors) compare false: keep them as candidates so the NaN propagates
        const bool want = key != 0u && !(approx < kth - kMsraMargin);
        unsigned todo = __ballot_sync(0xffffffffu, want);
        // a full list (KC == 16: 4 groups x 4 keys, the 4th key in lanes 3, 7, 11, 15) whose last key is within the margin
        const unsigned ovf_mask = (ovf.ctrl && KC == kMsraCand) ? msra_group_mask(__ballot_sync(0xffffffffu, want && (lane & 3) == 3)) : 0u;
        float val[KM]; int loc[KM];
#pragma unroll
        for (int i = 0; i < KM; ++i) { val[i] = -INFINITY; loc[i] = 0x7fffffff; }
        bool any_nan = false;
        while (todo) {
            const int j = __ffs(todo) - 1;
            todo &= todo - 1;
            const int l = (int)(__shfl_sync(0xffffffffu, key, j) & 0xFFFu);
            if (l >= HW) continue;
            const float rinv = 1.0f / ref_norm[(size_t)t * HW + l];
            const float *r = ref_t + (size_t)l * C;
            float s = 0.f;
            if ((C & 3) == 0) {
                for (int c = lane * 4; c < C; c += 128) {
                    const float4 v = ldg_f4(r + c);
                    const float4 qq = *reinterpret_cast<const float4 *>(q + c);
                    s = fmaf(qq.x, v.x * rinv, s); s = fmaf(qq.y, v.y * rinv, s);
                    s = fmaf(qq.z, v.z * rinv, s); s = fmaf(qq.w, v.w * rinv, s);
                }
            } else {
                for (int c = lane; c < C; c += 32) s = fmaf(q[c], __ldg(r + c) * rinv, s);
            }
            s = warp_sum(s);
            if (s != s) any_nan = true;            // torch.topk ranks NaN first: the reference row becomes NaN
            topk_insert<KM>(val, loc, s, l);  // identical on all lanes
        }
        if (any_nan) loc[0] = 0x7fffffff;          // msra_emit writes a NaN row
        if (ovf_mask && !any_nan && lane == 0)       // (a NaN row stays NaN whatever a re-scan would find)
            msra_flag_overflow(ovf, NP, row, t, ovf_mask, val[0], loc[0], KM > 1 ? val[1] : -INFINITY, KM > 1 ? loc[1] : 0x7fffffff);
        msra_emit<KM>(ref_t, out + ((size_t)t * NP + row) * C, idx_out ? idx_out + ((size_t)row * T + t) * k : nullptr,
                             val_out ? val_out + ((size_t)row * T + t) * k : nullptr, val, loc, k, C, lane);
    }
}

// Lean form of the re-score for the reference's configuration (k <= 2, C = 128 * NQ <= 512, 16 candidates):
//   * warp task = (RoI row, chunk of kRfFrames frames): 3x more, shorter tasks than warp-per-row -> no wave tail;
//   * the normalised RoI row lives in registers (lane owns channels lane*4 + 128*i), no shared memory;
//   * the k-th best approximate similarity is found on the integer value field of the keys with two warp REDUX
//     (the field is monotone in the similarity), the margin is 6 quanta of 2^-11 (>= kMsraMargin);
//   * same arithmetic as the generic kernel for the exact similarity, the softmax and the weighted sum.
// The generic kernel above needed ~1300 warp instructions per (row, frame); this one ~300.
#ifndef VOD_RF_FRAMES
#define VOD_RF_FRAMES 5
#endif
#ifndef VOD_RF_WARPS
#define VOD_RF_WARPS 4
#endif
constexpr int kRfFrames = VOD_RF_FRAMES;
constexpr int kRfWarps = VOD_RF_WARPS;
template <int NQ, bool TWO>
// (register budget, round 2: the compiler's own 80 registers = 6 CTAs of 4 warps per SM 205 us; 8 CTAs (64 registers, no spills)
// 188 us; 10 CTAs (48 registers, spills) 241 us)
#ifndef VOD_RF_MINB
#define VOD_RF_MINB 8
#endif
__global__ void __launch_bounds__(kRfWarps * 32, VOD_RF_MINB)
msra_rescore_fast_kernel(const float *__restrict__ roi, const float *__restrict__ ref, const float *__restrict__ roi_norm,
                         const float *__restrict__ ref_norm, const uint32_t *__restrict__ cand,
                         float *__restrict__ out, int *__restrict__ idx_out, float *__restrict__ val_out, int NP,
                         int T, int HW, int k, int nchunks, const MsraOvf ovf) {
    constexpr int C = 128 * NQ;
    __shared__ float4 s_rows[kRfWarps][2 * NQ * 32];      // per warp: two parked rows, [slot][i][lane]
    const int lane = threadIdx.x & 31;
    float4 *my_rows = s_rows[threadIdx.x >> 5];
    const long task = (long)blockIdx.x * kRfWarps + (threadIdx.x >> 5);
    if (task >= (long)NP * nchunks) return;
    const int row = (int)(task / nchunks), t_begin = (int)(task % nchunks) * kRfFrames, t_end = min(T, t_begin + kRfFrames);
    const float qinv = 1.0f / roi_norm[row];
    float4 q[NQ];
#pragma unroll
    for (int i = 0; i < NQ; ++i) {
        q[i] = ldg_f4(roi + (size_t)row * C + lane * 4 + 128 * i);
        q[i].x *= qinv; q[i].y *= qinv; q[i].z *= qinv; q[i].w *= qinv;
    }
    constexpr uint32_t kMarginQ = 6;   // ceil(kMsraMargin * 2048) = 6 quanta
    static_assert(kMsraMargin * 2048.0f <= 6.0f, "margin quanta");
    uint32_t key = lane < kMsraCand ? __ldg(cand + ((size_t)row * T + t_begin) * kMsraCand + lane) : 0u;
    for (int t = t_begin; t < t_end; ++t) {
        const uint32_t next = (t + 1 < t_end && lane < kMsraCand) ? __ldg(cand + ((size_t)row * T + t + 1) * kMsraCand + lane) : 0u;
        const float *ref_t = ref + (size_t)t * HW * C;
        // k-th largest value field (k <= 2); a NaN similarity has the largest field and is always re-scored
        const uint32_t vf = key >> 12;
        uint32_t kth = __reduce_max_sync(0xffffffffu, vf);
        if (k > 1) {
            const unsigned holders = __ballot_sync(0xffffffffu, vf == kth);
            const uint32_t rest = (lane == __ffs(holders) - 1) ? 0u : vf;
            kth = __reduce_max_sync(0xffffffffu, rest);
        }
        const bool want = key != 0u && vf + kMarginQ >= kth;
        unsigned todo = __ballot_sync(0xffffffffu, want);
        // overflow test per location group: its list is full (4th key, lanes 3/7/11/15, non-empty) and that key is within the margin
        const unsigned ovf_mask = ovf.ctrl ? msra_group_mask(todo & 0x8888u) : 0u;
        float v0 = -INFINITY, v1 = -INFINITY; int l0 = 0x7fffffff, l1 = 0x7fffffff;
        // The RAW reference rows of the best / second best location so far are parked in this warp's two shared-memory slots,
        // so the gather below does not fetch the two rows a second time (the kernel is L2-bandwidth bound: ~2 GB of L2 -> SM
        // traffic for 555 MB of algorithmic bytes with the second fetch).  Shared memory, not registers: parking them in
        // registers cost 80 registers and, at 12 resident warps per SM, ran 1.6x SLOWER than re-fetching.
        int best_slot = 0;            // slot holding the best row; the other one holds the second best
        bool any_nan = false;
        auto insert = [&](float s, int l, const float4 (&row)[NQ]) {
            if (s != s) any_nan = true;               // torch.topk ranks NaN first: the reference row becomes NaN
            if (s > v0 || (s == v0 && l < l0)) {
                v1 = v0; l1 = l0; v0 = s; l0 = l;
                best_slot ^= 1;                       // the old best becomes the second best; its slot-mate is overwritten
#pragma unroll
                for (int i = 0; i < NQ; ++i) my_rows[(best_slot * NQ + i) * 32 + lane] = row[i];
            } else if (s > v1 || (s == v1 && l < l1)) {
                v1 = s; l1 = l;
#pragma unroll
                for (int i = 0; i < NQ; ++i) my_rows[((best_slot ^ 1) * NQ + i) * 32 + lane] = row[i];
            }
        };
        // candidates are re-scored two at a time: both rows (and both norms) are in flight before the first reduction
        while (todo) {
            const int j0 = __ffs(todo) - 1;
            todo &= todo - 1;
            const bool two = TWO && todo != 0u;
            const int j1 = two ? __ffs(todo) - 1 : j0;
            if (two) todo &= todo - 1;
            const int la = min((int)(__shfl_sync(0xffffffffu, key, j0) & 0xFFFu), HW - 1);
            const int lb = min((int)(__shfl_sync(0xffffffffu, key, j1) & 0xFFFu), HW - 1);
            const float *pa = ref_t + (size_t)la * C + lane * 4, *pb = ref_t + (size_t)lb * C + lane * 4;
            float4 xa[NQ], xb[NQ];
#pragma unroll
            for (int i = 0; i < NQ; ++i) xa[i] = ldg_f4(pa + 128 * i);
            if (two) {
#pragma unroll
                for (int i = 0; i < NQ; ++i) xb[i] = ldg_f4(pb + 128 * i);
            }
            const float rinv_a = 1.0f / __ldg(ref_norm + (size_t)t * HW + la);
            const float rinv_b = 1.0f / __ldg(ref_norm + (size_t)t * HW + lb);
            float sa = 0.f, sb = 0.f;
#pragma unroll
            for (int i = 0; i < NQ; ++i) {
                sa = fmaf(q[i].x, xa[i].x * rinv_a, sa); sa = fmaf(q[i].y, xa[i].y * rinv_a, sa);
                sa = fmaf(q[i].z, xa[i].z * rinv_a, sa); sa = fmaf(q[i].w, xa[i].w * rinv_a, sa);
            }
            if (two) {
#pragma unroll
                for (int i = 0; i < NQ; ++i) {
                    sb = fmaf(q[i].x, xb[i].x * rinv_b, sb); sb = fmaf(q[i].y, xb[i].y * rinv_b, sb);
                    sb = fmaf(q[i].z, xb[i].z * rinv_b, sb); sb = fmaf(q[i].w, xb[i].w * rinv_b, sb);
                }
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {       // two warp sums, interleaved (same pairing order as warp_sum)
                sa += __shfl_xor_sync(0xffffffffu, sa, o);
                sb += __shfl_xor_sync(0xffffffffu, sb, o);
            }
            insert(sa, la, xa);
            if (two) insert(sb, lb, xb);
        }
        float *out_row = out + ((size_t)t * NP + row) * C + lane * 4;
        const bool bad = any_nan || l0 == 0x7fffffff || (k > 1 && l1 == 0x7fffffff);
        if (ovf_mask && !any_nan) {
            // Second look with the EXACT k-th best similarity now known: a location that fell off group g's full list has a key
            // value <= the list's 4th key, so its exact similarity is below value(4th key) + kMsraKeyErr; the group needs the
            // exact re-scan only if that bound reaches the k-th best.  (Cuts the flags on iid noise by ~2-3x.)
            const float vk = k > 1 ? v1 : v0;
            const float key_val = (float)(key >> 12) * (1.0f / 2048.0f) - 1.5f;
            const unsigned still = msra_group_mask(__ballot_sync(0xffffffffu, want && (lane & 3) == 3 && !(key_val + kMsraKeyErr < vk)));
            if (still && lane == 0) msra_flag_overflow(ovf, NP, row, t, still & ovf_mask, v0, l0, v1, l1);
        }
        float w0 = 1.f, w1 = 0.f;
        if (k > 1) {
            const float e1 = expf(v1 - v0), sum = 1.0f + e1;   // softmax over the k values (temporal_roi_align.py:155)
            w0 = 1.0f / sum; w1 = e1 / sum;
        }
        if (lane < k) {
            const float qnan = __int_as_float(0x7fc00000);
            if (idx_out) idx_out[((size_t)row * T + t) * k + lane] = bad ? 0 : (lane ? l1 : l0);
            if (val_out) val_out[((size_t)row * T + t) * k + lane] = bad ? qnan : (lane ? v1 : v0);
        }
        if (bad) {
            const float qnan = __int_as_float(0x7fc00000);
#pragma unroll
            for (int i = 0; i < NQ; ++i) stg_cs_f4(out_row + 128 * i, make_float4(qnan, qnan, qnan, qnan));
        } else {
#pragma unroll
            for (int i = 0; i < NQ; ++i) {
                // topk_feats * topk_weights summed over k (temporal_roi_align.py:170-172), same order as msra_emit
                const float4 ra = my_rows[(best_slot * NQ + i) * 32 + lane];
                float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
                acc.x += ra.x * w0; acc.y += ra.y * w0; acc.z += ra.z * w0; acc.w += ra.w * w0;
                if (k > 1) {
                    const float4 rb = my_rows[((best_slot ^ 1) * NQ + i) * 32 + lane];
                    acc.x += rb.x * w1; acc.y += rb.y * w1; acc.z += rb.z * w1; acc.w += rb.w * w1;
                }
                stg_cs_f4(out_row + 128 * i, acc);
            }
        }
        key = next;
    }
}

int msra_launch_scan(const float *roi, const float *ref, const float *roi_norm, const float *ref_norm, float *out,
                     int *idx_out, float *val_out, int NP, int C, int T, int HW, int k, cudaStream_t st) {
    size_t smem = sizeof(float) * kScanWarps * C;
    if (smem > 40 * 1024) cudaFuncSetAttribute(msra_scan_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    long tasks = (long)NP * T;
    msra_scan_kernel<<<(unsigned)ceil_div(tasks, (long)kScanWarps), kScanWarps * 32, smem, st>>>(
        roi, ref, roi_norm, ref_norm, out, idx_out, val_out, NP, C, T, HW, k); note_launch();
    return check_launch("msra_scan");
}

int msra_launch_rescore(const float *roi, const float *ref, const float *roi_norm, const float *ref_norm,
                        const uint32_t *cand, int KC, float *out, int *idx_out, float *val_out, int NP, int C, int T,
                        int HW, int k, const MsraOvf *ovf_in, cudaStream_t st) {
    MsraOvf ovf;
    memset(&ovf, 0, sizeof(ovf));
    if (ovf_in) ovf = *ovf_in;
    if (k <= 2 && KC == kMsraCand && (C & 127) == 0 && C <= 512 && HW <= 4096) {
        const int nchunks = ceil_div(T, kRfFrames);
        const unsigned g = (unsigned)ceil_div((long)NP * nchunks, (long)kRfWarps);
        auto go = [&](auto kern) {
            kern<<<g, kRfWarps * 32, 0, st>>>(roi, ref, roi_norm, ref_norm, cand, out, idx_out, val_out, NP, T, HW, k, nchunks, ovf);
        };
        // (two candidates in flight per pass -- msra_rescore_fast_kernel<NQ, true> -- measured 206 us against 197 us: the extra
        // 20 registers cost more resident warps than the second row in flight buys)
        switch (C >> 7) {
            case 1: go(msra_rescore_fast_kernel<1, false>); break;
            case 2: go(msra_rescore_fast_kernel<2, false>); break;
            case 3: go(msra_rescore_fast_kernel<3, false>); break;
            default: go(msra_rescore_fast_kernel<4, false>); break;
        }
        note_launch();
        return check_launch("msra_rescore_fast");
    }
    // generic shapes (C % 128 != 0): still k <= 2, the only k the candidate pass is used for (msra_gemm.cu)
    if (k > 2) return fail(VOD_E_UNSUPPORTED, "msra_rescore: k=%d > 2", k);
    size_t smem = sizeof(float) * kScanWarps * C;
    const unsigned grid = (unsigned)ceil_div(NP, kScanWarps);
    if (smem > 40 * 1024) cudaFuncSetAttribute(msra_rescore_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    msra_rescore_kernel<2><<<grid, kScanWarps * 32, smem, st>>>(roi, ref, roi_norm, ref_norm, cand, KC, out, idx_out, val_out, NP, C,
                                                              T, HW, k, ovf);
    note_launch();
    return check_launch("msra_rescore");
}

}  // namespace vod

using namespace vod;

static int tafa_launch(const float *x_all, const float *emb_all, const float *emb_bias, const float *logit_parts, int nparts,
                       float *out, int T1, int N, int P, int C, int heads, int out_layout, vod_stream_t stream);

extern "C" int vod_tafa_weighted_sum(const float *x_all, const float *emb_all, const float *emb_bias, float *out, int T1,
                                     int N, int P, int C, int heads, int out_layout, vod_stream_t stream) {
    return tafa_launch(x_all, emb_all, emb_bias, nullptr, 0, out, T1, N, P, C, heads, out_layout, stream);
}

extern "C" int vod_tafa_weighted_sum_logits(const float *x_all, const float *logit_parts, int nparts, float *out, int T1,
                                            int N, int P, int C, int heads, int out_layout, vod_stream_t stream) {
    VOD_REQUIRE(heads > 0 && nparts > 0 && (logit_parts || N == 0), "vod_tafa_weighted_sum_logits: heads, nparts, logit_parts required");
    return tafa_launch(x_all, nullptr, nullptr, logit_parts, nparts, out, T1, N, P, C, heads, out_layout, stream);
}

static int tafa_launch(const float *x_all, const float *emb_all, const float *emb_bias, const float *logit_parts, int nparts,
                       float *out, int T1, int N, int P, int C, int heads, int out_layout, vod_stream_t stream) {
    if (N == 0) return VOD_OK;
    VOD_REQUIRE(x_all && out, "vod_tafa_weighted_sum: null pointer");
    VOD_REQUIRE(T1 > 0 && T1 <= kTafaMaxT, "vod_tafa_weighted_sum: T1=%d not in [1,%d]", T1, kTafaMaxT);
    VOD_REQUIRE(N > 0 && P > 0 && C > 0, "vod_tafa_weighted_sum: bad dims");
    VOD_REQUIRE(out_layout == 0 || out_layout == 1, "vod_tafa_weighted_sum: out_layout");
    const int use_attn = heads > 0;
    VOD_REQUIRE(!use_attn || emb_all || logit_parts, "vod_tafa_weighted_sum: emb_all required when heads > 0");
    VOD_REQUIRE(!use_attn || C % heads == 0, "vod_tafa_weighted_sum: C=%d not divisible by heads=%d", C, heads);
    int hs = use_attn ? C / heads : min(C, 128);
    int groups = use_attn ? heads : ceil_div(C, hs);
    // reference: / float(c_embed / num_attention_blocks) ** 0.5 (temporal_roi_align.py:86-87)
    float scale = use_attn ? (float)(1.0 / sqrt((double)C / (double)heads)) : 1.f;
    size_t smem = out_layout == 0 ? sizeof(float) * (size_t)hs * P : 0;
    VOD_REQUIRE(smem <= 200 * 1024, "vod_tafa_weighted_sum: head size %d x %d bins too large", hs, P);
    const bool vec4 = (hs % 4 == 0) && (C % 4 == 0) && ((reinterpret_cast<uintptr_t>(x_all) & 15) == 0) &&
                      (!emb_all || (reinterpret_cast<uintptr_t>(emb_all) & 15) == 0) &&
                      (!emb_bias || (reinterpret_cast<uintptr_t>(emb_bias) & 15) == 0) &&
                      ((reinterpret_cast<uintptr_t>(out) & 15) == 0);
    dim3 grid(N, groups);
    if (out_layout != 0) grid = dim3((unsigned)ceil_div((long)N * P * groups, (long)kTafaWarps), 1);
    auto launch = [&](auto kern) {
        if (smem > 40 * 1024) cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        kern<<<grid, kTafaWarps * 32, smem, as_stream(stream)>>>(x_all, emb_all, emb_bias, out, T1, N, P, C, hs, scale, use_attn, out_layout,
                                                                       logit_parts, nparts);
        note_launch();
    };
    if (out_layout != 0) { if (vec4) launch(tafa_kernel<4, true>); else launch(tafa_kernel<1, true>); }
    else { if (vec4) launch(tafa_kernel<4, false>); else launch(tafa_kernel<1, false>); }
    return check_launch("vod_tafa_weighted_sum");
}
