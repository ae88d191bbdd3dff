// nms.cu -- (5) batched bitmask NMS, entirely on the device (no host round trip).
//
// Semantics restated from mmcv-full 1.2.x nms / batched_nms (un-vendored) as
// called at mmdetection/mmdet/core/post_processing/bbox_nms.py:84 and
// mmdetection/mmdet/models/dense_heads/rpn_head.py:233-235.
//
// Pipeline per call (all images of the call in the same launches; image = blockIdx.z / .y):
//   1. nms_rank_kernel     stable descending rank of every box by counting (score desc, index asc)
//                          + per-image max coordinate (for the class-offset trick)
//   2. nms_scatter_kernel  sorted boxes (+ label * (max + 1) offset, un-fused fp32 adds), ids, labels
//   3. nms_mask_kernel     64x64-tile IoU > thr bitmask, upper triangle only
//   4. nms_sweep_kernel    greedy sweep: one CTA per image, 64-box blocks resolved by a warp in
//                          registers, kept rows OR-ed into a shared-memory "removed" bitmap
// IoU arithmetic uses __fmul_rn/__fadd_rn/__fdiv_rn so nothing is contracted into an FMA: kept
// indices are bit-exact against an un-contracted CPU evaluation of the same formula.
#include "common.cuh"

namespace vod {

constexpr int kMaxImages = 256;   // images (segments) per call; the table travels as a kernel parameter (1 KB)
struct SegTable {
    int n_images;
    int off[kMaxImages + 1];
};

// Optional extensions used by the fixed-shape (sync-free, graph-capturable) multiclass path.
struct NmsOpt {
    const int *n_valid;      // [n_images] device or null: candidates whose score is -inf are invalid and the
                             // image's effective box count is n_valid[img] (they sort last)
    int split_thr;           // mode 3 (auto): coordinate-offset trick below split_thr valid boxes, per-class at/above
    float *dets_out;         // [n_images][keep_stride][5] or null: (x1,y1,x2,y2,score) of the survivors
    int64_t *labels_out;     // [n_images][keep_stride] or null
    const float *boxes, *scores;
    const int64_t *labels;
    int keep_stride;
};
__device__ __forceinline__ int eff_n(const NmsOpt &o, int img, int n) { return o.n_valid ? min(n, o.n_valid[img]) : n; }
__device__ __forceinline__ int eff_mode(const NmsOpt &o, int mode, int n_eff) {
    return mode == 3 ? (n_eff < o.split_thr ? 1 : 2) : mode;
}

struct NmsWorkspace {
    int *rank;            // [n_total]
    unsigned *segmax;     // [kMaxImages] order-preserving encoding of the max coordinate
    float4 *sboxes;       // [n_total] sorted boxes (offset applied in mode 1)
    int *sidx;            // [n_total] original index (relative to the image) of sorted position
    int *slab;            // [n_total] label of sorted position
    unsigned long long *mask;  // [n_total * words]
    size_t bytes;
};

static NmsWorkspace carve(void *ws, int n_total, int words) {
    NmsWorkspace w;
    char *p = reinterpret_cast<char *>(ws);
    size_t o = 0;
    w.rank = reinterpret_cast<int *>(p + o);           o = align_up(o + sizeof(int) * (size_t)n_total, 256);
    w.segmax = reinterpret_cast<unsigned *>(p + o);    o = align_up(o + sizeof(unsigned) * kMaxImages, 256);
    w.sboxes = reinterpret_cast<float4 *>(p + o);      o = align_up(o + sizeof(float4) * (size_t)n_total, 256);
    w.sidx = reinterpret_cast<int *>(p + o);           o = align_up(o + sizeof(int) * (size_t)n_total, 256);
    w.slab = reinterpret_cast<int *>(p + o);           o = align_up(o + sizeof(int) * (size_t)n_total, 256);
    w.mask = reinterpret_cast<unsigned long long *>(p + o);
    o = align_up(o + sizeof(unsigned long long) * (size_t)n_total * (size_t)words, 256);
    w.bytes = o;
    return w;
}

__device__ __forceinline__ unsigned enc_ordered(float f) {
    unsigned u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float dec_ordered(unsigned u) {
    return __uint_as_float((u & 0x80000000u) ? (u & 0x7fffffffu) : ~u);
}

// total order on scores: -inf < finite < +inf < NaN; +0 and -0 compare equal, as they do as floats
__device__ __forceinline__ unsigned score_key(float f) {
    if (f != f) return 0xFFFFFFFFu;
    return enc_ordered(f + 0.0f);      // -0 + 0 = +0
}

// grid (i-blocks, j-splits, images), 256 threads
constexpr int kRankThreads = 256;
__global__ void __launch_bounds__(kRankThreads)
nms_rank_kernel(const float *__restrict__ boxes, const float *__restrict__ scores, SegTable seg,
                int *__restrict__ rank, unsigned *__restrict__ segmax, int want_max, int skip_invalid) {
    const int img = blockIdx.z;
    const int beg = seg.off[img], n = seg.off[img + 1] - beg;
    const int i = blockIdx.x * kRankThreads + threadIdx.x;
    if (blockIdx.x * kRankThreads >= n) return;
    // Scores are compared through an order-preserving unsigned key with NaN mapped to the largest key (torch.sort(descending)
    // ranks NaN first): the order is total, so the ranks are a permutation for ANY input -- with float compares a NaN score
    // compares false everywhere, every NaN box would get rank 0 and the scatter would leave sorted slots unwritten.
    const float si_f = i < n ? scores[beg + i] : 0.f;
    const unsigned si = score_key(si_f);

    __shared__ __align__(16) unsigned sj[kRankThreads];
    const int per = ceil_div(n, (int)gridDim.y);
    const int j0 = blockIdx.y * per, j1 = min(n, j0 + per);
    int cnt = 0;
    const int i_lo = blockIdx.x * kRankThreads, i_hi = i_lo + kRankThreads - 1;
    for (int base = j0; base < j1; base += kRankThreads) {
        int j = base + threadIdx.x;
        sj[threadIdx.x] = j < j1 ? score_key(scores[beg + j]) : 0u;
        __syncthreads();
        const int lim = min(kRankThreads, j1 - base);
        // rank = #{s_j > s_i} + #{s_j == s_i, j < i}: for a tile entirely before (after) this block's rows the
        // tie term is constant, so the inner loop is a single compare
        if (base + lim - 1 < i_lo) {
            int t = 0;
            for (; t + 4 <= lim; t += 4) {
                const uint4 s4 = *reinterpret_cast<const uint4 *>(&sj[t]);
                cnt += (s4.x >= si) + (s4.y >= si) + (s4.z >= si) + (s4.w >= si);
            }
            for (; t < lim; ++t) cnt += sj[t] >= si;
        } else if (base > i_hi) {
            int t = 0;
            for (; t + 4 <= lim; t += 4) {
                const uint4 s4 = *reinterpret_cast<const uint4 *>(&sj[t]);
                cnt += (s4.x > si) + (s4.y > si) + (s4.z > si) + (s4.w > si);
            }
            for (; t < lim; ++t) cnt += sj[t] > si;
        } else {
            for (int t = 0; t < lim; ++t) {
                const unsigned s = sj[t];
                cnt += (s > si) || (s == si && (base + t) < i);
            }
        }
        __syncthreads();
    }
    if (i < n && cnt) atomicAdd(&rank[beg + i], cnt);

    if (want_max && blockIdx.y == 0) {
        float m = -INFINITY;
        if (i < n && !(skip_invalid && si_f == -INFINITY)) {   // boxes.max() is taken over the valid boxes only
            float4 b = reinterpret_cast<const float4 *>(boxes)[beg + i];
            m = fmaxf(fmaxf(b.x, b.y), fmaxf(b.z, b.w));
        }
        m = warp_max(m);
        if ((threadIdx.x & 31) == 0 && m > -INFINITY) atomicMax(&segmax[img], enc_ordered(m));
    }
}

// grid (i-blocks, images)
__global__ void __launch_bounds__(256)
nms_scatter_kernel(const float *__restrict__ boxes, const int64_t *__restrict__ labels, SegTable seg,
                   const int *__restrict__ rank, const unsigned *__restrict__ segmax, int mode_in, NmsOpt opt,
                   float4 *__restrict__ sboxes, int *__restrict__ sidx, int *__restrict__ slab) {
    const int img = blockIdx.y;
    const int beg = seg.off[img], n = seg.off[img + 1] - beg;
    const int mode = eff_mode(opt, mode_in, eff_n(opt, img, n));
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float4 b = reinterpret_cast<const float4 *>(boxes)[beg + i];
    int lab = (mode != 0 && labels) ? (int)labels[beg + i] : 0;
    if (mode == 1) {
        // offsets = idxs.to(boxes) * (boxes.max() + 1); boxes + offsets[:, None]   (mmcv batched_nms)
        float off = __fmul_rn((float)lab, __fadd_rn(dec_ordered(segmax[img]), 1.0f));
        b.x = __fadd_rn(b.x, off); b.y = __fadd_rn(b.y, off);
        b.z = __fadd_rn(b.z, off); b.w = __fadd_rn(b.w, off);
    }
    int r = rank[beg + i];
    sboxes[beg + r] = b;
    sidx[beg + r] = i;
    slab[beg + r] = lab;
}

__device__ __forceinline__ bool iou_gt(const float4 &a, float area_a, const float4 &b, float thr) {
    float area_b = __fmul_rn(__fsub_rn(b.z, b.x), __fsub_rn(b.w, b.y));
    float xx1 = fmaxf(a.x, b.x), yy1 = fmaxf(a.y, b.y);
    float xx2 = fminf(a.z, b.z), yy2 = fminf(a.w, b.w);
    float w = fmaxf(0.f, __fsub_rn(xx2, xx1)), h = fmaxf(0.f, __fsub_rn(yy2, yy1));
    float inter = __fmul_rn(w, h);
    // disjoint boxes: inter == 0 -> IoU is 0 (or 0/0 = NaN), never > thr; skipping the IEEE division here
    // avoids its slow path (zero numerator) for the vast majority of pairs without changing any result
    if (!(inter > 0.f)) return false;
    float ovr = __fdiv_rn(inter, __fsub_rn(__fadd_rn(area_a, area_b), inter));
    return ovr > thr;
}

// grid (col-blocks, row-blocks, images), 64 threads; thread = one row of the tile
__global__ void __launch_bounds__(64)
nms_mask_kernel(const float4 *__restrict__ sboxes, const int *__restrict__ slab, SegTable seg,
                float thr, int mode_in, NmsOpt opt, int words, unsigned long long *__restrict__ mask) {
    const int img = blockIdx.z;
    const int beg = seg.off[img];
    const int n = eff_n(opt, img, seg.off[img + 1] - beg);
    const int mode = eff_mode(opt, mode_in, n);
    const int rb = blockIdx.y, cb = blockIdx.x;
    if (cb < rb || rb * 64 >= n || cb * 64 >= n) return;
    __shared__ float4 cbox[64];
    __shared__ int clab[64];
    const int t = threadIdx.x;
    const int ncol = min(64, n - cb * 64);
    if (t < ncol) {
        cbox[t] = sboxes[beg + cb * 64 + t];
        clab[t] = slab[beg + cb * 64 + t];
    }
    __syncthreads();
    const int r = rb * 64 + t;
    if (r >= n) return;
    const float4 a = sboxes[beg + r];
    const int la = slab[beg + r];
    const float area_a = __fmul_rn(__fsub_rn(a.z, a.x), __fsub_rn(a.w, a.y));
    unsigned long long bits = 0;
    const int start = (rb == cb) ? t + 1 : 0;
    for (int c = start; c < ncol; ++c) {
        bool sup = iou_gt(a, area_a, cbox[c], thr);
        if (mode == 2) sup = sup && (clab[c] == la);
        if (sup) bits |= 1ull << c;
    }
    mask[(size_t)(beg + r) * words + cb] = bits;
}

// one CTA per image
constexpr int kSweepThreads = 1024;
__global__ void __launch_bounds__(kSweepThreads)
nms_sweep_kernel(const unsigned long long *__restrict__ mask, const int *__restrict__ sidx,
                 SegTable seg, int words, int max_keep, NmsOpt opt, int64_t *__restrict__ keep_out,
                 int *__restrict__ num_keep_out) {
    extern __shared__ unsigned long long remv[];  // [words]
    __shared__ unsigned long long s_keepbits;
    __shared__ int s_count;
    const int img = blockIdx.x;
    const int beg = seg.off[img];
    const int n = eff_n(opt, img, seg.off[img + 1] - beg);
    const int nblk = ceil_div(n, 64);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int w = tid; w < words; w += kSweepThreads) remv[w] = 0ull;
    if (tid == 0) s_count = 0;
    const int limit = max_keep > 0 ? max_keep : 0x7fffffff;

    for (int blk = 0; blk < nblk; ++blk) {
        __syncthreads();
        if (warp == 0) {
            const int r0 = blk * 64 + lane, r1 = r0 + 32;
            unsigned long long d_lo = r0 < n ? mask[(size_t)(beg + r0) * words + blk] : 0ull;
            unsigned long long d_hi = r1 < n ? mask[(size_t)(beg + r1) * words + blk] : 0ull;
            unsigned long long cur = remv[blk];
            const int valid = min(64, n - blk * 64);
            if (valid < 64) cur |= ~0ull << valid;
            unsigned long long kb = 0ull;
#pragma unroll
            for (int t = 0; t < 64; ++t) {
                unsigned long long dt = __shfl_sync(0xffffffffu, t < 32 ? d_lo : d_hi, t & 31);
                if (!((cur >> t) & 1ull)) { kb |= 1ull << t; cur |= dt; }
            }
            const int before = s_count;
            // lanes write the kept ids of this block (two candidate rows per lane)
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                int t = lane + 32 * h;
                if ((kb >> t) & 1ull) {
                    int pos = before + __popcll(kb & ((1ull << t) - 1ull));
                    if (pos < limit) {
                        const int orig = sidx[beg + blk * 64 + t];
                        if (keep_out) keep_out[beg + pos] = (int64_t)orig;
                        if (opt.dets_out) {
                            const float4 b = reinterpret_cast<const float4 *>(opt.boxes)[beg + orig];
                            float *d = opt.dets_out + ((size_t)img * opt.keep_stride + pos) * 5;
                            d[0] = b.x; d[1] = b.y; d[2] = b.z; d[3] = b.w; d[4] = opt.scores[beg + orig];
                        }
                        if (opt.labels_out)
                            opt.labels_out[(size_t)img * opt.keep_stride + pos] = opt.labels ? opt.labels[beg + orig] : 0;
                    }
                }
            }
            __syncwarp();
            if (lane == 0) {
                s_keepbits = kb;
                s_count = before + __popcll(kb);
            }
        }
        __syncthreads();
        const unsigned long long kb = s_keepbits;
        if (s_count >= limit) break;
        // OR the mask rows of this block's survivors into remv for all later column words.
        const int nw = words - (blk + 1);
        if (nw > 0 && kb) {
            // 4 row groups x 256 word lanes
            const int g = tid >> 8, wl = tid & 255;
            for (int w = blk + 1 + wl; w < words; w += 256) {
                unsigned long long acc = 0ull;
                unsigned long long bits = kb;
                while (bits) {
                    int t = __ffsll((long long)bits) - 1;
                    bits &= bits - 1ull;
                    if ((t & 3) == g) acc |= mask[(size_t)(beg + blk * 64 + t) * words + w];
                }
                if (acc) atomicOr(&remv[w], acc);
            }
        }
    }
    __syncthreads();
    if (tid == 0) num_keep_out[img] = min(s_count, limit);
}

// ---------------------------------------------------------------------------------------------------------
// Bounded-survivor path (0 < max_keep <= kLazyMaxKeep, what multiclass_nms (100) and the RPN (300) ask for): the greedy
// sweep only ever needs IoUs against the boxes ALREADY KEPT, so the n^2/2 bitmask is never built.  One CTA per image
// walks the score-sorted candidates 64 at a time:
//   (a) 16 threads per candidate test it against the kept list (shared memory),
//   (b) the 64x64 intra-block mask is computed by the same 1024 threads,
//   (c) warp 0 resolves the block serially in registers (identical recurrence to nms_sweep_kernel) and appends.
// It stops as soon as max_keep boxes survived.  Same IoU arithmetic, same order => bit-identical survivors.
constexpr int kLazyMaxKeep = 1024;
__global__ void __launch_bounds__(1024)
nms_lazy_kernel(const float4 *__restrict__ sboxes, const int *__restrict__ slab, const int *__restrict__ sidx,
                SegTable seg, float thr, int mode_in, int max_keep, NmsOpt opt, int64_t *__restrict__ keep_out,
                int *__restrict__ num_keep_out) {
    extern __shared__ float4 kept_box[];                 // [max_keep]
    int *kept_lab = reinterpret_cast<int *>(kept_box + max_keep);   // [max_keep]
    __shared__ float4 cbox[64];
    __shared__ int clab[64];
    __shared__ unsigned long long s_dead, s_d[64], s_keepbits;
    __shared__ int s_count;
    const int img = blockIdx.x;
    const int beg = seg.off[img];
    const int n = eff_n(opt, img, seg.off[img + 1] - beg);
    const int mode = eff_mode(opt, mode_in, n);
    const int nblk = ceil_div(n, 64);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) s_count = 0;
    __syncthreads();

    for (int blk = 0; blk < nblk; ++blk) {
        const int nc = min(64, n - blk * 64);
        if (tid < 64) {
            if (tid < nc) { cbox[tid] = sboxes[beg + blk * 64 + tid]; clab[tid] = slab[beg + blk * 64 + tid]; }
            s_d[tid] = 0ull;
        }
        if (tid == 0) s_dead = nc < 64 ? (~0ull << nc) : 0ull;
        __syncthreads();
        const int count = s_count;
        const int c = tid >> 4, h = tid & 15;           // candidate, helper lane
        // (a) candidate c against the kept list
        if (c < nc) {
            const float4 b = cbox[c];
            const int lb = clab[c];
            bool sup = false;
            for (int k = h; k < count && !sup; k += 16) {
                const float4 a = kept_box[k];
                const float area_a = __fmul_rn(__fsub_rn(a.z, a.x), __fsub_rn(a.w, a.y));
                sup = iou_gt(a, area_a, b, thr) && (mode != 2 || kept_lab[k] == lb);
            }
            const unsigned grp = __ballot_sync(0xffffffffu, sup) >> (lane & 16) & 0xffffu;
            if (h == 0 && grp) atomicOr(&s_dead, 1ull << c);
        } else {
            __ballot_sync(0xffffffffu, false);
        }
        // (b) intra-block mask: row r = c (higher score), columns 4h..4h+3
        {
            unsigned long long bits = 0ull;
            if (c < nc) {
                const float4 a = cbox[c];
                const int la = clab[c];
                const float area_a = __fmul_rn(__fsub_rn(a.z, a.x), __fsub_rn(a.w, a.y));
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const int cc = 4 * h + q;
                    if (cc > c && cc < nc) {
                        bool sup = iou_gt(a, area_a, cbox[cc], thr);
                        if (mode == 2) sup = sup && (clab[cc] == la);
                        if (sup) bits |= 1ull << cc;
                    }
                }
            }
            // OR-reduce over the 16 helper lanes of this candidate
#pragma unroll
            for (int o = 8; o > 0; o >>= 1) bits |= __shfl_xor_sync(0xffffffffu, bits, o);
            if (h == 0 && c < nc) s_d[c] = bits;
        }
        __syncthreads();
        // (c) serial resolve of the block by warp 0
        if (warp == 0) {
            const unsigned long long d_lo = s_d[lane], d_hi = s_d[lane + 32];
            unsigned long long cur = s_dead, kb = 0ull;
#pragma unroll
            for (int t = 0; t < 64; ++t) {
                const unsigned long long dt = __shfl_sync(0xffffffffu, t < 32 ? d_lo : d_hi, t & 31);
                if (!((cur >> t) & 1ull)) { kb |= 1ull << t; cur |= dt; }
            }
#pragma unroll
            for (int hh = 0; hh < 2; ++hh) {
                const int t = lane + 32 * hh;
                if ((kb >> t) & 1ull) {
                    const int pos = count + __popcll(kb & ((1ull << t) - 1ull));
                    if (pos < max_keep) {
                        kept_box[pos] = cbox[t];
                        kept_lab[pos] = clab[t];
                        const int orig = sidx[beg + blk * 64 + t];
                        if (keep_out) keep_out[beg + pos] = (int64_t)orig;
                        if (opt.dets_out) {
                            const float4 b = reinterpret_cast<const float4 *>(opt.boxes)[beg + orig];
                            float *d = opt.dets_out + ((size_t)img * opt.keep_stride + pos) * 5;
                            d[0] = b.x; d[1] = b.y; d[2] = b.z; d[3] = b.w; d[4] = opt.scores[beg + orig];
                        }
                        if (opt.labels_out)
                            opt.labels_out[(size_t)img * opt.keep_stride + pos] = opt.labels ? opt.labels[beg + orig] : 0;
                    }
                }
            }
            __syncwarp();
            if (lane == 0) { s_keepbits = kb; s_count = count + __popcll(kb); }
        }
        __syncthreads();
        if (s_count >= max_keep) break;
    }
    if (tid == 0) num_keep_out[img] = min(s_count, max_keep);
}

}  // namespace vod

using namespace vod;

extern "C" size_t vod_nms_workspace_bytes(int n_total, int max_seg) {
    if (n_total <= 0) return 256;
    return carve(nullptr, n_total, ceil_div(max_seg, 64)).bytes;
}

extern "C" int vod_batched_nms_ex(const float *boxes, const float *scores, const int64_t *labels,
                                  int n_total, const int *seg_offsets_host, int n_images, float iou_thr,
                                  int mode, int max_keep, const int *n_valid_dev, int split_thr,
                                  int64_t *keep_out, int *num_keep_out, float *dets_out, int64_t *labels_out,
                                  void *ws, size_t ws_bytes, vod_stream_t stream) {
    VOD_REQUIRE(n_images >= 1 && n_images <= kMaxImages, "vod_batched_nms: n_images=%d not in [1,%d]",
                n_images, kMaxImages);
    VOD_REQUIRE(seg_offsets_host && num_keep_out, "vod_batched_nms: null seg_offsets/num_keep_out");
    VOD_REQUIRE(mode >= 0 && mode <= 3, "vod_batched_nms: mode=%d", mode);
    VOD_REQUIRE(!(dets_out || labels_out) || max_keep > 0, "vod_batched_nms: dets_out/labels_out need max_keep > 0");
    VOD_REQUIRE(mode == 0 || labels, "vod_batched_nms: labels required for mode %d", mode);
    SegTable seg;
    seg.n_images = n_images;
    int max_seg = 0;
    for (int i = 0; i <= n_images; ++i) seg.off[i] = seg_offsets_host[i];
    for (int i = 0; i < n_images; ++i) {
        VOD_REQUIRE(seg.off[i + 1] >= seg.off[i], "vod_batched_nms: offsets not monotone");
        max_seg = max(max_seg, seg.off[i + 1] - seg.off[i]);
    }
    VOD_REQUIRE(seg.off[0] == 0 && seg.off[n_images] == n_total, "vod_batched_nms: offsets do not cover n_total");
    cudaStream_t st = as_stream(stream);
    if (n_total == 0 || max_seg == 0) {
        cudaMemsetAsync(num_keep_out, 0, sizeof(int) * n_images, st);
        return check_launch("vod_batched_nms(memset)");
    }
    VOD_REQUIRE(boxes && scores && (keep_out || dets_out) && ws, "vod_batched_nms: null pointer");
    NmsOpt opt;
    opt.n_valid = n_valid_dev; opt.split_thr = split_thr; opt.dets_out = dets_out; opt.labels_out = labels_out;
    opt.boxes = boxes; opt.scores = scores; opt.labels = labels; opt.keep_stride = max_keep;
    const int words = ceil_div(max_seg, 64);
    VOD_REQUIRE(words * 8 <= 200 * 1024, "vod_batched_nms: image with %d boxes too large", max_seg);
    NmsWorkspace w = carve(ws, n_total, words);
    if (ws_bytes < w.bytes) return fail(VOD_E_WORKSPACE, "vod_batched_nms: workspace %zu < %zu", ws_bytes, w.bytes);

    // rank + segmax are contiguous at the head of the workspace
    size_t zero_bytes = reinterpret_cast<char *>(w.sboxes) - reinterpret_cast<char *>(w.rank);
    cudaMemsetAsync(w.rank, 0, zero_bytes, st);

    const int iblocks = ceil_div(max_seg, kRankThreads);
    // enough j-splits to fill the machine (~4 CTAs/SM), each split at least 256 candidates
    int jsplits = max(1, min(ceil_div(max_seg, 256), ceil_div(4 * num_sms(), iblocks * n_images)));
    nms_rank_kernel<<<dim3(iblocks, jsplits, n_images), kRankThreads, 0, st>>>(
        boxes, scores, seg, w.rank, w.segmax, mode == 1 || mode == 3, n_valid_dev != nullptr); note_launch();
    nms_scatter_kernel<<<dim3(ceil_div(max_seg, 256), n_images), 256, 0, st>>>(
        boxes, labels, seg, w.rank, w.segmax, mode, opt, w.sboxes, w.sidx, w.slab); note_launch();
    if (max_keep > 0 && max_keep <= kLazyMaxKeep) {
        const size_t lsmem = (sizeof(float4) + sizeof(int)) * (size_t)max_keep;
        nms_lazy_kernel<<<n_images, 1024, lsmem, st>>>(w.sboxes, w.slab, w.sidx, seg, iou_thr, mode, max_keep, opt,
                                                       keep_out, num_keep_out); note_launch();
        return check_launch("vod_batched_nms(lazy)");
    }
    nms_mask_kernel<<<dim3(words, words, n_images), 64, 0, st>>>(w.sboxes, w.slab, seg, iou_thr, mode, opt,
                                                                 words, w.mask); note_launch();
    size_t smem = sizeof(unsigned long long) * words;
    if (smem > 40 * 1024)
        cudaFuncSetAttribute(nms_sweep_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    nms_sweep_kernel<<<n_images, kSweepThreads, smem, st>>>(w.mask, w.sidx, seg, words, max_keep, opt,
                                                            keep_out, num_keep_out); note_launch();
    return check_launch("vod_batched_nms");
}

extern "C" int vod_batched_nms(const float *boxes, const float *scores, const int64_t *labels, int n_total,
                               const int *seg_offsets_host, int n_images, float iou_thr, int mode, int max_keep,
                               int64_t *keep_out, int *num_keep_out, void *ws, size_t ws_bytes, vod_stream_t stream) {
    VOD_REQUIRE(mode >= 0 && mode <= 2, "vod_batched_nms: mode=%d", mode);
    return vod_batched_nms_ex(boxes, scores, labels, n_total, seg_offsets_host, n_images, iou_thr, mode, max_keep,
                              nullptr, 0, keep_out, num_keep_out, nullptr, nullptr, ws, ws_bytes, stream);
}
