// msra_gemm.cu -- (4) TemporalRoIAlign most-similar-location search on the tensor cores (sm_100a).
//
// Restates TemporalRoIAlign.most_similar_roi_align,
// mmtracking/mmtrack/models/roi_heads/roi_extractors/temporal_roi_align.py:99-181:
//   sim[row, t, loc] = <roi_unit[row, :], ref_unit[t, loc, :]>,  top-k over loc per (row, t).
// The [N*49, T*HW] similarity matrix (2.1 GB fp32 at N=300, T=15) is never written: the epilogue keeps a
// running top-4 per (row, frame, interleaved location group) in registers while the accumulator tile is read out of TMEM.  The
// candidates are then re-scored in exact fp32 (msra_rescore_kernel, tafa.cu) so that the selected
// locations match the fp32 reference; the bf16 GEMM is only a pre-filter.
//
// Persistent kernel, one CTA per SM, CTAs paired into clusters of 2 (tcgen05 cta_group::2), 576 threads per CTA;
// a work unit = (PAIR of adjacent 128-row tiles, frame t), i.e. a 256-row x HW similarity slab:
//   A tile   (128 RoI rows x C bf16 per CTA) lives in TENSOR MEMORY, not shared memory: row r in TMEM lane r, two
//            bf16 per 32-bit column (256 columns for C = 512), written with tcgen05.st by epilogue warps 0-3
//            whenever the row tile changes and consumed by the TS form of tcgen05.mma.  That frees all of the
//            shared memory for the B ring.
//   warp 16  TMA producer (both CTAs): each CTA loads HALF of every B stage (64 of the tile's 128 locations, all C
//            channels) with ONE 5-D TMA box (64 channels | 32 x 2 interleaved locations | frame | 8 K slices ->
//            eight SWIZZLE_128B sub-tiles, 64 KB) and signals the leader's mbarrier; 3 stages per CTA.  One copy
//            per 8 KB had left the TMA unit, not the tensor pipe, in charge of the pace (26 stages, 535 us).
//   warp 17  MMA issuer (leader CTA only): tcgen05.mma.cta_group::2 kind::f16 (bf16 in, fp32 accumulate), M = 256
//            over the pair, N = 128; accumulator double buffered in each CTA's TMEM (2 x 128 columns; TMEM total:
//            256 accumulator + 256 A = 512 columns); commits are multicast to both CTAs' barriers
//   warps 0-15 epilogue (both CTAs, own 128 rows): thread = row (four warps per TMEM lane quarter, each owning 32 of
//            the tile's 128 columns = every 4th location of the tile); every similarity is packed with its location into one order-preserving 32-bit
//            key (20 value bits | 12 location bits, built on the FMA pipe) and pushed through a branch-free min/max insertion network that
//            keeps the 4 largest keys in registers -- no divergence although the 32 lanes of a warp follow 32
//            different rows; one 16-byte store per (row, frame, column group) at the end
// Units are assigned to clusters in contiguous ranges so the A tile is reloaded only when the row-tile pair changes.
// History (profiles/r01_bench_history.md): 5-stage smem-A 1-CTA 1108 us -> TMEM-A 13-stage 1-CTA 641 us -> CTA pair 555 us ->
// whole-tile 64 KB TMA stages 395 us -> FMA-pipe key packing 380 us (+ interleaved 5-D boxes, same speed).
#include <cuda_bf16.h>

#include "common.cuh"
#include "msra.cuh"
#include "tc.cuh"

namespace vod {

constexpr int kMgBM = 128;           // rows per CTA (a CTA pair covers 256)
constexpr int kMgBN = 128;           // locations per accumulator tile
constexpr int kMgSlice = 64;         // bf16 elements per 128-byte K slice
constexpr int kMgMaxSlices = 8;      // C <= 512
constexpr int kMgHalf = 64 * 128;    // bytes of one K slice of a CTA's half of B: 64 locations x 128 B
constexpr int kMgSPS = 8;            // K slices per stage = per TMA instruction (3-D box): few, large bulk copies --
                                     // with one 8 KB copy per instruction the TMA unit, not the tensor pipe, set the pace
constexpr int kMgStageBytes = kMgSPS * kMgHalf;
constexpr int kMgStages = 3;         // 6 x 32 KB = 192 KB of B in flight per CTA
constexpr int kMgEpiWarps = 16;
constexpr int kMgThreads = (kMgEpiWarps + 2) * 32;
constexpr int kMgSmem = kMgStages * kMgStageBytes + 1024;
constexpr uint32_t kMgACol = 256;    // first TMEM column of the A tile

struct MgParams {
    uint32_t *cand; // [NP, T, kMsraCand] packed keys: round((1.5 + sim) * 2^11) << 12 | location
    const __nv_bfloat16 *a_rows;  // [NP, C] unit-norm RoI rows
    int NP, T, HW, nslices, nstg, row_tiles, ntiles;  // ntiles = ceil(HW / 128), nstg = ceil(nslices / kMgSPS)
    int units;      // ceil(row_tiles / 2) * T  (a unit = a PAIR of row tiles x one frame)
    int k4096, kone;   // the constants 4096 and 1, opaque to the compiler (see the epilogue's key packing)
};

// CTA pair (cluster of 2, cta_group::2): the two CTAs own two adjacent 128-row tiles, i.e. one 256 x 128 MMA tile.
// Each CTA keeps its own A rows in its own TMEM and TMA-loads only HALF of every B stage (64 of the 128 locations);
// the leader CTA (rank 0) issues tcgen05.mma.cta_group::2 for both.  A single-CTA tcgen05.mma runs at half the
// tensor-core rate of a CTA pair (measured: 1.09 PFLOP/s ceiling for the 1-CTA version of this kernel), and the pair
// also halves the L2 -> shared-memory operand traffic.
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kMgThreads, 1)
msra_gemm_topk_kernel(const __grid_constant__ CUtensorMap tm_b, const MgParams p) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t *sB = smem;                               // [stages][slices per stage][8 KB]
    __shared__ uint64_t a_full, b_full[kMgStages], b_empty[kMgStages], acc_full[2], acc_empty[2];
    __shared__ uint32_t tmem_slot;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t crank = tc::cluster_ctarank();
    // contiguous unit range of this CLUSTER (both CTAs walk the same units; the CTA rank picks the row tile)
    const int ncl = gridDim.x >> 1, cl = blockIdx.x >> 1;
    const int per = p.units / ncl, rem = p.units % ncl;
    const int u0 = cl * per + min(cl, rem);
    const int u1 = u0 + per + (cl < rem ? 1 : 0);

    if (threadIdx.x == 0) {
        // leader-side barriers are also initialised in the peer (unused there) to keep the code symmetric
        tc::mbar_init(&a_full, 8);                      // 4 A-writer warps in each CTA of the pair
        for (int i = 0; i < kMgStages; ++i) { tc::mbar_init(&b_full[i], 1); tc::mbar_init(&b_empty[i], 1); }
        for (int i = 0; i < 2; ++i) { tc::mbar_init(&acc_full[i], 1); tc::mbar_init(&acc_empty[i], 2 * kMgEpiWarps); }
        tc::fence_barrier_init();
    }
    if (warp == kMgEpiWarps + 1) tc::tmem_alloc_2sm(&tmem_slot, 512);
    tc::tcgen05_fence_before();
    __syncthreads();
    tc::cluster_sync_all();   // both CTAs' barriers and TMEM exist before anything crosses the pair
    tc::tcgen05_fence_after();
    const uint32_t tmem = tmem_slot;

    if (warp == kMgEpiWarps) {
        // ------------------------------------------------------------------ TMA producer (both CTAs: own half of B)
        if (tc::elect_one()) {
            tc::tma_prefetch_desc(&tm_b);
            uint32_t st = 0, ph = 0;
            for (int u = u0; u < u1; ++u) {
                const int t = u % p.T;
                for (int nt = 0; nt < p.ntiles; ++nt) {
                    for (int g = 0; g < p.nstg; ++g) {
                        tc::mbar_wait(&b_empty[st], ph ^ 1);   // the pair's MMAs have consumed this stage (multicast commit)
                        if (crank == 0) tc::mbar_arrive_expect_tx(&b_full[st], 2 * kMgStageBytes);   // both halves land on the leader's barrier
                        // box = 64 channels x (32 a x 2 g) locations x kMgSPS slices of frame t; this CTA's 64 B rows are the
                        // locations 128*nt + 4*a + g with g in {2*rank, 2*rank + 1} (slices past C / locations past the
                        // frame are zero-filled or masked in the epilogue)
                        tc::tma_load_5d_2sm(sB + st * kMgStageBytes, &tm_b, &b_full[st], 0, nt * (kMgBN / 4), 2 * (int)crank, t, g * kMgSPS);
                        if (++st == kMgStages) { st = 0; ph ^= 1; }
                    }
                }
            }
        }
    } else if (warp == kMgEpiWarps + 1) {
        // ------------------------------------------------------------------ MMA issuer (leader CTA only)
        if (crank == 0) {
            constexpr uint32_t idesc = tc::umma_idesc(tc::kFmtBF16, 2 * kMgBM, kMgBN);   // M = 256 over the pair
            int cur_pr = -1;
            uint32_t a_loads = 0, st = 0, ph = 0, tile = 0;
            for (int u = u0; u < u1; ++u) {
                const int pr = u / p.T;
                if (pr != cur_pr) {
                    tc::mbar_wait(&a_full, a_loads & 1);   // both CTAs have written their A rows into their TMEM
                    tc::tcgen05_fence_after();
                    cur_pr = pr;
                    ++a_loads;
                }
                for (int nt = 0; nt < p.ntiles; ++nt, ++tile) {
                    const uint32_t buf = tile & 1;
                    tc::mbar_wait(&acc_empty[buf], ((tile >> 1) & 1) ^ 1);   // drained by the epilogue warps of BOTH CTAs
                    tc::tcgen05_fence_after();
                    for (int g = 0; g < p.nstg; ++g) {
                        tc::mbar_wait(&b_full[st], ph);
                        tc::tcgen05_fence_after();
                        if (tc::elect_one()) {
                            const int ns = min(kMgSPS, p.nslices - g * kMgSPS);
                            for (int j = 0; j < ns; ++j) {
                                const int s = g * kMgSPS + j;
                                const uint32_t b0 = tc::smem_u32(sB + st * kMgStageBytes + j * kMgHalf);
#pragma unroll
                                for (int k = 0; k < 4; ++k)   // 16 bf16 of K per MMA: 8 TMEM columns of A, 32 bytes of the B slice
                                    tc::umma_f16_ts_2sm(tmem + buf * kMgBN, tmem + kMgACol + s * 32 + k * 8,
                                                        tc::umma_desc_k_sw128(b0 + k * 32), idesc, (s | k) != 0);
                            }
                            tc::umma_commit_2sm(&b_empty[st], (uint16_t)0x3);                       // frees the stage in both CTAs
                            if (g == p.nstg - 1) tc::umma_commit_2sm(&acc_full[buf], (uint16_t)0x3);   // wakes both epilogues
                        }
                        __syncwarp();
                        if (++st == kMgStages) { st = 0; ph ^= 1; }
                    }
                }
            }
        }
    } else {
        // ------------------------------------------------------------------ epilogue: running top-4 keys per row
        const int quarter = warp & 3, grp = warp >> 2;   // TMEM lane quarter, 32-column group of the tile
        const uint32_t tl = tmem + ((uint32_t)(quarter * 32) << 16) + grp * 32;
        uint32_t tile_it = 0;
        int cur_rt = -1;
        for (int u = u0; u < u1; ++u) {
            const int rt = (u / p.T) * 2 + (int)crank, t = u % p.T;
            if (rt != cur_rt) {
                cur_rt = rt;
                if (warp < 4) {
                    // New row tile: this thread's RoI row -> TMEM lane, two bf16 per column.  Every MMA that read the
                    // previous A tile has completed (this warp consumed that tile's last accumulator already).
                    const int arow = rt * kMgBM + warp * 32 + lane;
                    const uint32_t ta = tmem + ((uint32_t)(warp * 32) << 16) + kMgACol;
                    const int C = p.nslices * kMgSlice;
                    for (int c = 0; c < C / 2; c += 16) {
                        uint32_t v[16];
                        if (arow < p.NP) {
                            const uint4 *src = reinterpret_cast<const uint4 *>(p.a_rows + (size_t)arow * C + 2 * c);
#pragma unroll
                            for (int q = 0; q < 4; ++q) {
                                const uint4 w4 = __ldg(src + q);
                                v[4 * q] = w4.x; v[4 * q + 1] = w4.y; v[4 * q + 2] = w4.z; v[4 * q + 3] = w4.w;
                            }
                        } else {
#pragma unroll
                            for (int q = 0; q < 16; ++q) v[q] = 0u;
                        }
                        tc::tmem_st_32x16(ta + c, v);
                    }
                    tc::tmem_st_wait();
                    tc::tcgen05_fence_before();
                    __syncwarp();
                    if (lane == 0) tc::mbar_arrive_cluster(&a_full, 0);   // leader's barrier
                }
            }
            uint32_t k0 = 0, k1 = 0, k2 = 0, k3 = 0;   // descending; 0 = "nothing yet" (below every real key)
            for (int nt = 0; nt < p.ntiles; ++nt, ++tile_it) {
                const uint32_t buf = tile_it & 1;
                tc::mbar_wait(&acc_full[buf], (tile_it >> 1) & 1);
                tc::tcgen05_fence_after();
                uint32_t ra[32];
                tc::tmem_ld_32x32(tl + buf * kMgBN, ra);
                tc::tmem_ld_wait();
                tc::tcgen05_fence_before();
                __syncwarp();
                if (lane == 0) tc::mbar_arrive_cluster(&acc_empty[buf], 0);   // values are in registers: the pair's MMA may overwrite
                // accumulator column 32*grp + j holds location 128*nt + 4*j + grp (see the producer): the four column groups
                // interleave the locations, so a run of near-tied NEIGHBOURING pixels is spread over all four top-4 lists
                const uint32_t base = (uint32_t)(nt * kMgBN + grp);
                const int nvalid = (p.HW - (int)base + 3) >> 2;    // j < nvalid <=> location < HW (tail tile only)
                auto push = [&](uint32_t bits, uint32_t j, bool ok) {
                    // |cos| <= 1 (+ bf16 noise): t = sim + 6145.5 lies in [4096, 8192), where one ulp is 2^-11, so the FADD
                    // itself rounds the similarity to 11 fractional bits and bits(t) = 0x45800000 + (2049.5 + sim) * 2^11.
                    // Multiplying the word by 4096 (mod 2^32) drops the constant part and leaves (1.5 + sim) * 2^23, a
                    // positive, order-preserving 20-bit value field above 12 free location bits.  FADD and both IMADs run
                    // on the FMA pipe; the ALU pipe -- the busy one -- only sees the 7 min/max of the insertion network.
                    // (The multipliers 4096 and 1 are kernel parameters so that ptxas cannot turn the IMADs into ALU ops.)
                    // A NaN similarity gives the canonical 0x7FFFFFFF -> value field 0xFFFFF: ranked first, like torch.topk.
                    const uint32_t w = __float_as_uint(__uint_as_float(bits) + 6145.5f);
                    uint32_t key = tc::mad_lo(tc::mad_lo(w, (uint32_t)p.k4096, base), (uint32_t)p.kone, 4u * j);
                    if (!ok) key = 0u;
                    uint32_t hi;
                    hi = max(k0, key); key = min(k0, key); k0 = hi;
                    hi = max(k1, key); key = min(k1, key); k1 = hi;
                    hi = max(k2, key); key = min(k2, key); k2 = hi;
                    k3 = max(k3, key);
                };
                if (nvalid >= 32) {
#pragma unroll
                    for (int j = 0; j < 32; ++j) push(ra[j], j, true);
                } else {
#pragma unroll
                    for (int j = 0; j < 32; ++j) push(ra[j], j, j < nvalid);
                }
            }
            const int row = rt * kMgBM + quarter * 32 + lane;
            if (row < p.NP)
                *reinterpret_cast<uint4 *>(p.cand + ((size_t)row * p.T + t) * kMsraCand + grp * 4) = make_uint4(k0, k1, k2, k3);
        }
    }
    tc::tcgen05_fence_before();
    __syncthreads();
    tc::cluster_sync_all();   // neither CTA may free TMEM / exit while the pair's MMAs or remote arrivals are pending
    if (warp == kMgEpiWarps + 1) tc::tmem_dealloc_2sm(tmem, 512);
}

bool msra_gemm_supported(int NP, int C, int T, int HW) {
    // HW <= 4096: the location shares a 32-bit key with the similarity (12 bits)
    return NP > 0 && T > 0 && HW > 0 && HW <= 4096 && C % kMgSlice == 0 && C / kMgSlice <= kMgMaxSlices;
}

int msra_launch_gemm_topk(const void *roi_unit_bf16, const void *ref_unit_bf16, uint32_t *cand, int NP, int NP_pad, int C,
                          int T, int HW, cudaStream_t st) {
    (void)NP_pad;
    CUtensorMap tb;
    int rc;
    if ((reinterpret_cast<uintptr_t>(roi_unit_bf16) & 15) != 0) return fail(VOD_E_BADARG, "msra_gemm: A rows must be 16-byte aligned");
    // each CTA of a pair loads 64 (interleaved) locations x kMgSPS K slices per instruction
    if ((rc = make_tmap_msra_b(&tb, ref_unit_bf16, T, HW, C, kMgSPS))) return rc;
    MgParams p;
    p.cand = cand; p.NP = NP; p.T = T; p.HW = HW;
    p.a_rows = reinterpret_cast<const __nv_bfloat16 *>(roi_unit_bf16);
    p.nslices = C / kMgSlice;
    p.nstg = ceil_div(p.nslices, kMgSPS);
    p.row_tiles = ceil_div(NP, kMgBM);
    p.ntiles = ceil_div(HW, kMgBN);
    p.units = ceil_div(p.row_tiles, 2) * T;
    p.k4096 = 4096; p.kone = 1;
    cudaFuncSetAttribute(msra_gemm_topk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kMgSmem);
    const int grid = 2 * min(num_sms() / 2, p.units);   // clusters of 2 CTAs
    msra_gemm_topk_kernel<<<grid, kMgThreads, kMgSmem, st>>>(tb, p); note_launch();
    return check_launch("msra_gemm_topk");
}

}  // namespace vod

using namespace vod;

namespace {
struct MsraWs {
    size_t roi_norm, ref_norm, roi_unit, ref_unit, cand, ovf_ctrl, ovf_ctrl_ints, ovf_pairs, ovf_top2, ovf_bins, ovf_scan, ovf_split, bytes;
};
MsraWs msra_ws(int NP, int C, int T, int HW) {
    MsraWs w;
    size_t o = 0;
    w.roi_norm = o; o = align_up(o + sizeof(float) * (size_t)NP, 256);
    w.ref_norm = o; o = align_up(o + sizeof(float) * (size_t)T * HW, 256);
    w.roi_unit = o; o = align_up(o + 2 * (size_t)NP * C, 1024);
    w.ref_unit = o; o = align_up(o + 2 * ((size_t)T * HW + 4) * C, 1024);   // + the padding rows the TMA box may touch
    w.cand = o;     o = align_up(o + sizeof(int) * (size_t)NP * T * kMsraCand, 256);
    // overflow work lists of the exact-by-construction top-k (msra.cuh): sized for the worst case, every pair flagged
    const size_t pairs = (size_t)NP * T;
    w.ovf_ctrl_ints = 1 + 4 * (size_t)T + kMsraOvfSplitChunks;      // counters + the split scan's arrival counters: one memset
    w.ovf_ctrl = o;  o = align_up(o + sizeof(int) * w.ovf_ctrl_ints, 256);
    w.ovf_split = o; o = align_up(o + sizeof(float4) * kMsraOvfSplitChunks * 8 * 8, 256);
    w.ovf_pairs = o; o = align_up(o + sizeof(int4) * pairs, 256);
    w.ovf_top2 = o;  o = align_up(o + sizeof(float4) * pairs, 256);
    w.ovf_bins = o;  o = align_up(o + sizeof(int2) * 4 * pairs, 256);
    w.ovf_scan = o;  o = align_up(o + sizeof(float4) * 4 * pairs, 256);
    w.bytes = o;
    return w;
}
}  // namespace

extern "C" size_t vod_msra_workspace_bytes(int NP, int C, int T, int HW, int k) {
    (void)k;
    if (NP <= 0 || T <= 0) return 256;
    return msra_ws(NP, C, T, HW).bytes;
}

extern "C" int vod_msra_topk_sample(const float *roi_feats, const float *ref_nhwc, const float *ref_norm,
                                    const void *ref_unit_bf16, float *out, int *idx_out, float *val_out, int NP, int C,
                                    int T, int HW, int k, int impl, void *ws, size_t ws_bytes, vod_stream_t stream) {
    if (NP == 0 || T == 0) return VOD_OK;
    VOD_REQUIRE(roi_feats && ref_nhwc && out && ws, "vod_msra_topk_sample: null pointer");
    VOD_REQUIRE(NP > 0 && C > 0 && T > 0 && HW > 0, "vod_msra_topk_sample: bad dims");
    VOD_REQUIRE(k >= 1 && k <= kMsraMaxK && k <= HW, "vod_msra_topk_sample: k=%d not in [1,%d]", k, kMsraMaxK);
    VOD_REQUIRE(impl >= 0 && impl <= 2, "vod_msra_topk_sample: impl");
    VOD_REQUIRE((reinterpret_cast<uintptr_t>(ws) & 1023) == 0, "vod_msra_topk_sample: workspace must be 1024-byte aligned");
    const MsraWs w = msra_ws(NP, C, T, HW);
    if (ws_bytes < w.bytes) return fail(VOD_E_WORKSPACE, "vod_msra_topk_sample: workspace %zu < %zu", ws_bytes, w.bytes);
    uint8_t *wsb = reinterpret_cast<uint8_t *>(ws);
    // The candidate pass keeps the 4 best of four interleaved location groups: sized for the reference's k = 2 (and k = 1).
    // Larger k runs the exact scan (a true top-4 with three members in one group would leave no slack in that group).
    const bool tc_ok = msra_gemm_supported(NP, C, T, HW) && HW >= kMsraCand && k <= 2 && T <= 256 && vod_device_is_sm100();
    if (impl == 2 && !tc_ok)
        return fail(VOD_E_UNSUPPORTED, "vod_msra_topk_sample: tcgen05 path needs C %% 64 == 0, C <= 512, k <= 2, sm_100");
    const bool use_tc = impl == 2 || (impl == 0 && tc_ok);

    float *roi_norm = reinterpret_cast<float *>(wsb + w.roi_norm);
    void *roi_unit = use_tc ? wsb + w.roi_unit : nullptr;
    int rc = vod_rows_l2norm(roi_feats, roi_norm, roi_unit, NP, C, stream);
    if (rc) return rc;
    const float *rn = ref_norm;
    const void *ru = ref_unit_bf16;
    if (!rn || (use_tc && !ru)) {
        float *rn_ws = reinterpret_cast<float *>(wsb + w.ref_norm);
        void *ru_ws = use_tc ? wsb + w.ref_unit : nullptr;
        rc = vod_rows_l2norm(ref_nhwc, rn_ws, ru_ws, T * HW, C, stream);
        if (rc) return rc;
        rn = rn_ws;
        if (use_tc) ru = ru_ws;
    }
    cudaStream_t st = as_stream(stream);
    if (!use_tc) return msra_launch_scan(roi_feats, ref_nhwc, roi_norm, rn, out, idx_out, val_out, NP, C, T, HW, k, st);
    uint32_t *cand = reinterpret_cast<uint32_t *>(wsb + w.cand);
    rc = msra_launch_gemm_topk(roi_unit, ru, cand, NP, NP, C, T, HW, st);
    if (rc) return rc;
    // exact-by-construction top-k: the re-score flags every (row, frame, group) whose candidate list may have lost a member of
    // the exact top-k; msra_overflow_fix re-scans those groups in fp32 (usually none: two empty launches)
    MsraOvf ovf;
    ovf.ctrl = reinterpret_cast<int *>(wsb + w.ovf_ctrl);
    ovf.pair_list = reinterpret_cast<int4 *>(wsb + w.ovf_pairs);
    ovf.pair_top = reinterpret_cast<float4 *>(wsb + w.ovf_top2);
    ovf.bin_list = reinterpret_cast<int2 *>(wsb + w.ovf_bins);
    ovf.ovf_top = reinterpret_cast<float4 *>(wsb + w.ovf_scan);
    ovf.done = ovf.ctrl + 1 + 4 * (size_t)T;
    ovf.split_top = reinterpret_cast<float4 *>(wsb + w.ovf_split);
    const bool fix = 4 * T <= 1024;    // (more frames than the fix-up kernels index: T > 256 never reaches here in practice)
    if (fix) {
        cudaError_t e = cudaMemsetAsync(ovf.ctrl, 0, sizeof(int) * w.ovf_ctrl_ints, st);
        if (e != cudaSuccess) return fail(VOD_E_LAUNCH, "vod_msra_topk_sample: memset: %s", cudaGetErrorString(e));
    }
    rc = msra_launch_rescore(roi_feats, ref_nhwc, roi_norm, rn, cand, kMsraCand, out, idx_out, val_out, NP, C, T, HW, k,
                             fix ? &ovf : nullptr, st);
    if (rc || !fix) return rc;
    return msra_overflow_fix(roi_feats, ref_nhwc, roi_norm, rn, ovf, out, idx_out, val_out, NP, C, T, HW, k, st);
}

extern "C" size_t vod_msra_overflow_counter_offset(int NP, int C, int T, int HW) {
    return msra_ws(NP, C, T, HW).ovf_ctrl;
}

extern "C" int vod_msra_gemm_candidates(const void *roi_unit_bf16, const void *ref_unit_bf16, uint32_t *cand_out, int NP,
                                        int C, int T, int HW, vod_stream_t stream) {
    VOD_REQUIRE(roi_unit_bf16 && ref_unit_bf16 && cand_out, "vod_msra_gemm_candidates: null pointer");
    if (!msra_gemm_supported(NP, C, T, HW) || !vod_device_is_sm100())
        return fail(VOD_E_UNSUPPORTED, "vod_msra_gemm_candidates: needs C %% 64 == 0, C <= 512, HW <= 4096, sm_100");
    return msra_launch_gemm_topk(roi_unit_bf16, ref_unit_bf16, cand_out, NP, NP, C, T, HW, as_stream(stream));
}
