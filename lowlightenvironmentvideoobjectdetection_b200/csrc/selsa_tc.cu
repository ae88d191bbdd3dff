// selsa_tc.cu -- (3) SELSA similarity-softmax-weighted-sum as a fused tcgen05/TMEM kernel (sm_100a).
//
//   O_h = softmax(Q_h K_h^T * scale) V_h,   Q_h [N,64], K_h [M,64], V_h [M,64]   (per head h)
//   restates mmtracking/mmtrack/models/aggregators/selsa_aggregator.py:51-70 without the [heads,N,M] tensor.
//
// One CTA = (128 query rows, head, split of the reference axis).  192 threads:
//   warps 0-3  softmax: thread = query row; S chunk read from TMEM with tcgen05.ld, online softmax held
//              entirely in registers (row max / row sum are thread-local, no shuffles), probabilities
//              written to shared memory in the 128B-swizzled K-major operand layout.  O is NOT read back per
//              chunk: it stays in TMEM, exponentials are taken against a per-row reference maximum that is
//              raised (with a rescale of the row's O and row sum) only when a score outgrows it by 2^64.
//   warp 4     TMA producer: a ring of (K chunk, V^T chunk) tiles.  (Q never enters shared memory: the softmax warps load their
//              row from global memory once and park it in TMEM, the A operand of the TS-form S = Q K^T.)
//   warp 5     MMA issuer: S = Q K^T (kind::tf32 or kind::f16/bf16) into a double-buffered TMEM
//              accumulator, O += P V into a third TMEM accumulator (over all chunks); completion via tcgen05.commit.
// Reference rows are processed in chunks of 64.  When N*heads/128 CTAs cannot fill 148 SMs the reference
// axis is split across CTAs and the partial (acc, max, sum) triples are merged by selsa_merge_kernel.
//
// TMEM columns: kSBufs S / P tiles of 64 columns, then O (64 columns), then Q (fp32: 64 columns, bf16: 32).
#include <cuda_bf16.h>

#include "common.cuh"
#include "selsa.cuh"
#include "tc.cuh"

namespace vod {

constexpr int kBM = 128;
constexpr int kBN = 64;
constexpr int kHD = 64;
// K / V^T ring depth and CTAs per SM, measured with experiment builds (round 2, N x M = 300 x 4500 | 1000 x 31000, tf32, incl. the
// merge): 2 stages 56 | 765 us, 3 stages 37.9 | 406 us, 4 stages 35.8 | 360 us, 5 / 6 stages 35.8 | 364-367 us; 2 stages with two
// co-resident CTAs per SM (2 x 97 KB of shared memory, 2 x 256 TMEM columns, twice the splits) 39.9 | 407 us = the 3-stage
// single CTA: what counts is the number of K / V^T chunks in flight per SM, and beyond 4 the MMA <-> softmax handshake chain.
#ifndef VOD_SELSA_STAGES
#define VOD_SELSA_STAGES 4
#endif
#ifndef VOD_SELSA_OCC
#define VOD_SELSA_OCC 1
#endif
constexpr int kKvStages = VOD_SELSA_STAGES;
#ifndef VOD_SELSA_SBUFS
#define VOD_SELSA_SBUFS 2
#endif
constexpr int kSBufs = VOD_SELSA_SBUFS;                  // S / P tiles in TMEM (the MMA issuer runs kSBufs - 1 score tiles ahead)
constexpr int kSelsaThreads = 192;
constexpr int kSelsaTmemCols = (kSBufs + 2) * 64 <= 256 ? 256 : 512;
constexpr uint32_t kOCol = kSBufs * 64;                  // O accumulator: 64 TMEM columns after the score tiles
constexpr uint32_t kQCol = kOCol + 64;                   // then the CTA's 128 query rows (A operand of S = Q K^T)
static_assert((kSBufs + 2) * 64 <= 512 && kSBufs >= 2, "TMEM columns");
static_assert(VOD_SELSA_OCC * kSelsaTmemCols <= 512, "TMEM columns per SM");

template <bool BF16>
struct SelsaCfg {
    static constexpr int kElem = BF16 ? 2 : 4;
    static constexpr int kSliceElems = 128 / kElem;         // elements per 128-byte K slice
    static constexpr int kSlices = kHD / kSliceElems;       // slices along d (QK^T) and along refs (PV): kBN == kHD
    static constexpr int kKBytes = kSlices * kBN * 128;
    static constexpr int kVBytes = kSlices * kHD * 128;
    static constexpr int kSmem = kKvStages * (kKBytes + kVBytes) + 1024;   // neither Q nor P ever touches shared memory
};

struct SelsaParams {
    const void *q;     // [N, D] fp32 or bf16 (read directly by the softmax warps, which park their row in TMEM)
    float *out;        // [N, D]
    float *part_acc;   // [S][heads][Npad][64]
    float *part_m;     // [S][heads][Npad]
    float *part_l;     // [S][heads][Npad]
    int N, M, heads, D, Npad, splits, chunks_per_split, nchunks;
    float scale_log2;  // scale * log2(e)
};

__device__ __forceinline__ uint32_t f32_to_tf32(float x) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return r;
}

template <bool BF16>
__global__ void __launch_bounds__(kSelsaThreads, VOD_SELSA_OCC)
selsa_tc_kernel(const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_k,
                const __grid_constant__ CUtensorMap tm_v, const SelsaParams p) {
    using Cfg = SelsaCfg<BF16>;
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t *sK = smem;
    uint8_t *sV = sK + kKvStages * Cfg::kKBytes;
    __shared__ uint64_t q_full, kv_full[kKvStages], kv_empty[kKvStages], s_full[kSBufs], s_empty[kSBufs], p_full[kSBufs];
    __shared__ uint32_t tmem_slot;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int rt = blockIdx.x, h = blockIdx.y, split = blockIdx.z;
    const int c0 = split * p.chunks_per_split;
    const int n = min(p.chunks_per_split, p.nchunks - c0);  // chunks of this CTA (>= 1 by construction)

    if (threadIdx.x == 0) {
        tc::mbar_init(&q_full, 4);                         // the four softmax warps, once their Q rows are in TMEM
        for (int i = 0; i < kKvStages; ++i) { tc::mbar_init(&kv_full[i], 1); tc::mbar_init(&kv_empty[i], 1); }
        for (int i = 0; i < kSBufs; ++i) {
            tc::mbar_init(&s_full[i], 1); tc::mbar_init(&s_empty[i], 1);
            tc::mbar_init(&p_full[i], 128);
        }
        tc::fence_barrier_init();
    }
    if (warp == 5) tc::tmem_alloc(&tmem_slot, kSelsaTmemCols);
    tc::tcgen05_fence_before();
    __syncthreads();
    tc::tcgen05_fence_after();
    const uint32_t tmem = tmem_slot;

    if (warp == 4) {
        // ------------------------------------------------------------------ TMA producer
        if (tc::elect_one()) {
            tc::tma_prefetch_desc(&tm_k); tc::tma_prefetch_desc(&tm_v);
            for (int j = 0; j < n; ++j) {
                const int st = j % kKvStages;
                const uint32_t ph = (j / kKvStages) & 1;
                tc::mbar_wait(&kv_empty[st], ph ^ 1);
                tc::mbar_arrive_expect_tx(&kv_full[st], Cfg::kKBytes + Cfg::kVBytes);
                const int r0 = (c0 + j) * kBN;
                for (int sl = 0; sl < Cfg::kSlices; ++sl) {
                    tc::tma_load_2d(sK + st * Cfg::kKBytes + sl * kBN * 128, &tm_k, &kv_full[st],
                                    h * kHD + sl * Cfg::kSliceElems, r0);
                    tc::tma_load_2d(sV + st * Cfg::kVBytes + sl * kHD * 128, &tm_v, &kv_full[st],
                                    r0 + sl * Cfg::kSliceElems, h * kHD);
                }
            }
        }
    } else if (warp == 5) {
        // ------------------------------------------------------------------ MMA issuer
        constexpr uint32_t idesc = tc::umma_idesc(BF16 ? tc::kFmtBF16 : tc::kFmtTF32, kBM, kBN);  // M=128, N=64 for both GEMMs
        auto issue_s = [&](int j) {
            const int st = j % kKvStages, buf = j % kSBufs;
            tc::mbar_wait(&kv_full[st], (j / kKvStages) & 1);
            tc::mbar_wait(&s_empty[buf], ((j / kSBufs) & 1) ^ 1);
            tc::tcgen05_fence_after();
            if (tc::elect_one()) {
                // S = Q K^T with Q read from TENSOR MEMORY (TS form, as P below): from shared memory every one of these small
                // MMAs (M = 128, N = 64, K = 32 bytes) re-read a 4 KB slice of Q next to 2 KB of K -- 6 KB per 131 kFLOP against
                // 128 B/clk of shared-memory bandwidth, more than the MMA's own time
                const uint32_t ka = tc::smem_u32(sK + st * Cfg::kKBytes);
#pragma unroll
                for (int sl = 0; sl < Cfg::kSlices; ++sl)
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        const uint64_t kb = tc::umma_desc_k_sw128(ka + sl * kBN * 128 + k * 32);
                        if (BF16) tc::umma_f16_ts(tmem + buf * kBN, tmem + kQCol + (sl * 4 + k) * 8, kb, idesc, (sl | k) != 0);
                        else tc::umma_tf32_ts(tmem + buf * kBN, tmem + kQCol + (sl * 4 + k) * 8, kb, idesc, (sl | k) != 0);
                    }
                tc::umma_commit(&s_full[buf]);
            }
            __syncwarp();
        };
        tc::mbar_wait(&q_full, 0);
        tc::tcgen05_fence_after();
        for (int j = 0; j < kSBufs - 1 && j < n; ++j) issue_s(j);
        for (int j = 0; j < n; ++j) {
            if (j + kSBufs - 1 < n) issue_s(j + kSBufs - 1);
            const int st = j % kKvStages, buf = j % kSBufs;
            tc::mbar_wait(&p_full[buf], (j / kSBufs) & 1);   // also orders any rescale of O (done before the P store) before this PV
            tc::tcgen05_fence_after();
            if (tc::elect_one()) {
                // O += P V with P read from TENSOR MEMORY (TS form): the softmax warps wrote it over the S tile in place, so
                // the probabilities never pass through shared memory (that round trip was 64 KB of the 160 KB of
                // shared-memory traffic per 64-row chunk which bounded the first versions of this kernel)
                const uint32_t va = tc::smem_u32(sV + st * Cfg::kVBytes);
                constexpr int kSteps = 4 * Cfg::kSlices;   // MMAs of 32 bytes of K: 8 tf32 or 16 bf16 reference rows = 8 TMEM columns
#pragma unroll
                for (int i = 0; i < kSteps; ++i) {
                    const uint64_t vb = tc::umma_desc_k_sw128(va + (i >> 2) * kHD * 128 + (i & 3) * 32);
                    if (BF16) tc::umma_f16_ts(tmem + kOCol, tmem + buf * kBN + i * 8, vb, idesc, (j | i) != 0);
                    else tc::umma_tf32_ts(tmem + kOCol, tmem + buf * kBN + i * 8, vb, idesc, (j | i) != 0);   // O accumulates over all chunks
                }
                tc::umma_commit(&kv_empty[st]);
                // the S/P tile may be overwritten by S(j + kSBufs); the same barrier tells the softmax warps that PV(j) -- and, the
                // tensor pipe being in order, every earlier PV -- has landed (a per-chunk "O ready" barrier would alias in parity
                // once the softmax warps run more than one chunk ahead of the PV products)
                tc::umma_commit(&s_empty[buf]);
            }
            __syncwarp();
        }
    } else {
        // ------------------------------------------------------------------ softmax / accumulate (warps 0-3)
        const int r = warp * 32 + lane;                       // row within the tile == TMEM lane
        const uint32_t tl = tmem + ((uint32_t)(warp * 32) << 16);
        // O stays in TENSOR MEMORY for the whole pass (the PV MMAs accumulate over all chunks) and the exponentials are
        // taken against a per-row reference maximum m_ref that is only raised -- with a rescale of the row's O and l -- when
        // a new score exceeds it by more than 2^kLazy (FlashAttention-4 style lazy rescaling).  With that the softmax
        // warps never wait for a PV product in the steady state: the first version read every chunk's product back
        // (acc = acc * corr + O_chunk), a commit -> mbarrier -> tcgen05.ld round trip per 64 reference rows that made
        // the kernel latency-bound (1970 clk per chunk with the tensor pipe 19 % busy; removing 7/8 of the MMAs,
        // halving the instructions or doubling the softmax warps did not change its time).
        // Instruction diet as before: packed fp32x2 FMA/ADD, ex2.approx on the raw MUFU, scale folded into the FMA.
        {
            // This thread's query row -> TMEM lane r, columns [kQCol, kQCol + 64): one tf32 (fp32 bits, the MMA ignores the low 13)
            // or two bf16 per 32-bit column, the A-operand layout of the TS-form MMA.  Rows past N are zero.
            const int qrow = rt * kBM + r;
            constexpr int kWords = kHD * Cfg::kElem / 4;       // 32-bit words per row: 64 (fp32) or 32 (bf16)
            const uint4 *src = reinterpret_cast<const uint4 *>(reinterpret_cast<const uint8_t *>(p.q) +
                                                               ((size_t)qrow * p.D + (size_t)h * kHD) * Cfg::kElem);
#pragma unroll
            for (int c = 0; c < kWords; c += 16) {
                uint32_t v[16];
#pragma unroll
                for (int q4 = 0; q4 < 4; ++q4) {
                    const uint4 w4 = qrow < p.N ? __ldg(src + c / 4 + q4) : make_uint4(0u, 0u, 0u, 0u);
                    v[4 * q4] = w4.x; v[4 * q4 + 1] = w4.y; v[4 * q4 + 2] = w4.z; v[4 * q4 + 3] = w4.w;
                }
                if (!BF16) {
                    // an A operand read from TMEM is TRUNCATED to tf32 by the MMA (measured: parked as raw fp32 bits the output
                    // error at N x M = 1000 x 31000 doubled, 6e-4 -> 1.3e-3): round to nearest here, once per row
#pragma unroll
                    for (int i = 0; i < 16; ++i) v[i] = f32_to_tf32(__uint_as_float(v[i]));
                }
                tc::tmem_st_32x16(tl + kQCol + c, v);
            }
            tc::tmem_st_wait();
            tc::tcgen05_fence_before();
            __syncwarp();
            if (lane == 0) tc::mbar_arrive(&q_full);
        }
        constexpr float kLazy = 64.f;   // P <= 2^64: no overflow in fp32 / tf32 / bf16, same relative precision
        float m_ref = -INFINITY, l_run = 0.f;
        auto ex2 = [](float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; };

        for (int j = 0; j < n; ++j) {
            const int buf = j % kSBufs;
            tc::mbar_wait(&s_full[buf], (j / kSBufs) & 1);
            tc::tcgen05_fence_after();
            uint32_t s0[32], s1[32];
            tc::tmem_ld_32x32(tl + buf * kBN, s0);
            tc::tmem_ld_32x32(tl + buf * kBN + 32, s1);
            tc::tmem_ld_wait();

            const int valid = min(kBN, p.M - (c0 + j) * kBN);  // reference rows of this chunk that exist
            float s[kBN];
#pragma unroll
            for (int i = 0; i < 32; ++i) { s[i] = __uint_as_float(s0[i]); s[32 + i] = __uint_as_float(s1[i]); }
            if (valid < kBN) {   // tail chunk only (warp-uniform)
#pragma unroll
                for (int i = 0; i < kBN; ++i) s[i] = i < valid ? s[i] : -INFINITY;
            }
            // Row maximum as four independent chains (a single chain is 32 dependent 3-input FMNMX, ~150 clk that nothing hides:
            // each sub-partition holds ONE softmax warp).
            float mx4[4] = {s[0], s[1], s[2], s[3]};
#pragma unroll
            for (int i = 4; i < kBN; ++i) mx4[i & 3] = fmaxf(mx4[i & 3], s[i]);
            const float m_new = fmaxf(fmaxf(mx4[0], mx4[1]), fmaxf(mx4[2], mx4[3])) * p.scale_log2;   // scale > 0: max commutes with the scaling
            if (j == 0) m_ref = m_new;                             // O is still empty (PV(0) overwrites it)
            // The exponentials are taken against the CURRENT reference maximum right away; whether some row outgrew it
            // (m_new > m_ref + 2^kLazy: rare) is only looked at afterwards, so the maximum is off the critical path.  A row that
            // did outgrow it has produced garbage (possibly inf) here and is recomputed below.
            const float2 sc2 = make_float2(p.scale_log2, p.scale_log2);
            float e[kBN];
            float2 sum2 = make_float2(0.f, 0.f);
            {
                const float2 nm2 = make_float2(-m_ref, -m_ref);
#pragma unroll
                for (int i = 0; i < kBN; i += 2) {
                    const float2 t = __ffma2_rn(make_float2(s[i], s[i + 1]), sc2, nm2);   // raw * scale - m_ref  (-inf stays -inf)
                    e[i] = ex2(t.x); e[i + 1] = ex2(t.y);
                    sum2 = __fadd2_rn(sum2, make_float2(e[i], e[i + 1]));
                }
            }
            if (j > 0 && __any_sync(0xffffffffu, m_new > m_ref + kLazy)) {
                // rare: some row of this warp outgrew its reference maximum.  All PV products issued so far (chunks < j) must
                // have landed; PV(j) cannot start before this thread's p_full arrival below.  tcgen05.ld/st are warp-wide,
                // rows that do not need it rescale by 1.
                tc::mbar_wait(&s_empty[(j - 1) % kSBufs], ((j - 1) / kSBufs) & 1);   // PV(j - 1) and all before it
                tc::tcgen05_fence_after();
                const bool need = m_new > m_ref + kLazy;
                const float f = need ? ex2(m_ref - m_new) : 1.0f;
#pragma unroll
                for (int half = 0; half < 2; ++half) {
                    uint32_t o[32];
                    tc::tmem_ld_32x32(tl + kOCol + half * 32, o);
                    tc::tmem_ld_wait();
                    uint32_t lo[16], hi[16];
#pragma unroll
                    for (int i = 0; i < 16; ++i) {
                        lo[i] = __float_as_uint(__uint_as_float(o[i]) * f);
                        hi[i] = __float_as_uint(__uint_as_float(o[16 + i]) * f);
                    }
                    tc::tmem_st_32x16(tl + kOCol + half * 32, lo);
                    tc::tmem_st_32x16(tl + kOCol + half * 32 + 16, hi);
                }
                tc::tmem_st_wait();
                tc::tcgen05_fence_before();
                l_run *= f;
                if (need) m_ref = m_new;
                // the exponentials again, against the raised maximum (rows that did not need it reproduce the same values)
                const float2 nm2 = make_float2(-m_ref, -m_ref);
                sum2 = make_float2(0.f, 0.f);
#pragma unroll
                for (int i = 0; i < kBN; i += 2) {
                    const float2 t = __ffma2_rn(make_float2(s[i], s[i + 1]), sc2, nm2);
                    e[i] = ex2(t.x); e[i + 1] = ex2(t.y);
                    sum2 = __fadd2_rn(sum2, make_float2(e[i], e[i + 1]));
                }
            }
            l_run += sum2.x + sum2.y;
#pragma unroll
            for (int i = 0; i < kBN; ++i) s[i] = e[i];

            // P overwrites this row's S values in place (thread = row = TMEM lane): tf32 one value per column, bf16 two
            if (BF16) {
                uint32_t pk[32];
#pragma unroll
                for (int i = 0; i < 32; ++i) {
                    __nv_bfloat162 h2 = __floats2bfloat162_rn(s[2 * i], s[2 * i + 1]);   // .x (even reference row) in the low half
                    pk[i] = *reinterpret_cast<uint32_t *>(&h2);
                }
                uint32_t a16[16], b16[16];
#pragma unroll
                for (int i = 0; i < 16; ++i) { a16[i] = pk[i]; b16[i] = pk[16 + i]; }
                tc::tmem_st_32x16(tl + buf * kBN, a16);
                tc::tmem_st_32x16(tl + buf * kBN + 16, b16);
            } else {
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    uint32_t c16[16];
                    // + half a tf32 ulp: the MMA truncates the low 13 bits, together = cvt.rna.tf32 (one IADD instead of two ops)
#pragma unroll
                    for (int i = 0; i < 16; ++i) c16[i] = __float_as_uint(s[q * 16 + i]) + 0x1000u;
                    tc::tmem_st_32x16(tl + buf * kBN + q * 16, c16);
                }
            }
            tc::tmem_st_wait();
            tc::tcgen05_fence_before();
            tc::mbar_arrive(&p_full[buf]);
        }
        // the finished accumulator: all PV products have landed once the last commit fires
        tc::mbar_wait(&s_empty[(n - 1) % kSBufs], ((n - 1) / kSBufs) & 1);
        tc::tcgen05_fence_after();
        float acc[kHD];
        {
            uint32_t o0[32], o1[32];
            tc::tmem_ld_32x32(tl + kOCol, o0);
            tc::tmem_ld_32x32(tl + kOCol + 32, o1);
            tc::tmem_ld_wait();
            tc::tcgen05_fence_before();
#pragma unroll
            for (int i = 0; i < 32; ++i) { acc[i] = __uint_as_float(o0[i]); acc[32 + i] = __uint_as_float(o1[i]); }
        }
        const float m_run = m_ref;

        const int row = rt * kBM + r;
        if (row < p.N) {
            if (p.splits == 1) {
                const float inv = 1.0f / l_run;
                float *dst = p.out + (size_t)row * p.D + h * kHD;
#pragma unroll
                for (int i = 0; i < kHD; i += 4)
                    *reinterpret_cast<float4 *>(dst + i) =
                        make_float4(acc[i] * inv, acc[i + 1] * inv, acc[i + 2] * inv, acc[i + 3] * inv);
            } else {
                const size_t pr = ((size_t)split * p.heads + h) * p.Npad + row;
                float *dst = p.part_acc + pr * kHD;
#pragma unroll
                for (int i = 0; i < kHD; i += 4)
                    *reinterpret_cast<float4 *>(dst + i) = make_float4(acc[i], acc[i + 1], acc[i + 2], acc[i + 3]);
                p.part_m[pr] = m_run;
                p.part_l[pr] = l_run;
            }
        }
    }
    tc::tcgen05_fence_before();
    __syncthreads();
    if (warp == 5) tc::tmem_dealloc(tmem, kSelsaTmemCols);
}

// out[n, h*64 + c] = sum_s acc_s * 2^(m_s - m*) / sum_s l_s * 2^(m_s - m*)
__global__ void __launch_bounds__(256)
selsa_merge_kernel(const SelsaParams p) {
    const long idx = (long)blockIdx.x * 256 + threadIdx.x;  // (n, h, c4) with c4 = group of 4 channels
    const long total = (long)p.N * p.heads * (kHD / 4);
    if (idx >= total) return;
    const int c4 = (int)(idx % (kHD / 4));
    const int h = (int)((idx / (kHD / 4)) % p.heads);
    const int n = (int)(idx / ((kHD / 4) * p.heads));
    float mstar = -INFINITY;
    for (int s = 0; s < p.splits; ++s) mstar = fmaxf(mstar, p.part_m[((size_t)s * p.heads + h) * p.Npad + n]);
    float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
    float l = 0.f;
    for (int s = 0; s < p.splits; ++s) {
        const size_t pr = ((size_t)s * p.heads + h) * p.Npad + n;
        const float w = exp2f(p.part_m[pr] - mstar);
        const float4 v = *reinterpret_cast<const float4 *>(p.part_acc + pr * kHD + c4 * 4);
        a.x = fmaf(v.x, w, a.x); a.y = fmaf(v.y, w, a.y); a.z = fmaf(v.z, w, a.z); a.w = fmaf(v.w, w, a.w);
        l = fmaf(p.part_l[pr], w, l);
    }
    const float inv = 1.0f / l;
    *reinterpret_cast<float4 *>(p.out + (size_t)n * p.D + h * kHD + c4 * 4) = make_float4(a.x * inv, a.y * inv, a.z * inv, a.w * inv);
}

// [M, D] -> [D, ldv] transposition for callers that hand V row-major (v_layout 0)
__global__ void __launch_bounds__(256)
transpose_rows_kernel(const uint8_t *__restrict__ in, uint8_t *__restrict__ out, int M, int D, int ldv, int elem) {
    __shared__ uint32_t tile[32][33];
    const int m0 = blockIdx.x * 32, d0 = blockIdx.y * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    for (int i = ty; i < 32; i += 8) {
        const int m = m0 + i, d = d0 + tx;
        uint32_t v = 0;
        if (m < M && d < D) v = elem == 4 ? reinterpret_cast<const uint32_t *>(in)[(size_t)m * D + d]
                                          : reinterpret_cast<const uint16_t *>(in)[(size_t)m * D + d];
        tile[i][tx] = v;
    }
    __syncthreads();
    for (int i = ty; i < 32; i += 8) {
        const int d = d0 + i, m = m0 + tx;
        if (d < D && m < ldv) {
            const uint32_t v = m < M ? tile[tx][i] : 0u;
            if (elem == 4) reinterpret_cast<uint32_t *>(out)[(size_t)d * ldv + m] = v;
            else reinterpret_cast<uint16_t *>(out)[(size_t)d * ldv + m] = (uint16_t)v;
        }
    }
}

// Splits over the reference axis: as many as fill the machine once.  (Measured, round 2, 1000 x 31000 = 128 units of 485 chunks:
// one split leaves 20 of 148 SMs idle, 8 splits run 6.9 waves of 61 chunks -- 12 % fewer chunk-times per SM -- and take the same
// 351-353 us: a CTA's fixed cost (TMEM allocation, Q load, pipeline fill, partial write, the merge) is worth ~8 chunks.)
static int pick_splits(int N, int M, int heads) {
    const int units = ceil_div(N, kBM) * heads;
    const int nchunks = ceil_div(M, kBN);
    int s = max(1, VOD_SELSA_OCC * num_sms() / units);
    s = min(s, max(1, nchunks / 4));  // at least 4 chunks per split
    return min(s, 16);
}

struct SelsaWs {
    size_t vt_off, acc_off, m_off, l_off, bytes;
    int splits, Npad, ldv;
};
static SelsaWs selsa_ws_layout(int N, int M, int heads) {
    SelsaWs w;
    w.splits = pick_splits(N, M, heads);
    w.Npad = ceil_div(N, kBM) * kBM;
    w.ldv = (int)align_up((size_t)M, 8);
    size_t o = 0;
    w.vt_off = o;  o = align_up(o + (size_t)heads * kHD * w.ldv * 4, 1024);   // worst case fp32 V^T copy
    w.acc_off = o; o = align_up(o + (size_t)w.splits * heads * w.Npad * kHD * 4, 256);
    w.m_off = o;   o = align_up(o + (size_t)w.splits * heads * w.Npad * 4, 256);
    w.l_off = o;   o = align_up(o + (size_t)w.splits * heads * w.Npad * 4, 256);
    w.bytes = o;
    return w;
}

bool selsa_tc_supported(int N, int M, int heads, int d, int dtype, const void *q, const void *k, const void *v,
                        int v_layout, int ldv) {
    if (d != kHD || N <= 0 || M <= 0) return false;
    const int eb = dtype == VOD_DTYPE_BF16 ? 2 : 4;
    if (((size_t)heads * d * eb) & 15) return false;
    if ((reinterpret_cast<uintptr_t>(q) | reinterpret_cast<uintptr_t>(k) | reinterpret_cast<uintptr_t>(v)) & 15) return false;
    if (v_layout == 1 && (((size_t)ldv * eb) & 15)) return false;
    if (ceil_div(N, kBM) > 65535 || heads > 65535) return false;
    return true;
}

size_t selsa_tc_workspace_bytes(int N, int M, int heads, int d) {
    if (d != kHD || N <= 0 || M <= 0) return 256;
    return selsa_ws_layout(N, M, heads).bytes;
}

template <bool BF16>
static int launch_tc(const void *q, const void *k, const void *vt, int ldv, float *out, int N, int M, int heads,
                     float scale, const SelsaWs &w, uint8_t *ws, cudaStream_t st) {
    using Cfg = SelsaCfg<BF16>;
    const int D = heads * kHD;
    CUtensorMap tq, tk, tv;
    int rc;
    if ((rc = make_tmap_2d_sw128(&tq, q, Cfg::kElem, N, D, (uint64_t)D * Cfg::kElem, kBM))) return rc;
    if ((rc = make_tmap_2d_sw128(&tk, k, Cfg::kElem, M, D, (uint64_t)D * Cfg::kElem, kBN))) return rc;
    // V^T [D, M] with row stride ldv: columns >= M are out of bounds -> zero fill
    if ((rc = make_tmap_2d_sw128(&tv, vt, Cfg::kElem, D, M, (uint64_t)ldv * Cfg::kElem, kHD))) return rc;
    SelsaParams p;
    p.q = q;
    p.out = out;
    p.part_acc = reinterpret_cast<float *>(ws + w.acc_off);
    p.part_m = reinterpret_cast<float *>(ws + w.m_off);
    p.part_l = reinterpret_cast<float *>(ws + w.l_off);
    p.N = N; p.M = M; p.heads = heads; p.D = D; p.Npad = w.Npad;
    p.nchunks = ceil_div(M, kBN);
    p.chunks_per_split = ceil_div(p.nchunks, w.splits);
    p.splits = ceil_div(p.nchunks, p.chunks_per_split);  // no empty split
    p.scale_log2 = scale * 1.4426950408889634f;
    auto kern = selsa_tc_kernel<BF16>;
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmem);
    dim3 grid(ceil_div(N, kBM), heads, p.splits);
    kern<<<grid, kSelsaThreads, Cfg::kSmem, st>>>(tq, tk, tv, p); note_launch();
    if (p.splits > 1) {
        const long total = (long)N * heads * (kHD / 4);
        selsa_merge_kernel<<<(unsigned)ceil_div(total, 256L), 256, 0, st>>>(p); note_launch();
    }
    return check_launch("vod_selsa_attn(tcgen05)");
}

int selsa_tc_launch(const void *q, const void *k, const void *v, float *out, int N, int M, int heads, int d, float scale,
                    int dtype, int v_layout, int ldv, void *ws, size_t ws_bytes, cudaStream_t st) {
    (void)d;
    const SelsaWs w = selsa_ws_layout(N, M, heads);
    if (!ws || ws_bytes < w.bytes) return fail(VOD_E_WORKSPACE, "vod_selsa_attn: workspace %zu < %zu", ws_bytes, w.bytes);
    uint8_t *wsb = reinterpret_cast<uint8_t *>(ws);
    if ((reinterpret_cast<uintptr_t>(ws) & 1023) != 0) return fail(VOD_E_BADARG, "vod_selsa_attn: workspace must be 1024-byte aligned");
    const int eb = dtype == VOD_DTYPE_BF16 ? 2 : 4;
    const void *vt = v;
    int ld = ldv;
    if (v_layout == 0) {
        const int D = heads * kHD;
        dim3 grid(ceil_div(w.ldv, 32), ceil_div(D, 32));
        transpose_rows_kernel<<<grid, 256, 0, st>>>(reinterpret_cast<const uint8_t *>(v), wsb + w.vt_off, M, D, w.ldv, eb); note_launch();
        vt = wsb + w.vt_off;
        ld = w.ldv;
    }
    if (dtype == VOD_DTYPE_BF16) return launch_tc<true>(q, k, vt, ld, out, N, M, heads, scale, w, wsb, st);
    return launch_tc<false>(q, k, vt, ld, out, N, M, heads, scale, w, wsb, st);
}

}  // namespace vod
