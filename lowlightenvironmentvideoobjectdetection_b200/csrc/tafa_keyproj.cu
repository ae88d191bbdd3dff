// tafa_keyproj.cu -- TemporalRoIAlign's attention logits WITHOUT the reference-frame embeddings.
//
// The reference (mmtracking/mmtrack/models/roi_heads/roi_extractors/temporal_roi_align.py:72-93) runs its 3x3 embed conv over
// all (T+1)*N RoI patches (1.11 TFLOP at N=300, T=15) and then only ever uses the result through the per-head dot product with
// the KEY patch's embedding.  The conv is linear, so that dot product can be taken on the weight side first:
//
//   logit[t,n,p,h] = sum_{o in head h} ek[n,p,o] * (conv(x_t)[n,p,o] + b[o])
//                  = sum_{tap, c} x_t[n, p+tap, c] * G[h, (n,p), tap, c]  +  (a term constant in t, cancelled by the softmax)
//   G[h, (n,p), tap, c] = sum_{o in head h} ek[n,p,o] * W[o, c, tap]          (one batched library GEMM, 69 GFLOP)
//
// Only the key patches go through the conv (1/(T+1) of the work); this kernel contracts the raw RoI features x_all with G:
// 3.6 G fp32 FMA on the CUDA cores and one streaming read of G (1.08 GB) and x_all (0.48 GB) -- HBM-bound, not a GEMM.
//
//   Work unit ("tile") = (RoI n, chunk of 32 input channels, block of 16 frames): its x_all slice [16][P][32] (98 KB) arrives
//   in shared memory as ONE 3-D TMA box and every element is reused by up to 9 taps of the neighbouring output positions.
//   warp = output position p; lane = (tg, cg): tg owns 4 frames, cg owns 4 consecutive channels (128-bit accesses).
//   G has no reuse but is 70 % of the bytes: each consumer warp streams the four head rows (9 taps x 32 channels, 1152
//   contiguous bytes each) of its NEXT positions into a private shared-memory ring with cp.async.bulk + mbarrier (no
//   registers held while in flight, 129 KB of ring per SM).
//   Per (tap, 32 channels): 4 G vectors x 4 frame vectors from shared memory -> 32 packed FFMA2 per lane.
//   The (frame, head) partial sums are reduced over the 8 cg lanes by a transposing butterfly (14 shuffles) and written as
//   per-chunk partial logits [C/32][N][P][H][T1]; the weighting kernel (tafa.cu) sums the chunks -- deterministic, no atomics.
//
//   Shipped for more than 8 frames: the persistent kernel <16, 1> -- one CTA per SM walks a contiguous range of tiles, a
//   producer warp requests the next frame tile as soon as every consumer warp has left the current one, and the consumer warps'
//   position stream (and with it the G rings) runs on across tile boundaries, so the first ring fill is paid once per SM instead
//   of once per tile.  Up to 8 frames: one CTA per 8-frame tile, two CTAs per SM (HBM-bound at 5.6 TB/s).  The other variants
//   (one CTA per 16-frame tile, persistent double-buffered 8-frame tiles) stay selectable through VOD_KP_* for the record of the
//   measurements in DESIGN.md section 4.  Probes (VOD_KP_DBG) on the one-CTA-per-tile version: memory alone 255 us, FMAs alone
//   160 us, per-CTA start-up 175 us; the ncu source page puts ~500 of the ~750 warp instructions per position outside the
//   FFMA2 / LDS core (bulk-copy issue, mbarrier waits, butterfly, tap control) -- the next thing to cut.
#include <stdlib.h>

#include <cuda_bf16.h>

#include "common.cuh"
#include "tc.cuh"

namespace vod {

constexpr int kKpHeads = 4;
constexpr int kKpCC = 32;                         // channels per tile
constexpr int kKpRowFloats = 9 * kKpCC;           // ELEMENTS of one (head, position) row of G for this chunk
constexpr int kKpStageFloats = kKpHeads * kKpRowFloats;   // elements of one ring stage (4 head rows)
// G element size GE: 4 = fp32, 2 = bf16 (round 2: G is 70 % of the kernel's bytes and a library GEMM's output; when the caller
// allows reduced-precision library math it is produced and streamed in bf16 -- half the bytes here and in the GEMM that writes
// it -- and unpacked to fp32 on the fly: one shift / mask per element, amortised over the lane's 4 frames).
constexpr int kKpMaxWarps = 16;
constexpr int kKpMaxWarpsMma = 21;                // tensor-core form: up to 20 consumer warps + the producer at 96 registers

__device__ __forceinline__ void bulk_g2s(void *smem_dst, const void *gsrc, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(tc::smem_u32(smem_dst)), "l"(gsrc), "r"(bytes), "r"(tc::smem_u32(bar)) : "memory");
}

// 3x3 window, pad 1: bit (ky*3 + kx) of the mask is set when the tap of output position p falls inside the ph x pw patch
__device__ __forceinline__ unsigned kp_tap_mask(int p, int ph, int pw) {
    const int py = p / pw, px = p - py * pw;
    const unsigned rows_ok = (py > 0 ? 1u : 0u) | 2u | (py < ph - 1 ? 4u : 0u);
    const unsigned cols_ok = (px > 0 ? 1u : 0u) | 2u | (px < pw - 1 ? 4u : 0u);
    unsigned m = 0;
#pragma unroll
    for (int tap = 0; tap < 9; ++tap) m |= ((rows_ok >> (tap / 3)) & (cols_ok >> (tap % 3)) & 1u) << tap;
    return m;
}

// One output position: contract the 3x3 neighbourhood of the position at `xp` (this lane's (tg, cg) slice of it; row_floats /
// frame_floats = distance to the patch row below / to the next frame in the shared-memory tile) with the four G rows in `gs`; returns the NV/8 sums this lane owns after the butterfly: v = cg*NV/8 + k, frame = v / H, head = v % H.
// 4 consecutive G elements at element offset `e` of a ring stage (fp32: one 128-bit load; bf16: one 64-bit load + unpack)
template <int GE>
__device__ __forceinline__ float4 kp_load_g(const unsigned char *__restrict__ stage, int e) {
    if (GE == 4) return *reinterpret_cast<const float4 *>(stage + (size_t)e * 4);
    const uint2 w = *reinterpret_cast<const uint2 *>(stage + (size_t)e * 2);
    return make_float4(__uint_as_float(w.x << 16), __uint_as_float(w.x & 0xffff0000u),
                       __uint_as_float(w.y << 16), __uint_as_float(w.y & 0xffff0000u));
}

template <int FR, int GE>
__device__ __forceinline__ void kp_position(const float *__restrict__ xp, const unsigned char *__restrict__ gs, int row_floats,
                                            int frame_floats, unsigned taps, int cg, float (&vc)[FR * kKpHeads / 8]) {
    constexpr int H = kKpHeads, CC = kKpCC, NV = FR * H;

    // packed fp32x2 FMAs (FFMA2: the full-rate fp32 path of sm_100): even / odd channels accumulate separately
    float2 acc2[FR][H];
#pragma unroll
    for (int u = 0; u < FR; ++u)
#pragma unroll
        for (int h = 0; h < H; ++h) acc2[u][h] = make_float2(0.f, 0.f);
#pragma unroll
    for (int tap = 0; tap < 9; ++tap) {
        if (!((taps >> tap) & 1u)) continue;                // warp-uniform
        const float *xq = xp + (tap / 3 - 1) * row_floats + (tap % 3 - 1) * CC;
        float4 g[H];
#pragma unroll
        for (int h = 0; h < H; ++h) g[h] = kp_load_g<GE>(gs, h * kKpRowFloats + tap * CC + cg * 4);
#pragma unroll
        for (int u = 0; u < FR; ++u) {
            const float4 xv = *reinterpret_cast<const float4 *>(xq + u * frame_floats);
#pragma unroll
            for (int h = 0; h < H; ++h) {
                acc2[u][h] = __ffma2_rn(make_float2(xv.x, xv.y), make_float2(g[h].x, g[h].y), acc2[u][h]);
                acc2[u][h] = __ffma2_rn(make_float2(xv.z, xv.w), make_float2(g[h].z, g[h].w), acc2[u][h]);
            }
        }
    }
    float acc[FR][H];
#pragma unroll
    for (int u = 0; u < FR; ++u)
#pragma unroll
        for (int h = 0; h < H; ++h) acc[u][h] = acc2[u][h].x + acc2[u][h].y;

    // transposing butterfly over the 8 cg lanes: NV sums -> NV/8 per lane
    float va[NV / 2], vb[NV / 4];
    {
        const bool up = cg & 4;
#pragma unroll
        for (int k = 0; k < NV / 2; ++k) {
            const float lo = acc[k / H][k % H], hi = acc[(k + NV / 2) / H][(k + NV / 2) % H];
            const float send = up ? lo : hi, keep = up ? hi : lo;
            va[k] = keep + __shfl_xor_sync(0xffffffffu, send, 4);
        }
    }
    {
        const bool up = cg & 2;
#pragma unroll
        for (int k = 0; k < NV / 4; ++k) {
            const float send = up ? va[k] : va[k + NV / 4], keep = up ? va[k + NV / 4] : va[k];
            vb[k] = keep + __shfl_xor_sync(0xffffffffu, send, 2);
        }
    }
    {
        const bool up = cg & 1;
#pragma unroll
        for (int k = 0; k < NV / 8; ++k) {
            const float send = up ? vb[k] : vb[k + NV / 8], keep = up ? vb[k + NV / 8] : vb[k];
            vc[k] = keep + __shfl_xor_sync(0xffffffffu, send, 1);
        }
    }
}

// ------------------------------------------------------------------------------------------------ tensor-core form
// Reduced-precision mode only (bf16 G, i.e. the caller allowed tf32 library math).  Per output position and 32-channel chunk the
// contraction is a [4 heads] x [taps * 32 channels] x [16 frames] product whose BOTH operands change with the position -- a
// batched small product, not a GEMM tile, so tcgen05's 128-row atoms do not apply; the warp-level m16n8k8 tf32 MMA does
// (measured on B200: one per 8 clocks per sub-partition, 512 MAC/clk/SM -- scripts/micro/mma_rate.cu):
//   A (16 rows) = heads (rows 0-3 used: the other rows compute sums nobody reads), B (8 columns) = frames (two column groups
//   for 16 frames), K = 8 channels of one tap; 8 MMAs per tap.
// What it buys is instruction issue, not flops: ~22 warp instructions per tap (4 LDS.128 + 2 LDS.64 + 8 unpack + 8 HMMA) and
// no butterfly, against ~80 for the FFMA2 form (which ncu shows issue-bound at ~750 warp instructions per position).
//   x tile: TMA box per patch row [16 frames][PW][32 ch], 128-byte swizzled and rounded to tf32 by the TMA unit.  Lane
//   (g = lane / 4, t = lane % 4) reads channels 4t..4t+3 (+16 s) of frames f(g) and f(g) + 8 with 128-bit loads, where
//   f(g) = g/2 + 4 (g & 1): the two frames of a quarter-warp then sit on 128-byte lines that differ in bit 2 of the line
//   index, so its 8 loads cover all 32 banks.  MMA m of a 16-channel step uses K slot t <-> channel 4t + 2m and slot
//   t + 4 <-> channel 4t + 2m + 1: the B fragment of an MMA is an aligned register PAIR of the 128-bit load (frames as the
//   A operand would need 16 register moves per tap) and one 64-bit (4 x bf16) G load feeds the A fragments of four MMAs.
//   G ring stage: the 4 head rows (576 B each) are placed 608 B apart, so the 4 heads' 64-bit loads hit different banks.
constexpr int kKpMmaHeadPitch = kKpRowFloats * 2 + 32;

__device__ __forceinline__ void mma_tf32_m16n8k8(float (&d)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, float b0, float b1) {
    asm("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
        : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
        : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(__float_as_uint(b0)), "r"(__float_as_uint(b1)));
}
__device__ __forceinline__ float4 lds128(uint32_t addr) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
    return v;
}
__device__ __forceinline__ uint2 lds64(uint32_t addr) {
    uint2 v;
    asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(addr));
    return v;
}

// xrow = shared-memory address of the box of patch row py of the tile (1024-byte aligned), gs = of the G ring stage; returns
// for head g (lanes with g < 4) the sums of frames t, t + 4, t + 8, t + 12.
template <int PW>
__device__ __forceinline__ void kp_position_mma(uint32_t xrow, uint32_t gs, int px, unsigned taps, int lane, float (&c)[4]) {
    constexpr int kRowB = 16 * PW * 128;                  // bytes of one patch-row box
    const int g = lane >> 2, t = lane & 3;
    const int f0 = (g >> 1) + ((g & 1) << 2);              // B column g <-> frames f0 and f0 + 8 (second column group)
    uint32_t off[3][2];                                    // [dx][column group]: address of channels 4t.. in patch row py
#pragma unroll
    for (int dx = 0; dx < 3; ++dx)
#pragma unroll
        for (int hf = 0; hf < 2; ++hf) {
            const int line = (f0 + 8 * hf) * PW + px + dx - 1;       // (-1 only for a masked tap)
            off[dx][hf] = xrow + (uint32_t)(line * 128 + (((t ^ line) & 7) << 4));
        }
    const uint32_t ga = gs + (uint32_t)((g & 3) * kKpMmaHeadPitch + t * 8);
    float acc[2][2][4];                                    // [column group][m]: independent accumulation chains
#pragma unroll
    for (int i = 0; i < 16; ++i) (&acc[0][0][0])[i] = 0.f;
#pragma unroll
    for (int tap = 0; tap < 9; ++tap) {
        if (!((taps >> tap) & 1u)) continue;                // warp-uniform
        const int dyb = (tap / 3 - 1) * kRowB;
#pragma unroll
        for (int s = 0; s < 2; ++s) {                       // channels 16 s .. 16 s + 15: chunk index ^ 4 <=> address ^ 64
            const float4 v0 = lds128((off[tap % 3][0] ^ (uint32_t)(s << 6)) + dyb);
            const float4 v1 = lds128((off[tap % 3][1] ^ (uint32_t)(s << 6)) + dyb);
            const uint2 w = lds64(ga + tap * 64 + s * 32);
            const uint32_t a00 = w.x << 16, a01 = w.x & 0xffff0000u, a10 = w.y << 16, a11 = w.y & 0xffff0000u;
            mma_tf32_m16n8k8(acc[0][0], a00, 0u, a01, 0u, v0.x, v0.y);
            mma_tf32_m16n8k8(acc[1][0], a00, 0u, a01, 0u, v1.x, v1.y);
            mma_tf32_m16n8k8(acc[0][1], a10, 0u, a11, 0u, v0.z, v0.w);
            mma_tf32_m16n8k8(acc[1][1], a10, 0u, a11, 0u, v1.z, v1.w);
        }
    }
    // C[row = head g][columns 2t, 2t+1 <-> frames t, t + 4]
    c[0] = acc[0][0][0] + acc[0][1][0];
    c[1] = acc[0][0][1] + acc[0][1][1];
    c[2] = acc[1][0][0] + acc[1][1][0];
    c[3] = acc[1][0][1] + acc[1][1][1];
}

// ------------------------------------------------------------------------------------------------ persistent kernel
// blockDim = (nwc + 1) warps: warps 0..nwc-1 consume, warp nwc produces the frame tiles.  Requires nwc <= P.
// <TB = 8, NBUF = 2>: 8-frame tiles, double-buffered (a warp may run one tile ahead of the slowest).
// <TB = 16, NBUF = 1>: 16-frame tiles (twice the FMAs per shared-memory load and per G byte), one buffer.
//   ROWP = false: the tile is one TMA box, requested when every warp has left the previous tile (ncu, round 2: the consumer
//     warps then sit ~30 % of their time on that barrier -- 98 KB per tile cannot hide behind a one-tile buffer).
//   ROWP = true (shipped): the tile is laid out [patch row][frame][pw][32] and travels as `ph` boxes (one per patch row, all
//     16 frames), each with its own full / empty barrier pair.  A warp walks its positions in increasing order, so once every
//     warp is past the positions that read patch row r (rows r-1 .. r+1 of the outputs), row r of the NEXT tile is requested
//     while rows r+1.. of the current one are still being used: the single buffer behaves like a rolling window and the
//     load of tile j+1 overlaps the FMAs of tile j.
constexpr int kKpMaxRows = 16;                    // patch rows with their own barrier pair (ROWP)

//   MPW > 0 (with ROWP, bf16 G): the tensor-core form above, for patches MPW positions wide.
template <int TB, int NBUF, int GE, bool ROWP, int MPW>
__global__ void __launch_bounds__((MPW > 0 ? kKpMaxWarpsMma : kKpMaxWarps) * 32, 1)
tafa_keyproj_persist_kernel(const __grid_constant__ CUtensorMap tm_x, const unsigned char *__restrict__ G, float *__restrict__ parts,
                            int T1, int N, int ph, int pw, int C, int depth, int tiles_per_cta, int dbg) {
    constexpr int H = kKpHeads, CC = kKpCC, FR = TB / 4, NV = FR * H;
    constexpr bool MMA = MPW > 0;
    static_assert(!ROWP || NBUF == 1, "row pipelining replaces the double buffer");
    static_assert(!MMA || (ROWP && GE == 2 && TB == 16), "the mma form reads bf16 G and per-row swizzled 16-frame boxes");
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    const int nch = C / CC, ntb = ceil_div(T1, TB), P = ph * pw;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwc = (blockDim.x >> 5) - 1;
    const int tg = lane >> 3, cg = lane & 7;
    const int total_tiles = N * nch * ntb;                              // tile id = (n*nch + chunk)*ntb + tb (host checks < 2^31)
    const int tile0 = blockIdx.x * tiles_per_cta;
    const int my_tiles = max(0, min(tiles_per_cta, total_tiles - tile0));
    const int tile_floats = TB * P * CC;
    // shared-memory tile: [TB][P][CC] (one box) or [ph][TB][pw][CC] (one box per patch row)
    const int row_floats = ROWP ? TB * pw * CC : pw * CC, frame_floats = ROWP ? pw * CC : P * CC;
    float *xs = reinterpret_cast<float *>(smem_raw);
    // one ring stage = the 4 head rows of one position (kRowBytes each, copied separately), kHeadPitch apart
    constexpr int kRowBytes = kKpRowFloats * GE, kHeadPitch = MMA ? kKpMmaHeadPitch : kRowBytes;
    constexpr int kStageBytes = H * kHeadPitch, kStageTx = H * kRowBytes;
    unsigned char *ring = reinterpret_cast<unsigned char *>(xs + NBUF * (size_t)tile_floats);   // [nwc][depth][H][kHeadPitch]
    uint64_t *bars = reinterpret_cast<uint64_t *>(ring + (size_t)nwc * depth * kStageBytes);      // [nwc][depth], xfull[], xempty[]
    uint64_t *xfull = bars + nwc * depth, *xempty = xfull + kKpMaxRows;
    unsigned *tap_mask = reinterpret_cast<unsigned *>(xempty + kKpMaxRows);             // [P]: tap mask | patch row << 16
    const size_t NP = (size_t)N * P;
    const size_t head_stride = NP * 9 * (size_t)C * GE;   // bytes between the heads of G
    const int nxb = ROWP ? ph : NBUF;                     // frame-tile barrier pairs in use

    if (tid == 0) {
        if (MMA && (tc::smem_u32(xs) & 1023u)) __trap();   // the swizzle pattern is a function of the absolute address
        for (int b = 0; b < nxb; ++b) { tc::mbar_init(xfull + b, 1); tc::mbar_init(xempty + b, nwc); }
        for (int s = 0; s < nwc * depth; ++s) tc::mbar_init(bars + s, 1);
        tc::fence_barrier_init();
    }
    for (int q = tid; q < P; q += blockDim.x) tap_mask[q] = ((dbg & 1) ? 0u : kp_tap_mask(q, ph, pw)) | ((unsigned)(q / pw) << 16);
    __syncthreads();
    if (my_tiles == 0) return;

    if (warp == nwc) {
        // ---- producer (frames past T1 arrive as zeros)
        if (lane == 0) {
            for (int j = 0; j < my_tiles; ++j) {
                const unsigned tile = tile0 + j, nc = tile / ntb;
                const int c0 = (int)(nc % nch) * CC, r0 = (int)(nc / nch) * P, f0 = (int)(tile % ntb) * TB;
                if (ROWP) {
                    for (int r = 0; r < ph; ++r) {
                        if (j > 0) tc::mbar_wait(xempty + r, (uint32_t)((j - 1) & 1));   // every warp is past row r of tile j-1
                        if (!(dbg & 4)) {
                            tc::mbar_arrive_expect_tx(xfull + r, (uint32_t)(row_floats * sizeof(float)));
                            tc::tma_load_3d(xs + (size_t)r * row_floats, &tm_x, xfull + r, c0, r0 + r * pw, f0);
                        } else {
                            tc::mbar_arrive(xfull + r);
                        }
                    }
                    continue;
                }
                const int b = j % NBUF;
                if (j >= NBUF) tc::mbar_wait(xempty + b, (uint32_t)((j / NBUF - 1) & 1));   // consumers are done with tile j-NBUF
                if (!(dbg & 4)) {
                    tc::mbar_arrive_expect_tx(xfull + b, (uint32_t)(tile_floats * sizeof(float)));
                    tc::tma_load_3d(xs + (size_t)b * tile_floats, &tm_x, xfull + b, c0, r0, f0);
                    // one buffer: the consumers cannot overlap the NEXT tile's load with their FMAs, so at least its HBM leg is
                    // started now (L2 prefetch); the load issued when they leave this tile then comes from L2
                    if (NBUF == 1 && j + 1 < my_tiles && !(dbg & 8)) {
                        const unsigned t2 = tile + 1, nc2 = t2 / ntb;
                        tc::tma_prefetch_3d(&tm_x, (int)(nc2 % nch) * CC, (int)(nc2 / nch) * P, (int)(t2 % ntb) * TB);
                    }
                } else {
                    tc::mbar_arrive(xfull + b);
                }
            }
        }
        return;
    }

    // ---- consumers: warp w owns positions w, w + nwc, ... of the CTA's concatenated (tile, position) stream (nwc <= P)
    unsigned char *my_ring = ring + (size_t)warp * depth * kStageBytes;
    uint64_t *my_bars = bars + warp * depth;
    // lane 0 feeds the warp's own ring, `depth` positions ahead of the consumer: (ji, pi) = tile and position of the next
    // stage to request, si = its ring stage; gi_row = G row of (tile ji, position 0) for this chunk
    int ji = 0, pi = warp, si = 0, ji_cur = -1;
    const unsigned char *gi_row = nullptr;
    auto issue = [&]() {
        if (ji >= my_tiles) return;
        if (ji != ji_cur) {
            ji_cur = ji;
            const unsigned nc = (unsigned)(tile0 + ji) / ntb;            // n*nch + chunk
            gi_row = G + ((size_t)(nc / nch) * P * nch + (size_t)(nc % nch)) * kRowBytes;
        }
        tc::mbar_arrive_expect_tx(my_bars + si, kStageTx);
        const unsigned char *row = gi_row + (size_t)pi * nch * kRowBytes;
#pragma unroll
        for (int h = 0; h < H; ++h)
            bulk_g2s(my_ring + (size_t)si * kStageBytes + h * kHeadPitch, row + (size_t)h * head_stride, kRowBytes,
                     my_bars + si);
        if (++si == depth) si = 0;
        pi += nwc;
        if (pi >= P) { pi -= P; ++ji; }
    };
    if (lane == 0)
        for (int i = 0; i < depth; ++i) issue();

    int j = 0, p = warp, s = 0;                           // current tile (local index), position, ring stage
    uint32_t phase = 0;
    const uint32_t xs_addr = tc::smem_u32(xs), ring_addr = tc::smem_u32(my_ring);
    int cur = -1, t0 = 0, nt = 0;
    int rows_seen = 0, rows_freed = 0;                    // ROWP: patch rows of tile j waited for / handed back by this warp
    int refills = (dbg & 2) ? 0 : 0x7fffffff;
    float *out_tile = nullptr;
    const float *xlane = nullptr;
    while (j < my_tiles) {
        if (j != cur) {
            cur = j;
            const unsigned tile = tile0 + j, nc = tile / ntb;
            t0 = (int)(tile % ntb) * TB; nt = min(TB, T1 - t0);
            out_tile = parts + ((size_t)(nc % nch) * N + nc / nch) * P * H * T1 + t0;    // [chunk][n] + frame block
            if (ROWP) {
                rows_seen = rows_freed = 0;
                xlane = xs + (size_t)tg * FR * frame_floats + cg * 4;
            } else {
                tc::mbar_wait(xfull + j % NBUF, (uint32_t)((j / NBUF) & 1));
                xlane = xs + (size_t)(j % NBUF) * tile_floats + (size_t)tg * FR * frame_floats + cg * 4;
            }
        }
        const unsigned tm = tap_mask[p];
        const int py = (int)(tm >> 16);
        if (ROWP)                                          // rows py-1 .. py+1 of this tile have landed
            for (const int need = min(py + 2, ph); rows_seen < need; ++rows_seen) tc::mbar_wait(xfull + rows_seen, (uint32_t)(j & 1));
        tc::mbar_wait(my_bars + s, phase);
        if (MMA) {
            float c[4];
            kp_position_mma<MMA ? MPW : 1>(xs_addr + (uint32_t)(py * row_floats) * 4u, ring_addr + (uint32_t)(s * kStageBytes), p - py * pw,
                                           tm & 0x1ffu, lane, c);
            // the last (warp-synchronous) MMA has consumed every lane's loads of stage s and of the frame rows
            if (lane == 0) {
                if (refills > 0) issue();
                else tc::mbar_arrive(my_bars + s);
            }
            const int g = lane >> 2, t2 = lane & 3;
            if (g < 4) {                                   // head g: frames t2, t2 + 4, t2 + 8, t2 + 12
                float *o = out_tile + ((size_t)p * H + g) * T1 + t2;
#pragma unroll
                for (int i = 0; i < 4; ++i)
                    if (t2 + 4 * i < nt) o[4 * i] = c[i];
            }
        } else {
            float vc[NV / 8];
            kp_position<FR, GE>(ROWP ? xlane + py * row_floats + (p - py * pw) * CC : xlane + p * CC, my_ring + (size_t)s * kStageBytes,
                                row_floats, frame_floats, tm & 0x1ffu, cg, vc);
            // every lane's sums (hence its reads of stage s and of the frame tile) are complete once it has taken part in the
            // butterfly shuffles: lane 0 may hand the stage back to the copy engine
            if (lane == 0) {
                if (refills > 0) issue();
                else tc::mbar_arrive(my_bars + s);             // probe mode: keep the phases moving without a copy
            }
#pragma unroll
            for (int k = 0; k < NV / 8; ++k) {
                const int v = cg * (NV / 8) + k;
                const int t = tg * FR + v / H;
                if (t < nt) out_tile[((size_t)p * H + v % H) * T1 + t] = vc[k];
            }
        }
        if (++s == depth) { s = 0; phase ^= 1; }
        p += nwc;
        const bool last = p >= P;                          // this warp's last position in tile j
        if (ROWP) {                                        // rows no later position of this warp reads go back to the producer
            const int keep_from = last ? ph : max((int)(tap_mask[p] >> 16) - 1, 0);
            if (lane == 0)
                for (int r = rows_freed; r < keep_from; ++r) tc::mbar_arrive(xempty + r);
            rows_freed = max(rows_freed, keep_from);
        }
        if (last) {
            p -= P;
            if (!ROWP && lane == 0) tc::mbar_arrive(xempty + j % NBUF);
            ++j;
        }
    }
}

// ------------------------------------------------------------------------------------------------ one CTA per tile
// TB = frames per CTA: 8 -> [8][P][32] tile + 7 x 2 G stages = 112 KB, two CTAs per SM; 16 -> one CTA per SM, G read once.
template <int TB, int MAXW, int MINB, int GE>
__global__ void __launch_bounds__(MAXW * 32, MINB)
tafa_keyproj_logits_kernel(const __grid_constant__ CUtensorMap tm_x, const unsigned char *__restrict__ G, float *__restrict__ parts,
                           int T1, int N, int ph, int pw, int C, int depth, int dbg) {
    constexpr int H = kKpHeads, CC = kKpCC, FR = TB / 4, NV = FR * H;   // FR frames per lane group, NV sums per lane
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int tb = blockIdx.x, chunk = blockIdx.y, n = blockIdx.z, nch = C / CC;
    const int P = ph * pw;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = blockDim.x >> 5;
    const int tg = lane >> 3, cg = lane & 7;
    const int t0 = tb * TB, nt = min(TB, T1 - t0);
    float *xs = reinterpret_cast<float *>(smem_raw);                                   // [TB][P][CC]
    constexpr int kStageBytes = kKpStageFloats * GE, kRowBytes = kKpRowFloats * GE;
    unsigned char *ring = reinterpret_cast<unsigned char *>(xs + (size_t)TB * P * CC);  // [nwarps][depth][H][9][CC] of GE bytes
    uint64_t *bars = reinterpret_cast<uint64_t *>(ring + (size_t)nwarps * depth * kStageBytes);   // [nwarps][depth] + 1
    uint64_t *xbar = bars + nwarps * depth;
    const size_t NP = (size_t)N * P;
    const size_t head_stride = NP * 9 * (size_t)C * GE;   // bytes between the heads of G
    const unsigned char *g_roi = G + ((size_t)n * P * nch + chunk) * kRowBytes;   // + p * nch*9*CC*GE + h * head_stride

    unsigned char *my_ring = ring + (size_t)warp * depth * kStageBytes;
    uint64_t *my_bars = bars + warp * depth;
    auto issue = [&](int i) {
        const int p = warp + i * nwarps;
        if (p >= P) return;
        if ((dbg & 2) && i >= depth) return;
        const int s = i % depth;
        tc::mbar_arrive_expect_tx(my_bars + s, kStageBytes);
        const unsigned char *row = g_roi + (size_t)p * nch * kRowBytes;
#pragma unroll
        for (int h = 0; h < H; ++h)
            bulk_g2s(my_ring + (size_t)s * kStageBytes + h * kRowBytes, row + (size_t)h * head_stride, kRowBytes,
                     my_bars + s);
    };
    if (lane == 0) {
        if (warp == 0) {
            tc::mbar_init(xbar, 1);
            tc::fence_barrier_init();
            if (!(dbg & 4)) {
                tc::mbar_arrive_expect_tx(xbar, (uint32_t)(TB * P * CC * sizeof(float)));
                tc::tma_load_3d(xs, &tm_x, xbar, chunk * CC, n * P, t0);
            }
        }
        for (int s = 0; s < depth; ++s) tc::mbar_init(my_bars + s, 1);
        tc::fence_barrier_init();
        for (int i = 0; i < depth; ++i) issue(i);
    }
    __syncthreads();                 // barrier initialisations visible to every waiter
    if (!(dbg & 4)) tc::mbar_wait(xbar, 0);

    const float *xlane = xs + (size_t)tg * FR * P * CC + cg * 4;   // + u * P*CC + q * CC
    int i = 0;
    for (int p = warp; p < P; p += nwarps, ++i) {
        const int s = i % depth;
        if (!((dbg & 2) && i >= depth)) tc::mbar_wait(my_bars + s, (uint32_t)((i / depth) & 1));
        float vc[NV / 8];
        kp_position<FR, GE>(xlane + p * CC, my_ring + (size_t)s * kStageBytes, pw * CC, P * CC, (dbg & 1) ? 0u : kp_tap_mask(p, ph, pw), cg, vc);
        if (lane == 0) issue(i + depth);
#pragma unroll
        for (int k = 0; k < NV / 8; ++k) {
            const int v = cg * (NV / 8) + k;
            const int t = tg * FR + v / H;
            if (t < nt) parts[((((size_t)chunk * N + n) * P + p) * H + v % H) * T1 + t0 + t] = vc[k];
        }
    }
}

}  // namespace vod

using namespace vod;

static size_t kp_smem_bytes(int tile_frames, int P, int warps, int depth, int ge = 4, bool mma = false) {
    const size_t stage = mma ? (size_t)kKpHeads * kKpMmaHeadPitch : (size_t)kKpStageFloats * ge;
    return (size_t)tile_frames * P * kKpCC * sizeof(float) + (size_t)warps * depth * (stage + 8) +
           2 * kKpMaxRows * 8 + 32 + (size_t)P * 4;
}
constexpr size_t kKpSmemLimit = 227 * 1024;

// Channel-chunk width the logits kernel uses for these dims (the caller lays G out with it); 0 = shape not supported
// (the caller then takes the full-embedding path, vod_tafa_weighted_sum).
extern "C" int vod_tafa_keyproj_chunk(int T1, int P, int C, int heads) {
    if (heads != kKpHeads || T1 <= 0 || P <= 0 || C <= 0) return 0;
    if (C % kKpCC != 0) return 0;
    if (P > 256 || kp_smem_bytes(8, P, 4, 2) > kKpSmemLimit) return 0;   // TMA box dims <= 256
    return kKpCC;
}

extern "C" int vod_tafa_keyproj_logits(const float *x_all, const void *G_, int g_dtype, float *parts, int T1, int N, int ph, int pw,
                                       int C, int heads, int cc, vod_stream_t stream) {
    const unsigned char *G = reinterpret_cast<const unsigned char *>(G_);
    VOD_REQUIRE(g_dtype == VOD_DTYPE_F32 || g_dtype == VOD_DTYPE_BF16, "vod_tafa_keyproj_logits: g_dtype");
    const int ge = g_dtype == VOD_DTYPE_BF16 ? 2 : 4;
    if (N == 0) return VOD_OK;
    VOD_REQUIRE(x_all && G && parts, "vod_tafa_keyproj_logits: null pointer");
    VOD_REQUIRE(T1 > 0 && N > 0 && ph > 0 && pw > 0 && C > 0, "vod_tafa_keyproj_logits: bad dims");
    VOD_REQUIRE(cc != 0 && cc == vod_tafa_keyproj_chunk(T1, ph * pw, C, heads),
                "vod_tafa_keyproj_logits: unsupported shape (T1=%d P=%d C=%d heads=%d cc=%d)", T1, ph * pw, C, heads, cc);
    VOD_REQUIRE(((reinterpret_cast<uintptr_t>(x_all) | reinterpret_cast<uintptr_t>(G)) & 15) == 0,
                "vod_tafa_keyproj_logits: x_all and G must be 16-byte aligned");
    VOD_REQUIRE(N <= 65535 && C / cc <= 65535 && (long)N * (C / cc) * ceil_div(T1, 8) < (1L << 30),
                "vod_tafa_keyproj_logits: grid too large");
    const int P = ph * pw;
    // Measured at N=300, C=512 (B200).  16 stacked frames: persistent 16-frame tiles 381 us, one CTA per 16-frame tile 391 us,
    // persistent double-buffered 8-frame tiles 432 us, one CTA per 8-frame tile (two per SM) 458 us.  32 frames: persistent
    // 16-frame tiles 723 us, one CTA per tile 893 us.  8 frames: one CTA per 8-frame tile 241 us (5.6 TB/s).
    // (bf16 G halves the rings: room for a 15th consumer warp next to the 98 KB frame tile -- 361 vs 370 us)
    // Round 2, 16 frames: + L2 prefetch of the next tile 352 (bf16 G) / 384; tile as one box per patch row (rolling buffer)
    // 325 / 342; bf16 G with the tf32 mma.sync form 239-249 us (memory traffic alone, FMAs off: 193 us).
    int persist = T1 > 8, tb = T1 > 8 ? 16 : 8, warps = min(T1 > 8 ? (ge == 2 ? 15 : 14) : 7, P), depth = 2, dbg = 0;
    // Tuning / probe hooks (dbg: 1 = no FMAs, 2 = no G refills after the first ring fill, 4 = no frame tile loads) exist only
    // in a -DVOD_PROBES build: the production library ignores the environment, so a stray variable can never switch
    // parts of the computation off.
#ifdef VOD_PROBES
    if (const char *e = getenv("VOD_KP_PERSIST")) persist = atoi(e) != 0;
    if (const char *e = getenv("VOD_KP_TB")) tb = atoi(e) == 16 ? 16 : 8;
    if (const char *e = getenv("VOD_KP_WARPS")) warps = max(1, min(atoi(e), min(kKpMaxWarpsMma, P)));
    if (const char *e = getenv("VOD_KP_DEPTH")) depth = max(1, min(atoi(e), 8));
    if (const char *e = getenv("VOD_KP_DBG")) dbg = atoi(e);
#endif
    // persistent 16-frame tiles: one TMA box per patch row (rolling single buffer); with bf16 G and 7-wide patches the
    // tensor-core form (tf32 mma.sync; bf16 G already means the caller allowed reduced-precision math)
    int rowp = persist && tb == 16 && ph <= kKpMaxRows;
    int mma = rowp && ge == 2 && pw == 7;
#ifdef VOD_PROBES
    if (const char *e = getenv("VOD_KP_ROWP")) rowp = rowp && atoi(e) != 0;
    if (const char *e = getenv("VOD_KP_MMA")) mma = mma && atoi(e) != 0;
    mma = mma && rowp;
#endif
#ifdef VOD_PROBES
    if (!getenv("VOD_KP_WARPS"))
#endif
    if (mma) warps = min(kKpMaxWarpsMma - 1, P);           // few registers per thread: 20 consumer warps (measured 239-249 us vs 257-273 with 15)
    warps = min(warps, (mma ? kKpMaxWarpsMma : kKpMaxWarps) - (persist ? 1 : 0));      // persistent: + the producer warp
    const int tile_frames = persist ? 16 : tb;              // persistent: 2 x 8 frames (double buffer) or 1 x 16
    while (warps > 4 && kp_smem_bytes(tile_frames, P, warps, depth, ge, mma) > kKpSmemLimit) --warps;
    while (depth > 2 && kp_smem_bytes(tile_frames, P, warps, depth, ge, mma) > kKpSmemLimit) --depth;
    if (persist && kp_smem_bytes(tile_frames, P, warps, depth, ge, mma) > kKpSmemLimit) {   // large patches: one tile per CTA
        persist = rowp = mma = 0;
        tb = 8;
    }
    const size_t smem = kp_smem_bytes(persist ? 16 : tb, P, warps, depth, ge, mma);
    VOD_REQUIRE(smem <= kKpSmemLimit, "vod_tafa_keyproj_logits: tile does not fit shared memory");
    // x_all [T1][N*P][C] as a 3-D tensor, box = (32 channels, P positions or one patch row, tb frames)
    CUtensorMap tm_x;
    if (int rc = make_tmap_f32_3d(&tm_x, x_all, (uint64_t)C, (uint64_t)N * P, (uint64_t)T1, (uint64_t)C * 4,
                                  (uint64_t)N * P * C * 4, kKpCC, (uint32_t)(rowp ? pw : P), (uint32_t)tb, mma != 0))
        return rc;
    cudaStream_t st = as_stream(stream);
    // the opt-in to > 48 KB of dynamic shared memory is a per-device function attribute: set it on every call (cheap,
    // idempotent) rather than once per process, so a second device in the same process works too
    auto allow = [&](auto kern) { cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); return kern; };
    if (persist) {
        const long total_tiles = (long)N * (C / cc) * ceil_div(T1, tb);
        int sms = num_sms();
#ifdef VOD_PROBES
        if (const char *e = getenv("VOD_KP_CTAS")) sms = max(1, atoi(e));
#endif
        const int tiles_per_cta = (int)ceil_div(total_tiles, (long)sms);
        const int grid = (int)ceil_div(total_tiles, (long)tiles_per_cta);
#define VOD_KP_PERSIST(TB_, NB_, RP_)                                                                                              \
    do {                                                                                                                           \
        if (ge == 2) allow(tafa_keyproj_persist_kernel<TB_, NB_, 2, RP_, 0>)<<<grid, (warps + 1) * 32, smem, st>>>(tm_x, G, parts, T1, N, ph, pw, C, depth, tiles_per_cta, dbg); \
        else allow(tafa_keyproj_persist_kernel<TB_, NB_, 4, RP_, 0>)<<<grid, (warps + 1) * 32, smem, st>>>(tm_x, G, parts, T1, N, ph, pw, C, depth, tiles_per_cta, dbg);         \
    } while (0)
        if (mma) allow(tafa_keyproj_persist_kernel<16, 1, 2, true, 7>)<<<grid, (warps + 1) * 32, smem, st>>>(tm_x, G, parts, T1, N, ph, pw, C, depth, tiles_per_cta, dbg);
        else if (rowp) VOD_KP_PERSIST(16, 1, true);
        else if (tb == 16) VOD_KP_PERSIST(16, 1, false);
        else VOD_KP_PERSIST(8, 2, false);
#undef VOD_KP_PERSIST
    } else {
        dim3 grid(ceil_div(T1, tb), C / cc, N);
#define VOD_KP_TILE(TB_, MW_, MB_)                                                                                                 \
    do {                                                                                                                           \
        if (ge == 2) allow(tafa_keyproj_logits_kernel<TB_, MW_, MB_, 2>)<<<grid, warps * 32, smem, st>>>(tm_x, G, parts, T1, N, ph, pw, C, depth, dbg); \
        else allow(tafa_keyproj_logits_kernel<TB_, MW_, MB_, 4>)<<<grid, warps * 32, smem, st>>>(tm_x, G, parts, T1, N, ph, pw, C, depth, dbg);         \
    } while (0)
        if (tb == 16) VOD_KP_TILE(16, kKpMaxWarps, 1);
        else if (warps <= 8 && kp_smem_bytes(tb, P, warps, depth, ge) <= 113 * 1024) VOD_KP_TILE(8, 8, 2);
        else VOD_KP_TILE(8, kKpMaxWarps, 1);
#undef VOD_KP_TILE
    }
    note_launch();
    return check_launch("vod_tafa_keyproj_logits");
}
