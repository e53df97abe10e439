// fused_kernels.cuh -- the post-attention half of a transformer Block in ONE kernel per 128-token tile
// (model.py:71,85-87,102-103 + the next block's ln_1, model.py:102):
//
//     x1 = x + att @ Wproj^T                      (attn.c_proj + residual)
//     h  = gelu(LN2(x1) @ Wfc^T)                  (ln_2, mlp.c_fc, GELU) -- never leaves the SM
//     x' = x1 + h @ Wproj2^T                      (mlp.c_proj + residual)
//     xn = LN1_next(x')                           (A operand of the next block's QKV GEMM)
//
// HBM traffic per token: read att (2C) + x (4C), write x' (4C) + xn (2C) = 12C bytes (1920 B at C=160),
// against 5 separate kernels moving 38C bytes.  The 4C-wide hidden activation lives only in TMEM/SMEM.
//
// TMEM: [0,C) = running fp32 residual accumulator (proj result -> x1 -> x1 + MLP), [C, C+C/2) = FC chunk.
// The residual add is free: x1 is written back to TMEM and the mlp.c_proj UMMAs accumulate onto it.
// SMEM: A tile (att, then LN2(x1)) | 2 hidden-chunk buffers | weight-stage ring | reduction scratch.
// Weights arrive as a pre-packed stream of equal-size stage images (host: pack_post_attn_stream) in the
// exact order the UMMA issuer consumes them: proj k-steps, then FC(0), FC(1), P2(0), FC(2), P2(1), ...
//
// warps 0-7: workers (thread pair per row: TMEM lane quadrant = warp&3, column half = warp>>2)
// warp 8: weight-stage producer (bulk copies)      warp 9: UMMA issuer + TMEM owner
#pragma once
#include "gpt_kernels.cuh"

namespace mg {

struct PostAttnArgs {
    const __nv_bfloat16 *att;      // A_ti [MT][C/8][128][8]
    float *x;                      // X_ti [MT][C/4][128][4], updated in place
    const __nv_bfloat16 *wstream;  // stage images for this layer
    const __nv_bfloat16 *wstream_pair;   // host side only: the CTA-pair format of the same stream (the launcher swaps it in)
    const float *ln2_gain;         // [C] documentation only: ln_2's gain is folded into the c_fc rows of the stream
    const float *next_gain;        // [C] ln_1 gain of the next block (xn_out path; folded into c_attn when qkv_out is set)
    __nv_bfloat16 *xn_out;         // A_ti for the next block's QKV GEMM, or nullptr
    long long *timeline;           // test hook: clock64() stamps of CTA 0..3 ([cta][128]); nullptr in production
    // fused QKV projection of the NEXT block (c_attn, model.py:50): xn never goes to HBM.  The 3 x 2 extra weight stages
    // follow the 18 regular ones in `wstream`.  nullptr = write xn_out instead.
    __nv_bfloat16 *qkv_out;        // [seq][3][head][hs/8][256][8]
    int n_head, hs;
    // block 0 only: the residual tile comes straight from the (token, position) table (block0_lookup_kernel then writes q/k/v
    // only: x never makes the HBM round trip).  nullptr = read a.x.
    const uint8_t *tokens0;        // [MT * 128] token ids of this chunk
    const uint4 *tab0;             // [67][256][tab_nrec] records, x first
    int tab_nrec;
    // tile groups (of NT tiles) in this launch.  A CTA processes groups blockIdx.x, blockIdx.x + gridDim.x, ...: with
    // gridDim.x == n_groups every CTA owns one group; the persistent launch (fused-QKV layers) sizes the grid to the resident
    // CTAs and keeps barriers, TMEM and the weight ring alive across tiles.
    int n_groups;
    int stagger_ns;                // persistent launch: start offsets of the clusters are spread over this many ns (0 = none)
    // one-group-per-CTA launch: at its start a CTA asks L2 for the att / residual tiles of the CTA that will take over its slot
    // (group blockIdx.x + pf_dist, pf_dist = resident CTAs), so that CTA's first loads hit L2 instead of queueing behind the
    // stores of 295 other CTAs in the DRAM controllers.  0 = off.
    int pf_dist;
    // the block after this one is the pruned last block (only token 255 of each sequence goes on): of the updated residual and of
    // the next block's q rows only those of token 255 are stored (k and v are needed for every token) -- 6C of 16C bytes per token
    int tail_rows_only;
    int x_in_24, x_out_24;         // a.x is read / written in the 24-bit tile layout (pack24x16) instead of fp32
};
// steady-state tile groups (not the first wave, whose loads all hit DRAM at once), keyed by the group index `mg_group` in scope;
// single-tile launches have fewer groups and stamp nothing
#define MG_STAMP_CTA0 2048u
#define MG_STAMP(id)                                                                     \
    do {                                                                                 \
        if (a.timeline != nullptr && (mg_group - MG_STAMP_CTA0) < 4u) a.timeline[(mg_group - MG_STAMP_CTA0) * 128 + (id)] = clock64(); \
    } while (0)

// Phase profile (builds with -DMG_PHASE_PROF only, tools/phase_profile.py): cycles per phase summed over ALL CTAs of every
// launch into timeline[2048 + id] -- worker thread 0 (ids 0..15) and the UMMA issuer's lane 0 (ids 20..25).
#ifdef MG_PHASE_PROF
#define MG_LAP(id)                                                                                     \
    do {                                                                                               \
        if (a.timeline != nullptr) {                                                                   \
            const long long now_ = clock64();                                                          \
            atomicAdd(reinterpret_cast<unsigned long long *>(a.timeline) + 2048 + (id), (unsigned long long)(now_ - lap_t)); \
            lap_t = now_;                                                                              \
        }                                                                                              \
    } while (0)
#define MG_LAP_INIT long long lap_t = clock64()
#else
#define MG_LAP(id) do { } while (0)
#define MG_LAP_INIT do { } while (0)
#endif
// (v - mean) * rstd * gain for 8 consecutive columns -> 8 bf16 (one 16-byte store); a = rstd, b = -mean * rstd
__device__ __forceinline__ uint4 ln_pack8(const uint32_t *v, f32x2 a, f32x2 b, const float4 g0, const float4 g1)
{
    uint4 o;
    o.x = pack_bf16x2_p(mul2(fma2(pk2u(v[0], v[1]), a, b), pk2(g0.x, g0.y)));
    o.y = pack_bf16x2_p(mul2(fma2(pk2u(v[2], v[3]), a, b), pk2(g0.z, g0.w)));
    o.z = pack_bf16x2_p(mul2(fma2(pk2u(v[4], v[5]), a, b), pk2(g1.x, g1.y)));
    o.w = pack_bf16x2_p(mul2(fma2(pk2u(v[6], v[7]), a, b), pk2(g1.z, g1.w)));
    return o;
}
// (v - mean) * rstd for 8 consecutive columns -> 8 bf16; the LayerNorm gain is folded into the weights that consume it
__device__ __forceinline__ uint4 ln_pack8_ng(const uint32_t *v, f32x2 a, f32x2 b)
{
    uint4 o;
    o.x = pack_bf16x2_p(fma2(pk2u(v[0], v[1]), a, b));
    o.y = pack_bf16x2_p(fma2(pk2u(v[2], v[3]), a, b));
    o.z = pack_bf16x2_p(fma2(pk2u(v[4], v[5]), a, b));
    o.w = pack_bf16x2_p(fma2(pk2u(v[6], v[7]), a, b));
    return o;
}
// N (multiple of 16) consecutive fp32 columns of this thread's TMEM lane; the caller waits once for the whole batch
template <int N>
__device__ __forceinline__ void tmem_ld_n(uint32_t taddr, uint32_t (&v)[N])
{
    static_assert(N % 16 == 0, "tmem_ld_n: whole 16-column loads");
#pragma unroll
    for (int i = 0; i < N / 16; i++) tmem_ld16(taddr + 16 * i, *reinterpret_cast<uint32_t(*)[16]>(&v[16 * i]));
}
template <int V> struct IntC { static constexpr int value = V; };
// the thread's HALF columns in batches of <= 48: one TMEM round trip per batch instead of one per 16 columns
#define MG_COL_BATCHES(HALF_, F)                                                 \
    do {                                                                         \
        static_assert((HALF_) == 64 || (HALF_) == 80 || (HALF_) == 128, "column batches"); \
        if constexpr ((HALF_) == 64) {                                           \
            F(IntC<0>{}, IntC<32>{});                                            \
            F(IntC<32>{}, IntC<32>{});                                           \
        } else if constexpr ((HALF_) == 80) {                                    \
            F(IntC<0>{}, IntC<48>{});                                            \
            F(IntC<48>{}, IntC<32>{});                                           \
        } else {                                                                 \
            F(IntC<0>{}, IntC<48>{});                                            \
            F(IntC<48>{}, IntC<48>{});                                           \
            F(IntC<96>{}, IntC<32>{});                                           \
        }                                                                        \
    } while (0)
// sum of squared deviations of 16 values, packed
__device__ __forceinline__ f32x2 sqdev16(const uint32_t (&v)[16], f32x2 negmean, f32x2 acc)
{
#pragma unroll
    for (int j = 0; j < 16; j += 2) {
        const f32x2 d = add2(pk2u(v[j], v[j + 1]), negmean);
        acc = fma2(d, d, acc);
    }
    return acc;
}

// NT = 128-token tiles per CTA.  NT = 2 (one whole 256-token sequence per CTA, one CTA per SM) lets both tiles share
// every weight stage: half the L2->SMEM weight traffic per token and more bytes in flight per SM.
// Timeline measurements (tools/timeline.py, clock64 stamps) drove three choices:
//   * a weight stage is U k-step units (25.6 KB at C=160) -- every full-barrier wait + fence + commit costs the single
//     issuing thread ~300 cycles, so 5 KB stages made the issuer, not the tensor pipe, the bottleneck;
//   * x is pre-loaded into the TMEM accumulator by the workers while the att tile is still in flight, so the c_proj
//     UMMAs produce x1 = x + proj directly and the first epilogue has no global loads;
//   * LayerNorm statistics are single-pass (sum, sum of squares) in packed fp32x2.
template <int C, int NT, int UU = 0, int CL = 1>
struct PostAttnCfg {
    // WIDE (C = 256 only, -DMG_POST_WIDE256=1): the MLP in 4 chunks of C hidden columns instead of 8 of C/2 -- half the phase
    // hand-offs per tile (TMEM 256 + 256 = 512 columns, hidden buffer 64 KB); the c_attn tail keeps half n-tiles of HC columns.
    // Measured SLOWER (3.62 -> 3.70 ms per launch: coarser chunks overlap the tensor pipe and the workers less): off by default.
#ifndef MG_POST_WIDE256
#define MG_POST_WIDE256 0
#endif
    static constexpr bool WIDE = C == 256 && NT == 1 && MG_POST_WIDE256;
    static constexpr int HC = C / 2;                 // half n-tile of the fused c_attn; MLP chunk width unless WIDE
    static constexpr int HM = WIDE ? C : HC;         // hidden chunk (FC N, proj2 K per chunk)
    static constexpr int NCH = 4 * C / HM;           // 8 chunks (4 when WIDE)
    static constexpr int UNIT_BYTES = 32 * C;        // [2 kc][C][16B] == [4 kc][HC][16B]
    static constexpr int NPROJ = C / 16;             // units of the proj GEMM (one k-step each)
    static constexpr int NFCQ = C / 32;              // units per half n-tile of the fused c_attn (two k-steps each)
    static constexpr int NFC = WIDE ? C / 16 : C / 32;   // units per FC chunk (two k-steps of HC rows each; WIDE: one k-step of C rows)
    static constexpr int NP2 = HM / 16;              // units per proj2 chunk (one k-step each)
    static constexpr int U = UU ? UU : (C == 160 ? 5 : 4);   // units per stage
    static constexpr int STAGE_BYTES = U * UNIT_BYTES;
    static constexpr int TOTAL_STAGES = (NPROJ + NCH * (NFC + NP2)) / U;
    static constexpr int A_BYTES = C * 256;          // [C/8][128][16B] per tile
    static constexpr int H_BYTES = HM * 256;         // [HM/8][128][16B] per tile
    static constexpr int SLOT_BYTES = STAGE_BYTES / CL;   // CTA pair: each CTA holds its half (N/2 weight rows) of every stage
    static constexpr int STAGES = CL * (UU ? (NT == 2 ? 4 : (C <= 160 ? 2 : 3)) * (C == 160 ? 5 : 4) / UU : (NT == 2 ? 4 : (C <= 160 || WIDE ? 2 : 3)));
    static constexpr int CTAS_PER_SM = (NT == 1 && C <= 160) ? 2 : 1;
    static constexpr int TILE_COLS = (C + HM) <= 256 ? 256 : 512;          // TMEM columns per tile
    static constexpr uint32_t TMEM_COLS = TILE_COLS * NT;
    // NH column groups per tile: a tile's 128 rows are covered by 4 warps (TMEM lane quadrants) x NH groups of C / NH columns.
    // C = 256 runs one CTA per SM (TMEM / shared memory) at 16 % warps-active with its 8 worker warps; 16 worker warps (NH = 4,
    // -DMG_POST_NH256=4) were measured SLOWER, 3.69 -> 4.06 ms per launch: twice the barrier arrivals and TMEM round trips per
    // phase for half the work per thread (profiles/r02_experiments.md).  Default: 2 column groups everywhere.
#ifndef MG_POST_NH256
#define MG_POST_NH256 2
#endif
    static constexpr int NH = (C == 256 && NT == 1) ? MG_POST_NH256 : 2;
    static constexpr int NW = 4 * NH;                // worker warps per tile
    static constexpr int THREADS = 64 + 32 * NW * NT;
    static constexpr int QKV_STAGES = 6 * NFCQ / U;    // next block's c_attn: 6 half n-tiles of HC columns (narrow FC-chunk stage format)
    static constexpr int NBAR = 3 * STAGES + 2 + NT * 14;
    static constexpr int SMEM_BYTES = NT * (A_BYTES + H_BYTES) + STAGES * SLOT_BYTES + NT * 2 * NH * 128 * 4 + NBAR * 8 + 16 + (3 * C / 8) * 4;
    static_assert(C % 32 == 0 && C <= 256, "post_attn_kernel: C must be a multiple of 32, <= 256");
    static_assert(NPROJ % U == 0 && NFC % U == 0 && NP2 % U == 0 && NFCQ % U == 0, "stage size must divide every GEMM phase");
    static_assert(TMEM_COLS <= 512, "post_attn_kernel: TMEM budget");
    static_assert(CL == 1 || (CL == 2 && NT == 1 && C % 32 == 0), "CTA pairs: one tile per CTA, N/2 a multiple of 16");
};

// CL = 2: CTA pair (cta_group::2).  The two CTAs of a cluster own adjacent 128-token tiles and run every GEMM as ONE M = 256
// UMMA issued by the leader (rank 0): A = each CTA's own tile, B = N/2 weight rows from each CTA's ring, D = 128 rows in each
// CTA's TMEM.  Each SM therefore pulls only HALF of the weight stream through L2 -> SM (the kernel was bound there:
// 36 B/clk/SM of the 42 B/clk/SM the L2 slices deliver, profiles/r01c_launch_list_summary.md; multicasting full stages to
// both CTAs was measured and does not help -- the SM-side ingest is what counts).  Cross-CTA protocol:
//   * workers of both CTAs arrive (one elected lane per warp) on the LEADER's barriers; the leader's commits are multicast
//     to the barriers of both CTAs;
//   * the peer's otherwise idle UMMA warp relays "my half of stage i landed" (and "my att tile landed") to the leader.
// PERSIST: the CTA loops over tile groups blockIdx.x, blockIdx.x + gridDim.x, ... (barriers, TMEM, weight ring and the issuer's
// descriptors live across tiles; the next att tile is requested as soon as the A buffer is free, the next residual tile is
// prefetched into L2).  !PERSIST: one group per CTA, and the residual tile is requested BEFORE the rendezvous (inside a tile loop
// the compiler parks those 80 registers in local memory across the rendezvous -- a store that waits for the data).
// Measured (profiles/r02_post_attn_persistent.md): PERSIST wins where one CTA fits per SM (C = 256: 3.89 -> 3.65 ms per launch)
// and loses where two do (C = 160: 1.92 -> 1.98 ms; two co-resident CTAs started by the hardware at different times already
// hide each other's launch gaps, while resident CTAs run in lockstep), so the launcher uses it for C = 256 only.
#ifndef MG_GELU_MIX_BITS
#define MG_GELU_MIX_BITS 0
#endif
template <int C, int NT, int UU = 0, int CL = 1, bool PERSIST = false>
__global__ void __launch_bounds__(PostAttnCfg<C, NT, UU, CL>::THREADS, PostAttnCfg<C, NT, UU, CL>::CTAS_PER_SM)
post_attn_kernel(const PostAttnArgs a)
{
    using K = PostAttnCfg<C, NT, UU, CL>;
    constexpr int HC = K::HC, S = K::STAGES, U = K::U;
    constexpr bool PAIR = CL == 2;
    constexpr int NH = K::NH, NW = K::NW;     // column groups / worker warps per tile
    constexpr int WARR = NW * CL;             // arrivals on a worker -> issuer barrier: one per worker warp of the CTA group
    constexpr int CB = C / CL, HB = HC / CL;  // weight rows per CTA of a C-wide / HC-wide B operand
    constexpr int UNIT_B = K::UNIT_BYTES / CL;
    constexpr uint32_t UM = 128 * CL;         // UMMA M
    constexpr int PROD_WARP = NW * NT, MMA_WARP = NW * NT + 1;
    const uint32_t crank = PAIR ? cluster_ctarank() : 0u;
    const bool leader = crank == 0;
    extern __shared__ __align__(1024) uint8_t smem[];
    uint8_t *As = smem;                                   // [NT] A tiles
    uint8_t *Hs = As + NT * K::A_BYTES;                   // [NT] hidden-chunk buffers
    uint8_t *ring = Hs + NT * K::H_BYTES;
    float *red = reinterpret_cast<float *>(ring + S * K::SLOT_BYTES);    // [NT][2 kinds][NH column groups][128]
    uint64_t *full = reinterpret_cast<uint64_t *>(red + NT * 2 * NH * 128);
    uint64_t *empty = full + S;
    uint64_t *pfull = empty + S;         // leader only: the peer's half of the stage landed (relayed)
    uint64_t *bar_proj = pfull + S;      // proj UMMAs retired (all tiles)
    uint64_t *bar_done = bar_proj + 1;   // all UMMAs retired
    uint64_t *bar_att = bar_done + 1;    // [NT] A tile landed (tx)
    uint64_t *bar_x = bar_att + NT;      // [NT] x pre-loaded into the TMEM accumulator (256 arrivals)
    uint64_t *bar_ln2 = bar_x + NT;      // [NT] LN2(x1) in smem (256 arrivals)
    uint64_t *bar_a1f = bar_ln2 + NT;    // [NT] FC chunk accumulated
    uint64_t *bar_a1e = bar_a1f + NT;    // [NT] FC chunk drained to registers (256 arrivals)
    uint64_t *bar_hf = bar_a1e + NT;     // [NT] hidden chunk written to smem (256 arrivals)
    uint64_t *bar_he = bar_hf + NT;      // [NT] hidden chunk consumed by the proj2 UMMAs
    uint64_t *bar_qa = bar_he + NT;      // [NT] fused QKV: LN1_next(x') in smem, accumulator free (256 arrivals)
    uint64_t *bar_qf = bar_qa + NT;      // [NT][3] fused QKV: half n-tile accumulated in buffer b
    uint64_t *bar_qe = bar_qf + 3 * NT;  // [NT][3] fused QKV: buffer b drained to registers (256 arrivals)
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bar_qe + 3 * NT);
    uint32_t *qkv_off = tmem_slot + 2;   // uint4 offset of each 8-column group inside a sequence's q/k/v block
    const bool fuse_qkv = a.qkv_out != nullptr;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int mt0 = blockIdx.x * NT;              // first tile of the CTA's first group
    const int n_groups = a.n_groups, gstride = (int)gridDim.x;
    uint32_t mg_group = blockIdx.x;               // group being processed (timeline stamps)

    const int n_stages = K::TOTAL_STAGES + (fuse_qkv ? K::QKV_STAGES : 0);
    // clock probe (tools/clock_probe.py): SM cycles and wall nanoseconds over one mid-grid CTA -> the SM clock this kernel really
    // runs at inside a long, power-capped step
    const bool clk_probe = a.timeline != nullptr && threadIdx.x == 0 && blockIdx.x == ((gridDim.x / 2) & ~1u);
    if (clk_probe) {
        long long gt;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt));
        a.timeline[2112] = clock64();
        a.timeline[2113] = gt;
    }
    if (a.stagger_ns > 0) {
        // Persistent launch: identical CTAs started together stay in lockstep, so every HBM phase (residual / att loads, x' and
        // q/k/v stores) would be a device-wide burst with the memory system idle in between (measured: 8.5k cycles per tile
        // waiting for the residual).  Each cluster therefore starts at its own offset inside one tile period (golden-ratio
        // sequence over the cluster index); tiles take equal time, so the offsets persist for the whole launch.
        const uint32_t cid = blockIdx.x / CL;
        const uint32_t frac = (cid * 0x9E3779B1u) >> 16;                      // [0, 65536)
        const long long wait_ns = ((long long)a.stagger_ns * frac) >> 16;
        long long t0, t1;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
        do {
            __nanosleep(256);
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
        } while (t1 - t0 < wait_ns);
    }
    if (warp == PROD_WARP && lane == 0) {
        // the producer owns the barriers of its copies and starts them before the CTA-wide (and cluster-wide) rendezvous:
        // the att tile and the first S weight stages are in flight while TMEM is being allocated
        for (int s = 0; s < S; s++) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], 1);
            mbar_init(&pfull[s], 1);
        }
        for (int t = 0; t < NT; t++) mbar_init(&bar_att[t], 1);
        fence_barrier_init();
        fence_proxy_async_smem();
        for (int t = 0; t < NT; t++) {
            mbar_expect_tx(&bar_att[t], K::A_BYTES);
            bulk_g2s(As + t * K::A_BYTES, a.att + (size_t)(mt0 + t) * C * 128, K::A_BYTES, &bar_att[t]);
        }
        const uint8_t *src = reinterpret_cast<const uint8_t *>(a.wstream);
        for (int i = 0; i < S && i < n_stages; i++) {
            mbar_expect_tx(&full[i], K::SLOT_BYTES);
            bulk_g2s(ring + i * K::SLOT_BYTES, src + (size_t)i * K::STAGE_BYTES + crank * K::SLOT_BYTES, K::SLOT_BYTES, &full[i]);
        }
        if constexpr (!PERSIST) {
            const int nb = (int)blockIdx.x + a.pf_dist;
            if (a.pf_dist > 0 && nb < a.n_groups) {
                bulk_prefetch_l2(a.att + (size_t)nb * NT * C * 128, NT * C * 128 * 2);
                if (a.tab0 == nullptr) bulk_prefetch_l2(a.x + (size_t)nb * NT * C * 128, NT * C * 128 * (a.x_in_24 ? 3 : 4));
            }
        }
    }
    constexpr int HALF = C / NH;                          // residual columns handled by one worker thread
    float4 xv[HALF / 4];                                  // this thread's residual columns
    auto load_x = [&](int mtl) {
        const int row = (warp & 3) * 32 + lane, hh = (warp % NW) >> 2;
        if (a.tab0 != nullptr) {   // block 0: this thread's 2C contiguous bytes of its token's record (L2-resident table)
            const int tok = min((int)a.tokens0[(size_t)mtl * 128 + row], 66);
            const float4 *src = reinterpret_cast<const float4 *>(a.tab0 + ((size_t)tok * 256 + ((mtl & 1) << 7) + row) * a.tab_nrec) + hh * (HALF / 4);
#pragma unroll
            for (int j = 0; j < HALF / 4; j++) xv[j] = __ldg(src + j);
        } else if (a.x_in_24) {   // 3 pieces per 16 columns; the raw pieces wait in xv[0 .. 3 HALF/16) until the TMEM stores
            const float4 *Xl = reinterpret_cast<const float4 *>(a.x) + (size_t)mtl * (C / 4) * 128 + row;
#pragma unroll
            for (int j = 0; j < 3 * HALF / 16; j++) xv[j] = Xl[(size_t)(hh * (3 * HALF / 16) + j) * 128];
        } else {
            const float4 *Xl = reinterpret_cast<const float4 *>(a.x) + (size_t)mtl * (C / 4) * 128 + row;
#pragma unroll
            for (int j = 0; j < HALF / 4; j++) xv[j] = Xl[(size_t)(hh * (HALF / 4) + j) * 128];
        }
    };
    if constexpr (!PERSIST) {   // the residual tile is on its way to registers during the rendezvous
        if (warp < NW * NT) load_x(mt0 + warp / NW);
    }
    if (threadIdx.x == 0) {
        mbar_init(bar_proj, 1);
        mbar_init(bar_done, 1);
        for (int t = 0; t < NT; t++) {
            mbar_init(&bar_x[t], WARR + (PAIR ? 1 : 0));   // + the peer's "att tile landed" relay
            mbar_init(&bar_ln2[t], WARR);
            mbar_init(&bar_a1f[t], 1);
            mbar_init(&bar_a1e[t], WARR);
            mbar_init(&bar_hf[t], WARR);
            mbar_init(&bar_he[t], 1);
            mbar_init(&bar_qa[t], WARR);
            for (int b = 0; b < 3; b++) {
                mbar_init(&bar_qf[t * 3 + b], 1);
                mbar_init(&bar_qe[t * 3 + b], WARR);
            }
        }
        fence_barrier_init();
    }
    if (fuse_qkv && threadIdx.x < 3 * C / 8) {
        const int n = 8 * threadIdx.x;
        const int which = n / C, rem = n - which * C;
        const int head = rem / a.hs, d0 = rem - head * a.hs;
        qkv_off[threadIdx.x] = (uint32_t)(((which * a.n_head + head) * (a.hs / 8) + d0 / 8) * 256);
    }
    if (warp == MMA_WARP) {
        if (PAIR) tmem_alloc_pair<K::TMEM_COLS>(tmem_slot);
        else tmem_alloc<K::TMEM_COLS>(tmem_slot);
    }
    tc_fence_before();
    __syncthreads();
    if (PAIR) cluster_sync_all();         // the barriers of both CTAs exist before any remote arrive / multicast commit
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;
    // issuer -> workers (and ring slot release): when the UMMAs issued so far have retired
    auto commit = [&](uint64_t *bar) {
        if (PAIR) umma_commit_pair(bar, (uint16_t)3);
        else umma_commit(bar);
    };
    auto mma = [&](uint32_t d, uint64_t ad, uint64_t bd, uint32_t idesc, uint32_t acc) {
        if (PAIR) umma_ss_pair(d, ad, bd, idesc, acc);
        else umma_ss(d, ad, bd, idesc, acc);
    };
    // workers -> issuer: every lane has fenced its own writes; one elected lane arrives on the (leader's) barrier
    auto arrive_issuer = [&](uint64_t *bar) {
        __syncwarp();
        if (lane == 0) {
            if (!PAIR || leader) mbar_arrive(bar);
            else mbar_arrive_cluster(bar, 0);
        }
        __syncwarp();
    };

    if (warp == PROD_WARP) {
        // ------------------------------------------------------------------ producer
        if (lane == 0) {
            const uint8_t *src = reinterpret_cast<const uint8_t *>(a.wstream);
            int gi = S < n_stages ? S : n_stages;          // ring cursor over ALL tiles of this CTA (the first S stages are in flight)
            int it = 0;
            for (int g = blockIdx.x; g < n_groups; g += gstride, it++) {
                if (it > 0) {
                    // the A tile buffer is free once the previous tile's last c_attn UMMAs (half-tile 5 = second use of
                    // accumulator buffer 2 in that tile) have retired: wait for both of that tile's completions in order
                    for (int t = 0; t < NT; t++) {
                        mbar_wait(&bar_qf[t * 3 + 2], 0);
                        mbar_wait(&bar_qf[t * 3 + 2], 1);
                    }
                    for (int t = 0; t < NT; t++) {
                        mbar_expect_tx(&bar_att[t], K::A_BYTES);
                        bulk_g2s(As + t * K::A_BYTES, a.att + (size_t)(g * NT + t) * C * 128, K::A_BYTES, &bar_att[t]);
                    }
                }
                if (g + gstride < n_groups && a.tab0 == nullptr)   // the next tile's residual: on its way into L2 during this tile
                    bulk_prefetch_l2(a.x + (size_t)(g + gstride) * NT * C * 128, NT * C * 128 * (a.x_in_24 ? 3 : 4));
                for (int i = it == 0 ? gi : 0; i < n_stages; i++, gi++) {
                    const int s = gi % S;
                    mbar_wait(&empty[s], ((gi / S) & 1) ^ 1);
                    mbar_expect_tx(&full[s], K::SLOT_BYTES);
                    bulk_g2s(ring + s * K::SLOT_BYTES, src + (size_t)i * K::STAGE_BYTES + crank * K::SLOT_BYTES, K::SLOT_BYTES, &full[s]);
                }
                if constexpr (!PERSIST) break;
            }
        }
    } else if (warp == MMA_WARP && PAIR && !leader) {
        // ------------------------------------------------------------------ peer of a CTA pair: relay "landed" to the leader
        int gi = 0, it = 0;
        for (int g = blockIdx.x; g < n_groups; g += gstride, it++) {
            mbar_wait(&bar_att[0], it & 1);
            if (lane == 0) mbar_arrive_cluster(&bar_x[0], 0);
            __syncwarp();
            for (int i = 0; i < n_stages; i++, gi++) {
                const int s = gi % S;
                mbar_wait(&full[s], (gi / S) & 1);
                if (lane == 0) mbar_arrive_cluster(&pfull[s], 0);
                __syncwarp();
            }
            if constexpr (!PERSIST) break;
        }
    } else if (warp == MMA_WARP) {
        // ------------------------------------------------------------------ UMMA issuer
        // The whole warp runs the control flow (waits, stage cursor), so descriptors stay in uniform registers; one elected
        // lane issues.  (With everything inside `if (lane == 0)` ptxas rebuilt each descriptor in vector registers and moved
        // it over with R2UR: ~90 cycles of issue per UMMA, more than the UMMA itself takes.)
        {
            constexpr uint32_t idescC = umma_idesc_bf16(UM, C, 0, 0);
            constexpr uint32_t idescH = umma_idesc_bf16(UM, HC, 0, 0);
            const uint32_t a_addr = smem_u32(As), h_addr = smem_u32(Hs), r_addr = smem_u32(ring);
            int i = 0;  // ring cursor over ALL tiles of this CTA
            int it = 0;
            MG_LAP_INIT;
            auto stage_wait = [&](int idx) -> uint32_t {
                const int s = idx % S;
                if (lane == 0) MG_LAP(21);                   // issue + waits on the workers since the last lap
                mbar_wait(&full[s], (idx / S) & 1);
                if (PAIR) mbar_wait(&pfull[s], (idx / S) & 1);
                tc_fence_after();
                if (lane == 0) MG_LAP(20);                   // waiting for the weight ring
                return r_addr + s * K::SLOT_BYTES;
            };
            for (int g = blockIdx.x; g < n_groups; g += gstride, it++) {
            mg_group = g;
            const uint32_t ph = it & 1;                      // parity of the barriers that complete once per tile
            // proj: acc_main (pre-loaded with x) += att @ Wproj^T
            if (elect_one()) MG_STAMP(0);
            __syncwarp();
            for (int t = 0; t < NT; t++) {
                mbar_wait(&bar_att[t], ph);
                mbar_wait(&bar_x[t], ph);
            }
            tc_fence_after();
            if (lane == 0) MG_LAP(22);                       // att tile + residual pre-load
            if (elect_one()) MG_STAMP(1);
            __syncwarp();
            for (int st = 0; st < K::NPROJ / U; st++, i++) {
                const uint32_t b = stage_wait(i);
                if (elect_one()) {
#pragma unroll
                    for (int t = 0; t < NT; t++)
#pragma unroll
                        for (int u = 0; u < U; u++)
                            mma(tmem + t * K::TILE_COLS, umma_desc(a_addr + t * K::A_BYTES + (st * U + u) * 4096, 2048, 128),
                                umma_desc(b + u * UNIT_B, CB * 16, 128), idescC, 1u);
                    commit(&empty[i % S]);
                    if (st == K::NPROJ / U - 1) {
                        commit(bar_proj);
                        MG_STAMP(2);
                    }
                }
                __syncwarp();
            }
            auto fc = [&](int j) {
                for (int st = 0; st < K::NFC / U; st++, i++) {
                    const uint32_t b = stage_wait(i);
#pragma unroll
                    for (int t = 0; t < NT; t++) {
                        if (st == 0) {
                            if (j == 0) mbar_wait(&bar_ln2[t], ph);
                            else mbar_wait(&bar_a1e[t], (j - 1) & 1);
                            tc_fence_after();
                        }
                        if (elect_one()) {
                            if constexpr (K::WIDE) {   // one k-step of all C hidden columns of the chunk per unit (proj's unit format)
#pragma unroll
                                for (int u = 0; u < U; u++)
                                    mma(tmem + t * K::TILE_COLS + C, umma_desc(a_addr + t * K::A_BYTES + (st * U + u) * 4096, 2048, 128),
                                        umma_desc(b + u * UNIT_B, CB * 16, 128), idescC, (st | u) != 0);
                            } else {
#pragma unroll
                            for (int u = 0; u < U; u++)
#pragma unroll
                                for (int ks = 0; ks < 2; ks++)
                                    mma(tmem + t * K::TILE_COLS + C,
                                        umma_desc(a_addr + t * K::A_BYTES + ((st * U + u) * 2 + ks) * 4096, 2048, 128),
                                        umma_desc(b + u * UNIT_B + ks * 2 * (HB * 16), HB * 16, 128), idescH,
                                        (st | u | ks) != 0);
                            }
                            if (st == K::NFC / U - 1) commit(&bar_a1f[t]);
                            if (t == NT - 1) {
                                commit(&empty[i % S]);
                                if (st == K::NFC / U - 1) MG_STAMP(10 + 2 * j);
                            }
                        }
                        __syncwarp();
                    }
                }
            };
            auto p2 = [&](int j) {
                for (int st = 0; st < K::NP2 / U; st++, i++) {
                    const uint32_t b = stage_wait(i);
#pragma unroll
                    for (int t = 0; t < NT; t++) {
                        if (st == 0) {
                            mbar_wait(&bar_hf[t], j & 1);
                            tc_fence_after();
                        }
                        if (elect_one()) {
#pragma unroll
                            for (int u = 0; u < U; u++)
                                mma(tmem + t * K::TILE_COLS, umma_desc(h_addr + t * K::H_BYTES + (st * U + u) * 4096, 2048, 128),
                                    umma_desc(b + u * UNIT_B, CB * 16, 128), idescC, 1u);
                            if (st == K::NP2 / U - 1) commit(&bar_he[t]);
                            if (t == NT - 1) {
                                commit(&empty[i % S]);
                                if (st == K::NP2 / U - 1) {
                                    if (j == K::NCH - 1) commit(bar_done);
                                    MG_STAMP(11 + 2 * j);
                                }
                            }
                        }
                        __syncwarp();
                    }
                }
            };
            fc(0);
            for (int j = 0; j < K::NCH; j++) {
                if (j + 1 < K::NCH) fc(j + 1);
                p2(j);
            }
            if (fuse_qkv) {
                // next block's c_attn: [q|k|v] = LN1_next(x') @ Wqkv^T as six HC-wide half n-tiles rotating through three
                // accumulator buffers (the FC accumulator and the two halves of the main one, all free by now), so the
                // UMMAs of half-tile h+1 run while the workers drain half-tile h
                for (int t = 0; t < NT; t++) mbar_wait(&bar_qa[t], ph);
                tc_fence_after();
                for (int hh = 0; hh < 6; hh++) {
                    const int buf = hh % 3;
                    const uint32_t col = buf == 0 ? C : (buf == 1 ? 0 : HC);
                    for (int st = 0; st < K::NFCQ / U; st++, i++) {
                        const uint32_t b = stage_wait(i);
#pragma unroll
                        for (int t = 0; t < NT; t++) {
                            if (st == 0 && hh >= 3) {
                                mbar_wait(&bar_qe[t * 3 + buf], 0);
                                tc_fence_after();
                            }
                            if (elect_one()) {
#pragma unroll
                                for (int u = 0; u < U; u++)
#pragma unroll
                                    for (int ks = 0; ks < 2; ks++)
                                        mma(tmem + t * K::TILE_COLS + col,
                                            umma_desc(a_addr + t * K::A_BYTES + ((st * U + u) * 2 + ks) * 4096, 2048, 128),
                                            umma_desc(b + u * UNIT_B + ks * 2 * (HB * 16), HB * 16, 128), idescH,
                                            (st | u | ks) != 0);
                                if (st == K::NFCQ / U - 1) commit(&bar_qf[t * 3 + buf]);
                                if (t == NT - 1) commit(&empty[i % S]);
                            }
                            __syncwarp();
                        }
                    }
                }
                if (elect_one()) MG_STAMP(41);
                __syncwarp();
            }
            if constexpr (!PERSIST) break;
            }   // tile groups
            if (lane == 0) MG_LAP(21);
        }
    } else {
        // ------------------------------------------------------------------ workers (256 threads per tile)
        const int t = warp / NW;                          // tile of this worker
        const int q = warp & 3, h = (warp % NW) >> 2;     // TMEM lane quadrant, column group
        const int r = q * 32 + lane;
        const uint32_t trow = tmem + t * K::TILE_COLS + ((uint32_t)(q * 32) << 16);
        uint8_t *At = As + t * K::A_BYTES;
        float *red_s = red + t * (2 * NH * 128), *red_q = red_s + NH * 128;
        const float inv_c = 1.0f / (float)C;
        const uint32_t nb = 1 + t;                        // named barrier of this tile's 256 workers
        MG_LAP_INIT;
#define MG_WLAP(id) do { if (threadIdx.x == 0) MG_LAP(id); } while (0)
        int it = 0;
#pragma unroll 1
        for (int g = blockIdx.x; g < n_groups; g += gstride, it++) {
        mg_group = g;
        const uint32_t ph = it & 1;                       // parity of the barriers that complete once per tile
        const int mt = g * NT + t;
        float4 *Xg = reinterpret_cast<float4 *>(a.x) + (size_t)mt * (C / 4) * 128 + r;
        const bool keep_row = !a.tail_rows_only || ((mt & 1) && r == 127);   // token 255 = row 127 of the sequence's second tile

        // ---- x -> TMEM accumulator (overlaps the att tile load); c_proj then accumulates onto it.
        // All loads are issued before the first TMEM store so the thread pays ONE memory round trip.
        {
            if constexpr (PERSIST) load_x(mt);            // all loads in flight, then the stores (L2 prefetch sent a tile ago)
            if (a.x_in_24 && a.tab0 == nullptr) {
#pragma unroll
                for (int i = 0; i < HALF / 16; i++) {
                    uint32_t v[16];
                    unpack24x16(xv[3 * i], xv[3 * i + 1], xv[3 * i + 2], v);
                    tmem_st16(trow + h * HALF + 16 * i, v);
                }
            } else {
#pragma unroll
            for (int i = 0; i < HALF / 16; i++) {
                uint32_t v[16];
#pragma unroll
                for (int j = 0; j < 4; j++) {
                    v[4 * j + 0] = __float_as_uint(xv[4 * i + j].x); v[4 * j + 1] = __float_as_uint(xv[4 * i + j].y);
                    v[4 * j + 2] = __float_as_uint(xv[4 * i + j].z); v[4 * j + 3] = __float_as_uint(xv[4 * i + j].w);
                }
                tmem_st16(trow + h * HALF + 16 * i, v);
            }
            }
        }
        tmem_wait_st();
        tc_fence_before();
        arrive_issuer(&bar_x[t]);
        MG_WLAP(0);      // residual tile -> TMEM

        // ---- epilogue 1: LN2(x1) -> A tile (x1 = x + proj stays in TMEM)
        mbar_wait(bar_proj, ph);
        tc_fence_after();
        MG_WLAP(1);      // waiting for att tile + c_proj
        if (threadIdx.x == 0) MG_STAMP(50);
        {
            f32x2 sum2 = pk2(0.f, 0.f), sq2 = pk2(0.f, 0.f);
            auto stats = [&](auto c0_, auto n_) {
                constexpr int C0 = decltype(c0_)::value, N = decltype(n_)::value;
                uint32_t v[N];
                tmem_ld_n<N>(trow + h * HALF + C0, v);
                tmem_wait_ld();
#pragma unroll
                for (int j = 0; j < N; j += 2) {
                    const f32x2 e = pk2u(v[j], v[j + 1]);
                    sum2 = add2(sum2, e);
                    sq2 = fma2(e, e, sq2);
                }
            };
            MG_COL_BATCHES(HALF, stats);
            if (threadIdx.x == 0) MG_STAMP(51);
            {
                float s0, s1, q0, q1;
                upk2(sum2, s0, s1);
                upk2(sq2, q0, q1);
                red_s[h * 128 + r] = s0 + s1;
                red_q[h * 128 + r] = q0 + q1;
            }
            named_bar_sync(nb, 128 * NH);
            float ssum = 0.f, qsum = 0.f;
#pragma unroll
            for (int g2 = 0; g2 < NH; g2++) {
                ssum += red_s[g2 * 128 + r];
                qsum += red_q[g2 * 128 + r];
            }
            const float mean = ssum * inv_c;
            const float var = fmaxf(qsum * inv_c - mean * mean, 0.f);
            const float rstd = rsqrtf(var + 1e-5f);
            const f32x2 la = pk2(rstd, rstd), lb = pk2(-mean * rstd, -mean * rstd);
            auto norm = [&](auto c0_, auto n_) {
                constexpr int C0 = decltype(c0_)::value, N = decltype(n_)::value;
                uint32_t v[N];
                tmem_ld_n<N>(trow + h * HALF + C0, v);
                tmem_wait_ld();
#pragma unroll
                for (int j = 0; j < N / 8; j++)
                    *reinterpret_cast<uint4 *>(At + (((h * HALF + C0) / 8 + j) * 128 + r) * 16) = ln_pack8_ng(&v[8 * j], la, lb);
            };
            MG_COL_BATCHES(HALF, norm);
            tc_fence_before();
            fence_proxy_async_smem();
            arrive_issuer(&bar_ln2[t]);
            MG_WLAP(2);  // LN2
            if (threadIdx.x == 0) MG_STAMP(52);
        }

        // ---- MLP chunks: acc1 -> GELU -> hidden chunk in smem
        constexpr int HH = K::HM / NH;        // hidden columns per thread per chunk
        constexpr int SB = HH > 64 ? 64 : HH; // ... drained / activated / stored in sub-batches of SB columns (registers)
        constexpr int NSUB = HH / SB;
        constexpr int NV = SB / 8;            // 16-byte groups per sub-batch
        static_assert(HH % SB == 0 && SB % 8 == 0, "MLP chunk sub-batches");
        uint8_t *Hb = Hs + t * K::H_BYTES;
        // (Issuing the TMEM reads of chunk j+1 before chunk j's hidden values are stored -- a register ping-pong software
        // pipeline -- was measured SLOWER, 2.18 vs 1.80 ms per launch: it needs 40 more live registers under the 96-register
        // cap and delays the hidden-chunk hand-off whenever FC(j+1) is late.  Same for the q/k/v drain below.)
#pragma unroll 1
        for (int j = 0; j < K::NCH; j++) {
            mbar_wait(&bar_a1f[t], j & 1);
            tc_fence_after();
            MG_WLAP(3);  // waiting for the FC chunk
            if (threadIdx.x == 0) MG_STAMP(60 + 3 * j);
#pragma unroll
            for (int sub = 0; sub < NSUB; sub++) {
                uint32_t v[NV][8];
#pragma unroll
                for (int g = 0; g < NV; g++) tmem_ld8(trow + C + h * HH + sub * SB + g * 8, v[g]);
                tmem_wait_ld();
                if (sub == NSUB - 1) {            // the whole chunk is in registers / already stored: FC(j+1) may overwrite it
                    tc_fence_before();
                    arrive_issuer(&bar_a1e[t]);
                    MG_WLAP(4);  // FC chunk -> registers
                    if (threadIdx.x == 0) MG_STAMP(61 + 3 * j);
                }
                uint4 o[NV];
#pragma unroll
                for (int g = 0; g < NV; g++) {
                    // MG_GELU_MIX_BITS: which of every 8 consecutive pairs take the FMA-only form instead of the MUFU.TANH one
                    auto act = [&](int k) {
                        const float x0 = __uint_as_float(v[g][2 * k]), x1 = __uint_as_float(v[g][2 * k + 1]);
                        return pack_bf16x2_p(((MG_GELU_MIX_BITS >> ((4 * g + k) & 7)) & 1) ? gelu2_fma(x0, x1) : gelu2(x0, x1));
                    };
                    o[g].x = act(0);
                    o[g].y = act(1);
                    o[g].z = act(2);
                    o[g].w = act(3);
                }
                if (sub == NSUB - 1) MG_WLAP(5);  // GELU
                if (sub == 0 && j >= 1) {
                    mbar_wait(&bar_he[t], (j - 1) & 1);   // proj2(j-1) has finished reading the buffer
                    MG_WLAP(6);  // waiting for the hidden buffer
                }
#pragma unroll
                for (int g = 0; g < NV; g++) *reinterpret_cast<uint4 *>(Hb + ((h * (HH / 8) + sub * NV + g) * 128 + r) * 16) = o[g];
            }
            fence_proxy_async_smem();
            arrive_issuer(&bar_hf[t]);
            MG_WLAP(7);  // hidden chunk -> smem
            if (threadIdx.x == 0) MG_STAMP(62 + 3 * j);
        }

        // ---- final epilogue: x' -> HBM, xn = LN1_next(x') -> HBM
        mbar_wait(bar_done, ph);
        tc_fence_after();
        MG_WLAP(8);      // waiting for the last proj2
        if (threadIdx.x == 0) MG_STAMP(90);
        {
            f32x2 sum2 = pk2(0.f, 0.f), sq2 = pk2(0.f, 0.f);
            auto store_x = [&](auto c0_, auto n_) {
                constexpr int C0 = decltype(c0_)::value, N = decltype(n_)::value;
                uint32_t v[N];
                tmem_ld_n<N>(trow + h * HALF + C0, v);
                tmem_wait_ld();
                if (a.x_out_24) {
#pragma unroll
                    for (int gq = 0; gq < N / 16; gq++) {
                        uint4 o[3];
                        pack24x16(&v[16 * gq], o);
#pragma unroll
                        for (int k = 0; k < 3; k++)
                            __stcs(reinterpret_cast<uint4 *>(&Xg[(size_t)(((h * HALF + C0) / 16 + gq) * 3 + k) * 128]), o[k]);
                    }
                }
#pragma unroll
                for (int j = 0; j < N / 4; j++) {
                    if (keep_row && !a.x_out_24)
                    __stcs(&Xg[(size_t)((h * HALF + C0) / 4 + j) * 128], make_float4(__uint_as_float(v[4 * j + 0]), __uint_as_float(v[4 * j + 1]),
                                                                                      __uint_as_float(v[4 * j + 2]), __uint_as_float(v[4 * j + 3])));
                    const f32x2 e0 = pk2u(v[4 * j + 0], v[4 * j + 1]), e1 = pk2u(v[4 * j + 2], v[4 * j + 3]);
                    sum2 = add2(sum2, add2(e0, e1));
                    sq2 = fma2(e0, e0, sq2);
                    sq2 = fma2(e1, e1, sq2);
                }
            };
            MG_COL_BATCHES(HALF, store_x);
            MG_WLAP(9);  // x' -> HBM + statistics
            if (threadIdx.x == 0) MG_STAMP(92);
            if (a.xn_out != nullptr || fuse_qkv) {
                // the scratch was last read in the LN2 epilogue; every worker has long passed that point (the mbarrier chain of
                // the MLP phase orders it), the barrier states it in a form compute-sanitizer's racecheck can follow
                named_bar_sync(nb, 128 * NH);
                {
                    float s0, s1, q0, q1;
                    upk2(sum2, s0, s1);
                    upk2(sq2, q0, q1);
                    red_s[h * 128 + r] = s0 + s1;
                    red_q[h * 128 + r] = q0 + q1;
                }
                named_bar_sync(nb, 128 * NH);
                float ssum = 0.f, qsum = 0.f;
#pragma unroll
                for (int g2 = 0; g2 < NH; g2++) {
                    ssum += red_s[g2 * 128 + r];
                    qsum += red_q[g2 * 128 + r];
                }
                const float mean = ssum * inv_c;
                const float var = fmaxf(qsum * inv_c - mean * mean, 0.f);
                const float rstd = rsqrtf(var + 1e-5f);
                const f32x2 la = pk2(rstd, rstd), lb = pk2(-mean * rstd, -mean * rstd);
                if (fuse_qkv) {
                    auto norm = [&](auto c0_, auto n_) {
                        constexpr int C0 = decltype(c0_)::value, N = decltype(n_)::value;
                        uint32_t v[N];
                        tmem_ld_n<N>(trow + h * HALF + C0, v);
                        tmem_wait_ld();
#pragma unroll
                        for (int j = 0; j < N / 8; j++)
                            *reinterpret_cast<uint4 *>(At + (((h * HALF + C0) / 8 + j) * 128 + r) * 16) = ln_pack8_ng(&v[8 * j], la, lb);
                    };
                    MG_COL_BATCHES(HALF, norm);
                } else {
                    const float4 *g4 = reinterpret_cast<const float4 *>(a.next_gain);
                    uint4 *O = reinterpret_cast<uint4 *>(a.xn_out) + (size_t)mt * (C / 8) * 128 + r;
#pragma unroll 1
                    for (int c0 = h * HALF; c0 < (h + 1) * HALF; c0 += 16) {
                        uint32_t v[16];
                        tmem_ld16(trow + c0, v);
                        tmem_wait_ld();
#pragma unroll
                        for (int j = 0; j < 2; j++)
                            O[(size_t)(c0 / 8 + j) * 128] =
                                ln_pack8(&v[8 * j], la, lb, __ldg(g4 + c0 / 4 + 2 * j), __ldg(g4 + c0 / 4 + 2 * j + 1));
                    }
                }
            }
        }
        if (fuse_qkv) {
            tc_fence_before();
            fence_proxy_async_smem();
            arrive_issuer(&bar_qa[t]);
            MG_WLAP(10); // LN1_next -> smem
            if (threadIdx.x == 0) MG_STAMP(93);
            const int seq = mt >> 1, tok = ((mt & 1) << 7) + r;
            uint4 *Oseq = reinterpret_cast<uint4 *>(a.qkv_out) + (size_t)seq * (3 * C / 8) * 256 + tok;
#pragma unroll 1
            for (int hh = 0; hh < 6; hh++) {
                const int buf = hh % 3;
                const uint32_t col = (buf == 0 ? C : (buf == 1 ? 0 : HC)) + h * (HC / NH);
                mbar_wait(&bar_qf[t * 3 + buf], (hh / 3) & 1);
                tc_fence_after();
                MG_WLAP(11);  // waiting for a q/k/v half-tile
                if (threadIdx.x == 0) MG_STAMP(94 + hh);
                uint32_t v[HC / NH];
#pragma unroll
                for (int c = 0; c < HC / NH; c += 8) tmem_ld8(trow + col + c, *reinterpret_cast<uint32_t(*)[8]>(&v[c]));
                tmem_wait_ld();
                tc_fence_before();
                arrive_issuer(&bar_qe[t * 3 + buf]);   // values are in registers: the issuer may reuse the buffer
#pragma unroll
                for (int j = 0; j < HC / NH / 8; j++) {
                    uint4 o;
                    o.x = pack_bf16x2(__uint_as_float(v[8 * j + 0]), __uint_as_float(v[8 * j + 1]));
                    o.y = pack_bf16x2(__uint_as_float(v[8 * j + 2]), __uint_as_float(v[8 * j + 3]));
                    o.z = pack_bf16x2(__uint_as_float(v[8 * j + 4]), __uint_as_float(v[8 * j + 5]));
                    o.w = pack_bf16x2(__uint_as_float(v[8 * j + 6]), __uint_as_float(v[8 * j + 7]));
                    if (hh >= 2 || keep_row)   // hh 0, 1 = the q columns
                    __stcs(&Oseq[qkv_off[(hh * HC + h * (HC / NH)) / 8 + j]], o);   // streaming: 3.4 GB per launch pass through L2 once
                }
                MG_WLAP(12);  // q/k/v half-tile -> HBM
            }
        }
        // the next tile's residual goes into accumulator columns that OTHER workers of this tile have just drained
        if constexpr (!PERSIST) {
            MG_WLAP(13);
            break;
        }
        if (g + gstride < n_groups) named_bar_sync(nb, 128 * NH);
        MG_WLAP(13);
        }   // tile groups
    }
    if (threadIdx.x == 0) MG_STAMP(91);
    if (clk_probe) {
        long long gt;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt));
        a.timeline[2114] = clock64();
        a.timeline[2115] = gt;
    }
    tc_fence_before();
    __syncthreads();
    if (PAIR) cluster_sync_all();         // no CTA leaves (or frees TMEM) while the pair's UMMAs / commits may still target it
    if (warp == MMA_WARP) {
        if (PAIR) tmem_dealloc_pair<K::TMEM_COLS>(tmem);
        else tmem_dealloc<K::TMEM_COLS>(tmem);
    }
}

// embedding + ln_1 of block 0 in one pass (model.py:171-175 + model.py:102): X_ti and A_ti out.
// thread == row.  wte (67 x C fp32) is staged in shared memory with a padded row pitch (random token rows would
// otherwise hit the same banks); wpe arrives pre-tiled ([2][C/4][128][4]) so position rows are coalesced loads.
__global__ void __launch_bounds__(128) embed_ln_kernel(const uint8_t *__restrict__ tokens, const float *__restrict__ wte,
                                                       const float *__restrict__ wpe_ti, const float *__restrict__ gain,
                                                       float *__restrict__ X, __nv_bfloat16 *__restrict__ XN, int C)
{
    extern __shared__ __align__(16) float wte_s[];      // [67][C + 4]
    const int pitch = C + 4;
    for (int i = threadIdx.x; i < 67 * (C / 4); i += 128) {
        const int row = i / (C / 4), c4 = i - row * (C / 4);
        *reinterpret_cast<float4 *>(wte_s + row * pitch + c4 * 4) = __ldg(reinterpret_cast<const float4 *>(wte) + i);
    }
    __syncthreads();
    const int mt = blockIdx.x, r = threadIdx.x;
    const int tok = min((int)tokens[(size_t)mt * 128 + r], 66);   // ids are validated on the host; clamp like block0_lookup
    const float4 *te = reinterpret_cast<const float4 *>(wte_s + tok * pitch);
    const float4 *pe = reinterpret_cast<const float4 *>(wpe_ti) + (size_t)(mt & 1) * (C / 4) * 128 + r;
    float4 *Xo = reinterpret_cast<float4 *>(X) + (size_t)mt * (C / 4) * 128 + r;
    float s = 0.f;
    for (int c4 = 0; c4 < C / 4; c4++) {
        const float4 t = te[c4], p = __ldg(pe + (size_t)c4 * 128);
        const float4 v = make_float4(t.x + p.x, t.y + p.y, t.z + p.z, t.w + p.w);
        Xo[(size_t)c4 * 128] = v;
        s += (v.x + v.y) + (v.z + v.w);
    }
    const float mean = s / (float)C;
    float q = 0.f;
    for (int c4 = 0; c4 < C / 4; c4++) {
        const float4 t = te[c4], p = __ldg(pe + (size_t)c4 * 128);
        const float a0 = t.x + p.x - mean, a1 = t.y + p.y - mean, a2 = t.z + p.z - mean, a3 = t.w + p.w - mean;
        q += (a0 * a0 + a1 * a1) + (a2 * a2 + a3 * a3);
    }
    const float rstd = rsqrtf(q / (float)C + 1e-5f);
    uint4 *O = reinterpret_cast<uint4 *>(XN) + (size_t)mt * (C / 8) * 128 + r;
    const float4 *g4 = reinterpret_cast<const float4 *>(gain);
    for (int c8 = 0; c8 < C / 8; c8++) {
        const float4 t0 = te[2 * c8], p0 = __ldg(pe + (size_t)(2 * c8) * 128);
        const float4 t1 = te[2 * c8 + 1], p1 = __ldg(pe + (size_t)(2 * c8 + 1) * 128);
        const float4 g0 = __ldg(g4 + 2 * c8), g1 = __ldg(g4 + 2 * c8 + 1);
        uint4 o;
        o.x = pack_bf16x2((t0.x + p0.x - mean) * rstd * g0.x, (t0.y + p0.y - mean) * rstd * g0.y);
        o.y = pack_bf16x2((t0.z + p0.z - mean) * rstd * g0.z, (t0.w + p0.w - mean) * rstd * g0.w);
        o.z = pack_bf16x2((t1.x + p1.x - mean) * rstd * g1.x, (t1.y + p1.y - mean) * rstd * g1.y);
        o.w = pack_bf16x2((t1.z + p1.z - mean) * rstd * g1.z, (t1.w + p1.w - mean) * rstd * g1.w);
        O[(size_t)c8 * 128] = o;
    }
}

// ------------------------------------------------------------------------------------------------
// Block 0 as a lookup.  x = wte[token] + wpe[position] (model.py:171-175) depends only on (token id, position), and so do
// ln_1(x) and q/k/v = ln_1(x) @ Wqkv^T of block 0 (model.py:102,50): 67 x 256 combinations.  At model load the engine runs
// embed_ln_kernel + the QKV GEMM once on 67 synthetic sequences (sequence i = token i at every position), i.e. through the
// SAME kernels, so the looked-up values are bit-identical to the computed ones; the table (one record of C fp32 + 3C bf16 per
// combination, 27 MB at C = 160) stays L2-resident.  Per timestep block 0 is then a pure HBM-write-bound gather:
// 4C + 6C bytes per token out, no GEMM, no xn round trip.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) block0_table_kernel(const float *__restrict__ X, const __nv_bfloat16 *__restrict__ QKV,
                                                           uint4 *__restrict__ tab, int C)
{
    const int mt = blockIdx.x, r = threadIdx.x, tok = mt >> 1, t = ((mt & 1) << 7) + r;
    const int nx = C / 4, nq = 3 * C / 8;
    uint4 *rec = tab + ((size_t)tok * 256 + t) * (nx + nq);
    const uint4 *X4 = reinterpret_cast<const uint4 *>(X) + (size_t)mt * nx * 128 + r;
    const uint4 *Q4 = reinterpret_cast<const uint4 *>(QKV) + (size_t)tok * nq * 256 + t;
    for (int g = 0; g < nx; g++) rec[g] = X4[(size_t)g * 128];
    for (int g = 0; g < nq; g++) rec[nx + g] = Q4[(size_t)g * 256];
}
// The table must stay in L2 while 3.4 GB of output stream through it: table reads carry an evict_last L2 policy, output
// stores are streaming (evict_first).
#ifndef MG_LOOKUP_PLAIN
__device__ __forceinline__ uint64_t l2_policy_evict_last()
{
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ uint4 ld_keep_l2(const uint4 *p, uint64_t policy)
{
    uint4 v;
    asm volatile("ld.global.nc.L2::cache_hint.v4.u32 {%0,%1,%2,%3}, [%4], %5;"
                 : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
                 : "l"(p), "l"(policy));
    return v;
}
__device__ __forceinline__ void st_stream(uint4 *p, const uint4 v) { __stcs(p, v); }
#else
__device__ __forceinline__ uint64_t l2_policy_evict_last() { return 0; }
__device__ __forceinline__ uint4 ld_keep_l2(const uint4 *p, uint64_t) { return __ldg(p); }
__device__ __forceinline__ void st_stream(uint4 *p, const uint4 v) { *p = v; }
#endif
// One block per 128-token tile.  Records are read the way they lie in memory (consecutive threads read consecutive 16-byte
// groups of one record: a warp-level load touches 2-4 contiguous segments instead of 32 lines), transposed through shared
// memory (row pitch CH + 1 groups: conflict-free 16-byte accesses), and written token-major (512 contiguous bytes per warp).
template <int C>
__global__ void __launch_bounds__(128) block0_lookup_kernel(const uint8_t *__restrict__ tokens, const uint4 *__restrict__ tab,
                                                            float *__restrict__ X, __nv_bfloat16 *__restrict__ QKV)   // X == nullptr: q/k/v only
{
    constexpr int NX = C / 4, NQ = 3 * C / 8, NREC = NX + NQ;
    constexpr int CH = C == 160 ? 10 : 8, PITCH = CH + 1;       // groups per pass; NX and NQ are multiples of CH
    static_assert(NX % CH == 0 && NQ % CH == 0, "block0_lookup_kernel: chunking");
    __shared__ uint4 S[128 * PITCH];
    __shared__ int rec_of[128];
    const uint64_t keep = l2_policy_evict_last();
    const int mt = blockIdx.x, r = threadIdx.x, seq = mt >> 1, t = ((mt & 1) << 7) + r;
    rec_of[r] = (min((int)tokens[(size_t)mt * 128 + r], 66) * 256 + t) * NREC;
    uint4 *X4 = reinterpret_cast<uint4 *>(X) + (size_t)mt * NX * 128 + r;
    uint4 *Q4 = reinterpret_cast<uint4 *>(QKV) + (size_t)seq * NQ * 256 + t;
    __syncthreads();
#pragma unroll 1
    for (int g0 = X != nullptr ? 0 : NX; g0 < NREC; g0 += CH) {
        uint4 v[CH];
#pragma unroll
        for (int k = 0; k < CH; k++) {
            const int f = k * 128 + r, tk = f / CH, gg = f - tk * CH;
            v[k] = ld_keep_l2(tab + rec_of[tk] + g0 + gg, keep);
        }
#pragma unroll
        for (int k = 0; k < CH; k++) {
            const int f = k * 128 + r, tk = f / CH, gg = f - tk * CH;
            S[tk * PITCH + gg] = v[k];
        }
        __syncthreads();
        if (g0 < NX) {
#pragma unroll
            for (int gg = 0; gg < CH; gg++) st_stream(X4 + (size_t)(g0 + gg) * 128, S[r * PITCH + gg]);
        } else {
#pragma unroll
            for (int gg = 0; gg < CH; gg++) st_stream(Q4 + (size_t)(g0 - NX + gg) * 256, S[r * PITCH + gg]);
        }
        __syncthreads();
    }
}

// ------------------------------------------------------------------------------------------------
// Last block pruning: only logits[255][0:5] are consumed (model.py:186,249-252), so in the LAST
// block K and V are needed for all 256 tokens but Q, attention output, c_proj and the MLP only for token 255.
// last_attn_kernel: one warp per (sequence, head) on CUDA cores (1 query x 256 keys x hs): softmax(q K^T / sqrt(hs)) V,
// written into a COMPACT A tile image (row = sequence), and the residual row of token 255 gathered next to it.
// post_attn_kernel then runs on n_seq rows instead of 256 * n_seq.
// ------------------------------------------------------------------------------------------------
template <int HS>
__global__ void __launch_bounds__(512) last_attn_kernel(const __nv_bfloat16 *__restrict__ qkv, const float *__restrict__ X,
                                                        __nv_bfloat16 *__restrict__ att_c, float *__restrict__ x_c, int n_head,
                                                        int C, float scale)
{
    const int seq = blockIdx.x, head = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int mtc = seq >> 7, rc = seq & 127;
    // gather x[seq*256 + 255] -> compact row
    {
        const float4 *src = reinterpret_cast<const float4 *>(X) + (size_t)(seq * 2 + 1) * (C / 4) * 128 + 127;
        float4 *dst = reinterpret_cast<float4 *>(x_c) + (size_t)mtc * (C / 4) * 128 + rc;
        for (int c4 = threadIdx.x; c4 < C / 4; c4 += blockDim.x) dst[(size_t)c4 * 128] = src[(size_t)c4 * 128];
    }
    if (head >= n_head) return;
    const size_t blk = (size_t)(HS / 8) * 256 * 8;
    const __nv_bfloat16 *Qg = qkv + (((size_t)seq * 3 + 0) * n_head + head) * blk;
    const uint4 *Kg = reinterpret_cast<const uint4 *>(qkv + (((size_t)seq * 3 + 1) * n_head + head) * blk);
    const __nv_bfloat16 *Vg = qkv + (((size_t)seq * 3 + 2) * n_head + head) * blk;
    float q[HS];
#pragma unroll
    for (int c = 0; c < HS / 8; c++) {
        const uint4 u = *reinterpret_cast<const uint4 *>(Qg + ((size_t)c * 256 + 255) * 8);
        const __nv_bfloat162 *h2 = reinterpret_cast<const __nv_bfloat162 *>(&u);
#pragma unroll
        for (int j = 0; j < 4; j++) {
            const float2 f = __bfloat1622float2(h2[j]);
            q[8 * c + 2 * j] = f.x; q[8 * c + 2 * j + 1] = f.y;
        }
    }
    float s[8];
    float mx = -INFINITY;
#pragma unroll
    for (int i = 0; i < 8; i++) {
        const int key = lane + 32 * i;
        float acc = 0.f;
#pragma unroll
        for (int c = 0; c < HS / 8; c++) {
            const uint4 u = Kg[(size_t)c * 256 + key];
            const __nv_bfloat162 *h2 = reinterpret_cast<const __nv_bfloat162 *>(&u);
#pragma unroll
            for (int j = 0; j < 4; j++) {
                const float2 f = __bfloat1622float2(h2[j]);
                acc = fmaf(q[8 * c + 2 * j], f.x, acc);
                acc = fmaf(q[8 * c + 2 * j + 1], f.y, acc);
            }
        }
        s[i] = acc * scale;
        mx = fmaxf(mx, s[i]);
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    float sum = 0.f;
#pragma unroll
    for (int i = 0; i < 8; i++) {
        s[i] = __expf(s[i] - mx);
        sum += s[i];
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    const float inv = 1.0f / sum;
    // o[d] = sum_j p_j v[j][d]: every lane accumulates all HS dims over ITS OWN 8 keys from 16-byte V loads (a warp-level load
    // is 512 contiguous bytes), then the 32 partial vectors are summed across the warp by a transposing butterfly: after step
    // k a lane holds HS >> k dims, after five steps lane l holds dims {l} (hs 32) or {2l, 2l+1} (hs 64) summed over all keys.
    float o[HS];
#pragma unroll
    for (int d = 0; d < HS; d++) o[d] = 0.f;
    const uint4 *Vg4 = reinterpret_cast<const uint4 *>(Vg);
#pragma unroll
    for (int i = 0; i < 8; i++) {
        const int key = lane + 32 * i;
#pragma unroll
        for (int c = 0; c < HS / 8; c++) {
            const uint4 u = Vg4[(size_t)c * 256 + key];
            const __nv_bfloat162 *h2 = reinterpret_cast<const __nv_bfloat162 *>(&u);
#pragma unroll
            for (int j = 0; j < 4; j++) {
                const float2 f = __bfloat1622float2(h2[j]);
                o[8 * c + 2 * j] = fmaf(s[i], f.x, o[8 * c + 2 * j]);
                o[8 * c + 2 * j + 1] = fmaf(s[i], f.y, o[8 * c + 2 * j + 1]);
            }
        }
    }
    // butterfly: at step `off` a lane keeps the half of its dims selected by bit `off` of its lane id and adds the partner's
#pragma unroll
    for (int off = 16, n = HS; off >= 1; off >>= 1, n >>= 1) {
        const bool up = (lane & off) != 0;
#pragma unroll
        for (int d = 0; d < n / 2; d++) {
            const float keep = up ? o[d + n / 2] : o[d];
            const float give = up ? o[d] : o[d + n / 2];
            o[d] = keep + __shfl_xor_sync(0xffffffffu, give, off);
        }
    }
    // bit 4 of the lane id chose the top half first, bit 0 the last: the lane owns dims [lane * HS/32, (lane + 1) * HS/32)
    __nv_bfloat16 *dst = att_c + (size_t)mtc * C * 128;
    const int base = lane * (HS / 32);
#pragma unroll
    for (int k = 0; k < HS / 32; k++) {
        const int col = head * HS + base + k;
        dst[((size_t)(col >> 3) * 128 + rc) * 8 + (col & 7)] = __float2bfloat16(o[k] * inv);
    }
}

// head on the compact residual (row = sequence): ln_f + 5 logits
__global__ void __launch_bounds__(128) head_compact_kernel(const float *__restrict__ Xc, const float *__restrict__ gain,
                                                           const float *__restrict__ wte, float *__restrict__ logits, int C,
                                                           int n_seq)
{
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int seq = blockIdx.x * 4 + warp;
    if (seq >= n_seq) return;
    const float4 *Xi = reinterpret_cast<const float4 *>(Xc) + (size_t)(seq >> 7) * (C / 4) * 128 + (seq & 127);
    float s = 0.f;
    for (int c4 = lane; c4 < C / 4; c4 += 32) {
        const float4 v = Xi[(size_t)c4 * 128];
        s += (v.x + v.y) + (v.z + v.w);
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    const float mean = s / (float)C;
    float q = 0.f;
    for (int c4 = lane; c4 < C / 4; c4 += 32) {
        const float4 v = Xi[(size_t)c4 * 128];
        const float a = v.x - mean, b = v.y - mean, c = v.z - mean, d = v.w - mean;
        q += (a * a + b * b) + (c * c + d * d);
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
    const float rstd = rsqrtf(q / (float)C + 1e-5f);
    float acc[5] = {0.f, 0.f, 0.f, 0.f, 0.f};
    for (int c4 = lane; c4 < C / 4; c4 += 32) {
        const float4 v = Xi[(size_t)c4 * 128];
        const float4 g = __ldg(reinterpret_cast<const float4 *>(gain) + c4);
        const float y0 = (v.x - mean) * rstd * g.x, y1 = (v.y - mean) * rstd * g.y;
        const float y2 = (v.z - mean) * rstd * g.z, y3 = (v.w - mean) * rstd * g.w;
#pragma unroll
        for (int k = 0; k < 5; k++) {
            const float4 w = __ldg(reinterpret_cast<const float4 *>(wte + (size_t)k * C) + c4);
            acc[k] += y0 * w.x + y1 * w.y + y2 * w.z + y3 * w.w;
        }
    }
#pragma unroll
    for (int k = 0; k < 5; k++)
#pragma unroll
        for (int o = 16; o; o >>= 1) acc[k] += __shfl_xor_sync(0xffffffffu, acc[k], o);
    if (lane < 5) {
        float v = acc[0];
        if (lane == 1) v = acc[1];
        if (lane == 2) v = acc[2];
        if (lane == 3) v = acc[3];
        if (lane == 4) v = acc[4];
        logits[(size_t)seq * 8 + lane] = v;
    }
}

}  // namespace mg
