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
    const float *ln2_gain;         // [C]
    const float *next_gain;        // [C] ln_1 gain of the next block, or nullptr
    __nv_bfloat16 *xn_out;         // A_ti for the next block's QKV GEMM, or nullptr
};

// erf-GELU on a PAIR of values with packed fp32x2 math (FFMA2): erf(z) ~ z * P(z^2), odd degree-17 polynomial on
// |z| <= 3 (z clamped by one saturating FFMA per element), FMA-only, no MUFU.  Max |gelu error| 5e-5 over all x;
// the result is rounded to bf16 (rel. 4e-3) right after, see DESIGN.md "Tolerance".
__device__ __forceinline__ f32x2 gelu2(float x0, float x1)
{
    const float w0 = __saturatef(fmaf(x0, 0.70710678118654752440f / 6.0f, 0.5f));
    const float w1 = __saturatef(fmaf(x1, 0.70710678118654752440f / 6.0f, 0.5f));
    const f32x2 z = fma2(pk2(w0, w1), pk2(6.0f, 6.0f), pk2(-3.0f, -3.0f));
    const f32x2 u = mul2(z, z);
    f32x2 p = pk2(3.9138299712249136e-08f, 3.9138299712249136e-08f);
    p = fma2(p, u, pk2(-1.8835556829799316e-06f, -1.8835556829799316e-06f));
    p = fma2(p, u, pk2(4.0097045712172985e-05f, 4.0097045712172985e-05f));
    p = fma2(p, u, pk2(-0.0005030000465922058f, -0.0005030000465922058f));
    p = fma2(p, u, pk2(0.004197265952825546f, 0.004197265952825546f));
    p = fma2(p, u, pk2(-0.02500014565885067f, -0.02500014565885067f));
    p = fma2(p, u, pk2(0.11093290150165558f, 0.11093290150165558f));
    p = fma2(p, u, pk2(-0.3752213716506958f, -0.3752213716506958f));
    p = fma2(p, u, pk2(1.128251075744629f, 1.128251075744629f));
    const f32x2 hx = mul2(pk2(x0, x1), pk2(0.5f, 0.5f));
    return fma2(hx, mul2(z, p), hx);
}

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

template <int C>
struct PostAttnCfg {
    static constexpr int HC = C / 2;                 // hidden chunk (FC N, proj2 K per chunk)
    static constexpr int NCH = 4 * C / HC;           // 8 chunks
    static constexpr int STAGE_BYTES = 32 * C;       // [2 kc][C][16B] == [4 kc][HC][16B]
    static constexpr int NPROJ = C / 16;             // stages of the proj GEMM (one k-step each)
    static constexpr int NFC = C / 32;               // stages per FC chunk (two k-steps each)
    static constexpr int NP2 = HC / 16;              // stages per proj2 chunk (one k-step each)
    static constexpr int TOTAL_STAGES = NPROJ + NCH * (NFC + NP2);
    static constexpr int A_BYTES = C * 256;          // [C/8][128][16B]
    static constexpr int H_BYTES = HC * 256;         // [HC/8][128][16B]
    static constexpr int STAGES = C <= 160 ? 5 : 8;
    static constexpr int CTAS_PER_SM = C <= 160 ? 2 : 1;
    static constexpr uint32_t TMEM_COLS = (C + HC) <= 256 ? 256 : 512;
    static constexpr int SMEM_BYTES = A_BYTES + 2 * H_BYTES + STAGES * STAGE_BYTES + 4 * 128 * 4 + (2 * STAGES + 12) * 8 + 16;
    static_assert(C % 32 == 0 && C <= 256, "post_attn_kernel: C must be a multiple of 32, <= 256");
};

template <int C>
__global__ void __launch_bounds__(320, PostAttnCfg<C>::CTAS_PER_SM) post_attn_kernel(const PostAttnArgs a)
{
    using K = PostAttnCfg<C>;
    constexpr int HC = K::HC, S = K::STAGES;
    extern __shared__ __align__(1024) uint8_t smem[];
    uint8_t *As = smem;
    uint8_t *Hs = As + K::A_BYTES;                       // 2 buffers
    uint8_t *ring = Hs + 2 * K::H_BYTES;
    float *red = reinterpret_cast<float *>(ring + S * K::STAGE_BYTES);   // [2 kinds][2 halves][128]
    uint64_t *full = reinterpret_cast<uint64_t *>(red + 4 * 128);
    uint64_t *empty = full + S;
    uint64_t *bar_att = empty + S;       // A tile landed (tx)
    uint64_t *bar_proj = bar_att + 1;    // proj UMMAs retired
    uint64_t *bar_ln2 = bar_proj + 1;    // LN2(x1) in smem, x1 in TMEM (256 arrivals)
    uint64_t *bar_a1f = bar_ln2 + 1;     // FC chunk accumulated
    uint64_t *bar_a1e = bar_a1f + 1;     // FC chunk drained to registers (256 arrivals)
    uint64_t *bar_hf = bar_a1e + 1;      // [2] hidden chunk written to smem (256 arrivals)
    uint64_t *bar_he = bar_hf + 2;       // [2] hidden chunk consumed by proj2 UMMAs
    uint64_t *bar_done = bar_he + 2;     // all UMMAs retired
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bar_done + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int mt = blockIdx.x;

    if (threadIdx.x == 0) {
        for (int s = 0; s < S; s++) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], 1);
        }
        mbar_init(bar_att, 1);
        mbar_init(bar_proj, 1);
        mbar_init(bar_ln2, 256);
        mbar_init(bar_a1f, 1);
        mbar_init(bar_a1e, 256);
        mbar_init(&bar_hf[0], 256);
        mbar_init(&bar_hf[1], 256);
        mbar_init(&bar_he[0], 1);
        mbar_init(&bar_he[1], 1);
        mbar_init(bar_done, 1);
        fence_barrier_init();
    }
    if (warp == 9) tmem_alloc<K::TMEM_COLS>(tmem_slot);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;

    if (warp == 8) {
        // ------------------------------------------------------------------ producer
        if (lane == 0) {
            mbar_expect_tx(bar_att, K::A_BYTES);
            bulk_g2s(As, a.att + (size_t)mt * C * 128, K::A_BYTES, bar_att);
            bulk_prefetch_l2(a.x + (size_t)mt * C * 128, C * 128 * 4);
            const uint8_t *src = reinterpret_cast<const uint8_t *>(a.wstream);
            for (int i = 0; i < K::TOTAL_STAGES; i++) {
                const int s = i % S;
                mbar_wait(&empty[s], ((i / S) & 1) ^ 1);
                mbar_expect_tx(&full[s], K::STAGE_BYTES);
                bulk_g2s(ring + s * K::STAGE_BYTES, src + (size_t)i * K::STAGE_BYTES, K::STAGE_BYTES, &full[s]);
            }
        }
    } else if (warp == 9) {
        // ------------------------------------------------------------------ UMMA issuer
        if (lane == 0) {
            constexpr uint32_t idescC = umma_idesc_bf16(128, C, 0, 0);
            constexpr uint32_t idescH = umma_idesc_bf16(128, HC, 0, 0);
            const uint32_t a_addr = smem_u32(As), h_addr = smem_u32(Hs), r_addr = smem_u32(ring);
            int i = 0;  // stage cursor
            auto stage_wait = [&](int idx) -> uint32_t {
                const int s = idx % S;
                mbar_wait(&full[s], (idx / S) & 1);
                tc_fence_after();
                return r_addr + s * K::STAGE_BYTES;
            };
            // proj: acc_main = att @ Wproj^T
            mbar_wait(bar_att, 0);
            tc_fence_after();
            for (int ks = 0; ks < K::NPROJ; ks++, i++) {
                const uint32_t b = stage_wait(i);
                umma_ss(tmem, umma_desc(a_addr + ks * 4096, 2048, 128), umma_desc(b, C * 16, 128), idescC, ks != 0);
                umma_commit(&empty[i % S]);
            }
            umma_commit(bar_proj);
            mbar_wait(bar_ln2, 0);
            tc_fence_after();
            auto fc = [&](int j) {
                if (j > 0) {
                    mbar_wait(bar_a1e, (j - 1) & 1);
                    tc_fence_after();
                }
                for (int kb = 0; kb < K::NFC; kb++, i++) {
                    const uint32_t b = stage_wait(i);
#pragma unroll
                    for (int ks = 0; ks < 2; ks++)
                        umma_ss(tmem + C, umma_desc(a_addr + (kb * 2 + ks) * 4096, 2048, 128),
                                umma_desc(b + ks * 2 * (HC * 16), HC * 16, 128), idescH, (kb | ks) != 0);
                    umma_commit(&empty[i % S]);
                }
                umma_commit(bar_a1f);
            };
            auto p2 = [&](int j) {
                const int hb = j & 1;
                mbar_wait(&bar_hf[hb], (j >> 1) & 1);
                tc_fence_after();
                for (int ks = 0; ks < K::NP2; ks++, i++) {
                    const uint32_t b = stage_wait(i);
                    umma_ss(tmem, umma_desc(h_addr + hb * K::H_BYTES + ks * 4096, 2048, 128), umma_desc(b, C * 16, 128),
                            idescC, 1u);
                    umma_commit(&empty[i % S]);
                }
                umma_commit(&bar_he[hb]);
            };
            fc(0);
            for (int j = 0; j < K::NCH; j++) {
                if (j + 1 < K::NCH) fc(j + 1);
                p2(j);
            }
            umma_commit(bar_done);
        }
    } else {
        // ------------------------------------------------------------------ workers (256 threads)
        const int q = warp & 3, h = warp >> 2;
        const int r = q * 32 + lane;
        const uint32_t trow = tmem + ((uint32_t)(q * 32) << 16);
        constexpr int HALF = C / 2;                       // columns of the residual handled by this thread
        float *red_s = red, *red_q = red + 256;
        const float inv_c = 1.0f / (float)C;

        // ---- epilogue 1: x1 = x + proj, LN2 -> A tile, x1 -> TMEM
        mbar_wait(bar_proj, 0);
        tc_fence_after();
        {
            const float4 *Xg = reinterpret_cast<const float4 *>(a.x) + (size_t)mt * (C / 4) * 128 + r;
            float sum = 0.f;
            f32x2 sum2 = pk2(0.f, 0.f);
#pragma unroll 1
            for (int c0 = h * HALF; c0 < (h + 1) * HALF; c0 += 16) {
                uint32_t v[16];
                tmem_ld16(trow + c0, v);
                float4 xv[4];
#pragma unroll
                for (int j = 0; j < 4; j++) xv[j] = Xg[(size_t)(c0 / 4 + j) * 128];
                tmem_wait_ld();
#pragma unroll
                for (int j = 0; j < 4; j++) {
                    const f32x2 e0 = add2(pk2u(v[4 * j + 0], v[4 * j + 1]), pk2(xv[j].x, xv[j].y));
                    const f32x2 e1 = add2(pk2u(v[4 * j + 2], v[4 * j + 3]), pk2(xv[j].z, xv[j].w));
                    upk2u(e0, v[4 * j + 0], v[4 * j + 1]);
                    upk2u(e1, v[4 * j + 2], v[4 * j + 3]);
                    sum2 = add2(sum2, add2(e0, e1));
                }
                tmem_st16(trow + c0, v);
            }
            tmem_wait_st();
            {
                float s0, s1;
                upk2(sum2, s0, s1);
                sum = s0 + s1;
            }
            red_s[h * 128 + r] = sum;
            named_bar_sync(1, 256);
            const float mean = (red_s[r] + red_s[128 + r]) * inv_c;
            f32x2 sq2 = pk2(0.f, 0.f);
            const f32x2 negmean = pk2(-mean, -mean);
#pragma unroll 1
            for (int c0 = h * HALF; c0 < (h + 1) * HALF; c0 += 16) {
                uint32_t v[16];
                tmem_ld16(trow + c0, v);
                tmem_wait_ld();
                sq2 = sqdev16(v, negmean, sq2);
            }
            {
                float s0, s1;
                upk2(sq2, s0, s1);
                red_q[h * 128 + r] = s0 + s1;
            }
            named_bar_sync(1, 256);
            const float rstd = rsqrtf((red_q[r] + red_q[128 + r]) * inv_c + 1e-5f);
            const f32x2 la = pk2(rstd, rstd), lb = pk2(-mean * rstd, -mean * rstd);
            const float4 *g4 = reinterpret_cast<const float4 *>(a.ln2_gain);
#pragma unroll 1
            for (int c0 = h * HALF; c0 < (h + 1) * HALF; c0 += 16) {
                uint32_t v[16];
                tmem_ld16(trow + c0, v);
                tmem_wait_ld();
#pragma unroll
                for (int j = 0; j < 2; j++)
                    *reinterpret_cast<uint4 *>(As + ((c0 / 8 + j) * 128 + r) * 16) =
                        ln_pack8(&v[8 * j], la, lb, __ldg(g4 + c0 / 4 + 2 * j), __ldg(g4 + c0 / 4 + 2 * j + 1));
            }
            tc_fence_before();
            fence_proxy_async_smem();
            mbar_arrive(bar_ln2);
        }

        // ---- MLP chunks: acc1 -> GELU -> hidden chunk in smem
        constexpr int HH = HC / 2;            // hidden columns per thread per chunk
        constexpr int NV = HH / 8;            // 16-byte groups
#pragma unroll 1
        for (int j = 0; j < K::NCH; j++) {
            const int hb = j & 1;
            mbar_wait(bar_a1f, j & 1);
            tc_fence_after();
            uint32_t v[NV][8];
#pragma unroll
            for (int g = 0; g < NV; g++) tmem_ld8(trow + C + h * HH + g * 8, v[g]);
            tmem_wait_ld();
            tc_fence_before();
            mbar_arrive(bar_a1e);
            if (j >= 2) mbar_wait(&bar_he[hb], ((j >> 1) - 1) & 1);
            uint8_t *Hb = Hs + hb * K::H_BYTES;
#pragma unroll
            for (int g = 0; g < NV; g++) {
                uint4 o;
                o.x = pack_bf16x2_p(gelu2(__uint_as_float(v[g][0]), __uint_as_float(v[g][1])));
                o.y = pack_bf16x2_p(gelu2(__uint_as_float(v[g][2]), __uint_as_float(v[g][3])));
                o.z = pack_bf16x2_p(gelu2(__uint_as_float(v[g][4]), __uint_as_float(v[g][5])));
                o.w = pack_bf16x2_p(gelu2(__uint_as_float(v[g][6]), __uint_as_float(v[g][7])));
                *reinterpret_cast<uint4 *>(Hb + ((h * NV + g) * 128 + r) * 16) = o;
            }
            fence_proxy_async_smem();
            mbar_arrive(&bar_hf[hb]);
        }

        // ---- final epilogue: x' -> HBM, xn = LN1_next(x') -> HBM
        mbar_wait(bar_done, 0);
        tc_fence_after();
        {
            float4 *Xg = reinterpret_cast<float4 *>(a.x) + (size_t)mt * (C / 4) * 128 + r;
            float sum = 0.f;
#pragma unroll 1
            for (int c0 = h * HALF; c0 < (h + 1) * HALF; c0 += 16) {
                uint32_t v[16];
                tmem_ld16(trow + c0, v);
                tmem_wait_ld();
#pragma unroll
                for (int j = 0; j < 4; j++) {
                    const float4 o = make_float4(__uint_as_float(v[4 * j + 0]), __uint_as_float(v[4 * j + 1]),
                                                 __uint_as_float(v[4 * j + 2]), __uint_as_float(v[4 * j + 3]));
                    Xg[(size_t)(c0 / 4 + j) * 128] = o;
                    sum += (o.x + o.y) + (o.z + o.w);
                }
            }
            if (a.xn_out != nullptr) {
                red_s[h * 128 + r] = sum;
                named_bar_sync(1, 256);
                const float mean = (red_s[r] + red_s[128 + r]) * inv_c;
                f32x2 sq2 = pk2(0.f, 0.f);
                const f32x2 negmean = pk2(-mean, -mean);
#pragma unroll 1
                for (int c0 = h * HALF; c0 < (h + 1) * HALF; c0 += 16) {
                    uint32_t v[16];
                    tmem_ld16(trow + c0, v);
                    tmem_wait_ld();
                    sq2 = sqdev16(v, negmean, sq2);
                }
                {
                    float s0, s1;
                    upk2(sq2, s0, s1);
                    red_q[h * 128 + r] = s0 + s1;
                }
                named_bar_sync(1, 256);
                const float rstd = rsqrtf((red_q[r] + red_q[128 + r]) * inv_c + 1e-5f);
                const f32x2 la = pk2(rstd, rstd), lb = pk2(-mean * rstd, -mean * rstd);
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
    tc_fence_before();
    __syncthreads();
    if (warp == 9) tmem_dealloc<K::TMEM_COLS>(tmem);
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
    const int tok = tokens[(size_t)mt * 128 + r];
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

}  // namespace mg
