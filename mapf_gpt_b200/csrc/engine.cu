// engine.cu -- host side of libmapf_gpt_b200.so: device state, launch orchestration, C ABI.
// See include/mapf_gpt_b200.h for the contract and the reference interfaces each entry replaces.
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <cstdlib>
#include <string>
#include <vector>

#include "../../include/mapf_gpt_b200.h"
#include "env_kernels.cuh"
#include "gpt_kernels.cuh"
#include "fused_kernels.cuh"
#include "precise_kernels.cuh"

using namespace mg;

// ------------------------------------------------------------------------------------------- errors
static thread_local std::string g_err;
static int fail(int code, const char *fmt, ...)
{
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    g_err = buf;
    return code;
}
#define CU(call)                                                                                           \
    do {                                                                                                   \
        cudaError_t _e = (call);                                                                           \
        if (_e != cudaSuccess)                                                                             \
            return fail(MG_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(_e), __FILE__, __LINE__); \
    } while (0)


// cudaFuncSetAttribute / occupancy results are PER DEVICE: a process may drive several GPUs (MAPFGPTInference on cuda:0 and
// cuda:1), so "done once" flags are kept per device index.  The attribute is set BEFORE the flag, so a racing thread at worst
// sets it twice.
#include <atomic>
static const int MG_MAX_DEVICES = 64;
struct PerDevice {
    std::atomic<int> v[MG_MAX_DEVICES];
    PerDevice() { for (auto &x : v) x.store(0); }
};
static int cur_device()
{
    int d = 0;
    cudaGetDevice(&d);
    return (d < 0 || d >= MG_MAX_DEVICES) ? 0 : d;
}

// ------------------------------------------------------------------------------------------- model
enum KernelClass { KC_BFS = 0, KC_OBSERVE, KC_EMBED, KC_LN, KC_QKV, KC_ATTN, KC_PROJ, KC_FC, KC_PROJ2, KC_HEAD, KC_STEP, KC_POST, KC_ATTN_LAST, KC_POST_LAST, KC_COUNT };

struct Layer {
    float *ln1 = nullptr, *ln2 = nullptr;
    __nv_bfloat16 *wqkv = nullptr, *wproj = nullptr, *wfc = nullptr, *wproj2 = nullptr;
    __nv_bfloat16 *wqkv_p = nullptr, *wproj_p = nullptr, *wfc_p = nullptr, *wproj2_p = nullptr;   // CTA-pair packing (generic path, BN = 256)
    float *cs_qkv = nullptr, *cs_fc = nullptr;   // LayerNorm folded into the GEMMs: column sums of the gain-folded bf16 weights
    __nv_bfloat16 *wstream = nullptr;   // fused post-attention kernel: stage images in consumption order
    __nv_bfloat16 *wstream_pair = nullptr;   // the same stream for CTA pairs: every stage split into two N/2-row halves
};
struct Model {
    bool loaded = false;
    mg_model_config cfg{};
    int BN = 0, BK = 0, hs = 0;
    bool fused = false;   // C in {160, 256}: post_attn_kernel replaces proj / ln_2 / fc / proj2 / next ln_1
    bool fuse_qkv = false;  // ... and the next block's c_attn
    float q_fold = 1.f;     // log2(e) / sqrt(hs), multiplied into the q rows of every c_attn weight at load time
    bool ln_fused = false;  // generic path on CTA-pair GEMMs: ln_1 / ln_2 have no pass of their own (GemmArgs: xb_out / stats / colsum)
    float *wte = nullptr, *wpe = nullptr, *lnf = nullptr;
    float *wpe_ti = nullptr;   // wpe re-tiled [2][C/4][128][4] for embed_ln_kernel
    uint4 *tab0 = nullptr;     // block 0 as a lookup: [67 tokens][256 positions][C/4 + 3C/8] 16-byte groups (x, then q|k|v)
    std::vector<Layer> layers;
};
struct Workspace {
    int chunk_seqs = 0;
    float *X = nullptr;
    __nv_bfloat16 *XN = nullptr, *QKV = nullptr, *ATT = nullptr, *HID = nullptr;
    float *STATS = nullptr;           // LayerNorm folded into the GEMMs: [M/128][C/128][2][128] partial row sums / sums of squares
    float *X24 = nullptr;             // generic path: the residual in its 24-bit layout between the residual GEMMs (same tile pitch)
    float *Xc = nullptr;              // compact residual rows of token 255 (last-block pruning)
    __nv_bfloat16 *ATTc = nullptr;    // compact attention output of token 255
    uint8_t *tok = nullptr;   // staging for forward_tokens
    float *logits = nullptr;
};

struct EvPair { cudaEvent_t a, b; int kc; };
#ifndef MG_LANES_DEFAULT
#define MG_LANES_DEFAULT(C) 1
#endif

// Stream lanes (fused path, C = 160): the forward chunks of one timestep alternate between two lanes, each with its own
// workspace and two streams -- attention on a HIGH-priority stream, everything else on a low-priority one -- so that the
// MUFU-bound attention CTAs of one chunk (grid = one per SM) share the SMs with the tensor/FMA-bound post_attn CTAs of the
// other chunk instead of running back to back (one CTA of each fits per SM: 2 x 115 KB of shared memory, 2 x 256 TMEM columns).
struct Lane {
    Workspace ws;
    cudaStream_t lo = nullptr, hi = nullptr;
    int *ctr = nullptr;               // work counter of this lane's persistent attention launches
    cudaEvent_t ev_attn = nullptr, ev_post = nullptr;
};

struct mg_engine {
    int device = 0;
    cudaStream_t stream = nullptr;
    mg_params params{};
    EnvState s{};
    int n_envs = 0;  // highest reset slot + 1
    Model model;
    Workspace ws;
    // staging (device) + pinned host
    int32_t *d_pos_in = nullptr, *d_goal_in = nullptr, *d_act_in = nullptr, *d_step_act = nullptr;
    float *d_q = nullptr;
    double *d_metrics = nullptr;
    uint64_t seed = 0;
    int env_offset = 0;
    int max_episode_steps = 0;
    long long launches = 0;
    bool profiling = false;
    bool prune_last = true;   // last-block pruning (MAPF_GPT_B200_NO_PRUNE=1 disables it for A/B tests)
    std::vector<EvPair> evs;
    size_t ev_used = 0;
    cudaEvent_t ev_t0 = nullptr, ev_t1 = nullptr, ev_p[4] = {nullptr, nullptr, nullptr, nullptr};
    float last_total_ms = 0.f, last_phase_ms[3] = {0.f, 0.f, 0.f};
    float kc_ms[KC_COUNT] = {0};
    long long *d_timeline = nullptr;   // test hook (mg_test_timeline)
    int *d_attn_ctr = nullptr;         // work counter of the persistent attention kernel (self-resetting)
    Lane lanes[2];                     // stream lanes (see Lane); lane workspaces are allocated on first use
    bool lanes_ready = false;
    int n_lanes = 1;                   // MAPF_GPT_B200_LANES (default: 2 for the C = 160 fused path, else 1)
    int lane_attn_grid = 0;            // attention CTAs per launch in lane mode (MAPF_GPT_B200_LANE_ATTN_GRID, default n_sms)
    int **attn_ctr_slot = nullptr;     // which counter launch_attn_persistent uses (a lane's or d_attn_ctr)
    int attn_grid_cap = 0;             // > 0: grid of the persistent attention kernel (lane mode)
    cudaEvent_t ev_fork = nullptr;
    CUtensorMap map_c2g{}, map_loc{};  // TMA descriptors of the FOV-window fields (observe_tma_kernel)
    std::vector<std::vector<uint8_t>> map_grids;   // large maps: host copies of the distinct grids (precompute tables per map)
    std::vector<uint16_t *> map_pre;               // device K x K tables
    bool use_tma = false;
    int n_sms = 148;
    long long kc_n[KC_COUNT] = {0};
    std::vector<int32_t> h_nag;        // host mirror of s.nag (range checks of host-supplied positions / goals)
    // attention softmax: the max-free kernels (scores pre-scaled into the log2 domain by the folded Wq) run until a step
    // produces a non-finite logit; that step is redone with the max-subtracting kernels, which then stay on
    bool safe_softmax = false;
    int precise = 0;                   // MAPF_GPT_B200_PRECISION=fp32: fp32 CUDA-core verification forward (precise_kernels.cuh)
    float *pw = nullptr;               // precise path: the raw fp32 checkpoint buffer on the device
    float *p_ws = nullptr;             // precise path: workspace
    size_t p_ws_floats = 0;
};

static void prof_begin(mg_engine *e, int kc)
{
    e->launches++;
    if (!e->profiling) return;
    if (e->ev_used == e->evs.size()) {
        EvPair p;
        cudaEventCreate(&p.a);
        cudaEventCreate(&p.b);
        e->evs.push_back(p);
    }
    e->evs[e->ev_used].kc = kc;
    cudaEventRecord(e->evs[e->ev_used].a, e->stream);
}
static void prof_end(mg_engine *e)
{
    if (!e->profiling) return;
    cudaEventRecord(e->evs[e->ev_used].b, e->stream);
    e->ev_used++;
}
static void prof_collect(mg_engine *e)
{   // caller has synchronized the stream
    for (size_t i = 0; i < e->ev_used; i++) {
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, e->evs[i].a, e->evs[i].b) == cudaSuccess) {
            e->kc_ms[e->evs[i].kc] += ms;
            e->kc_n[e->evs[i].kc]++;
        }
    }
    e->ev_used = 0;
}

// ------------------------------------------------------------------------------------------- helpers
template <typename T>
static cudaError_t dalloc(T **p, size_t n) { return cudaMalloc(reinterpret_cast<void **>(p), n * sizeof(T)); }

static uint16_t f2bf(float f)
{   // round-to-nearest-even, like __float2bfloat16_rn
    uint32_t u;
    memcpy(&u, &f, 4);
    if ((u & 0x7FFFFFFFu) > 0x7F800000u) return (uint16_t)((u >> 16) | 0x40);
    u += 0x7FFFu + ((u >> 16) & 1u);
    return (uint16_t)(u >> 16);
}

// W[N][K] fp32 (torch Linear weight) -> bf16 tile image [N/BN][K/8][BN][8]
static int upload_packed(const float *W, int N, int K, int BN, __nv_bfloat16 **out)
{
    std::vector<uint16_t> h((size_t)N * K);
    for (int n = 0; n < N; n++)
        for (int k = 0; k < K; k++) {
            const int nt = n / BN, nn = n % BN;
            h[(((size_t)nt * (K / 8) + k / 8) * BN + nn) * 8 + (k & 7)] = f2bf(W[(size_t)n * K + k]);
        }
    CU(dalloc(out, (size_t)N * K));
    CU(cudaMemcpy(*out, h.data(), h.size() * 2, cudaMemcpyHostToDevice));
    return MG_OK;
}
static int upload_f32(const float *src, size_t n, float **out)
{
    CU(dalloc(out, n));
    CU(cudaMemcpy(*out, src, n * 4, cudaMemcpyHostToDevice));
    return MG_OK;
}

// the same for gemm_pair_persistent_kernel: [N/BN][2 halves][K/8][BN/2][8], CTA r of a pair streams rows [r * BN/2, (r+1) * BN/2) of the tile
static int upload_packed_pair(const float *W, int N, int K, int BN, __nv_bfloat16 **out)
{
    std::vector<uint16_t> h((size_t)N * K);
    const int HB = BN / 2;
    for (int n = 0; n < N; n++)
        for (int k = 0; k < K; k++) {
            const int nt = n / BN, nn = n % BN, half = nn / HB, hn = nn % HB;
            h[((((size_t)nt * 2 + half) * (K / 8) + k / 8) * HB + hn) * 8 + (k & 7)] = f2bf(W[(size_t)n * K + k]);
        }
    CU(dalloc(out, (size_t)N * K));
    CU(cudaMemcpy(*out, h.data(), h.size() * 2, cudaMemcpyHostToDevice));
    return MG_OK;
}

// Stage stream of post_attn_kernel<C> (fused_kernels.cuh): proj k-steps, then FC(0), FC(1), P2(0), FC(2), ...
// Wproj[C][C], Wfc[4C][C], Wproj2[C][4C] are torch Linear weights (row = output feature).
// split = 2 (CTA pairs, post_attn_kernel<.., CL = 2>): a stage of U units is stored as [half][unit][kc][rows/2][8], so that
// CTA r of a pair copies one contiguous half-stage holding rows [r * rows/2, (r+1) * rows/2) of every unit.
// LayerNorm gains are folded into the weights that consume the normalised activations (the kernel then writes plain
// (x - mean) * rstd): ln_2's gain g2[k] scales column k of Wfc, the next block's ln_1 gain gn[k] scales column k of Wqkv_next.
// qscale: the attention scale folded into the q rows (rows < C) of the next block's c_attn (Model::q_fold).
static int upload_post_attn_stream(const float *Wproj, const float *Wfc, const float *Wproj2, const float *Wqkv_next,
                                   const float *g2, const float *gn, int C, int split, float qscale, __nv_bfloat16 **out)
{
    // PostAttnCfg<C>::WIDE (C = 256): the MLP runs in 4 chunks of C hidden columns (FC units = one k-step of all C rows, like
    // proj) instead of 8 chunks of C/2; the c_attn tail keeps the narrow half-n-tile format either way
    const bool wide = PostAttnCfg<256, 1, 0, 1>::WIDE && C == 256;
    const int HC = C / 2, HM = wide ? C : HC, NCH = 4 * C / HM, NPROJ = C / 16, NFC = C / 32, NFCM = wide ? C / 16 : C / 32, NP2 = HM / 16;
    const int U = C == 160 ? 5 : 4;              // units per stage (PostAttnCfg::U)
    const size_t stage_elems = (size_t)16 * C;   // 32*C bytes per unit
    const size_t total = (size_t)(NPROJ + NCH * (NFCM + NP2) + (Wqkv_next ? 6 * NFC : 0)) * stage_elems;
    std::vector<uint16_t> h(total, 0);
    size_t st = 0;
    auto put_kn = [&](size_t base, int kc, int rows, int n, int k8, float v) {   // [kc][rows][8]; base = unit index * stage_elems
        if (split == 1) {
            h[base + ((size_t)kc * rows + n) * 8 + k8] = f2bf(v);
            return;
        }
        const size_t unit = base / stage_elems, rh = rows / 2;
        const size_t half = n / rh, nn = n % rh;
        h[(unit / U) * (U * stage_elems) + half * (U * stage_elems / 2) + (unit % U) * (stage_elems / 2) + ((size_t)kc * rh + nn) * 8 + k8] = f2bf(v);
    };
    for (int ks = 0; ks < NPROJ; ks++, st++)
        for (int n = 0; n < C; n++)
            for (int k = 0; k < 16; k++) put_kn(st * stage_elems, k / 8, C, n, k % 8, Wproj[(size_t)n * C + ks * 16 + k]);
    auto put_fc = [&](int j) {
        if (wide) {   // one k-step (16 k) of the chunk's C hidden rows per unit
            for (int ks = 0; ks < NFCM; ks++, st++)
                for (int n = 0; n < C; n++)
                    for (int k = 0; k < 16; k++)
                        put_kn(st * stage_elems, k / 8, C, n, k % 8, Wfc[(size_t)(j * C + n) * C + ks * 16 + k] * g2[ks * 16 + k]);
            return;
        }
        for (int kb = 0; kb < NFC; kb++, st++)
            for (int n = 0; n < HC; n++)
                for (int k = 0; k < 32; k++)
                    put_kn(st * stage_elems, k / 8, HC, n, k % 8, Wfc[(size_t)(j * HC + n) * C + kb * 32 + k] * g2[kb * 32 + k]);
    };
    auto put_p2 = [&](int j) {
        for (int ks = 0; ks < NP2; ks++, st++)
            for (int n = 0; n < C; n++)
                for (int k = 0; k < 16; k++)
                    put_kn(st * stage_elems, k / 8, C, n, k % 8, Wproj2[(size_t)n * 4 * C + j * HM + ks * 16 + k]);
    };
    put_fc(0);
    for (int j = 0; j < NCH; j++) {
        if (j + 1 < NCH) put_fc(j + 1);
        put_p2(j);
    }
    if (Wqkv_next)   // next block's c_attn [3C][C]: six half n-tiles of HC rows in the FC-chunk stage format
        for (int hh = 0; hh < 6; hh++)
            for (int kb = 0; kb < NFC; kb++, st++)
                for (int n = 0; n < HC; n++)
                    for (int k = 0; k < 32; k++)
                        put_kn(st * stage_elems, k / 8, HC, n, k % 8,
                               Wqkv_next[(size_t)(hh * HC + n) * C + kb * 32 + k] * gn[kb * 32 + k] * (hh * HC + n < C ? qscale : 1.f));
    CU(dalloc(out, total));
    CU(cudaMemcpy(*out, h.data(), total * 2, cudaMemcpyHostToDevice));
    return MG_OK;
}

#ifndef MG_POST_PF_DEFAULT
#define MG_POST_PF_DEFAULT(slots) ((slots) / 4)
#endif
#ifndef MG_POST_CL_DEFAULT
#define MG_POST_CL_DEFAULT 2
#endif
template <int C, int NT, int UU = 0, int CL = 1, bool PERSIST = false>
static int launch_post_attn_c(mg_engine *e, PostAttnArgs a, int MT, int kc)
{
    using K = PostAttnCfg<C, NT, UU, CL>;
    static PerDevice attr_set;
    if (const int d = cur_device(); !attr_set.v[d].load()) {
        CU(cudaFuncSetAttribute(post_attn_kernel<C, NT, UU, CL, PERSIST>, cudaFuncAttributeMaxDynamicSharedMemorySize, K::SMEM_BYTES));
        attr_set.v[d].store(1);
    }
    a.n_groups = MT / NT;
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(a.n_groups);
    cfg.blockDim = dim3(K::THREADS);
    cfg.dynamicSmemBytes = K::SMEM_BYTES;
    cfg.stream = e->stream;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = CL;
    at[0].val.clusterDim.y = 1;
    at[0].val.clusterDim.z = 1;
    cfg.attrs = at;
    cfg.numAttrs = CL == 1 ? 0 : 1;
    if (CL == 2) a.wstream = a.wstream_pair;
    if (PERSIST) {
        // as many CTAs (CTA pairs) as the device keeps resident, each looping over tile groups g, g + grid, ...
        static const int grid_override = getenv("MAPF_GPT_B200_POST_GRID") ? atoi(getenv("MAPF_GPT_B200_POST_GRID")) : 0;
        static const int stagger_override = getenv("MAPF_GPT_B200_POST_STAGGER_NS") ? atoi(getenv("MAPF_GPT_B200_POST_STAGGER_NS")) : 0;
        int cap = grid_override > 0 ? grid_override : K::CTAS_PER_SM * e->n_sms;   // all resident CTA slots (a late CTA only costs tail time)
        cap -= cap % CL;
        if (a.n_groups > cap) {
            cfg.gridDim = dim3(cap);
            a.stagger_ns = stagger_override;   // spreading the cluster start times over a tile period measured neutral: off by default
        }
    }
    if (!PERSIST) {
        // successor prefetch distance (MAPF_GPT_B200_POST_PF overrides; 0 = off): a quarter of the device's CTA slots measured best
        // (32-96 of 296: post_attn<160> 1.80 -> 1.75 ms per launch; 148: 1.74-1.79; the full 296 and beyond: 1.82-1.84, slower than none)
        static const int pf_override = getenv("MAPF_GPT_B200_POST_PF") ? atoi(getenv("MAPF_GPT_B200_POST_PF")) : -1;
        a.pf_dist = pf_override >= 0 ? pf_override : MG_POST_PF_DEFAULT(K::CTAS_PER_SM * e->n_sms);
    }
    prof_begin(e, kc);
    CU((cudaLaunchKernelEx(&cfg, post_attn_kernel<C, NT, UU, CL, PERSIST>, a)));
    prof_end(e);
    CU(cudaGetLastError());
    return MG_OK;
}
static int launch_post_attn(mg_engine *e, int C, const PostAttnArgs &a, int MT, bool single_tiles)
{
    // NT = 1 (two CTAs per SM) measured faster than NT = 2 (one CTA per SM, shared weight stages): with one CTA per SM
    // the HBM phases (tile load / store) and the compute phase of an SM do not overlap.  MAPF_GPT_B200_POST_NT=2 selects it.
    // Default: CTA pairs (cta_group::2 UMMAs, each SM streams half of the weights) whenever the tile count is even; measured
    // +1 % on the whole step over single CTAs (less L2 traffic -> higher clocks under the power cap).  MAPF_GPT_B200_POST_CL=1
    // selects single CTAs.
    static const int nt_override = getenv("MAPF_GPT_B200_POST_NT") ? atoi(getenv("MAPF_GPT_B200_POST_NT")) : 0;
    static const int cl_req = getenv("MAPF_GPT_B200_POST_CL") ? atoi(getenv("MAPF_GPT_B200_POST_CL")) : MG_POST_CL_DEFAULT;
    const int kc = single_tiles ? KC_POST_LAST : KC_POST;
    static const int u_override = getenv("MAPF_GPT_B200_POST_U") ? atoi(getenv("MAPF_GPT_B200_POST_U")) : 0;
    const bool pair = cl_req == 2 && MT % 2 == 0 && a.wstream_pair != nullptr;
    if (C == 160 && u_override == 1) return launch_post_attn_c<160, 1, 1>(e, a, MT, kc);
    // Persistent CTAs (tile loop inside the kernel) for the layers that emit the next block's q/k/v: default where one CTA fits
    // per SM (C = 256: +6 % on the kernel), opt-in where two do (C = 160: -3 %).  MAPF_GPT_B200_POST_PERSIST=0/1 overrides.
    static const int persist_req = getenv("MAPF_GPT_B200_POST_PERSIST") ? atoi(getenv("MAPF_GPT_B200_POST_PERSIST")) : -1;
    const bool persist = a.qkv_out != nullptr && pair && (persist_req >= 0 ? persist_req == 1 : C == 256);
    if (C == 160) {
        if (nt_override == 2 && !single_tiles && MT % 2 == 0) return launch_post_attn_c<160, 2>(e, a, MT, kc);
        if (persist) return launch_post_attn_c<160, 1, 0, 2, true>(e, a, MT, kc);
        return pair ? launch_post_attn_c<160, 1, 0, 2>(e, a, MT, kc) : launch_post_attn_c<160, 1>(e, a, MT, kc);
    }
    if (C == 256) {
        if (persist) return launch_post_attn_c<256, 1, 0, 2, true>(e, a, MT, kc);
        return pair ? launch_post_attn_c<256, 1, 0, 2>(e, a, MT, kc) : launch_post_attn_c<256, 1>(e, a, MT, kc);
    }
    return fail(MG_ERR_ARG, "post_attn: unsupported width %d", C);
}

// ------------------------------------------------------------------------------------------- GEMM dispatch
template <int BN, int BK, int STAGES, int EPI>
static int launch_gemm_cfg(mg_engine *e, const GemmArgs &a, int kc)
{
    constexpr int smem = gemm_smem_bytes<BN, BK, STAGES>();
    static PerDevice attr_set;
    if (const int d = cur_device(); !attr_set.v[d].load()) {
        CU(cudaFuncSetAttribute(gemm_kernel<BN, BK, STAGES, EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        attr_set.v[d].store(1);
    }
    const int tiles = (a.M / 128) * (a.N / BN);
    if (e) prof_begin(e, kc);
    gemm_kernel<BN, BK, STAGES, EPI><<<tiles, 192, smem, e ? e->stream : 0>>>(a);
    if (e) prof_end(e);
    CU(cudaGetLastError());
    return MG_OK;
}
// persistent CTA-pair GEMM (gemm_pair_persistent_kernel): one pair per SM pair, two TMEM accumulators
template <int EPI>
static int launch_gemm_pair_persistent(mg_engine *e, const GemmArgs &a, int kc)
{
    constexpr int BN = 256, BK = 64, STAGES = 6;
    const int smem = gemm_pair_persistent_smem_bytes<BN, BK, STAGES>(a.N);
    constexpr int SMEM_MAX = 224 * 1024;
    if (smem > SMEM_MAX) return fail(MG_ERR_ARG, "gemm: N=%d too wide for the persistent pair kernel", a.N);
    static PerDevice max_clusters_dev;
    const int dev = cur_device();
    int max_clusters = max_clusters_dev.v[dev].load();
    cudaLaunchConfig_t cfg{};
    cfg.blockDim = dim3(320);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = e ? e->stream : 0;   // e == nullptr: test hook (mg_test_gemm)
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = 2;
    at[0].val.clusterDim.y = 1;
    at[0].val.clusterDim.z = 1;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    if (!max_clusters) {
        CU(cudaFuncSetAttribute(gemm_pair_persistent_kernel<BN, BK, STAGES, EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_MAX));
        cfg.gridDim = dim3(2 * 74);
        int n = 0;
        CU(cudaOccupancyMaxActiveClusters(&n, gemm_pair_persistent_kernel<BN, BK, STAGES, EPI>, &cfg));
        max_clusters = n > 0 ? n : 1;
        max_clusters_dev.v[dev].store(max_clusters);
    }
    const int n_tiles = (a.M / 256) * (a.N / BN);
    cfg.gridDim = dim3(2 * std::min(max_clusters, n_tiles));
    if (e) prof_begin(e, kc);
    CU(cudaLaunchKernelEx(&cfg, gemm_pair_persistent_kernel<BN, BK, STAGES, EPI>, a));
    if (e) prof_end(e);
    CU(cudaGetLastError());
    return MG_OK;
}
template <int EPI>
static int launch_gemm(mg_engine *e, int BN, const GemmArgs &a, int kc)
{
    if (a.M % 128) return fail(MG_ERR_ARG, "gemm: M=%d not a multiple of 128", a.M);
    // CTA pairs (gemm_pair_persistent_kernel) whenever the pair packing exists: 256-wide tiles, an even number of 128-row tiles.
    // MAPF_GPT_B200_GEMM_PAIR=0 selects the single-CTA kernel (A/B).
    const bool pair_off = getenv("MAPF_GPT_B200_GEMM_PAIR") && getenv("MAPF_GPT_B200_GEMM_PAIR")[0] == '0';   // per launch: tests flip it
    if (a.Wp && !pair_off && BN == 256 && a.N % 256 == 0 && a.K % 64 == 0 && (a.M / 128) % 2 == 0)
        return launch_gemm_pair_persistent<EPI>(e, a, kc);
    if (a.x_in_24 || a.x_out_24) return fail(MG_ERR_STATE, "gemm: the 24-bit residual layout needs the CTA-pair kernel");
    // (a single-stage K=160 variant <160,160,1> measured SLOWER, 0.68 vs 0.60 ms: no load/UMMA overlap inside the CTA)
    if (BN == 160 && a.N % 160 == 0 && a.K % 32 == 0) return launch_gemm_cfg<160, 32, 4, EPI>(e, a, kc);
    if (BN == 256 && a.N % 256 == 0 && a.K % 64 == 0) return launch_gemm_cfg<256, 64, 2, EPI>(e, a, kc);
    if (BN == 128 && a.N % 128 == 0 && a.K % 64 == 0) return launch_gemm_cfg<128, 64, 3, EPI>(e, a, kc);
    if constexpr (EPI == EPI_STORE_F32) {   // microbenchmark shapes (mg_test_gemm_time)
        if (BN == 80 && a.N % 80 == 0 && a.K % 64 == 0) return launch_gemm_cfg<80, 64, 4, EPI>(e, a, kc);
        if (BN == 240 && a.N % 240 == 0 && a.K % 64 == 0) return launch_gemm_cfg<240, 64, 3, EPI>(e, a, kc);
    }
    return fail(MG_ERR_ARG, "gemm: unsupported shape N=%d K=%d for BN=%d", a.N, a.K, BN);
}
static int pick_bn(int C)
{
    if (C % 256 == 0) return 256;
    if (C % 160 == 0) return 160;
    if (C % 128 == 0) return 128;
    return 0;
}

template <int HS>
static int launch_attn_hs(mg_engine *e, const AttnArgs &a, int n_seq, cudaStream_t st)
{
    constexpr int smem = attn_smem_bytes<HS>();
    static PerDevice attr_set;
    if (const int d = cur_device(); !attr_set.v[d].load()) {
        CU(cudaFuncSetAttribute(attn_kernel<HS>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        attr_set.v[d].store(1);
    }
    if (e) prof_begin(e, KC_ATTN);
    attn_kernel<HS><<<n_seq * a.n_head, 288, smem, st>>>(a);
    if (e) prof_end(e);
    CU(cudaGetLastError());
    return MG_OK;
}
template <bool FAST>
static int launch_attn_persistent(mg_engine *e, const AttnArgs &a, int n_seq, cudaStream_t st)
{
    constexpr int smem = attn_persistent_smem_bytes();
    static PerDevice attr_set, n_sms_dev;
    const int dev = cur_device();
    if (!attr_set.v[dev].load()) {
        CU(cudaFuncSetAttribute(attn_persistent_kernel<FAST>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        int n = 148;
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
        n_sms_dev.v[dev].store(n);
        attr_set.v[dev].store(1);
    }
    const int n_sms = n_sms_dev.v[dev].load();
    const int n_items = n_seq * a.n_head;
    AttnArgs aa = a;
    int *tmp_ctr = nullptr;
    if (e) {
        int **slot = e->attn_ctr_slot ? e->attn_ctr_slot : &e->d_attn_ctr;
        if (!*slot) {
            CU(dalloc(slot, 2));
            CU(cudaMemsetAsync(*slot, 0, 8, st));
        }
        aa.work_counter = *slot;
    } else {   // test hook without an engine: a counter of its own, freed after the (synchronous) call
        CU(dalloc(&tmp_ctr, 2));
        CU(cudaMemsetAsync(tmp_ctr, 0, 8, st));
        aa.work_counter = tmp_ctr;
    }
    if (e) prof_begin(e, KC_ATTN);
    const int grid = e && e->attn_grid_cap > 0 ? e->attn_grid_cap : 2 * n_sms;
    attn_persistent_kernel<FAST><<<std::min(n_items, grid), 320, smem, st>>>(aa, n_items);
    if (e) prof_end(e);
    CU(cudaGetLastError());
    if (tmp_ctr) {
        CU(cudaStreamSynchronize(st));
        cudaFree(tmp_ctr);
    }
    return MG_OK;
}
template <bool FAST>
static int launch_attn_ts64(mg_engine *e, const AttnArgs &a, int n_seq, cudaStream_t st)
{   // head size 64: probabilities in TMEM, two CTAs per SM (attn_ts_kernel)
    constexpr int smem = attn_ts_smem_bytes<64>();
    static PerDevice attr_set;
    if (const int d = cur_device(); !attr_set.v[d].load()) {
        CU(cudaFuncSetAttribute(attn_ts_kernel<64, FAST>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        attr_set.v[d].store(1);
    }
    if (e) prof_begin(e, KC_ATTN);
    attn_ts_kernel<64, FAST><<<n_seq * a.n_head, 288, smem, st>>>(a);
    if (e) prof_end(e);
    CU(cudaGetLastError());
    return MG_OK;
}
// `fast`: the scores arrive in the log2 domain (scale folded into Wq) and the max-free kernels may be used; otherwise the
// max-subtracting kernels run with a.scale_log2e (1.0 when the scale is folded).
static int launch_attn(mg_engine *e, const AttnArgs &a, int hs, int n_seq, cudaStream_t st, bool fast)
{
    const bool classic = getenv("MAPF_GPT_B200_ATTN_CLASSIC") != nullptr;   // read per launch: tests flip it between engines
    if (hs == 32 && !classic) return fast ? launch_attn_persistent<true>(e, a, n_seq, st) : launch_attn_persistent<false>(e, a, n_seq, st);
    if (hs == 32) return launch_attn_hs<32>(e, a, n_seq, st);
    if (hs == 64 && !classic) return fast ? launch_attn_ts64<true>(e, a, n_seq, st) : launch_attn_ts64<false>(e, a, n_seq, st);
    if (hs == 64) return launch_attn_hs<64>(e, a, n_seq, st);
    return fail(MG_ERR_ARG, "attention: head size %d unsupported (32 or 64)", hs);
}

// ------------------------------------------------------------------------------------------- forward
static int ensure_workspace(mg_engine *e, Workspace &w, int want_seqs)
{
    const int C = e->model.cfg.n_embd;
    // sequences per forward chunk (workspace = 22 * chunk * 256 * C bytes); MAPF_GPT_B200_CHUNK_SEQS overrides (multiple of 128)
    // (read when the workspace is first sized: an engine keeps its chunk size)
    const int chunk_max = getenv("MAPF_GPT_B200_CHUNK_SEQS") ? std::max(128, atoi(getenv("MAPF_GPT_B200_CHUNK_SEQS")) / 4 * 4) : 8192;
    int chunk = std::min(want_seqs, chunk_max);
    if (chunk <= w.chunk_seqs) return MG_OK;
    cudaFree(w.X); cudaFree(w.XN); cudaFree(w.QKV); cudaFree(w.ATT); cudaFree(w.HID); cudaFree(w.Xc); cudaFree(w.ATTc); cudaFree(w.STATS);
    cudaFree(w.X24);
    w.X = w.Xc = w.STATS = w.X24 = nullptr; w.XN = w.QKV = w.ATT = w.HID = w.ATTc = nullptr; w.chunk_seqs = 0;
    const size_t M = (size_t)chunk * 256;
    if (e->model.ln_fused) CU(dalloc(&w.STATS, M * 2 * (C / 128)));
    if (e->model.ln_fused) CU(dalloc(&w.X24, M * C));
    CU(dalloc(&w.X, M * C));
    CU(dalloc(&w.XN, M * C));
    CU(dalloc(&w.QKV, M * 3 * C));
    CU(dalloc(&w.ATT, M * C));
    CU(dalloc(&w.HID, M * 4 * C));
    const size_t Mc = ((size_t)chunk + 127) / 128 * 128;
    CU(dalloc(&w.Xc, Mc * C));
    CU(dalloc(&w.ATTc, Mc * C));
    CU(cudaMemsetAsync(w.Xc, 0, Mc * C * 4, e->stream));
    CU(cudaMemsetAsync(w.ATTc, 0, Mc * C * 2, e->stream));
    w.chunk_seqs = chunk;
    return MG_OK;
}

static void launch_ln(mg_engine *e, const float *X, const float *gain, __nv_bfloat16 *out, int C, int MT)
{
    if (C == 768) ln_rows_kernel<24><<<MT * 4, 256, 0, e->stream>>>(X, gain, out);
    else if (C == 512) ln_rows_kernel<16><<<MT * 4, 256, 0, e->stream>>>(X, gain, out);
    else if (C == 256) ln_rows_kernel<8><<<MT * 4, 256, 0, e->stream>>>(X, gain, out);
    else if (C == 128) ln_rows_kernel<4><<<MT * 4, 256, 0, e->stream>>>(X, gain, out);
    else ln_kernel<<<MT, 128, 0, e->stream>>>(X, gain, out, C);
}

// ------------------------------------------------------------------------------------------- fp32 verification forward
static int g_precision = 0;   // mg_set_precision
// tokens (device, uint8 [n_seq][256]) -> logits (device, fp32 [n_seq][8]) on the CUDA cores in fp32 (precise_kernels.cuh)
static int forward_precise(mg_engine *e, const uint8_t *tokens, int n_seq, float *logits)
{
    const Model &m = e->model;
    const int C = m.cfg.n_embd, H = m.cfg.n_head, hs = m.hs, L = m.cfg.n_layer;
    const size_t CC = (size_t)C * C;
    const int chunk = std::min(n_seq, 32);
    const size_t Mc = (size_t)chunk * 256, need = Mc * C * 10;
    if (need > e->p_ws_floats) {
        cudaFree(e->p_ws);
        e->p_ws = nullptr; e->p_ws_floats = 0;
        CU(dalloc(&e->p_ws, need));
        e->p_ws_floats = need;
    }
    float *X = e->p_ws, *XN = X + Mc * C, *QKV = XN + Mc * C, *ATT = QKV + Mc * 3 * C, *HID = ATT + Mc * C;
    const float *wte = e->pw, *wpe = wte + (size_t)MG_VOCAB * C, *lw = wpe + (size_t)256 * C;
    const float *lnf = lw + (size_t)L * (2 * C + 12 * CC);
    if (hs == 32) CU(cudaFuncSetAttribute(p_attn_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, p_attn_smem_bytes<32>()));
    else CU(cudaFuncSetAttribute(p_attn_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, p_attn_smem_bytes<64>()));
    auto gemm = [&](int epi, const float *A, const float *W, float *Cm, int M, int N, int K) {
        const dim3 grid((N + 63) / 64, M / 64);
        e->launches++;
        if (epi == P_EPI_STORE) p_gemm_kernel<P_EPI_STORE><<<grid, 256, 0, e->stream>>>(A, W, Cm, M, N, K);
        else if (epi == P_EPI_RESID) p_gemm_kernel<P_EPI_RESID><<<grid, 256, 0, e->stream>>>(A, W, Cm, M, N, K);
        else p_gemm_kernel<P_EPI_GELU><<<grid, 256, 0, e->stream>>>(A, W, Cm, M, N, K);
    };
    for (int s0 = 0; s0 < n_seq; s0 += chunk) {
        const int ns = std::min(chunk, n_seq - s0), M = ns * 256;
        e->launches += 2 + 3 * L;
        p_embed_kernel<<<(M + 7) / 8, 256, 0, e->stream>>>(tokens + (size_t)s0 * 256, wte, wpe, X, C, M);
        for (int l = 0; l < L; l++) {
            const float *ln1 = lw + (size_t)l * (2 * C + 12 * CC), *wqkv = ln1 + C, *wproj = wqkv + 3 * CC, *ln2 = wproj + CC,
                        *wfc = ln2 + C, *wproj2 = wfc + 4 * CC;
            p_ln_kernel<<<(M + 7) / 8, 256, 0, e->stream>>>(X, ln1, XN, C, M);
            gemm(P_EPI_STORE, XN, wqkv, QKV, M, 3 * C, C);
            if (hs == 32) p_attn_kernel<32><<<ns * H, 256, p_attn_smem_bytes<32>(), e->stream>>>(QKV, ATT, H, C);
            else p_attn_kernel<64><<<ns * H, 256, p_attn_smem_bytes<64>(), e->stream>>>(QKV, ATT, H, C);
            gemm(P_EPI_RESID, ATT, wproj, X, M, C, C);
            p_ln_kernel<<<(M + 7) / 8, 256, 0, e->stream>>>(X, ln2, XN, C, M);
            gemm(P_EPI_GELU, XN, wfc, HID, M, 4 * C, C);
            gemm(P_EPI_RESID, HID, wproj2, X, M, C, 4 * C);
        }
        p_head_kernel<<<(ns + 3) / 4, 128, 0, e->stream>>>(X, lnf, wte, logits + (size_t)s0 * 8, C, ns);
    }
    CU(cudaGetLastError());
    return MG_OK;
}

// tokens (device, uint8 [n_seq][256]) -> logits (device, fp32 [n_seq][8])
// optional per-sequence validation outputs of forward_device (mg_engine_eval_tokens): device pointers, all three or none
struct EvalOut {
    const int8_t *targets = nullptr;   // [n_seq] ground-truth action of each row, -1 = ignore
    float *loss = nullptr;             // [n_seq] cross-entropy over the 67 logits of position 255
    int32_t *pred = nullptr;           // [n_seq] arg-max action
};
static void launch_head_loss(mg_engine *e, const float *X, int compact, const EvalOut &ev, int s0, int ns)
{
    const Model &m = e->model;
    const int C = m.cfg.n_embd;
    e->launches++;
    head_loss_kernel<<<(ns + 3) / 4, 128, 4 * C * sizeof(float), e->stream>>>(X, compact, m.lnf, m.wte, ev.targets + s0, ev.loss + s0,
                                                                             ev.pred + s0, C, ns);
}

// fork: both lanes start behind everything queued on the engine's stream (the tokens of this timestep)
static int lanes_begin(mg_engine *e, int chunk_seqs)
{
    if (!e->lanes_ready) {
        int least = 0, greatest = 0;
        CU(cudaDeviceGetStreamPriorityRange(&least, &greatest));
        for (Lane &ln : e->lanes) {
            CU(cudaStreamCreateWithPriority(&ln.lo, cudaStreamNonBlocking, least));
            CU(cudaStreamCreateWithPriority(&ln.hi, cudaStreamNonBlocking, greatest));
            CU(cudaEventCreateWithFlags(&ln.ev_attn, cudaEventDisableTiming));
            CU(cudaEventCreateWithFlags(&ln.ev_post, cudaEventDisableTiming));
        }
        CU(cudaEventCreateWithFlags(&e->ev_fork, cudaEventDisableTiming));
        e->lanes_ready = true;
    }
    cudaStream_t main = e->stream;
    const int rc = ensure_workspace(e, e->lanes[1].ws, chunk_seqs);   // (its memsets are queued on the engine's stream, before the fork)
    if (rc) return rc;
    CU(cudaEventRecord(e->ev_fork, main));
    for (Lane &ln : e->lanes) CU(cudaStreamWaitEvent(ln.lo, e->ev_fork, 0));
    return MG_OK;
}

static int forward_device(mg_engine *e, const uint8_t *tokens, int n_seq, float *logits, const EvalOut *ev = nullptr)
{
    Model &m = e->model;
    if (!m.loaded) return fail(MG_ERR_STATE, "forward: no model loaded (mg_engine_load_model)");
    if (e->precise) {
        if (ev) return fail(MG_ERR_STATE, "validation loss is not available in the fp32 verification mode");
        return forward_precise(e, tokens, n_seq, logits);
    }
    int rc = ensure_workspace(e, e->ws, n_seq);
    if (rc) return rc;
    const int C = m.cfg.n_embd, H = m.cfg.n_head, hs = m.hs;
    // Stream lanes (struct Lane): chunks alternate between two workspaces / stream pairs; per-kernel event timing (profiling),
    // the validation outputs and the max-subtracting fallback keep the single-stream order.
    const int chunk_seqs = e->ws.chunk_seqs;
    const bool lanes = m.fused && hs == 32 && e->n_lanes == 2 && !e->profiling && !ev && n_seq > chunk_seqs &&
                       getenv("MAPF_GPT_B200_ATTN_CLASSIC") == nullptr;
    struct LaneScope {   // launches below go to e->stream: point it at the lane's streams and restore on every exit path
        mg_engine *e; cudaStream_t main;
        ~LaneScope() { e->stream = main; e->attn_ctr_slot = nullptr; e->attn_grid_cap = 0; }
    } scope{e, e->stream};
    if (lanes) {
        if ((rc = lanes_begin(e, chunk_seqs))) return rc;
    }
    for (int s0 = 0, ci = 0; s0 < n_seq; s0 += chunk_seqs, ci++) {
        const int ns = std::min(chunk_seqs, n_seq - s0);
        const int M = ns * 256, MT = M / 128;
        Lane *ln = lanes ? &e->lanes[ci & 1] : nullptr;
        Workspace &w = (ln && (ci & 1)) ? ln->ws : e->ws;   // lane 0 works in the engine's own workspace
        if (ln) {
            e->stream = ln->lo;
            e->attn_ctr_slot = &ln->ctr;
            e->attn_grid_cap = e->lane_attn_grid > 0 ? e->lane_attn_grid : e->n_sms;
        }
        if (m.fused) {
            // with >= 2 blocks the first post_attn takes its residual tile from the table itself: the lookup writes q/k/v only
            static const bool x_via_hbm = getenv("MAPF_GPT_B200_BLOCK0_X_VIA_HBM") != nullptr;
            const bool x_from_tab = m.tab0 && m.cfg.n_layer >= 2 && !x_via_hbm;
            // ... and block 0's attention gathers q/k/v from the table itself: no lookup kernel, no q/k/v round trip through HBM
            // (MAPF_GPT_B200_NO_BLOCK0_GATHER=1 or the classic attention kernel keep the lookup kernel)
            const bool no_gather = getenv("MAPF_GPT_B200_NO_BLOCK0_GATHER") != nullptr;   // per forward: tests flip it between engines
            const bool gather0 = x_from_tab && hs == 32 && !no_gather && getenv("MAPF_GPT_B200_ATTN_CLASSIC") == nullptr &&
                                 !(m.cfg.n_layer == 1 && e->prune_last);
            if (!gather0) prof_begin(e, KC_EMBED);
            if (gather0) {
                // nothing to launch
            } else if (m.tab0 && C == 160)
                block0_lookup_kernel<160><<<MT, 128, 0, e->stream>>>(tokens + (size_t)s0 * 256, m.tab0, x_from_tab ? nullptr : w.X, w.QKV);
            else if (m.tab0 && C == 256)
                block0_lookup_kernel<256><<<MT, 128, 0, e->stream>>>(tokens + (size_t)s0 * 256, m.tab0, x_from_tab ? nullptr : w.X, w.QKV);
            else
                embed_ln_kernel<<<MT, 128, 67 * (C + 4) * 4, e->stream>>>(tokens + (size_t)s0 * 256, m.wte, m.wpe_ti, m.layers[0].ln1,
                                                                          w.X, w.XN, C);
            if (!gather0) prof_end(e);
            for (int l = 0; l < m.cfg.n_layer; l++) {
                const Layer &L = m.layers[l];
                if ((l == 0 && !m.tab0) || !m.fuse_qkv) {   // blocks >= 1 get q/k/v from the previous block's fused kernel
                    GemmArgs g{};
                    g.A = w.XN; g.W = L.wqkv; g.out = w.QKV; g.M = M; g.N = 3 * C; g.K = C; g.C = C; g.n_head = H; g.hs = hs;
                    if ((rc = launch_gemm<EPI_QKV>(e, m.BN, g, KC_QKV))) return rc;
                }
                const bool last = l + 1 == m.cfg.n_layer;
                if (last && e->prune_last) {
                    // last block: Q / attention / c_proj / MLP only for token 255 of each sequence (App. D.2)
                    prof_begin(e, KC_ATTN_LAST);
                    // q carries log2(e) / sqrt(hs) (Model::q_fold): ln(2) brings q k^T back to natural-log units
                    if (hs == 32)
                        last_attn_kernel<32><<<ns, 32 * H, 0, e->stream>>>(w.QKV, w.X, w.ATTc, w.Xc, H, C, 0.6931471805599453f);
                    else
                        last_attn_kernel<64><<<ns, 32 * H, 0, e->stream>>>(w.QKV, w.X, w.ATTc, w.Xc, H, C, 0.6931471805599453f);
                    prof_end(e);
                    PostAttnArgs pa{};
                    pa.att = w.ATTc; pa.x = w.Xc; pa.wstream = L.wstream; pa.wstream_pair = L.wstream_pair; pa.ln2_gain = L.ln2;
                    if ((rc = launch_post_attn(e, C, pa, (ns + 127) / 128, true))) return rc;
                    prof_begin(e, KC_HEAD);
                    head_compact_kernel<<<(ns + 3) / 4, 128, 0, e->stream>>>(w.Xc, m.lnf, m.wte, logits + (size_t)s0 * 8, C, ns);
                    prof_end(e);
                    if (ev) launch_head_loss(e, w.Xc, 1, *ev, s0, ns);
                    break;
                }
                AttnArgs at{};
                at.qkv = w.QKV; at.out = w.ATT; at.n_head = H; at.C = C;
                at.scale_log2e = 1.0f;   // folded into Wq (Model::q_fold)
                at.timeline = e->d_timeline;
                static const int stamp_item = getenv("MAPF_GPT_B200_STAMP_ITEM") ? atoi(getenv("MAPF_GPT_B200_STAMP_ITEM")) : 40;
                at.dbg_variant = stamp_item;   // which item of a persistent attention CTA tools/timeline.py stamps
                if (l == 0 && gather0) { at.tokens0 = tokens + (size_t)s0 * 256; at.tab0 = m.tab0; at.tab_nrec = C / 4 + 3 * C / 8; at.tab_qkv0 = C / 4; }
                if (ln) {   // attention on the lane's high-priority stream, behind everything the low-priority one has queued
                    CU(cudaEventRecord(ln->ev_post, ln->lo));
                    CU(cudaStreamWaitEvent(ln->hi, ln->ev_post, 0));
                    e->stream = ln->hi;
                }
                rc = launch_attn(e, at, hs, ns, e->stream, !e->safe_softmax);
                if (ln) {
                    e->stream = ln->lo;
                    if (!rc) {
                        CU(cudaEventRecord(ln->ev_attn, ln->hi));
                        CU(cudaStreamWaitEvent(ln->lo, ln->ev_attn, 0));
                    }
                }
                if (rc) return rc;
                PostAttnArgs pa{};
                pa.att = w.ATT; pa.x = w.X; pa.wstream = L.wstream; pa.wstream_pair = L.wstream_pair; pa.ln2_gain = L.ln2;
                pa.next_gain = last ? nullptr : m.layers[l + 1].ln1;
                pa.xn_out = (last || m.fuse_qkv) ? nullptr : w.XN;
                pa.qkv_out = (!last && m.fuse_qkv) ? w.QKV : nullptr;
                pa.n_head = H; pa.hs = hs;
                pa.timeline = e->d_timeline;
                if (l == 0 && x_from_tab) { pa.tokens0 = tokens + (size_t)s0 * 256; pa.tab0 = m.tab0; pa.tab_nrec = C / 4 + 3 * C / 8; }
                // the next block is the pruned last one: it reads the residual and q of token 255 only (last_attn_kernel)
                const bool full_tail = getenv("MAPF_GPT_B200_FULL_TAIL_STORES") != nullptr;   // per forward: tests flip it between engines
                pa.tail_rows_only = (e->prune_last && l + 2 == m.cfg.n_layer && m.fuse_qkv && !full_tail) ? 1 : 0;
                // 24-bit residual stream between consecutive post_attn launches (pack24x16); the rows last_attn_kernel / head_kernel
                // read stay fp32.  MAPF_GPT_B200_X24=0 keeps fp32 everywhere.
                const bool x24 = !(getenv("MAPF_GPT_B200_X24") && getenv("MAPF_GPT_B200_X24")[0] == '0');
                pa.x_in_24 = (x24 && l > 0) ? 1 : 0;
                pa.x_out_24 = (x24 && !last && !pa.tail_rows_only && !(e->prune_last && l + 2 == m.cfg.n_layer)) ? 1 : 0;
                if ((rc = launch_post_attn(e, C, pa, MT, false))) return rc;
            }
            if (e->prune_last) continue;
            prof_begin(e, KC_HEAD);
            head_kernel<<<(ns + 3) / 4, 128, 0, e->stream>>>(w.X, m.lnf, m.wte, logits + (size_t)s0 * 8, C, ns);
            prof_end(e);
            if (ev) launch_head_loss(e, w.X, 0, *ev, s0, ns);
            continue;
        }
        const bool lnf = m.ln_fused;    // ln_1 / ln_2 live in the epilogues of the GEMMs around them; w.XN holds the RAW bf16 residual
        bool pruned_tail = false;
        // 24-bit residual stream between the residual GEMMs of the CTA-pair path (GemmArgs::x_in_24 / x_out_24): everything between
        // embed_kernel (writes fp32) and the last full-size residual GEMM (writes fp32 for last_attn_kernel / head_kernel)
        const bool x24_env = !(getenv("MAPF_GPT_B200_X24") && getenv("MAPF_GPT_B200_X24")[0] == '0');
        const bool x24 = x24_env && lnf;
        const bool tail_pruned = e->prune_last && (hs == 32 || hs == 64) && 32 * H <= 512;
        const int last_full = tail_pruned ? m.cfg.n_layer - 2 : m.cfg.n_layer - 1;   // block whose mlp c_proj is the last full-size one
        prof_begin(e, KC_EMBED);
        const bool embed_rows = getenv("MAPF_GPT_B200_EMBED_ROWS") != nullptr;   // A/B and tests: the lane-per-row kernel
        if (lnf && C == 768 && !embed_rows) {
            static PerDevice attr_set;
            if (const int d = cur_device(); !attr_set.v[d].load()) {
                CU(cudaFuncSetAttribute(embed_tile_kernel<768>, cudaFuncAttributeMaxDynamicSharedMemorySize, embed_tile_smem_bytes()));
                attr_set.v[d].store(1);
            }
            embed_tile_kernel<768><<<MT, 128, embed_tile_smem_bytes(), e->stream>>>(tokens + (size_t)s0 * 256, m.wte, m.wpe, w.X, w.XN, w.STATS);
        } else
        embed_kernel<<<MT, 128, 0, e->stream>>>(tokens + (size_t)s0 * 256, m.wte, m.wpe, w.X, C, lnf ? w.XN : nullptr, w.STATS);
        prof_end(e);
        for (int l = 0; l < m.cfg.n_layer; l++) {
            const Layer &L = m.layers[l];
            if (!lnf) {
                prof_begin(e, KC_LN);
                launch_ln(e, w.X, L.ln1, w.XN, C, MT);
                prof_end(e);
            }
            GemmArgs g{};
            g.A = w.XN; g.W = L.wqkv; g.Wp = L.wqkv_p; g.out = w.QKV; g.M = M; g.N = 3 * C; g.K = C; g.C = C; g.n_head = H; g.hs = hs;
            if (lnf) { g.stats_in = w.STATS; g.colsum = L.cs_qkv; }
            const bool full_tail = getenv("MAPF_GPT_B200_FULL_TAIL_STORES") != nullptr;   // per forward: tests flip it between engines
            g.tail_rows_only = (tail_pruned && l + 1 == m.cfg.n_layer && !full_tail) ? 1 : 0;   // q of the pruned block: token 255 only
            if ((rc = launch_gemm<EPI_QKV>(e, m.BN, g, KC_QKV))) return rc;
            if (l + 1 == m.cfg.n_layer && e->prune_last && (hs == 32 || hs == 64) && 32 * H <= 512) {
                // Last block pruned (as on the fused path): only logits[255][0:5] are consumed, so K and V are
                // needed for all tokens but attention, c_proj and the MLP for token 255 alone.  The tail runs on COMPACT tiles (one
                // row per sequence) through the single-CTA GEMM with the plain weights and the LayerNorm kernel.
                prof_begin(e, KC_ATTN_LAST);
                if (hs == 32) last_attn_kernel<32><<<ns, 32 * H, 0, e->stream>>>(w.QKV, w.X, w.ATTc, w.Xc, H, C, 0.6931471805599453f);
                else last_attn_kernel<64><<<ns, 32 * H, 0, e->stream>>>(w.QKV, w.X, w.ATTc, w.Xc, H, C, 0.6931471805599453f);
                prof_end(e);
                const int Mc = (ns + 127) / 128 * 128;
                GemmArgs t{};
                t.A = w.ATTc; t.W = L.wproj; t.out = w.Xc; t.M = Mc; t.N = C; t.K = C;
                if ((rc = launch_gemm<EPI_RESID>(e, m.BN, t, KC_POST_LAST))) return rc;
                prof_begin(e, KC_POST_LAST);
                launch_ln(e, w.Xc, L.ln2, w.XN, C, Mc / 128);
                prof_end(e);
                t = GemmArgs{};
                t.A = w.XN; t.W = L.wfc; t.out = w.HID; t.M = Mc; t.N = 4 * C; t.K = C;
                if ((rc = launch_gemm<EPI_GELU>(e, m.BN, t, KC_POST_LAST))) return rc;
                t = GemmArgs{};
                t.A = w.HID; t.W = L.wproj2; t.out = w.Xc; t.M = Mc; t.N = C; t.K = 4 * C;
                if ((rc = launch_gemm<EPI_RESID>(e, m.BN, t, KC_POST_LAST))) return rc;
                prof_begin(e, KC_HEAD);
                head_compact_kernel<<<(ns + 3) / 4, 128, 0, e->stream>>>(w.Xc, m.lnf, m.wte, logits + (size_t)s0 * 8, C, ns);
                prof_end(e);
                if (ev) launch_head_loss(e, w.Xc, 1, *ev, s0, ns);
                pruned_tail = true;
                break;
            }
            AttnArgs at{};
            at.qkv = w.QKV; at.out = w.ATT; at.n_head = H; at.C = C;
            at.scale_log2e = 1.0f;   // folded into Wq (Model::q_fold)
            if ((rc = launch_attn(e, at, hs, ns, e->stream, !e->safe_softmax))) return rc;
            g = GemmArgs{};
            g.A = w.ATT; g.W = L.wproj; g.Wp = L.wproj_p; g.out = w.X; g.M = M; g.N = C; g.K = C;
            if (lnf) { g.xb_out = w.XN; g.stats_out = w.STATS; }                       // operand + statistics of ln_2
            if (x24 && g.Wp) {   // block 0 changes the layout: fp32 from w.X -> 24-bit into w.X24; later blocks stay in w.X24
                g.x_in_24 = l > 0; g.x_out_24 = 1;
                g.resid_in = l > 0 ? w.X24 : w.X; g.out = w.X24;
            }
            if ((rc = launch_gemm<EPI_RESID>(e, m.BN, g, KC_PROJ))) return rc;
            if (!lnf) {
                prof_begin(e, KC_LN);
                launch_ln(e, w.X, L.ln2, w.XN, C, MT);
                prof_end(e);
            }
            g = GemmArgs{};
            g.A = w.XN; g.W = L.wfc; g.Wp = L.wfc_p; g.out = w.HID; g.M = M; g.N = 4 * C; g.K = C;
            if (lnf) { g.stats_in = w.STATS; g.colsum = L.cs_fc; }
            if ((rc = launch_gemm<EPI_GELU>(e, m.BN, g, KC_FC))) return rc;
            g = GemmArgs{};
            g.A = w.HID; g.W = L.wproj2; g.Wp = L.wproj2_p; g.out = w.X; g.M = M; g.N = C; g.K = 4 * C;
            if (lnf && l + 1 < m.cfg.n_layer) { g.xb_out = w.XN; g.stats_out = w.STATS; }   // ... of the next block's ln_1
            if (x24 && g.Wp) {   // the last full-size one changes the layout back: 24-bit from w.X24 -> fp32 into w.X
                g.x_in_24 = 1; g.x_out_24 = l < last_full;
                g.resid_in = w.X24; g.out = l < last_full ? w.X24 : w.X;
            }
            // last_attn_kernel reads the residual of token 255 only (with a LayerNorm kernel of its own the next block reads every row)
            g.tail_rows_only = (lnf && tail_pruned && l == last_full && !full_tail) ? 1 : 0;
            if ((rc = launch_gemm<EPI_RESID>(e, m.BN, g, KC_PROJ2))) return rc;
        }
        if (pruned_tail) continue;
        prof_begin(e, KC_HEAD);
        head_kernel<<<(ns + 3) / 4, 128, 0, e->stream>>>(w.X, m.lnf, m.wte, logits + (size_t)s0 * 8, C, ns);
        prof_end(e);
        if (ev) launch_head_loss(e, w.X, 0, *ev, s0, ns);
    }
    if (lanes) {   // join: the engine's stream continues after both lanes
        for (int i = 0; i < 2; i++) {
            CU(cudaEventRecord(e->lanes[i].ev_post, e->lanes[i].lo));
            CU(cudaStreamWaitEvent(scope.main, e->lanes[i].ev_post, 0));
        }
    }
    CU(cudaGetLastError());
    return MG_OK;
}

// ------------------------------------------------------------------------------------------- env helpers
static size_t bfs_smem(const EnvState &s) { return (size_t)s.H * s.P * 7 + 16; }
static size_t step_smem(const EnvState &s) { return (size_t)s.N * 5 + 16; }
static size_t partial_smem(const EnvState &s)
{
    const size_t W = 2 * s.gs + 1, Q = W * W + 4 * W + 8;
    auto al = [](size_t b) { return (b + 15) & ~(size_t)15; };
    return al(W * W * 2) + al((size_t)(s.gs + 1) * (s.gs + 1) * 2) + 2 * al(Q * 2) + al((4 * W + 4) * 4) + al((4 * W + 4) * 2) + 16;
}
static const int MG_MAX_MAPS = 64;

// (re)compute cost-to-go fields: whole-grid BFS for small grids, windowed partials for large ones
static int launch_fields(mg_engine *e, int first_env, int n_envs, int only_dirty)
{
    EnvState &s = e->s;
    prof_begin(e, KC_BFS);
    if (s.large) partial_kernel<<<n_envs * s.N, 256, partial_smem(s), e->stream>>>(s, first_env, only_dirty);
    else bfs_kernel<<<n_envs * s.N, 128, bfs_smem(s), e->stream>>>(s, first_env, only_dirty);
    prof_end(e);
    CU(cudaGetLastError());
    return MG_OK;
}

// precompute_cost2go (cpp:43-113) for one distinct large map; returns its slot
static int add_large_map(mg_engine *e, const uint8_t *grid_pitched, int *slot_out)
{
    EnvState &s = e->s;
    const size_t cells = (size_t)s.H * s.P;
    for (size_t m = 0; m < e->map_grids.size(); m++)
        if (memcmp(e->map_grids[m].data(), grid_pitched, cells) == 0) { *slot_out = (int)m; return MG_OK; }
    if ((int)e->map_grids.size() >= MG_MAX_MAPS) return fail(MG_ERR_ARG, "more than %d distinct large maps in one engine", MG_MAX_MAPS);
    const int slot = (int)e->map_grids.size();
    std::vector<int32_t> cidx(cells, -1), list;
    for (int i = 0; i < s.H; i += s.gs)
        for (int j = 0; j < s.W; j++)
            if (!grid_pitched[(size_t)i * s.P + j] && cidx[(size_t)i * s.P + j] < 0) { cidx[(size_t)i * s.P + j] = (int)list.size(); list.push_back(i * s.P + j); }
    for (int i = 0; i < s.H; i++)
        for (int j = 0; j < s.W; j += s.gs)
            if (!grid_pitched[(size_t)i * s.P + j] && cidx[(size_t)i * s.P + j] < 0) { cidx[(size_t)i * s.P + j] = (int)list.size(); list.push_back(i * s.P + j); }
    const int K = (int)list.size();
    uint16_t *pre = nullptr;
    CU(dalloc(&pre, (size_t)std::max(K, 1) * std::max(K, 1)));
    // save_cost2go (cpp:62-80): "precomputed_cost2go.bin" in the working directory replaces the computation when present.
    // Format: size_t rows, size_t cols, rows x cols uint16 (row = source cell, in precomputed_cells_map order).  The
    // reference loads whatever it finds; a table whose shape does not fit this map is rejected here instead of being used.
    bool loaded = false;
    if (e->params.save_cost2go && K > 0) {
        if (FILE *f = fopen("precomputed_cost2go.bin", "rb")) {
            size_t dims[2] = {0, 0};
            std::vector<uint16_t> tab((size_t)K * K);
            const bool ok = fread(dims, sizeof(size_t), 2, f) == 2 && dims[0] == (size_t)K && dims[1] == (size_t)K &&
                            fread(tab.data(), 2, tab.size(), f) == tab.size();
            fclose(f);
            if (!ok) {
                cudaFree(pre);
                return fail(MG_ERR_ARG, "precomputed_cost2go.bin holds a %zu x %zu table, this map needs %d x %d: stale cache "
                                        "(the reference would use it unchecked, cpp:62-80)", dims[0], dims[1], K, K);
            }
            CU(cudaMemcpy(pre, tab.data(), tab.size() * 2, cudaMemcpyHostToDevice));
            loaded = true;
        }
    }
    uint8_t *d_grid = nullptr, *scratch = nullptr;
    int32_t *d_list = nullptr;
    CU(dalloc(&d_grid, cells));
    CU(dalloc(&d_list, (size_t)std::max(K, 1)));
    CU(cudaMemcpy(d_grid, grid_pitched, cells, cudaMemcpyHostToDevice));
    if (K > 0 && !loaded) {
        CU(cudaMemcpy(d_list, list.data(), (size_t)K * 4, cudaMemcpyHostToDevice));
        const int blocks = std::min(K, 2 * e->n_sms);
        CU(dalloc(&scratch, (size_t)blocks * (cells * 10 + 64)));
        e->launches++;
        precompute_kernel<<<blocks, 256, 0, e->stream>>>(d_grid, s.H, s.W, s.P, d_list, K, pre, scratch);
        CU(cudaStreamSynchronize(e->stream));
        if (e->params.save_cost2go) {   // cpp:114-131
            std::vector<uint16_t> tab((size_t)K * K);
            CU(cudaMemcpy(tab.data(), pre, tab.size() * 2, cudaMemcpyDeviceToHost));
            if (FILE *f = fopen("precomputed_cost2go.bin", "wb")) {   // like the reference, a file that cannot be opened is skipped
                const size_t dims[2] = {(size_t)K, (size_t)K};
                fwrite(dims, sizeof(size_t), 2, f);
                fwrite(tab.data(), 2, tab.size(), f);
                fclose(f);
            }
        }
    }
    cudaFree(scratch); cudaFree(d_grid); cudaFree(d_list);
    CU(cudaMemcpy(s.cell_idx + (size_t)slot * cells, cidx.data(), cells * 4, cudaMemcpyHostToDevice));
    CU(cudaMemcpy(s.pre + slot, &pre, sizeof pre, cudaMemcpyHostToDevice));
    CU(cudaMemcpy(s.preK + slot, &K, 4, cudaMemcpyHostToDevice));
    e->map_grids.emplace_back(grid_pitched, grid_pitched + cells);
    e->map_pre.push_back(pre);
    *slot_out = slot;
    return MG_OK;
}

// cuTensorMapEncodeTiled through the runtime's driver entry point (no -lcuda link dependency)
typedef CUresult (*encode_tiled_fn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                    const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static int make_window_maps(mg_engine *e)
{
    void *fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) != cudaSuccess || !fn)
        return fail(MG_ERR_CUDA, "cuTensorMapEncodeTiled is not available");
    encode_tiled_fn enc = reinterpret_cast<encode_tiled_fn>(fn);
    const EnvState &s = e->s;
    const cuuint32_t box[3] = {24, 11, 1}, estr[3] = {1, 1, 1};   // 24 cols: 8-aligned start + 11-wide window (see kernel)
    {   // cost-to-go fields: [E*N planes][H rows][P cols] u16
        const cuuint64_t dims[3] = {(cuuint64_t)s.FP, (cuuint64_t)s.FR, (cuuint64_t)s.E * s.N};
        const cuuint64_t strides[2] = {(cuuint64_t)s.FP * 2, (cuuint64_t)s.FR * s.FP * 2};
        if (enc(&e->map_c2g, CU_TENSOR_MAP_DATA_TYPE_UINT16, 3, s.c2g, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
            return fail(MG_ERR_CUDA, "cuTensorMapEncodeTiled(c2g) failed");
    }
    {   // agent-id maps: [E planes][H][P] i16
        const cuuint64_t dims[3] = {(cuuint64_t)s.P, (cuuint64_t)s.H, (cuuint64_t)s.E};
        const cuuint64_t strides[2] = {(cuuint64_t)s.P * 2, (cuuint64_t)s.H * s.P * 2};
        if (enc(&e->map_loc, CU_TENSOR_MAP_DATA_TYPE_UINT16, 3, s.loc, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
            return fail(MG_ERR_CUDA, "cuTensorMapEncodeTiled(loc) failed");
    }
    return MG_OK;
}

static int launch_observe(mg_engine *e, bool update, bool tokens)
{
    if (e->n_envs == 0) return fail(MG_ERR_STATE, "no environment has been reset");
    if (tokens && e->use_tma) {
        prof_begin(e, KC_OBSERVE);
        if (update) observe_tma_kernel<true><<<e->n_envs, 256, 0, e->stream>>>(e->s, e->map_c2g, e->map_loc);
        else observe_tma_kernel<false><<<e->n_envs, 256, 0, e->stream>>>(e->s, e->map_c2g, e->map_loc);
        prof_end(e);
        CU(cudaGetLastError());
        return MG_OK;
    }
    prof_begin(e, KC_OBSERVE);
    if (update && tokens) observe_kernel<true, true><<<e->n_envs, 256, 0, e->stream>>>(e->s);
    else if (update) observe_kernel<true, false><<<e->n_envs, 256, 0, e->stream>>>(e->s);
    else observe_kernel<false, true><<<e->n_envs, 256, 0, e->stream>>>(e->s);
    prof_end(e);
    CU(cudaGetLastError());
    return MG_OK;
}
static int launch_step(mg_engine *e, int mode, int do_step, const float *q, const int32_t *override_act)
{
    StepArgs a{};
    a.mode = mode; a.do_step = do_step; a.q = q; a.seed = e->seed; a.act_override = override_act;
    a.env_offset = e->env_offset; a.max_episode_steps = e->max_episode_steps;
    prof_begin(e, KC_STEP);
    sample_step_kernel<<<e->n_envs, 256, step_smem(e->s), e->stream>>>(e->s, a);
    prof_end(e);
    CU(cudaGetLastError());
    return MG_OK;
}
static int check_vocab(mg_engine *e)
{
    int32_t v = 0;
    CU(cudaMemcpyAsync(&v, e->s.vocab_err, 4, cudaMemcpyDeviceToHost, e->stream));
    CU(cudaStreamSynchronize(e->stream));
    if (v) {
        CU(cudaMemsetAsync(e->s.vocab_err, 0, 4, e->stream));
        return fail(MG_ERR_VOCAB, "a relative position left the token vocabulary (agents outside the FOV window?)");
    }
    return MG_OK;
}
// Reads and clears both device flags after the stream has drained.  *numeric = 1 when sample_step_kernel saw a non-finite logit.
static int read_flags(mg_engine *e, int *vocab, int *numeric)
{
    int32_t v[2] = {0, 0};
    CU(cudaMemcpyAsync(v, e->s.vocab_err, 8, cudaMemcpyDeviceToHost, e->stream));
    CU(cudaStreamSynchronize(e->stream));
    if (v[0] || v[1]) CU(cudaMemsetAsync(e->s.vocab_err, 0, 8, e->stream));
    *vocab = v[0]; *numeric = v[1];
    return MG_OK;
}
// A non-finite logit after a device-resident rollout cannot be redone (the envs have moved on): the engine switches to the
// max-subtracting softmax kernels for every later call and reports the rollout as failed.
static int numeric_failure(mg_engine *e)
{
    if (!e->safe_softmax && !e->precise) {
        e->safe_softmax = true;
        return fail(MG_ERR_NUMERIC, "a logit was not finite: an attention score left the range of the max-free softmax "
                                    "(|s| < ~69 nats); the engine now uses the max-subtracting kernels -- reset and rerun "
                                    "(MAPF_GPT_B200_SAFE_SOFTMAX=1 selects them from the start)");
    }
    return fail(MG_ERR_NUMERIC, "a logit was not finite (inf/nan weights?): torch.multinomial would raise here (model.py:257)");
}

// Pure UMMA issue/execute rate: operands stay in smem (zero-filled), one thread issues `iters` x 4 UMMAs (M128 x N x K16,
// SS operands, no-swizzle K-major), then commits; out[0] = clock64 cycles from first issue to completion.
template <int N>
__global__ void __launch_bounds__(128) umma_rate_kernel(int iters, long long *out, int ts)
{
    extern __shared__ __align__(1024) uint8_t smem[];
    uint8_t *As = smem;                       // 4 k-steps: [8 kc][128][16B] = 16 KB
    uint8_t *Bs = smem + 16384;               // [8 kc][N][16B]
    __shared__ uint64_t bar;
    __shared__ uint32_t slot;
    for (int i = threadIdx.x; i < (16384 + 8 * N * 16) / 16; i += blockDim.x) reinterpret_cast<uint4 *>(smem)[i] = make_uint4(0, 0, 0, 0);
    if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
    if (threadIdx.x < 32) tmem_alloc<256>(&slot);
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = slot;
    if (threadIdx.x == 0) {
        constexpr uint32_t idesc = umma_idesc_bf16(128, N, 0, 0);
        const uint32_t a = smem_u32(As), b = smem_u32(Bs);
        const long long t0 = clock64();
        for (int it = 0; it < iters; it++) {
#pragma unroll
            for (int ks = 0; ks < 4; ks++) {
                if (ts) umma_ts(tmem, tmem + 256 - 32 + ks * 8, umma_desc(b + ks * 2 * N * 16, N * 16, 128), idesc, 1u);   // A in TMEM
                else umma_ss(tmem, umma_desc(a + ks * 4096, 2048, 128), umma_desc(b + ks * 2 * N * 16, N * 16, 128), idesc, 1u);
            }
        }
        const long long t1 = clock64();
        umma_commit(&bar);
        mbar_wait(&bar, 0);
        const long long t2 = clock64();
        if (blockIdx.x == 0) { out[0] = t2 - t0; out[1] = t1 - t0; }
    }
    tc_fence_before();
    __syncthreads();
    if (threadIdx.x < 32) tmem_dealloc<256>(tmem);
}
template <int N>
static int run_umma_rate(int iters, int ctas, long long *d_out, long long *h_out, int ts)
{
    const int smem = 16384 + 8 * N * 16;
    CU(cudaFuncSetAttribute(umma_rate_kernel<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    umma_rate_kernel<N><<<ctas, 128, smem>>>(iters, d_out, ts);
    CU(cudaDeviceSynchronize());
    CU(cudaMemcpy(h_out, d_out, 16, cudaMemcpyDeviceToHost));
    return MG_OK;
}

// host-supplied positions / goals of every reset slot, padded [n_envs][N][2]: the same rule as mg_engine_reset (the kernels
// index the grid and seed the BFS with them unchecked)
static int check_xy(const mg_engine *e, const int32_t *pos_xy, const int32_t *goal_xy)
{
    const EnvState &s = e->s;
    for (int k = 0; k < e->n_envs; k++)
        for (int a = 0; a < e->h_nag[k]; a++) {
            const size_t i = ((size_t)k * s.N + a) * 2;
            if (pos_xy && (pos_xy[i] < 5 || pos_xy[i + 1] < 5 || pos_xy[i] >= s.H - 5 || pos_xy[i + 1] >= s.W - 5))
                return fail(MG_ERR_ARG, "env %d agent %d: position (%d,%d) outside the padded grid %dx%d", k, a, pos_xy[i],
                            pos_xy[i + 1], s.H, s.W);
            if (goal_xy && (goal_xy[i] < 0 || goal_xy[i + 1] < 0 || goal_xy[i] >= s.H || goal_xy[i + 1] >= s.W))
                return fail(MG_ERR_ARG, "env %d agent %d: goal (%d,%d) outside the grid %dx%d", k, a, goal_xy[i], goal_xy[i + 1],
                            s.H, s.W);
        }
    return MG_OK;
}
__global__ void scale_bf16_kernel(__nv_bfloat16 *x, size_t n, float f)
{
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) x[i] = __float2bfloat16(__bfloat162float(x[i]) * f);
}

// =========================================================================================== C ABI
extern "C" {

int mg_version(void) { return 100; }
const char *mg_last_error(void) { return g_err.c_str(); }
void mg_default_params(mg_params *p)
{
    p->cost2go_value_limit = 20; p->num_agents = 13; p->num_previous_actions = 5; p->context_size = 256;
    p->obs_radius = 5; p->agents_radius = 5; p->grid_step = 64; p->save_cost2go = 0;
}
int mg_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) return 0;
    return n;
}

mg_engine *mg_engine_create(int device, int max_envs, int max_agents, int H, int W, const mg_params *params)
{
    mg_params p;
    mg_default_params(&p);
    if (params) p = *params;
    if (p.cost2go_value_limit != 20 || p.num_agents != 13 || p.num_previous_actions != 5 || p.context_size != 256 ||
        p.obs_radius != 5 || p.agents_radius != 5) {
        fail(MG_ERR_ARG, "only the trained observation shape is supported (limit 20, 13 agents, 5 previous actions, "
                         "context 256, radius 5); the checkpoints fix it (inference.py:14-22)");
        return nullptr;
    }
    if (max_envs < 1 || max_agents < 1 || max_agents > 32767) { fail(MG_ERR_ARG, "bad capacity"); return nullptr; }
    if (H < 11 || W < 11 || H > 512 || W > 512) {
        fail(MG_ERR_ARG, "padded grid %dx%d unsupported: need 11..512 cells per side", H, W);
        return nullptr;
    }
    if (p.grid_step != 64) { fail(MG_ERR_ARG, "grid_step must be 64 (inference.py:28)"); return nullptr; }
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        fail(MG_ERR_CUDA, "no CUDA device: this library has no CPU fallback");
        return nullptr;
    }
    if (cudaSetDevice(device) != cudaSuccess) { fail(MG_ERR_CUDA, "cudaSetDevice(%d) failed", device); return nullptr; }
    mg_engine *e = new mg_engine();
    e->device = device;
    e->params = p;
    EnvState &s = e->s;
    s.E = max_envs; s.N = max_agents; s.H = H; s.W = W; s.P = (W + 7) / 8 * 8;
    s.gs = p.grid_step;
    // H, W <= 74: every agent's window is the whole grid and the field is one goal BFS; beyond that the
    // windowed machinery of observation_generator.cpp:43-286 runs (precompute tables + per-agent partial fields)
    s.large = (H > 74 || W > 74) ? 1 : 0;
    s.FR = s.large ? 2 * s.gs + 1 : s.H;
    s.FP = s.large ? (2 * s.gs + 1 + 7) / 8 * 8 : s.P;
    const size_t cells = (size_t)s.H * s.P, EN = (size_t)s.E * s.N;
    bool ok = cudaStreamCreateWithFlags(&e->stream, cudaStreamNonBlocking) == cudaSuccess;
    ok = ok && dalloc(&s.obst, s.E * cells) == cudaSuccess;
    ok = ok && dalloc(&s.loc, s.E * cells) == cudaSuccess;
    ok = ok && dalloc(&s.c2g, EN * s.FR * s.FP) == cudaSuccess && dalloc(&s.bounds, EN) == cudaSuccess;
    ok = ok && dalloc(&s.map_of_env, (size_t)s.E) == cudaSuccess;
    if (s.large) {
        ok = ok && dalloc(&s.cell_idx, (size_t)MG_MAX_MAPS * cells) == cudaSuccess && dalloc(&s.pre, MG_MAX_MAPS) == cudaSuccess;
        ok = ok && dalloc(&s.preK, MG_MAX_MAPS) == cudaSuccess;
    }
    ok = ok && dalloc(&s.pos, EN) == cudaSuccess && dalloc(&s.goal, EN) == cudaSuccess;
    ok = ok && dalloc(&s.hist, EN * 8) == cudaSuccess && dalloc(&s.nextb, EN) == cudaSuccess;
    ok = ok && dalloc(&s.act, EN) == cudaSuccess && dalloc(&s.nag, (size_t)s.E) == cudaSuccess;
    ok = ok && dalloc(&s.dirty, EN) == cudaSuccess && dalloc(&s.tokens, EN * 256) == cudaSuccess;
    ok = ok && dalloc(&s.active, (size_t)s.E) == cudaSuccess;
    ok = ok && dalloc(&s.logits, EN * 8) == cudaSuccess;
    ok = ok && dalloc(&s.steps, (size_t)s.E) == cudaSuccess && dalloc(&s.done, (size_t)s.E) == cudaSuccess;
    ok = ok && dalloc(&s.arrive, EN) == cudaSuccess && dalloc(&s.agent_steps, (size_t)s.E) == cudaSuccess;
    ok = ok && dalloc(&s.vocab_err, 2) == cudaSuccess;
    ok = ok && dalloc(&s.density_sum, (size_t)s.E) == cudaSuccess && dalloc(&s.density_n, (size_t)s.E) == cudaSuccess;
    ok = ok && dalloc(&e->d_pos_in, EN * 2) == cudaSuccess && dalloc(&e->d_goal_in, EN * 2) == cudaSuccess;
    ok = ok && dalloc(&e->d_act_in, EN) == cudaSuccess && dalloc(&e->d_step_act, EN) == cudaSuccess;
    ok = ok && dalloc(&e->d_q, EN * 5) == cudaSuccess && dalloc(&e->d_metrics, (size_t)s.E * MG_METRIC_COLS) == cudaSuccess;
    if (ok) {
        cudaDeviceGetAttribute(&e->n_sms, cudaDevAttrMultiProcessorCount, device);
        cudaMemset(s.nag, 0, s.E * 4);
        cudaMemset(s.active, 1, s.E);
        cudaMemset(s.vocab_err, 0, 8);
        cudaMemset(s.density_sum, 0, s.E * 4);
        cudaMemset(s.density_n, 0, s.E * 4);
        e->h_nag.assign((size_t)s.E, 0);
        cudaMemset(s.tokens, 66, EN * 256);
        cudaMemset(s.logits, 0, EN * 8 * 4);
        cudaMemset(s.dirty, 0, EN);
        cudaMemset(s.done, 0, s.E);
        cudaMemset(s.steps, 0, s.E * 4);
        cudaMemset(s.agent_steps, 0, s.E * 8);
        cudaEventCreate(&e->ev_t0); cudaEventCreate(&e->ev_t1);
        for (auto &ev : e->ev_p) cudaEventCreate(&ev);
        if (s.large) cudaFuncSetAttribute(partial_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)partial_smem(s));
        else cudaFuncSetAttribute(bfs_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bfs_smem(s));
        cudaMemset(s.map_of_env, 0, (size_t)s.E * 4);
        cudaFuncSetAttribute(sample_step_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)step_smem(s));
        ok = cudaDeviceSynchronize() == cudaSuccess;
    }
    if (!ok) {
        fail(MG_ERR_CUDA, "engine allocation failed: %s", cudaGetErrorString(cudaGetLastError()));
        mg_engine_destroy(e);
        return nullptr;
    }
    {   // FOV windows by TMA tiled loads (default); MAPF_GPT_B200_NO_TMA_WINDOWS=1 selects the plain-load tokenizer
        const char *no_tma = getenv("MAPF_GPT_B200_NO_TMA_WINDOWS");
        if (!(no_tma && no_tma[0] == '1')) {
            if (make_window_maps(e) != MG_OK) {
                mg_engine_destroy(e);
                return nullptr;
            }
            e->use_tma = true;
        }
    }
    return e;
}

void mg_engine_destroy(mg_engine *e)
{
    if (!e) return;
    cudaSetDevice(e->device);
    cudaDeviceSynchronize();
    EnvState &s = e->s;
    for (auto p : e->map_pre) cudaFree(p);
    cudaFree(e->d_attn_ctr);
    cudaFree(s.bounds); cudaFree(s.map_of_env); cudaFree(s.cell_idx); cudaFree(s.pre); cudaFree(s.preK);
    cudaFree(s.obst); cudaFree(s.loc); cudaFree(s.c2g); cudaFree(s.pos); cudaFree(s.goal); cudaFree(s.hist);
    cudaFree(s.nextb); cudaFree(s.act); cudaFree(s.nag); cudaFree(s.dirty); cudaFree(s.tokens); cudaFree(s.logits);
    cudaFree(s.active); cudaFree(s.steps); cudaFree(s.done); cudaFree(s.arrive); cudaFree(s.agent_steps); cudaFree(s.vocab_err);
    cudaFree(s.density_sum); cudaFree(s.density_n); cudaFree(e->pw); cudaFree(e->p_ws);
    cudaFree(e->d_pos_in); cudaFree(e->d_goal_in); cudaFree(e->d_act_in); cudaFree(e->d_step_act); cudaFree(e->d_q);
    cudaFree(e->d_metrics);
    Model &m = e->model;
    cudaFree(m.wte); cudaFree(m.wpe); cudaFree(m.lnf); cudaFree(m.wpe_ti); cudaFree(m.tab0);
    for (auto &L : m.layers) { cudaFree(L.ln1); cudaFree(L.ln2); cudaFree(L.wqkv); cudaFree(L.wproj); cudaFree(L.wfc); cudaFree(L.wproj2); cudaFree(L.wstream); cudaFree(L.wstream_pair);
                                cudaFree(L.wqkv_p); cudaFree(L.wproj_p); cudaFree(L.wfc_p); cudaFree(L.wproj2_p); cudaFree(L.cs_qkv); cudaFree(L.cs_fc); }
    for (Workspace *wp : {&e->ws, &e->lanes[0].ws, &e->lanes[1].ws}) {
        Workspace &w = *wp;
        cudaFree(w.X); cudaFree(w.XN); cudaFree(w.QKV); cudaFree(w.ATT); cudaFree(w.HID); cudaFree(w.tok); cudaFree(w.logits);
        cudaFree(w.Xc); cudaFree(w.ATTc); cudaFree(w.STATS); cudaFree(w.X24);
    }
    for (Lane &ln : e->lanes) {
        cudaFree(ln.ctr);
        if (ln.ev_attn) cudaEventDestroy(ln.ev_attn);
        if (ln.ev_post) cudaEventDestroy(ln.ev_post);
        if (ln.lo) cudaStreamDestroy(ln.lo);
        if (ln.hi) cudaStreamDestroy(ln.hi);
    }
    if (e->ev_fork) cudaEventDestroy(e->ev_fork);
    for (auto &p : e->evs) { cudaEventDestroy(p.a); cudaEventDestroy(p.b); }
    if (e->ev_t0) cudaEventDestroy(e->ev_t0);
    if (e->ev_t1) cudaEventDestroy(e->ev_t1);
    for (auto &ev : e->ev_p) if (ev) cudaEventDestroy(ev);
    if (e->stream) cudaStreamDestroy(e->stream);
    delete e;
}

size_t mg_model_num_floats(const mg_model_config *c)
{
    const size_t C = c->n_embd, V = c->vocab_size, T = c->block_size, L = c->n_layer;
    return V * C + T * C + L * (C + 3 * C * C + C * C + C + 4 * C * C + 4 * C * C) + C;
}

int mg_engine_load_model(mg_engine *e, const mg_model_config *cfg, const float *w, size_t n_floats)
{
    if (!e || !cfg || !w) return fail(MG_ERR_ARG, "null argument");
    CU(cudaSetDevice(e->device));
    if (cfg->block_size != 256 || cfg->vocab_size != MG_VOCAB)
        return fail(MG_ERR_ARG, "block_size must be 256 and vocab_size 67 (got %d, %d)", cfg->block_size, cfg->vocab_size);
    const int C = cfg->n_embd, H = cfg->n_head;
    if (H < 1 || C % H) return fail(MG_ERR_ARG, "n_embd %% n_head != 0");
    const int hs = C / H, BN = pick_bn(C);
    if (hs != 32 && hs != 64) return fail(MG_ERR_ARG, "head size %d unsupported (32 or 64)", hs);
    if (!BN) return fail(MG_ERR_ARG, "n_embd %d unsupported (multiple of 128 or 160)", C);
    if (n_floats != mg_model_num_floats(cfg))
        return fail(MG_ERR_ARG, "weight buffer has %zu floats, expected %zu", n_floats, mg_model_num_floats(cfg));
    Model &m = e->model;
    if (m.loaded) return fail(MG_ERR_STATE, "a model is already loaded in this engine");
    m.cfg = *cfg; m.BN = BN; m.hs = hs;
    {
        const char *ss = getenv("MAPF_GPT_B200_SAFE_SOFTMAX"), *pr = getenv("MAPF_GPT_B200_PRECISION");
        e->safe_softmax = ss && ss[0] == '1';
        e->precise = (g_precision == 1 || (pr && strcmp(pr, "fp32") == 0)) ? 1 : 0;
        if (pr && strcmp(pr, "fp32") != 0 && strcmp(pr, "bf16") != 0)
            return fail(MG_ERR_ARG, "MAPF_GPT_B200_PRECISION=%s: expected bf16 (default) or fp32", pr);
        if (e->precise) {   // the verification forward reads the checkpoint as it is
            CU(dalloc(&e->pw, n_floats));
            CU(cudaMemcpy(e->pw, w, n_floats * 4, cudaMemcpyHostToDevice));
        }
    }
    const size_t CC = (size_t)C * C;
    int rc;
    if ((rc = upload_f32(w, (size_t)MG_VOCAB * C, &m.wte))) return rc;
    w += (size_t)MG_VOCAB * C;
    if ((rc = upload_f32(w, (size_t)256 * C, &m.wpe))) return rc;
    {
        std::vector<float> ti((size_t)256 * C);
        for (int pos = 0; pos < 256; pos++)
            for (int c = 0; c < C; c++)
                ti[(((size_t)(pos >> 7) * (C / 4) + c / 4) * 128 + (pos & 127)) * 4 + (c & 3)] = w[(size_t)pos * C + c];
        if ((rc = upload_f32(ti.data(), ti.size(), &m.wpe_ti))) return rc;
        CU(cudaFuncSetAttribute(embed_ln_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 67 * (768 + 4) * 4));
    }
    w += (size_t)256 * C;
    m.layers.resize(cfg->n_layer);
    const char *force_generic = getenv("MAPF_GPT_B200_GENERIC");
    m.fused = (C == 160 || C == 256) && !(force_generic && force_generic[0] == '1');
    {   // stream lanes (struct Lane): only where an attention CTA and a post_attn CTA fit one SM together (C = 160)
        const char *ln = getenv("MAPF_GPT_B200_LANES");
        e->n_lanes = ln ? (atoi(ln) == 2 ? 2 : 1) : MG_LANES_DEFAULT(C);
        const char *lg = getenv("MAPF_GPT_B200_LANE_ATTN_GRID");
        e->lane_attn_grid = lg ? atoi(lg) : 0;
    }
    {
        const char *np = getenv("MAPF_GPT_B200_NO_PRUNE");
        e->prune_last = !(np && np[0] == '1');
    }
    const bool pair_gemm = !m.fused && BN == 256 && !(getenv("MAPF_GPT_B200_GEMM_PAIR") && getenv("MAPF_GPT_B200_GEMM_PAIR")[0] == '0');
    // LayerNorm folded into the CTA-pair GEMMs (no LayerNorm kernel, no normalised copy of the residual in HBM): needs every GEMM
    // of the block on the pair kernel.  MAPF_GPT_B200_LN_FUSED=0 keeps the separate one-pass LayerNorm kernel (A/B, fallback).
    m.ln_fused = pair_gemm && C % 128 == 0 && !(getenv("MAPF_GPT_B200_LN_FUSED") && getenv("MAPF_GPT_B200_LN_FUSED")[0] == '0');
    auto bf_round = [](float f) { const uint32_t u = (uint32_t)f2bf(f) << 16; float r; memcpy(&r, &u, 4); return r; };
    // rows of W (N x K) scaled per column by the LayerNorm gain; returns the column sums of the bf16-ROUNDED result (what the
    // tensor core multiplies), so that rstd * (x W^T - mean * colsum) is exact for a constant row
    auto fold_gain = [&](std::vector<float> &W, int N, int K, const float *gain, float **cs_dev) -> int {
        std::vector<float> cs((size_t)N);
        for (int n = 0; n < N; n++) {
            double acc = 0.0;
            for (int k = 0; k < K; k++) {
                float &v = W[(size_t)n * K + k];
                v *= gain[k];
                acc += (double)bf_round(v);
            }
            cs[n] = (float)acc;
        }
        return upload_f32(cs.data(), cs.size(), cs_dev);
    };
    // attention scale folded into Wq: q' = (log2(e) / sqrt(hs)) q, so S = Q' K^T is already the exponent of 2 (gpt_kernels.cuh,
    // "max-free softmax"); the fold happens in fp32, before the one bf16 rounding of the weight
    m.q_fold = (float)(1.4426950408889634 / std::sqrt((double)hs));
    std::vector<float> wq_s(3 * CC);
    for (auto &L : m.layers) {
        if ((rc = upload_f32(w, C, &L.ln1))) return rc;
        w += C;
        for (size_t i = 0; i < 3 * CC; i++) wq_s[i] = i < CC ? w[i] * m.q_fold : w[i];
        if ((rc = upload_packed(wq_s.data(), 3 * C, C, BN, &L.wqkv))) return rc;
        if (m.ln_fused && (rc = fold_gain(wq_s, 3 * C, C, w - C, &L.cs_qkv))) return rc;     // w - C: this block's ln_1 gain
        if (pair_gemm && (rc = upload_packed_pair(wq_s.data(), 3 * C, C, BN, &L.wqkv_p))) return rc;
        w += 3 * CC;
        const float *wproj = w;
        if ((rc = upload_packed(w, C, C, BN, &L.wproj))) return rc;
        if (pair_gemm && (rc = upload_packed_pair(w, C, C, BN, &L.wproj_p))) return rc;
        w += CC;
        const float *g2 = w;
        if ((rc = upload_f32(w, C, &L.ln2))) return rc;
        w += C;
        const float *wfc = w;
        if ((rc = upload_packed(w, 4 * C, C, BN, &L.wfc))) return rc;
        if (m.ln_fused) {   // ln_2's gain folded into the c_fc columns of the pair packing
            std::vector<float> wfc_s(w, w + 4 * CC);
            if ((rc = fold_gain(wfc_s, 4 * C, C, g2, &L.cs_fc))) return rc;
            if ((rc = upload_packed_pair(wfc_s.data(), 4 * C, C, BN, &L.wfc_p))) return rc;
        } else if (pair_gemm && (rc = upload_packed_pair(w, 4 * C, C, BN, &L.wfc_p))) return rc;
        w += 4 * CC;
        const float *wproj2 = w;
        if ((rc = upload_packed(w, C, 4 * C, BN, &L.wproj2))) return rc;
        if (pair_gemm && (rc = upload_packed_pair(w, C, 4 * C, BN, &L.wproj2_p))) return rc;
        w += 4 * CC;
        if (m.fused) {
            // the fused kernel of block l also computes block l+1's c_attn; its weights sit one block further in the buffer
            const bool has_next = (&L != &m.layers.back());
            const float *wqkv_next = has_next ? w + C : nullptr;     // skip ln_1[C] of the next block
            m.fuse_qkv = getenv("MAPF_GPT_B200_NO_QKV_FUSION") == nullptr;
            const float *gn = has_next ? w : nullptr;                 // ln_1 gain of the next block
            if ((rc = upload_post_attn_stream(wproj, wfc, wproj2, m.fuse_qkv ? wqkv_next : nullptr, g2, gn, C, 1, m.q_fold, &L.wstream))) return rc;
            if ((rc = upload_post_attn_stream(wproj, wfc, wproj2, m.fuse_qkv ? wqkv_next : nullptr, g2, gn, C, 2, m.q_fold, &L.wstream_pair))) return rc;
        }
    }
    if ((rc = upload_f32(w, C, &m.lnf))) return rc;
    if (m.fused && m.fuse_qkv && getenv("MAPF_GPT_B200_NO_BLOCK0_TABLE") == nullptr) {
        // block 0 as a lookup (fused_kernels.cuh): run the production embedding + QKV kernels on 67 synthetic sequences
        const int nseq = 67, M = nseq * 256, nrec = C / 4 + 3 * C / 8;
        std::vector<uint8_t> tk((size_t)M);
        for (int i = 0; i < nseq; i++) memset(tk.data() + (size_t)i * 256, i, 256);
        uint8_t *d_tk = nullptr;
        float *X = nullptr;
        __nv_bfloat16 *XN = nullptr, *QKV = nullptr;
        CU(dalloc(&d_tk, (size_t)M));
        CU(dalloc(&X, (size_t)M * C));
        CU(dalloc(&XN, (size_t)M * C));
        CU(dalloc(&QKV, (size_t)M * 3 * C));
        CU(dalloc(&m.tab0, (size_t)M * nrec));
        CU(cudaMemcpyAsync(d_tk, tk.data(), tk.size(), cudaMemcpyHostToDevice, e->stream));
        embed_ln_kernel<<<M / 128, 128, 67 * (C + 4) * 4, e->stream>>>(d_tk, m.wte, m.wpe_ti, m.layers[0].ln1, X, XN, C);
        GemmArgs g{};
        g.A = XN; g.W = m.layers[0].wqkv; g.out = QKV; g.M = M; g.N = 3 * C; g.K = C; g.C = C; g.n_head = cfg->n_head; g.hs = m.hs;
        if ((rc = launch_gemm<EPI_QKV>(e, m.BN, g, KC_QKV))) return rc;
        block0_table_kernel<<<M / 128, 128, 0, e->stream>>>(X, QKV, m.tab0, C);
        CU(cudaStreamSynchronize(e->stream));
        CU(cudaGetLastError());
        cudaFree(d_tk); cudaFree(X); cudaFree(XN); cudaFree(QKV);
    }
    m.loaded = true;
    return MG_OK;
}

int mg_engine_num_envs(const mg_engine *e) { return e ? e->n_envs : 0; }

int mg_engine_clear(mg_engine *e)
{
    if (!e) return fail(MG_ERR_ARG, "null engine");
    CU(cudaSetDevice(e->device));
    EnvState &s = e->s;
    CU(cudaStreamSynchronize(e->stream));
    CU(cudaMemsetAsync(s.nag, 0, (size_t)s.E * 4, e->stream));
    CU(cudaMemsetAsync(s.active, 1, (size_t)s.E, e->stream));
    CU(cudaMemsetAsync(s.vocab_err, 0, 8, e->stream));
    CU(cudaStreamSynchronize(e->stream));
    std::fill(e->h_nag.begin(), e->h_nag.end(), 0);
    e->n_envs = 0;
    return MG_OK;
}

int mg_engine_reset(mg_engine *e, int first_env, int n_envs, int n_agents, const uint8_t *obstacles,
                    const int32_t *pos_xy, const int32_t *goal_xy)
{
    if (!e || !obstacles || !pos_xy || !goal_xy) return fail(MG_ERR_ARG, "null argument");
    EnvState &s = e->s;
    if (first_env < 0 || n_envs < 1 || first_env + n_envs > s.E) return fail(MG_ERR_ARG, "env range outside capacity %d", s.E);
    if (n_agents < 1 || n_agents > s.N) return fail(MG_ERR_ARG, "n_agents %d outside capacity %d", n_agents, s.N);
    CU(cudaSetDevice(e->device));
    const size_t cells = (size_t)s.H * s.P;
    const size_t EN = (size_t)n_envs * s.N;
    std::vector<uint8_t> ob((size_t)n_envs * cells, 1);
    std::vector<int16_t> loc((size_t)n_envs * cells, -1);
    std::vector<short2> pos(EN, make_short2(0, 0)), goal(EN, make_short2(0, 0));
    std::vector<int32_t> arrive(EN, -1);
    for (int k = 0; k < n_envs; k++) {
        for (int i = 0; i < s.H; i++)
            for (int j = 0; j < s.W; j++)
                ob[k * cells + (size_t)i * s.P + j] = obstacles[((size_t)k * s.H + i) * s.W + j] ? 1 : 0;
        for (int a = 0; a < n_agents; a++) {
            const int32_t *p = pos_xy + ((size_t)k * n_agents + a) * 2, *g = goal_xy + ((size_t)k * n_agents + a) * 2;
            // the tokenizer reads pos +- 5 without bounds checks (cpp:492-495): enforce the padding contract
            if (p[0] < 5 || p[1] < 5 || p[0] >= s.H - 5 || p[1] >= s.W - 5 || g[0] < 0 || g[1] < 0 || g[0] >= s.H || g[1] >= s.W)
                return fail(MG_ERR_ARG, "env %d agent %d: position (%d,%d) / goal (%d,%d) outside the padded grid %dx%d",
                            first_env + k, a, p[0], p[1], g[0], g[1], s.H, s.W);
            pos[(size_t)k * s.N + a] = make_short2((short)p[0], (short)p[1]);
            goal[(size_t)k * s.N + a] = make_short2((short)g[0], (short)g[1]);
            loc[k * cells + (size_t)p[0] * s.P + p[1]] = (int16_t)a;
            arrive[(size_t)k * s.N + a] = (p[0] == g[0] && p[1] == g[1]) ? 0 : -1;
        }
    }
    std::vector<int32_t> nag(n_envs, n_agents), map_of(n_envs, 0);
    if (s.large)
        for (int k = 0; k < n_envs; k++) {
            int rc = add_large_map(e, ob.data() + (size_t)k * cells, &map_of[k]);
            if (rc) return rc;
        }
    const size_t eo = (size_t)first_env;
    CU(cudaMemcpyAsync(s.map_of_env + eo, map_of.data(), n_envs * 4, cudaMemcpyHostToDevice, e->stream));
    CU(cudaMemcpyAsync(s.obst + eo * cells, ob.data(), ob.size(), cudaMemcpyHostToDevice, e->stream));
    CU(cudaMemcpyAsync(s.loc + eo * cells, loc.data(), loc.size() * 2, cudaMemcpyHostToDevice, e->stream));
    CU(cudaMemcpyAsync(s.pos + eo * s.N, pos.data(), EN * sizeof(short2), cudaMemcpyHostToDevice, e->stream));
    CU(cudaMemcpyAsync(s.goal + eo * s.N, goal.data(), EN * sizeof(short2), cudaMemcpyHostToDevice, e->stream));
    CU(cudaMemcpyAsync(s.arrive + eo * s.N, arrive.data(), EN * 4, cudaMemcpyHostToDevice, e->stream));
    CU(cudaMemcpyAsync(s.nag + eo, nag.data(), n_envs * 4, cudaMemcpyHostToDevice, e->stream));
    CU(cudaMemsetAsync(s.hist + eo * s.N * 8, 44, EN * 8, e->stream));            // "n" x 5 (cpp:402-405)
    CU(cudaMemsetAsync(s.nextb + eo * s.N, 0, EN, e->stream));
    CU(cudaMemsetAsync(s.act + eo * s.N, 0xFF, EN * 4, e->stream));               // -1 (inference.py:140)
    CU(cudaMemsetAsync(s.dirty + eo * s.N, 0, EN, e->stream));
    CU(cudaMemsetAsync(s.steps + eo, 0, n_envs * 4, e->stream));
    CU(cudaMemsetAsync(s.done + eo, 0, n_envs, e->stream));
    CU(cudaMemsetAsync(s.agent_steps + eo, 0, n_envs * 8, e->stream));
    CU(cudaMemsetAsync(s.density_sum + eo, 0, n_envs * 4, e->stream));
    CU(cudaMemsetAsync(s.density_n + eo, 0, n_envs * 4, e->stream));
    for (int k = 0; k < n_envs; k++) e->h_nag[first_env + k] = n_agents;
    CU(cudaStreamSynchronize(e->stream));  // host vectors go out of scope
    e->n_envs = std::max(e->n_envs, first_env + n_envs);
    return launch_fields(e, first_env, n_envs, 0);
}

int mg_engine_update_agents(mg_engine *e, const int32_t *pos_xy, const int32_t *goal_xy, const int32_t *actions)
{
    if (!e) return fail(MG_ERR_ARG, "null engine");
    if (e->n_envs == 0) return fail(MG_ERR_STATE, "no environment has been reset");
    CU(cudaSetDevice(e->device));
    EnvState &s = e->s;
    const size_t EN = (size_t)e->n_envs * s.N;
    if (int rc = check_xy(e, pos_xy, goal_xy)) return rc;
    if (pos_xy) CU(cudaMemcpyAsync(e->d_pos_in, pos_xy, EN * 8, cudaMemcpyHostToDevice, e->stream));
    if (goal_xy) CU(cudaMemcpyAsync(e->d_goal_in, goal_xy, EN * 8, cudaMemcpyHostToDevice, e->stream));
    if (actions) CU(cudaMemcpyAsync(e->d_act_in, actions, EN * 4, cudaMemcpyHostToDevice, e->stream));
    if (pos_xy || goal_xy || actions) {
        e->launches++;
        set_state_kernel<<<e->n_envs, 256, 0, e->stream>>>(s, pos_xy ? e->d_pos_in : nullptr, goal_xy ? e->d_goal_in : nullptr,
                                                           actions ? e->d_act_in : nullptr);
    }
    if (s.large && !pos_xy && !goal_xy) {   // positions moved on the device (mg_engine_env_step) may have left their windows
        int rc = launch_fields(e, 0, e->n_envs, 1);
        if (rc) return rc;
    }
    if (goal_xy || (pos_xy && s.large)) {  // changed goals / FOV left the window -> recompute those fields (cpp:464-481)
        int rc = launch_fields(e, 0, e->n_envs, 1);
        if (rc) return rc;
    }
    return launch_observe(e, true, false);
}

int mg_engine_generate_observations(mg_engine *e, int8_t *out_tokens)
{
    if (!e) return fail(MG_ERR_ARG, "null engine");
    CU(cudaSetDevice(e->device));
    int rc = launch_observe(e, false, true);
    if (rc) return rc;
    if (out_tokens)
        CU(cudaMemcpyAsync(out_tokens, e->s.tokens, (size_t)e->n_envs * e->s.N * 256, cudaMemcpyDeviceToHost, e->stream));
    return check_vocab(e);
}

int mg_engine_set_active(mg_engine *e, const uint8_t *mask)
{
    if (!e) return fail(MG_ERR_ARG, "null engine");
    CU(cudaSetDevice(e->device));
    if (mask) CU(cudaMemcpyAsync(e->s.active, mask, (size_t)e->n_envs, cudaMemcpyHostToDevice, e->stream));
    else CU(cudaMemsetAsync(e->s.active, 1, (size_t)e->s.E, e->stream));
    CU(cudaStreamSynchronize(e->stream));
    return MG_OK;
}

int mg_engine_set_seed(mg_engine *e, uint64_t seed)
{
    if (!e) return fail(MG_ERR_ARG, "null engine");
    e->seed = seed;
    return MG_OK;
}

int mg_engine_act(mg_engine *e, int mode, const float *q_exp, int32_t *actions_out, float *logits_out)
{
    if (!e) return fail(MG_ERR_ARG, "null engine");
    if (mode < 0 || mode > 2) return fail(MG_ERR_ARG, "mode must be 0 (greedy), 1 (philox) or 2 (supplied q)");
    if (mode == 2 && !q_exp) return fail(MG_ERR_ARG, "mode 2 needs q_exp");
    if (e->n_envs == 0) return fail(MG_ERR_STATE, "no environment has been reset");
    CU(cudaSetDevice(e->device));
    EnvState &s = e->s;
    const size_t EN = (size_t)e->n_envs * s.N;
    if (mode == 2) CU(cudaMemcpyAsync(e->d_q, q_exp, EN * 5 * 4, cudaMemcpyHostToDevice, e->stream));
    int rc;
    for (int attempt = 0;; attempt++) {
        if ((rc = forward_device(e, s.tokens, (int)EN, s.logits))) return rc;
        if ((rc = launch_step(e, mode, 0, e->d_q, nullptr))) return rc;
        int vocab = 0, numeric = 0;
        if ((rc = read_flags(e, &vocab, &numeric))) return rc;
        if (!numeric) break;
        if (attempt == 0 && !e->safe_softmax && !e->precise) { e->safe_softmax = true; continue; }   // redo with the max-subtracting kernels
        return numeric_failure(e);
    }
    if (actions_out) CU(cudaMemcpyAsync(actions_out, s.act, EN * 4, cudaMemcpyDeviceToHost, e->stream));
    if (logits_out) {
        std::vector<float> tmp(EN * 8);
        CU(cudaMemcpyAsync(tmp.data(), s.logits, EN * 8 * 4, cudaMemcpyDeviceToHost, e->stream));
        CU(cudaStreamSynchronize(e->stream));
        for (size_t i = 0; i < EN; i++) memcpy(logits_out + i * 5, tmp.data() + i * 8, 20);
    }
    CU(cudaStreamSynchronize(e->stream));
    if (e->profiling) prof_collect(e);
    return MG_OK;
}

int mg_engine_forward_tokens(mg_engine *e, const int8_t *tokens, int n_rows, float *logits_out)
{
    if (!e || !tokens || !logits_out || n_rows < 1) return fail(MG_ERR_ARG, "bad argument");
    for (size_t i = 0; i < (size_t)n_rows * 256; i++)   // nn.Embedding raises IndexError on ids outside the table (model.py:174)
        if (tokens[i] < 0 || tokens[i] >= MG_VOCAB)
            return fail(MG_ERR_VOCAB, "row %zu token %zu: id %d outside the vocabulary [0, %d)", i / 256, i % 256, (int)tokens[i], MG_VOCAB);
    CU(cudaSetDevice(e->device));
    Workspace &w = e->ws;
    cudaFree(w.tok); cudaFree(w.logits);
    w.tok = nullptr; w.logits = nullptr;
    CU(dalloc(&w.tok, (size_t)n_rows * 256));
    CU(dalloc(&w.logits, (size_t)n_rows * 8));
    CU(cudaMemcpyAsync(w.tok, tokens, (size_t)n_rows * 256, cudaMemcpyHostToDevice, e->stream));
    int rc = forward_device(e, w.tok, n_rows, w.logits);
    if (rc) return rc;
    std::vector<float> tmp((size_t)n_rows * 8);
    CU(cudaMemcpyAsync(tmp.data(), w.logits, tmp.size() * 4, cudaMemcpyDeviceToHost, e->stream));
    CU(cudaStreamSynchronize(e->stream));
    for (int i = 0; i < n_rows; i++) memcpy(logits_out + (size_t)i * 5, tmp.data() + (size_t)i * 8, 20);
    if (e->profiling) prof_collect(e);
    return MG_OK;
}

int mg_engine_eval_tokens(mg_engine *e, const int8_t *tokens, const int8_t *targets, int n_rows, float *loss_out, int32_t *pred_out)
{
    if (!e || !tokens || !targets || !loss_out || !pred_out || n_rows < 1) return fail(MG_ERR_ARG, "bad argument");
    for (size_t i = 0; i < (size_t)n_rows * 256; i++)
        if (tokens[i] < 0 || tokens[i] >= MG_VOCAB)
            return fail(MG_ERR_VOCAB, "row %zu token %zu: id %d outside the vocabulary [0, %d)", i / 256, i % 256, (int)tokens[i], MG_VOCAB);
    for (int i = 0; i < n_rows; i++)
        if (targets[i] < -1 || targets[i] >= MG_VOCAB) return fail(MG_ERR_VOCAB, "row %d: target %d outside [-1, %d)", i, (int)targets[i], MG_VOCAB);
    CU(cudaSetDevice(e->device));
    Workspace &w = e->ws;
    cudaFree(w.tok); cudaFree(w.logits);
    w.tok = nullptr; w.logits = nullptr;
    CU(dalloc(&w.tok, (size_t)n_rows * 256));
    CU(dalloc(&w.logits, (size_t)n_rows * 8));
    int8_t *d_tgt = nullptr;
    float *d_loss = nullptr;
    int32_t *d_pred = nullptr;
    CU(dalloc(&d_tgt, (size_t)n_rows));
    CU(dalloc(&d_loss, (size_t)n_rows));
    CU(dalloc(&d_pred, (size_t)n_rows));
    CU(cudaMemcpyAsync(w.tok, tokens, (size_t)n_rows * 256, cudaMemcpyHostToDevice, e->stream));
    CU(cudaMemcpyAsync(d_tgt, targets, (size_t)n_rows, cudaMemcpyHostToDevice, e->stream));
    EvalOut ev;
    ev.targets = d_tgt; ev.loss = d_loss; ev.pred = d_pred;
    int rc = forward_device(e, w.tok, n_rows, w.logits, &ev);
    if (!rc) {
        cudaMemcpyAsync(loss_out, d_loss, (size_t)n_rows * 4, cudaMemcpyDeviceToHost, e->stream);
        cudaMemcpyAsync(pred_out, d_pred, (size_t)n_rows * 4, cudaMemcpyDeviceToHost, e->stream);
    }
    cudaError_t err = cudaStreamSynchronize(e->stream);
    cudaFree(d_tgt); cudaFree(d_loss); cudaFree(d_pred);
    if (rc) return rc;
    if (err != cudaSuccess) return fail(MG_ERR_CUDA, "eval_tokens: %s", cudaGetErrorString(err));
    return MG_OK;
}

int mg_engine_env_step(mg_engine *e, const int32_t *actions, int32_t *pos_out)
{
    if (!e) return fail(MG_ERR_ARG, "null engine");
    if (e->n_envs == 0) return fail(MG_ERR_STATE, "no environment has been reset");
    CU(cudaSetDevice(e->device));
    EnvState &s = e->s;
    const size_t EN = (size_t)e->n_envs * s.N;
    if (actions) CU(cudaMemcpyAsync(e->d_step_act, actions, EN * 4, cudaMemcpyHostToDevice, e->stream));
    int rc = launch_step(e, 3, 1, nullptr, actions ? e->d_step_act : nullptr);
    if (rc) return rc;
    if (pos_out) return mg_engine_get_positions(e, pos_out);
    return MG_OK;
}

int mg_engine_rollout(mg_engine *e, int n_steps, int mode)
{
    if (!e) return fail(MG_ERR_ARG, "null engine");
    if (mode < 0 || mode > 1) return fail(MG_ERR_ARG, "rollout mode must be 0 (greedy) or 1 (philox sampling)");
    if (e->n_envs == 0) return fail(MG_ERR_STATE, "no environment has been reset");
    CU(cudaSetDevice(e->device));
    EnvState &s = e->s;
    const int EN = e->n_envs * s.N;
    CU(cudaEventRecord(e->ev_t0, e->stream));
    for (int t = 0; t < n_steps; t++) {
        int rc;
        const bool ph = e->profiling && t == n_steps - 1;
        if (ph) cudaEventRecord(e->ev_p[0], e->stream);
        if (s.large && (rc = launch_fields(e, 0, e->n_envs, 1))) return rc;   // agents whose FOV left their window
        if ((rc = launch_observe(e, true, true))) return rc;
        if (ph) cudaEventRecord(e->ev_p[1], e->stream);
        if ((rc = forward_device(e, s.tokens, EN, s.logits))) return rc;
        if (ph) cudaEventRecord(e->ev_p[2], e->stream);
        if ((rc = launch_step(e, mode, 1, nullptr, nullptr))) return rc;
        if (ph) cudaEventRecord(e->ev_p[3], e->stream);
    }
    CU(cudaEventRecord(e->ev_t1, e->stream));
    return MG_OK;
}

int mg_engine_act_host(mg_engine *e, const int32_t *pos_xy, const int32_t *goal_xy, int mode, const float *q_exp,
                       int32_t *actions_out)
{
    if (!e || !actions_out) return fail(MG_ERR_ARG, "null argument");
    if (mode < 0 || mode > 2) return fail(MG_ERR_ARG, "bad mode");
    if (mode == 2 && !q_exp) return fail(MG_ERR_ARG, "mode 2 needs q_exp");
    if (e->n_envs == 0) return fail(MG_ERR_STATE, "no environment has been reset");
    CU(cudaSetDevice(e->device));
    EnvState &s = e->s;
    const size_t EN = (size_t)e->n_envs * s.N;
    if (int rc0 = check_xy(e, pos_xy, goal_xy)) return rc0;
    CU(cudaEventRecord(e->ev_t0, e->stream));
    if (pos_xy) CU(cudaMemcpyAsync(e->d_pos_in, pos_xy, EN * 8, cudaMemcpyHostToDevice, e->stream));
    if (goal_xy) CU(cudaMemcpyAsync(e->d_goal_in, goal_xy, EN * 8, cudaMemcpyHostToDevice, e->stream));
    if (mode == 2) CU(cudaMemcpyAsync(e->d_q, q_exp, EN * 5 * 4, cudaMemcpyHostToDevice, e->stream));
    if (pos_xy || goal_xy) {
        e->launches++;
        set_state_kernel<<<e->n_envs, 256, 0, e->stream>>>(s, pos_xy ? e->d_pos_in : nullptr,
                                                           goal_xy ? e->d_goal_in : nullptr, nullptr);
    }
    int rc;
    // large maps: agents moved on the device (mg_engine_env_step) may have left their windows even when no host array is given
    if ((goal_xy || s.large) && (rc = launch_fields(e, 0, e->n_envs, 1))) return rc;
    if ((rc = launch_observe(e, true, true))) return rc;
    int vocab = 0, numeric = 0;
    for (int attempt = 0;; attempt++) {
        if ((rc = forward_device(e, s.tokens, (int)EN, s.logits))) return rc;
        if ((rc = launch_step(e, mode, 0, e->d_q, nullptr))) return rc;
        CU(cudaMemcpyAsync(actions_out, s.act, EN * 4, cudaMemcpyDeviceToHost, e->stream));
        CU(cudaEventRecord(e->ev_t1, e->stream));
        if ((rc = read_flags(e, &vocab, &numeric))) return rc;
        if (!numeric) break;
        if (attempt == 0 && !e->safe_softmax && !e->precise) { e->safe_softmax = true; continue; }   // redo with the max-subtracting kernels
        return numeric_failure(e);
    }
    if (e->profiling) prof_collect(e);
    if (vocab) return fail(MG_ERR_VOCAB, "a relative position left the token vocabulary");
    return MG_OK;
}

int mg_engine_get_positions(mg_engine *e, int32_t *out)
{
    if (!e || !out) return fail(MG_ERR_ARG, "null argument");
    CU(cudaSetDevice(e->device));
    const size_t EN = (size_t)e->n_envs * e->s.N;
    std::vector<short2> p(EN);
    CU(cudaMemcpyAsync(p.data(), e->s.pos, EN * sizeof(short2), cudaMemcpyDeviceToHost, e->stream));
    CU(cudaStreamSynchronize(e->stream));
    for (size_t i = 0; i < EN; i++) { out[2 * i] = p[i].x; out[2 * i + 1] = p[i].y; }
    return MG_OK;
}

int mg_engine_get_tokens(mg_engine *e, int8_t *out)
{
    if (!e || !out) return fail(MG_ERR_ARG, "null argument");
    CU(cudaSetDevice(e->device));
    CU(cudaMemcpyAsync(out, e->s.tokens, (size_t)e->n_envs * e->s.N * 256, cudaMemcpyDeviceToHost, e->stream));
    CU(cudaStreamSynchronize(e->stream));
    return MG_OK;
}

int mg_engine_get_cost2go(mg_engine *e, int env, int agent, uint16_t *out_hw)
{
    if (!e || !out_hw) return fail(MG_ERR_ARG, "null argument");
    EnvState &s = e->s;
    if (env < 0 || env >= e->n_envs || agent < 0 || agent >= s.N) return fail(MG_ERR_ARG, "index out of range");
    CU(cudaSetDevice(e->device));
    if (s.large) return fail(MG_ERR_ARG, "large maps hold windowed fields: use mg_engine_get_partial");
    const size_t cells = (size_t)s.H * s.P;
    std::vector<uint16_t> t(cells);
    CU(cudaMemcpyAsync(t.data(), s.c2g + ((size_t)env * s.N + agent) * cells, cells * 2, cudaMemcpyDeviceToHost, e->stream));
    CU(cudaStreamSynchronize(e->stream));
    for (int i = 0; i < s.H; i++) memcpy(out_hw + (size_t)i * s.W, t.data() + (size_t)i * s.P, s.W * 2);
    return MG_OK;
}

int mg_engine_get_partial(mg_engine *e, int env, int agent, int32_t *bounds4, uint16_t *out, int cap)
{
    if (!e || !bounds4) return fail(MG_ERR_ARG, "null argument");
    EnvState &s = e->s;
    if (env < 0 || env >= e->n_envs || agent < 0 || agent >= s.N) return fail(MG_ERR_ARG, "index out of range");
    CU(cudaSetDevice(e->device));
    short4 b;
    CU(cudaMemcpyAsync(&b, s.bounds + (size_t)env * s.N + agent, sizeof b, cudaMemcpyDeviceToHost, e->stream));
    std::vector<uint16_t> t((size_t)s.FR * s.FP);
    CU(cudaMemcpyAsync(t.data(), s.c2g + ((size_t)env * s.N + agent) * s.FR * s.FP, t.size() * 2, cudaMemcpyDeviceToHost, e->stream));
    CU(cudaStreamSynchronize(e->stream));
    bounds4[0] = b.x; bounds4[1] = b.y; bounds4[2] = b.z; bounds4[3] = b.w;
    const int rows = b.y - b.x + 1, cols = b.w - b.z + 1;
    if (out) {
        if (cap < rows * cols) return fail(MG_ERR_ARG, "buffer too small for a %dx%d window", rows, cols);
        for (int i = 0; i < rows; i++) memcpy(out + (size_t)i * cols, t.data() + (size_t)i * s.FP, (size_t)cols * 2);
    }
    return rows * cols;
}

int mg_engine_get_metrics(mg_engine *e, double *out)
{
    if (!e || !out) return fail(MG_ERR_ARG, "null argument");
    if (e->n_envs == 0) return fail(MG_ERR_STATE, "no environment has been reset");
    CU(cudaSetDevice(e->device));
    e->launches++;
    metrics_kernel<<<(e->s.E + 127) / 128, 128, 0, e->stream>>>(e->s, e->d_metrics);
    CU(cudaMemcpyAsync(out, e->d_metrics, (size_t)e->n_envs * MG_METRIC_COLS * 8, cudaMemcpyDeviceToHost, e->stream));
    CU(cudaStreamSynchronize(e->stream));
    return MG_OK;
}

int mg_engine_synchronize(mg_engine *e)
{
    if (!e) return fail(MG_ERR_ARG, "null engine");
    CU(cudaSetDevice(e->device));
    int vocab = 0, numeric = 0, rc;
    if ((rc = read_flags(e, &vocab, &numeric))) return rc;
    if (e->profiling) prof_collect(e);
    if (numeric) return numeric_failure(e);
    if (vocab) return fail(MG_ERR_VOCAB, "a relative position left the token vocabulary");
    return MG_OK;
}

int mg_engine_set_profiling(mg_engine *e, int on)
{
    if (!e) return fail(MG_ERR_ARG, "null engine");
    e->profiling = on != 0;
    if (on) { memset(e->kc_ms, 0, sizeof e->kc_ms); memset(e->kc_n, 0, sizeof e->kc_n); e->ev_used = 0; }
    return MG_OK;
}

int mg_engine_last_timing(mg_engine *e, float *total_ms, float *phases_ms3)
{
    if (!e) return fail(MG_ERR_ARG, "null engine");
    CU(cudaSetDevice(e->device));
    CU(cudaStreamSynchronize(e->stream));
    if (total_ms) CU(cudaEventElapsedTime(total_ms, e->ev_t0, e->ev_t1));
    if (phases_ms3) {
        for (int k = 0; k < 3; k++) {
            phases_ms3[k] = 0.f;
            if (e->profiling) cudaEventElapsedTime(&phases_ms3[k], e->ev_p[k], e->ev_p[k + 1]);
        }
        cudaGetLastError();
    }
    return MG_OK;
}

long long mg_engine_launch_count(const mg_engine *e) { return e ? e->launches : 0; }

int mg_engine_num_lanes(const mg_engine *e) { return e ? e->n_lanes : 0; }
int mg_engine_kernel_times(mg_engine *e, float *ms_out, int n)
{   // [2*k] = total ms, [2*k+1] = launches, k in KernelClass order
    if (!e || !ms_out) return fail(MG_ERR_ARG, "null argument");
    for (int k = 0; k < KC_COUNT && 2 * k + 1 < n; k++) { ms_out[2 * k] = e->kc_ms[k]; ms_out[2 * k + 1] = (float)e->kc_n[k]; }
    return KC_COUNT;
}

// ---- engine knobs not in the reference surface
int mg_engine_set_env_offset(mg_engine *e, int off) { if (!e) return MG_ERR_ARG; e->env_offset = off; return MG_OK; }
int mg_engine_set_max_episode_steps(mg_engine *e, int n) { if (!e) return MG_ERR_ARG; e->max_episode_steps = n; return MG_OK; }

// ------------------------------------------------------------------------------------------- single-env twin
struct mg_gen {
    int H, W;
    std::vector<uint8_t> grid;
    mg_params p;
    mg_engine *e = nullptr;
    int n = 0;
    int device = 0;
};

mg_gen *mg_gen_create(const int32_t *grid, int H, int W, const mg_params *params)
{
    if (!grid) { fail(MG_ERR_ARG, "null grid"); return nullptr; }
    mg_gen *g = new mg_gen();
    g->H = H; g->W = W;
    g->grid.resize((size_t)H * W);
    for (size_t i = 0; i < g->grid.size(); i++) g->grid[i] = grid[i] != 0;
    mg_default_params(&g->p);
    if (params) g->p = *params;
    cudaGetDevice(&g->device);
    return g;
}
int mg_gen_create_agents(mg_gen *g, const int32_t *pos_xy, const int32_t *goal_xy, int n)
{
    if (!g) return fail(MG_ERR_ARG, "null generator");
    if (g->e) { mg_engine_destroy(g->e); g->e = nullptr; }
    g->e = mg_engine_create(g->device, 1, n, g->H, g->W, &g->p);
    if (!g->e) return MG_ERR_ARG;
    g->n = n;
    return mg_engine_reset(g->e, 0, 1, n, g->grid.data(), pos_xy, goal_xy);
}
int mg_gen_update_agents(mg_gen *g, const int32_t *pos_xy, const int32_t *goal_xy, const int32_t *actions, int n)
{
    if (!g || !g->e) return fail(MG_ERR_STATE, "create_agents has not been called");
    if (n != g->n) return fail(MG_ERR_ARG, "agent count changed (%d != %d)", n, g->n);
    return mg_engine_update_agents(g->e, pos_xy, goal_xy, actions);
}
int mg_gen_generate_observations(mg_gen *g, int32_t *out)
{
    if (!g || !g->e) return fail(MG_ERR_STATE, "create_agents has not been called");
    std::vector<int8_t> t((size_t)g->n * 256);
    int rc = mg_engine_generate_observations(g->e, t.data());
    if (rc) return rc;
    for (size_t i = 0; i < t.size(); i++) out[i] = t[i];
    return MG_OK;
}
void mg_gen_destroy(mg_gen *g)
{
    if (!g) return;
    if (g->e) mg_engine_destroy(g->e);
    delete g;
}

// ------------------------------------------------------------------------------------------- test hooks
// clock64() stamps of the first 4 CTAs of the LAST post_attn launch: out[4][128] (see MG_STAMP ids in fused_kernels.cuh)
int mg_test_timeline(mg_engine *e, int enable, long long *out)
{
    if (!e) return fail(MG_ERR_ARG, "null engine");
    CU(cudaSetDevice(e->device));
    if (enable && !e->d_timeline) {
        CU(dalloc(&e->d_timeline, 17 * 128));
        CU(cudaMemset(e->d_timeline, 0, 17 * 128 * 8));
    }
    if (out && e->d_timeline) {
        CU(cudaStreamSynchronize(e->stream));
        CU(cudaMemcpy(out, e->d_timeline, 17 * 128 * 8, cudaMemcpyDeviceToHost));
    }
    if (!enable && e->d_timeline) { cudaFree(e->d_timeline); e->d_timeline = nullptr; }
    return MG_OK;
}

int mg_test_gemm(int device, const void *A, const void *B, float *C, int M, int N, int K, int variant)
{
    CU(cudaSetDevice(device));
    const int cfg = variant & 0xF;
    const int BN = cfg == 0 ? 160 : (cfg == 1 ? 256 : 128);
    if (M % 128 || N % BN || ((K % 64) && !(cfg == 0 && K % 32 == 0))) return fail(MG_ERR_ARG, "test gemm: bad shape");
    __nv_bfloat16 *At = nullptr, *Bt = nullptr;
    CU(dalloc(&At, (size_t)M * K));
    CU(dalloc(&Bt, (size_t)N * K));
    const int th = 256;
    pack_rows_kernel<<<(unsigned)(((size_t)M * K + th - 1) / th), th>>>((const __nv_bfloat16 *)A, At, M, K, 128);
    pack_rows_kernel<<<(unsigned)(((size_t)N * K + th - 1) / th), th>>>((const __nv_bfloat16 *)B, Bt, N, K, BN);
    GemmArgs g{};
    g.A = At; g.W = Bt; g.out = C; g.M = M; g.N = N; g.K = K; g.dbg_swap_lbo_sbo = (variant >> 8) & 1;
    __nv_bfloat16 *Bp = nullptr;
    if (variant & 0x10) {   // the persistent CTA-pair kernel (the 85M path's GEMM): B in the pair packing
        if (cfg != 1 || (M / 128) % 2) { cudaFree(At); cudaFree(Bt); return fail(MG_ERR_ARG, "test gemm: the pair kernel needs cfg 1 and M % 256 == 0"); }
        std::vector<uint16_t> hb((size_t)N * K);
        CU(cudaMemcpy(hb.data(), B, hb.size() * 2, cudaMemcpyDeviceToHost));
        std::vector<float> hf(hb.size());
        for (size_t i = 0; i < hb.size(); i++) { const uint32_t u = (uint32_t)hb[i] << 16; memcpy(&hf[i], &u, 4); }
        if (int rc = upload_packed_pair(hf.data(), N, K, BN, &Bp)) { cudaFree(At); cudaFree(Bt); return rc; }
        g.Wp = Bp;
    }
    int rc = launch_gemm<EPI_STORE_F32>(nullptr, BN, g, 0);
    cudaError_t err = cudaDeviceSynchronize();
    cudaFree(At); cudaFree(Bt); cudaFree(Bp);
    if (rc) return rc;
    if (err != cudaSuccess) return fail(MG_ERR_CUDA, "test gemm: %s", cudaGetErrorString(err));
    return MG_OK;
}

int mg_test_umma_rate(int device, int N, int iters, int ctas, long long *cycles2)
{
    CU(cudaSetDevice(device));
    long long *d = nullptr;
    CU(dalloc(&d, 2));
    int rc = MG_ERR_ARG;
    const int ts = N >= 1000;   // N + 1000: A operand from tensor memory (TS form), N <= 160
    if (ts) N -= 1000;
    if (N == 48) rc = run_umma_rate<48>(iters, ctas, d, cycles2, ts);
    else if (N == 80) rc = run_umma_rate<80>(iters, ctas, d, cycles2, ts);
    else if (N == 128) rc = run_umma_rate<128>(iters, ctas, d, cycles2, ts);
    else if (N == 160) rc = run_umma_rate<160>(iters, ctas, d, cycles2, ts);
    else if (N == 256 && !ts) rc = run_umma_rate<256>(iters, ctas, d, cycles2, ts);
    cudaFree(d);
    return rc;
}

// UMMA-rate microbenchmark: `iters` launches of the production GEMM on uninitialised tile images; ms = average per launch
int mg_test_gemm_time(int device, int M, int N, int K, int BN, int iters, float *ms)
{
    CU(cudaSetDevice(device));
    __nv_bfloat16 *At = nullptr, *Bt = nullptr;
    float *Cd = nullptr;
    CU(dalloc(&At, (size_t)M * K));
    CU(dalloc(&Bt, (size_t)N * K));
    CU(dalloc(&Cd, (size_t)M * N));
    CU(cudaMemset(At, 0, (size_t)M * K * 2));
    CU(cudaMemset(Bt, 0, (size_t)N * K * 2));
    GemmArgs g{};
    g.A = At; g.W = Bt; g.out = Cd; g.M = M; g.N = N; g.K = K;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    int rc = launch_gemm<EPI_STORE_F32>(nullptr, BN, g, 0);
    if (!rc) {
        cudaEventRecord(e0, 0);
        for (int i = 0; i < iters && !rc; i++) rc = launch_gemm<EPI_STORE_F32>(nullptr, BN, g, 0);
        cudaEventRecord(e1, 0);
    }
    cudaError_t err = cudaDeviceSynchronize();
    if (!rc && err == cudaSuccess) { cudaEventElapsedTime(ms, e0, e1); *ms /= iters; }
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    cudaFree(At); cudaFree(Bt); cudaFree(Cd);
    if (rc) return rc;
    if (err != cudaSuccess) return fail(MG_ERR_CUDA, "gemm_time: %s", cudaGetErrorString(err));
    return MG_OK;
}

// q,k,v,out: bf16 [n_seq][n_head][256][hs] row-major
int mg_set_precision(int mode)
{
    const int prev = g_precision;
    if (mode == 0 || mode == 1) g_precision = mode;
    return prev;
}
int mg_test_attention(int device, const void *q, const void *k, const void *v, void *out, int n_seq, int n_head, int hs)
{
    return mg_test_attention_ex(device, q, k, v, out, n_seq, n_head, hs, 0);
}
int mg_test_attention_ex(int device, const void *q, const void *k, const void *v, void *out, int n_seq, int n_head, int hs, int variant)
{
    if (variant < 0 || variant > 2) return fail(MG_ERR_ARG, "test attention: variant must be 0, 1 or 2");
    CU(cudaSetDevice(device));
    const int C = n_head * hs;
    const size_t per = (size_t)n_seq * n_head * 256 * hs;
    __nv_bfloat16 *qkv = nullptr, *att = nullptr;
    CU(dalloc(&qkv, per * 3));
    CU(dalloc(&att, per));
    const int th = 256;
    // [seq][3][head][hs/8][256][8]: pack each (seq, head) block of 256 rows with tile_rows = 256
    for (int s = 0; s < n_seq; s++)
        for (int w = 0; w < 3; w++) {
            const __nv_bfloat16 *src = (const __nv_bfloat16 *)(w == 0 ? q : (w == 1 ? k : v)) + (size_t)s * n_head * 256 * hs;
            __nv_bfloat16 *dst = qkv + ((size_t)s * 3 + w) * n_head * 256 * hs;
            pack_rows_kernel<<<(unsigned)(((size_t)n_head * 256 * hs + th - 1) / th), th>>>(src, dst, n_head * 256, hs, 256);
        }
    AttnArgs a{};
    a.qkv = qkv; a.out = att; a.n_head = n_head; a.C = C;
    a.scale_log2e = (float)(1.4426950408889634 / std::sqrt((double)hs));
    if (variant == 1) {   // the engine folds the scale into Wq; here q itself is scaled (one extra bf16 rounding of q)
        for (int s = 0; s < n_seq; s++) {
            const size_t nq = (size_t)n_head * 256 * hs;
            scale_bf16_kernel<<<(unsigned)((nq + th - 1) / th), th>>>(qkv + (size_t)s * 3 * nq, nq, a.scale_log2e);
        }
        a.scale_log2e = 1.0f;
    }
    if (variant == 2) setenv("MAPF_GPT_B200_ATTN_CLASSIC", "1", 1);
    int rc = launch_attn(nullptr, a, hs, n_seq, 0, variant == 1);
    if (variant == 2) unsetenv("MAPF_GPT_B200_ATTN_CLASSIC");
    // att is A_ti [M/128][C/8][128][8] with column = head*hs + d -> out [seq][head][256][hs]
    __nv_bfloat16 *rm = nullptr;
    CU(dalloc(&rm, per));
    unpack_rows_kernel<<<(unsigned)((per + th - 1) / th), th>>>(att, rm, n_seq * 256, C, 128);
    cudaError_t err = cudaDeviceSynchronize();
    if (err == cudaSuccess) {
        // rm is [seq*256][C]; permute to [seq][head][256][hs] on the host side of the test (strided copy)
        for (int s = 0; s < n_seq && err == cudaSuccess; s++)
            for (int h = 0; h < n_head && err == cudaSuccess; h++)
                err = cudaMemcpy2D((__nv_bfloat16 *)out + ((size_t)s * n_head + h) * 256 * hs, (size_t)hs * 2,
                                   rm + (size_t)s * 256 * C + (size_t)h * hs, (size_t)C * 2, (size_t)hs * 2, 256,
                                   cudaMemcpyDeviceToDevice);
    }
    cudaFree(qkv); cudaFree(att); cudaFree(rm);
    if (rc) return rc;
    if (err != cudaSuccess) return fail(MG_ERR_CUDA, "test attention: %s", cudaGetErrorString(err));
    return MG_OK;
}

}  // extern "C"
