// precise_kernels.cuh -- fp32 verification forward (MAPF_GPT_B200_PRECISION=fp32 / mg_set_precision(1)).
//
// The production path multiplies bf16 operands on the tensor cores; its logits differ from the reference's fp32 logits
// (model.py runs SGEMM, TF32 off) by ~1e-2, so a sampled action flips whenever the decision margin of argmax(p / q) falls
// inside that error and a free-running trajectory then diverges for good.  A kind::tf32 variant (10-bit mantissa) would
// only shrink the error to ~3e-3: still several expected flips over the 4096 decisions of one 32-agent episode.  Whole-
// episode equality with the fp32 reference needs fp32 products, so the verification mode below is plain fp32 FFMA on the
// CUDA cores -- the same arithmetic as the reference up to summation order (~1e-6 relative).  It is a checker for small
// batches (config C1: 32 sequences per step), not a throughput path: row-major fp32 activations, no tile images.
//
//   p_embed_kernel   model.py:171-175      x = wte[tok] + wpe[pos]
//   p_ln_kernel      model.py:11-20        two-pass LayerNorm, gain only, eps 1e-5
//   p_gemm_kernel    model.py:50,71,85-87  C = A W^T (+ residual | exact erf GELU), 64 x 64 x 16 tiles, 4 x 4 per thread
//   p_attn_kernel    model.py:58-60        non-causal softmax(q k^T / sqrt(hs)) v, block per (sequence, head)
//   p_head_kernel    model.py:178,186      ln_f on the last token + the 5 tied lm_head rows
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace mg {

__global__ void __launch_bounds__(256) p_embed_kernel(const uint8_t *__restrict__ tokens, const float *__restrict__ wte,
                                                      const float *__restrict__ wpe, float *__restrict__ X, int C, int M)
{
    const int row = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (row >= M) return;
    const int tok = min((int)tokens[row], 66), pos = row & 255;
    for (int c = lane; c < C; c += 32) X[(size_t)row * C + c] = wte[(size_t)tok * C + c] + wpe[(size_t)pos * C + c];
}

// warp per row
__global__ void __launch_bounds__(256) p_ln_kernel(const float *__restrict__ X, const float *__restrict__ gain,
                                                   float *__restrict__ out, int C, int M)
{
    const int row = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (row >= M) return;
    const float *x = X + (size_t)row * C;
    float s = 0.f;
    for (int c = lane; c < C; c += 32) s += x[c];
#pragma unroll
    for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    const float mean = s / (float)C;
    float q = 0.f;
    for (int c = lane; c < C; c += 32) {
        const float d = x[c] - mean;
        q += d * d;
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
    const float rstd = 1.0f / sqrtf(q / (float)C + 1e-5f);
    for (int c = lane; c < C; c += 32) out[(size_t)row * C + c] = (x[c] - mean) * rstd * gain[c];
}

enum { P_EPI_STORE = 0, P_EPI_RESID = 1, P_EPI_GELU = 2 };

// C[M][N] (op)= A[M][K] * W[N][K]^T.  M % 64 == 0, N % 32 == 0 (edge columns guarded), K % 16 == 0.
template <int EPI>
__global__ void __launch_bounds__(256) p_gemm_kernel(const float *__restrict__ A, const float *__restrict__ W,
                                                     float *__restrict__ Cm, int M, int N, int K)
{
    __shared__ float As[16][64 + 4];
    __shared__ float Ws[16][64 + 4];
    const int m0 = blockIdx.y * 64, n0 = blockIdx.x * 64;
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
    float acc[4][4] = {};
    for (int k0 = 0; k0 < K; k0 += 16) {
        for (int i = threadIdx.x; i < 64 * 16; i += 256) {
            const int r = i >> 4, k = i & 15;
            As[k][r] = A[(size_t)(m0 + r) * K + k0 + k];
            Ws[k][r] = (n0 + r < N) ? W[(size_t)(n0 + r) * K + k0 + k] : 0.f;
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < 16; k++) {
            float a[4], b[4];
#pragma unroll
            for (int i = 0; i < 4; i++) { a[i] = As[k][ty * 4 + i]; b[i] = Ws[k][tx * 4 + i]; }
#pragma unroll
            for (int i = 0; i < 4; i++)
#pragma unroll
                for (int j = 0; j < 4; j++) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 4; i++)
#pragma unroll
        for (int j = 0; j < 4; j++) {
            const int m = m0 + ty * 4 + i, n = n0 + tx * 4 + j;
            if (n >= N) continue;
            float v = acc[i][j];
            if (EPI == P_EPI_RESID) v += Cm[(size_t)m * N + n];
            if (EPI == P_EPI_GELU) v = 0.5f * v * (1.0f + erff(v * 0.70710678118654752440f));   // exact erf GELU (model.py:80)
            Cm[(size_t)m * N + n] = v;
        }
}

// block per (sequence, head), 8 warps, a warp per query at a time.  K (padded pitch) and V of the head live in shared memory.
template <int HS>
__global__ void __launch_bounds__(256) p_attn_kernel(const float *__restrict__ QKV, float *__restrict__ out, int n_head, int C)
{
    extern __shared__ float p_sm[];
    float *Ks = p_sm;                       // [256][HS + 1]
    float *Vs = Ks + 256 * (HS + 1);       // [256][HS]
    float *Ps = Vs + 256 * HS;             // [8 warps][256]
    float *Qs = Ps + 8 * 256;              // [8 warps][HS]
    const int head = blockIdx.x % n_head, seq = blockIdx.x / n_head;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const float *base = QKV + (size_t)seq * 256 * 3 * C + head * HS;
    for (int i = threadIdx.x; i < 256 * HS; i += 256) {
        const int j = i / HS, d = i - j * HS;
        Ks[j * (HS + 1) + d] = base[(size_t)j * 3 * C + C + d];
        Vs[j * HS + d] = base[(size_t)j * 3 * C + 2 * C + d];
    }
    __syncthreads();
    const float scale = 1.0f / sqrtf((float)HS);
    float *p = Ps + warp * 256, *q = Qs + warp * HS;
    for (int qi = warp; qi < 256; qi += 8) {
        for (int d = lane; d < HS; d += 32) q[d] = base[(size_t)qi * 3 * C + d];
        __syncwarp();
        float s[8], mx = -INFINITY;
#pragma unroll
        for (int i = 0; i < 8; i++) {
            const int j = lane + 32 * i;
            float acc = 0.f;
#pragma unroll 8
            for (int d = 0; d < HS; d++) acc = fmaf(q[d], Ks[j * (HS + 1) + d], acc);
            s[i] = acc * scale;
            mx = fmaxf(mx, s[i]);
        }
#pragma unroll
        for (int o = 16; o; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
        float sum = 0.f;
#pragma unroll
        for (int i = 0; i < 8; i++) {
            s[i] = expf(s[i] - mx);
            sum += s[i];
        }
#pragma unroll
        for (int o = 16; o; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
        const float inv = 1.0f / sum;
#pragma unroll
        for (int i = 0; i < 8; i++) p[lane + 32 * i] = s[i] * inv;
        __syncwarp();
        for (int d = lane; d < HS; d += 32) {
            float acc = 0.f;
#pragma unroll 8
            for (int j = 0; j < 256; j++) acc = fmaf(p[j], Vs[j * HS + d], acc);
            out[(size_t)(seq * 256 + qi) * C + head * HS + d] = acc;
        }
        __syncwarp();
    }
}
template <int HS>
constexpr int p_attn_smem_bytes() { return (256 * (HS + 1) + 256 * HS + 8 * 256 + 8 * HS) * 4; }

// warp per sequence: ln_f on token 255 + logits[0:5] (tied lm_head rows), written as [seq][8]
__global__ void __launch_bounds__(128) p_head_kernel(const float *__restrict__ X, const float *__restrict__ gain,
                                                     const float *__restrict__ wte, float *__restrict__ logits, int C, int n_seq)
{
    const int seq = blockIdx.x * 4 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (seq >= n_seq) return;
    const float *x = X + ((size_t)seq * 256 + 255) * C;
    float s = 0.f;
    for (int c = lane; c < C; c += 32) s += x[c];
#pragma unroll
    for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    const float mean = s / (float)C;
    float q = 0.f;
    for (int c = lane; c < C; c += 32) {
        const float d = x[c] - mean;
        q += d * d;
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
    const float rstd = 1.0f / sqrtf(q / (float)C + 1e-5f);
    float acc[5] = {0.f, 0.f, 0.f, 0.f, 0.f};
    for (int c = lane; c < C; c += 32) {
        const float y = (x[c] - mean) * rstd * gain[c];
#pragma unroll
        for (int k = 0; k < 5; k++) acc[k] = fmaf(y, wte[(size_t)k * C + c], acc[k]);
    }
#pragma unroll
    for (int k = 0; k < 5; k++)
#pragma unroll
        for (int o = 16; o; o >>= 1) acc[k] += __shfl_xor_sync(0xffffffffu, acc[k], o);
    if (lane == 0) {
#pragma unroll
        for (int k = 0; k < 5; k++) logits[(size_t)seq * 8 + k] = acc[k];
    }
}

}  // namespace mg
