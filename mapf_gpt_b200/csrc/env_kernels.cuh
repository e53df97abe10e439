// env_kernels.cuh -- integer/byte kernels of the rollout path (HBM-bound, no tensor cores):
//   bfs_kernel         cost-to-go field per (env, agent)   observation_generator.cpp:134-176,200-220
//   set_state_kernel   positions/goals/actions from host   observation_generator.cpp:432-478
//   observe_kernel     history + greedy bits + tokenizer   observation_generator.cpp:412-430,288-311,487-528,352-389
//   sample_step_kernel GPT.act sampling (model.py:249-259) + POGEMA `soft` step
//
// HBM layout (capacity E envs x N agents, H x W grid, row pitch P = roundup(W, 8)):
//   obst  u8  [E][H][P]      loc  i16 [E][H][P] (agent id or -1)     c2g u16 [E][N][H][P]
//   pos, goal  short2 [E][N] (x = row, y = col)                       hist u8 [E][N][8] (5 used)
//   nextb u8 [E][N]          act i32 [E][N]                           tokens u8 [E*N][256]
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "ptx.cuh"

namespace mg {

struct EnvState {
    int E, N, H, W, P;
    int FR, FP;          // rows / pitch of one agent's cost-to-go field: (H, P) when the window is the whole grid
                         // (H, W <= 74), else (129, 136) = two grid_step blocks + 1 (cpp:207-210)
    int large;           // general windowed machinery on (some side > 74 cells)
    int gs;              // grid_step (64)
    short4 *bounds;      // [E][N] inclusive window (x=left, y=right, z=top, w=bottom), Cost2GoPartial h:67-82
    int32_t *map_of_env; // [E] slot of the per-map precompute tables (large only)
    int32_t *cell_idx;   // [maps][H*P] index of a precomputed cell or -1 (precomputed_cells_map, cpp:45-60)
    uint16_t **pre;      // [maps] -> K x K all-pairs distances between precomputed cells (cpp:82-113)
    int32_t *preK;       // [maps] K
    uint8_t *obst;
    int16_t *loc;
    uint16_t *c2g;
    short2 *pos, *goal;
    uint8_t *hist;
    uint8_t *nextb;
    int32_t *act;
    int32_t *nag;        // agents per env (0 = slot unused)
    uint8_t *active;     // [E] slots taking part in the current call (act_batch may address a subset of the slots)
    uint8_t *dirty;      // [E][N] cost-to-go must be recomputed
    uint8_t *tokens;
    float *logits;       // [E*N][8]
    // episode bookkeeping
    int32_t *steps;      // [E]
    uint8_t *done;       // [E]
    int32_t *arrive;     // [E][N] step at which the agent last arrived on its goal, -1 = not on goal
    unsigned long long *agent_steps;  // [E]
    float *density_sum;  // [E] sum over observations of the mean agents-per-traversable-FOV-cell (avg_agents_density)
    int32_t *density_n;  // [E] observations summed (reset observation + one per executed step)
    int32_t *vocab_err;  // [2]: [0] a relative position left the vocabulary, [1] a logit was not finite
};

__constant__ int c_moves[5][2] = {{0, 0}, {-1, 0}, {1, 0}, {0, -1}, {0, 1}};

// ------------------------------------------------------------------------------------------------
// Level-synchronous BFS from the goal over the whole padded grid, one block per (env, agent).
// For H,W <= 74 this IS compute_cost2go_partial: window = grid, only seed = goal.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) bfs_kernel(EnvState s, int first_env, int only_dirty)
{
    extern __shared__ __align__(16) uint8_t sm[];
    const int cells = s.H * s.P;
    uint16_t *dist = reinterpret_cast<uint16_t *>(sm);
    uint16_t *q0 = dist + cells;
    uint16_t *q1 = q0 + cells;
    uint8_t *ob = reinterpret_cast<uint8_t *>(q1 + cells);
    __shared__ int n_next;

    const int e = first_env + blockIdx.x / s.N, a = blockIdx.x % s.N;
    if (a >= s.nag[e]) return;
    if (only_dirty && !s.dirty[e * s.N + a]) return;
    const uint8_t *og = s.obst + (size_t)e * cells;
    for (int i = threadIdx.x; i < cells; i += blockDim.x) {
        dist[i] = 0xFFFF;
        ob[i] = (i % s.P) < s.W ? og[i] : 1;
    }
    const short2 g = s.goal[e * s.N + a];
    __syncthreads();
    if (threadIdx.x == 0) {
        dist[g.x * s.P + g.y] = 0;
        q0[0] = (uint16_t)(g.x * s.P + g.y);
        n_next = 0;
    }
    __syncthreads();
    int n_cur = 1, level = 0;
    uint16_t *qc = q0, *qn = q1;
    while (n_cur > 0) {
        for (int i = threadIdx.x; i < n_cur; i += blockDim.x) {
            const int c = qc[i];
            const int ci = c / s.P, cj = c - ci * s.P;
#pragma unroll
            for (int m = 1; m < 5; m++) {
                const int ni = ci + c_moves[m][0], nj = cj + c_moves[m][1];
                if (ni < 0 || nj < 0 || ni >= s.H || nj >= s.W) continue;
                const int nc = ni * s.P + nj;
                if (ob[nc]) continue;
                if (atomicCAS(reinterpret_cast<unsigned short *>(&dist[nc]), (unsigned short)0xFFFF,
                              (unsigned short)(level + 1)) == 0xFFFF)
                    qn[atomicAdd(&n_next, 1)] = (uint16_t)nc;
            }
        }
        __syncthreads();
        n_cur = n_next;
        __syncthreads();
        if (threadIdx.x == 0) n_next = 0;
        uint16_t *t = qc; qc = qn; qn = t;
        level++;
        __syncthreads();
    }
    uint16_t *out = s.c2g + ((size_t)e * s.N + a) * cells;
    for (int i = threadIdx.x; i < cells / 2; i += blockDim.x)
        reinterpret_cast<uint32_t *>(out)[i] = reinterpret_cast<uint32_t *>(dist)[i];
    if (threadIdx.x == 0) {
        s.dirty[e * s.N + a] = 0;
        s.bounds[e * s.N + a] = make_short4(0, (short)(s.H - 1), 0, (short)(s.W - 1));
    }
}

// ------------------------------------------------------------------------------------------------
// General cost-to-go machinery for maps wider than the window (some side > 74 cells), SURVEY 8(f).1.
//
// precompute_kernel: precompute_cost2go (cpp:43-113): one full-grid BFS per precomputed cell (free cells on rows/cols
// = 0 mod grid_step); pre[a][b] = distance between precomputed cells a and b.  Persistent blocks, each with its own
// global scratch (dist u16 + two u32 queues), level-synchronous.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) precompute_kernel(const uint8_t *__restrict__ obst, int H, int W, int P,
                                                         const int32_t *__restrict__ cells_list, int K,
                                                         uint16_t *__restrict__ pre, uint8_t *scratch)
{
    const int cells = H * P;
    uint16_t *dist = reinterpret_cast<uint16_t *>(scratch + (size_t)blockIdx.x * ((size_t)cells * 10 + 64));
    uint32_t *q0 = reinterpret_cast<uint32_t *>(dist + cells + (cells & 1));
    uint32_t *q1 = q0 + cells;
    __shared__ int n_next;
    for (int a = blockIdx.x; a < K; a += gridDim.x) {
        for (int i = threadIdx.x; i < cells; i += blockDim.x) dist[i] = 0xFFFF;
        __syncthreads();
        if (threadIdx.x == 0) {
            dist[cells_list[a]] = 0;
            q0[0] = (uint32_t)cells_list[a];
            n_next = 0;
        }
        __syncthreads();
        int n_cur = 1, level = 0;
        uint32_t *qc = q0, *qn = q1;
        while (n_cur > 0) {
            for (int i = threadIdx.x; i < n_cur; i += blockDim.x) {
                const int c = (int)qc[i];
                const int ci = c / P, cj = c - ci * P;
#pragma unroll
                for (int m = 1; m < 5; m++) {
                    const int ni = ci + c_moves[m][0], nj = cj + c_moves[m][1];
                    if (ni < 0 || nj < 0 || ni >= H || nj >= W) continue;
                    const int nc = ni * P + nj;
                    if (obst[nc]) continue;
                    if (atomicCAS(reinterpret_cast<unsigned short *>(&dist[nc]), (unsigned short)0xFFFF,
                                  (unsigned short)(level + 1)) == 0xFFFF)
                        qn[atomicAdd(&n_next, 1)] = (uint32_t)nc;
                }
            }
            __syncthreads();
            n_cur = n_next;
            __syncthreads();
            if (threadIdx.x == 0) n_next = 0;
            uint32_t *t = qc; qc = qn; qn = t;
            level++;
            __syncthreads();
        }
        for (int b = threadIdx.x; b < K; b += blockDim.x) pre[(size_t)a * K + b] = dist[cells_list[b]];
        __syncthreads();
    }
}

// ------------------------------------------------------------------------------------------------
// partial_kernel: compute_cost2go_partial (cpp:200-286) for one (env, agent) per block:
//   window = two grid_step blocks from the block holding pos-5, clipped (cpp:207-210);
//   goal-block BFS (get_goal_border_and_cost2go, cpp:134-176);
//   seeds on the window border lines: min over goal-block border cells of gcm[g] + pre[g][cell] (cpp:224-239), plus the
//   goal itself when it lies inside the window (cpp:241-245);
//   multi-source BFS over the window with seeds injected when the frontier reaches their cost (cpp:246-279).  The
//   reference's FIFO is reproduced level by level: all seeds of cost c are injected before level c expands -- and, as in
//   the reference (cpp:259,276), an injected seed OVERWRITES a smaller cost already on its cell.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) partial_kernel(EnvState s, int first_env, int only_dirty)
{
    extern __shared__ __align__(16) uint8_t sm[];
    const int e = first_env + blockIdx.x / s.N, a = blockIdx.x % s.N;
    if (a >= s.nag[e]) return;
    const int idx = e * s.N + a;
    if (only_dirty && !s.dirty[idx]) return;
    const int H = s.H, W = s.W, P = s.P, gs = s.gs;
    const int WMAX = 2 * gs + 1;                       // 129
    const int QCAP = WMAX * WMAX + 4 * WMAX + 8;       // every window cell once + re-pushed seeds
    size_t off = 0;
    auto carve = [&](size_t bytes) { uint8_t *p = sm + off; off += (bytes + 15) & ~(size_t)15; return p; };
    uint16_t *cm = reinterpret_cast<uint16_t *>(carve((size_t)WMAX * WMAX * 2));           // window cost matrix
    uint16_t *gcm = reinterpret_cast<uint16_t *>(carve((size_t)(gs + 1) * (gs + 1) * 2));  // goal-block cost matrix
    uint16_t *q0 = reinterpret_cast<uint16_t *>(carve((size_t)QCAP * 2));                  // frontier queues (window-local ids)
    uint16_t *q1 = reinterpret_cast<uint16_t *>(carve((size_t)QCAP * 2));
    int *seed_cost = reinterpret_cast<int *>(carve((size_t)(4 * WMAX + 4) * 4));
    uint16_t *seed_cell = reinterpret_cast<uint16_t *>(carve((size_t)(4 * WMAX + 4) * 2));
    __shared__ int n_next, n_seeds, next_cost;

    const uint8_t *ob = s.obst + (size_t)e * H * P;
    const short2 pos = s.pos[idx], goal = s.goal[idx];
    const int left = max(pos.x - 5, 0) / gs * gs, right = min(left + 2 * gs, H - 1);
    const int top = max(pos.y - 5, 0) / gs * gs, bottom = min(top + 2 * gs, W - 1);
    const int wr = right - left + 1, wc = bottom - top + 1;
    const int gL = goal.x / gs * gs, gR = min(gL + gs, H - 1), gT = goal.y / gs * gs, gB = min(gT + gs, W - 1);
    const int gc = gB - gT + 1, gr = gR - gL + 1;

    // ---- goal-block BFS
    for (int i = threadIdx.x; i < gr * gc; i += blockDim.x) gcm[i] = 0xFFFF;
    __syncthreads();
    if (threadIdx.x == 0) {
        gcm[(goal.x - gL) * gc + (goal.y - gT)] = 0;
        q0[0] = (uint16_t)((goal.x - gL) * gc + (goal.y - gT));
        n_next = 0;
    }
    __syncthreads();
    {
        int n_cur = 1, level = 0;
        uint16_t *qc = q0, *qn = q1;
        while (n_cur > 0) {
            for (int i = threadIdx.x; i < n_cur; i += blockDim.x) {
                const int c = qc[i];
                const int ci = c / gc + gL, cj = c % gc + gT;
#pragma unroll
                for (int m = 1; m < 5; m++) {
                    const int ni = ci + c_moves[m][0], nj = cj + c_moves[m][1];
                    if (ni < gL || nj < gT || ni > gR || nj > gB) continue;
                    if (ob[ni * P + nj]) continue;
                    const int nc = (ni - gL) * gc + (nj - gT);
                    if (atomicCAS(reinterpret_cast<unsigned short *>(&gcm[nc]), (unsigned short)0xFFFF,
                                  (unsigned short)(level + 1)) == 0xFFFF)
                        qn[atomicAdd(&n_next, 1)] = (uint16_t)nc;
                }
            }
            __syncthreads();
            n_cur = n_next;
            __syncthreads();
            if (threadIdx.x == 0) n_next = 0;
            uint16_t *t = qc; qc = qn; qn = t;
            level++;
            __syncthreads();
        }
    }
    // ---- seeds on the window border lines (get_cells_on_border, cpp:178-198: unclipped far lines, only if inside the grid)
    const int map = s.map_of_env[e];
    const int32_t *cidx = s.cell_idx + (size_t)map * H * P;
    const uint16_t *pre = s.pre[map];
    const int K = s.preK[map];
    const int far_r = left + 2 * gs, far_c = top + 2 * gs;
    const int n_i = min(far_r, H) - left, n_j = min(far_c, W) - top;
    if (threadIdx.x == 0) n_seeds = 0;
    for (int i = threadIdx.x; i < wr * wc; i += blockDim.x) cm[i] = 0xFFFF;
    __syncthreads();
    for (int t = threadIdx.x; t < 2 * n_i + 2 * n_j; t += blockDim.x) {
        int ci, cj;
        if (t < n_i) { ci = left + t; cj = top; }
        else if (t < 2 * n_i) { if (far_c >= W) continue; ci = left + (t - n_i); cj = far_c; }
        else if (t < 2 * n_i + n_j) { ci = left; cj = top + (t - 2 * n_i); }
        else { if (far_r >= H) continue; ci = far_r; cj = top + (t - 2 * n_i - n_j); }
        if (ob[ci * P + cj]) continue;
        const int ccol = cidx[ci * P + cj];
        int min_cost = 65535;
        // goal-block border cells (cpp:141-152)
        for (int u = 0; u < 2 * gr + 2 * gc; u++) {
            int gi, gj;
            if (u < gr) { gi = gL + u; gj = gT; }
            else if (u < 2 * gr) { if (gT + gs >= W) continue; gi = gL + (u - gr); gj = gB; }
            else if (u < 2 * gr + gc) { gi = gL; gj = gT + (u - 2 * gr); }
            else { if (gL + gs >= H) continue; gi = gR; gj = gT + (u - 2 * gr - gc); }
            if (ob[gi * P + gj]) continue;
            const int gcost = gcm[(gi - gL) * gc + (gj - gT)];
            if (gcost == 0xFFFF) continue;
            const int nc = gcost + (int)pre[(size_t)cidx[gi * P + gj] * K + ccol];
            if (min_cost > nc) min_cost = nc;
        }
        if (min_cost != 65535) {
            const int k = atomicAdd(&n_seeds, 1);
            seed_cost[k] = min_cost;
            seed_cell[k] = (uint16_t)((ci - left) * wc + (cj - top));
        }
    }
    __syncthreads();
    if (threadIdx.x == 0 && goal.x >= left && goal.x <= right && goal.y >= top && goal.y <= bottom) {
        const int k = n_seeds++;
        seed_cost[k] = 0;
        seed_cell[k] = (uint16_t)((goal.x - left) * wc + (goal.y - top));
    }
    __syncthreads();
    // ---- multi-source BFS, level by level
    const int ns = n_seeds;
    {
        int n_cur = 0, c = 0;
        uint16_t *qc = q0, *qn = q1;
        while (true) {
            if (n_cur == 0) {   // fringe empty: jump to the cheapest pending seed (cpp:247-249, 273-278); seeds < c are done
                if (threadIdx.x == 0) next_cost = 0x7FFFFFFF;
                __syncthreads();
                int best = 0x7FFFFFFF;
                for (int k = threadIdx.x; k < ns; k += blockDim.x)
                    if (seed_cost[k] >= c) best = min(best, seed_cost[k]);
                if (best != 0x7FFFFFFF) atomicMin(&next_cost, best);
                __syncthreads();
                if (next_cost == 0x7FFFFFFF) break;
                c = next_cost;
                if (threadIdx.x == 0) n_next = 0;
                __syncthreads();
            }
            // inject the seeds of cost c (overwriting, cpp:259,276)
            if (threadIdx.x == 0) n_next = n_cur;
            __syncthreads();
            for (int k = threadIdx.x; k < ns; k += blockDim.x)
                if (seed_cost[k] == c) {
                    cm[seed_cell[k]] = (uint16_t)c;
                    qc[atomicAdd(&n_next, 1)] = seed_cell[k];
                }
            __syncthreads();
            n_cur = n_next;
            __syncthreads();
            if (threadIdx.x == 0) n_next = 0;
            __syncthreads();
            for (int i = threadIdx.x; i < n_cur; i += blockDim.x) {
                const int cell = qc[i];
                const int ci = cell / wc + left, cj = cell % wc + top;
#pragma unroll
                for (int m = 1; m < 5; m++) {
                    const int ni = ci + c_moves[m][0], nj = cj + c_moves[m][1];
                    if (ni < left || ni > right || nj < top || nj > bottom) continue;
                    if (ob[ni * P + nj]) continue;
                    const int nc = (ni - left) * wc + (nj - top);
                    if (atomicCAS(reinterpret_cast<unsigned short *>(&cm[nc]), (unsigned short)0xFFFF,
                                  (unsigned short)(c + 1)) == 0xFFFF)
                        qn[atomicAdd(&n_next, 1)] = (uint16_t)nc;
                }
            }
            __syncthreads();
            n_cur = n_next;
            __syncthreads();
            uint16_t *t = qc; qc = qn; qn = t;
            c++;
        }
    }
    // ---- store the partial field and its window
    uint16_t *out = s.c2g + (size_t)idx * s.FR * s.FP;
    for (int i = threadIdx.x; i < wr * wc; i += blockDim.x) out[(i / wc) * s.FP + (i % wc)] = cm[i];
    if (threadIdx.x == 0) {
        s.bounds[idx] = make_short4((short)left, (short)right, (short)top, (short)bottom);
        s.dirty[idx] = 0;
    }
}

// ------------------------------------------------------------------------------------------------
// update_agents, first half (cpp:432-478): take positions / goals / actions from staging buffers.
// One block per env.  loc is cleared for ALL old cells before any new cell is written (cpp:434-435).
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) set_state_kernel(EnvState s, const int32_t *__restrict__ pos_in,
                                                        const int32_t *__restrict__ goal_in,
                                                        const int32_t *__restrict__ act_in)
{
    const int e = blockIdx.x;
    const int n = s.nag[e];
    if (n == 0 || !s.active[e]) return;
    const int cells = s.H * s.P;
    int16_t *loc = s.loc + (size_t)e * cells;
    if (pos_in) {
        for (int i = threadIdx.x; i < n; i += blockDim.x) {
            const short2 p = s.pos[e * s.N + i];
            loc[p.x * s.P + p.y] = -1;
        }
        __syncthreads();
        for (int i = threadIdx.x; i < n; i += blockDim.x) {
            const int x = pos_in[(e * s.N + i) * 2], y = pos_in[(e * s.N + i) * 2 + 1];
            s.pos[e * s.N + i] = make_short2((short)x, (short)y);
        }
        __syncthreads();
        // agents_locations[pos] = i in index order: on a (never legal) shared cell the highest id wins
        for (int i = threadIdx.x; i < n; i += blockDim.x) {
            const short2 p = s.pos[e * s.N + i];
            loc[p.x * s.P + p.y] = (int16_t)i;
        }
    }
    if (goal_in) {
        for (int i = threadIdx.x; i < n; i += blockDim.x) {
            const int x = goal_in[(e * s.N + i) * 2], y = goal_in[(e * s.N + i) * 2 + 1];
            const short2 g = s.goal[e * s.N + i];
            if (g.x != x || g.y != y) {
                s.goal[e * s.N + i] = make_short2((short)x, (short)y);
                s.dirty[e * s.N + i] = 1;
            }
        }
    }
    if (pos_in && s.large) {   // FOV leaves the window -> recompute the partial field (cpp:469-477)
        __syncthreads();
        for (int i = threadIdx.x; i < n; i += blockDim.x) {
            const short2 p = s.pos[e * s.N + i];
            const short4 b = s.bounds[e * s.N + i];
            if (p.x - 5 < b.x || p.x + 5 > b.y || p.y - 5 < b.z || p.y + 5 > b.w) s.dirty[e * s.N + i] = 1;
        }
    }
    if (act_in) {
        for (int i = threadIdx.x; i < n; i += blockDim.x) s.act[e * s.N + i] = act_in[e * s.N + i];
    }
}

// ------------------------------------------------------------------------------------------------
// observe_kernel: one block per env.
//   phase A (kUpdate): history push (cpp:441-462) + greedy-direction bits (update_next_action, cpp:412-430)
//   phase B (kTokens): cost-to-go window (cpp:288-311), 13 nearest agents by (manhattan, id)
//                      (cpp:487-514), Encoder::encode (cpp:352-389) -> 256 uint8 tokens per agent.
// Trained shape only: radius 5, 13 agents, 5 previous actions, limit 20, context 256.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ int field_dist(const uint16_t *F, int x, int y, const short4 b, int FP)
{   // get_distance, cpp:313-319: -1 outside [left,right) x [top,bottom) -- upper bounds EXCLUSIVE
    if (x < b.x || x >= b.y || y < b.z || y >= b.w) return -1;
    return F[(x - b.x) * FP + (y - b.z)];
}

template <bool kUpdate, bool kTokens>
__global__ void __launch_bounds__(256) observe_kernel(EnvState s)
{
    const int e = blockIdx.x;
    const int n = s.nag[e];
    if (n == 0 || !s.active[e]) return;
    const int cells = s.H * s.P;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    __shared__ __align__(16) uint8_t tokbuf[8][256];
    __shared__ int sel[8][16];

    if (kUpdate) {
        for (int i = threadIdx.x; i < n; i += blockDim.x) {
            const int idx = e * s.N + i;
            const int a = s.act[idx];
            uint8_t *h = s.hist + (size_t)idx * 8;
            h[0] = h[1]; h[1] = h[2]; h[2] = h[3]; h[3] = h[4];
            h[4] = (a >= 0 && a <= 4) ? (uint8_t)(45 + a) : (uint8_t)44;
            const uint16_t *F = s.c2g + (size_t)idx * s.FR * s.FP;
            const short2 p = s.pos[idx];
            const short4 bd = s.bounds[idx];
            const int cur = field_dist(F, p.x, p.y, bd, s.FP);
            int bits = 0;
#pragma unroll
            for (int m = 1; m < 5; m++) {
                const int nb = field_dist(F, p.x + c_moves[m][0], p.y + c_moves[m][1], bd, s.FP);
                bits = (bits << 1) | ((nb >= 0 && cur > nb) ? 1 : 0);
            }
            s.nextb[idx] = (uint8_t)bits;
        }
        __syncthreads();
    }
    if (!kTokens) return;

    const int16_t *loc = s.loc + (size_t)e * cells;
    for (int i = warp; i < n; i += (blockDim.x >> 5)) {
        const int idx = e * s.N + i;
        const short2 p = s.pos[idx];
        const short4 bd = s.bounds[idx];
        const uint16_t *F = s.c2g + (size_t)idx * s.FR * s.FP + (p.x - 5 - bd.x) * s.FP + (p.y - 5 - bd.z);   // window origin
        const int mid = F[5 * s.FP + 5];
        uint8_t *tb = tokbuf[warp];
        unsigned key[4];
#pragma unroll
        for (int t = 0; t < 4; t++) {
            const int w = lane + 32 * t;
            key[t] = 0xFFFFFFFFu;
            if (w < 121) {
                const int wi = w / 11, wj = w - wi * 11;
                const int c = (p.x - 5 + wi) * s.P + (p.y - 5 + wj);
                const int v = F[wi * s.FP + wj];
                int tok;
                if (v == 0xFFFF) tok = 41;
                else {
                    const int d = v - mid;
                    tok = d > 20 ? 43 : (d < -20 ? 42 : d + 20);
                }
                tb[w] = (uint8_t)tok;
                const int id = loc[c];
                if (id >= 0) key[t] = ((unsigned)(abs(wi - 5) + abs(wj - 5)) << 16) | (unsigned)id;
            }
        }
        // 13 smallest keys (distance, then agent id), warp-wide
        int count = 0;
#pragma unroll 1
        for (int k = 0; k < 13; k++) {
            const unsigned lm = min(min(key[0], key[1]), min(key[2], key[3]));
            const unsigned m = __reduce_min_sync(0xffffffffu, lm);
            if (m == 0xFFFFFFFFu) break;
#pragma unroll
            for (int t = 0; t < 4; t++)
                if (key[t] == m) key[t] = 0xFFFFFFFFu;
            if (lane == 0) sel[warp][k] = (int)(m & 0xFFFF);
            count++;
        }
        __syncwarp();
        for (int t = lane; t < 135; t += 32) {  // 130 slot tokens + 5 tail pads
            uint8_t tok = 66;
            const int slot = t / 10, f = t - slot * 10;
            if (slot < count) {
                const int j = e * s.N + sel[warp][slot];
                if (f < 4) {
                    const short2 q = (f < 2) ? s.pos[j] : s.goal[j];
                    int d = ((f & 1) ? q.y - p.y : q.x - p.x);
                    if (f >= 2) d = max(-20, min(20, d));               // relative goal is clamped (cpp:360-361)
                    else if (d < -20 || d > 20) atomicExch(s.vocab_err, 1);  // int_vocab.at() would throw
                    tok = (uint8_t)(d + 20);
                } else if (f < 9) tok = s.hist[(size_t)j * 8 + (f - 4)];
                else tok = (uint8_t)(50 + s.nextb[j]);
            }
            tb[121 + t] = tok;
        }
        __syncwarp();
        reinterpret_cast<uint2 *>(s.tokens + (size_t)idx * 256)[lane] = reinterpret_cast<const uint2 *>(tb)[lane];
        __syncwarp();
    }
}

// ------------------------------------------------------------------------------------------------
// observe_tma_kernel: same results as observe_kernel<kUpdate, true>, but each agent's two 11x11 FOV windows (its own
// cost-to-go field and the env's agent-id map) are pulled into shared memory by TMA tiled loads: one 3-D box
// {24 cols, 11 rows, 1 plane} of u16/i16 per field.  The box must START on a 16-byte boundary in the innermost dimension
// (measured: an unaligned start coordinate traps with "illegal instruction"), so the box starts at the 8-column boundary
// below the window and is 24 columns wide (22-byte window + up to 14 bytes of lead-in, rounded to a multiple of 16 B);
// columns past the grid pitch are zero-filled by the TMA unit and never read.  One block per env, one warp per agent at a
// time, double-buffered: the boxes of the warp's NEXT agent are in flight while it tokenizes the current one.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void tma_load_3d(void *smem_dst, const CUtensorMap *map, int x, int y, int z, uint64_t *bar)
{
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
        ::"r"(smem_u32(smem_dst)), "l"(map), "r"(x), "r"(y), "r"(z), "r"(smem_u32(bar))
        : "memory");
}

template <bool kUpdate>
__global__ void __launch_bounds__(256) observe_tma_kernel(EnvState s, const __grid_constant__ CUtensorMap map_c2g,
                                                          const __grid_constant__ CUtensorMap map_loc)
{
    const int e = blockIdx.x;
    const int n = s.nag[e];
    if (n == 0 || !s.active[e]) return;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    __shared__ __align__(128) uint16_t win[8][2][2][320];   // [warp][buffer][field: 0 c2g, 1 loc][11 rows x 24 cols], 640-byte blocks (128-byte aligned TMA destinations)
    __shared__ __align__(16) uint8_t tokbuf[8][256];
    __shared__ int sel[8][16];
    __shared__ uint64_t bars[8][2];

    if (lane == 0) {
        mbar_init(&bars[warp][0], 1);
        mbar_init(&bars[warp][1], 1);
        fence_barrier_init();
    }
    if (kUpdate) {
        for (int i = threadIdx.x; i < n; i += blockDim.x) {
            const int idx = e * s.N + i;
            const int a = s.act[idx];
            uint8_t *h = s.hist + (size_t)idx * 8;
            h[0] = h[1]; h[1] = h[2]; h[2] = h[3]; h[3] = h[4];
            h[4] = (a >= 0 && a <= 4) ? (uint8_t)(45 + a) : (uint8_t)44;
            const uint16_t *F = s.c2g + (size_t)idx * s.FR * s.FP;
            const short2 p = s.pos[idx];
            const short4 bd = s.bounds[idx];
            const int cur = field_dist(F, p.x, p.y, bd, s.FP);
            int bits = 0;
#pragma unroll
            for (int m = 1; m < 5; m++) {
                const int nb = field_dist(F, p.x + c_moves[m][0], p.y + c_moves[m][1], bd, s.FP);
                bits = (bits << 1) | ((nb >= 0 && cur > nb) ? 1 : 0);
            }
            s.nextb[idx] = (uint8_t)bits;
        }
    }
    __syncthreads();

    auto issue = [&](int i, int buf) {
        if (lane == 0) {
            const short2 p = s.pos[e * s.N + i];
            const short4 bd = s.bounds[e * s.N + i];
            mbar_expect_tx(&bars[warp][buf], 2 * 11 * 24 * 2);
            tma_load_3d(&win[warp][buf][0][0], &map_c2g, (p.y - 5 - bd.z) & ~7, p.x - 5 - bd.x, e * s.N + i, &bars[warp][buf]);
            tma_load_3d(&win[warp][buf][1][0], &map_loc, (p.y - 5) & ~7, p.x - 5, e, &bars[warp][buf]);
        }
    };
    const int nw = blockDim.x >> 5;
    if (warp < n) issue(warp, 0);
    int it = 0;
    for (int i = warp; i < n; i += nw, it++) {
        const int buf = it & 1;
        if (i + nw < n) issue(i + nw, buf ^ 1);
        mbar_wait(&bars[warp][buf], (it >> 1) & 1);
        const int idx = e * s.N + i;
        const short2 p = s.pos[idx];
        const short4 bd = s.bounds[idx];
        const uint16_t *wc = &win[warp][buf][0][0] + ((p.y - 5 - bd.z) & 7);      // window column 0 inside the 8-aligned box
        const int16_t *wl = reinterpret_cast<const int16_t *>(&win[warp][buf][1][0]) + ((p.y - 5) & 7);
        const int mid = wc[5 * 24 + 5];
        uint8_t *tb = tokbuf[warp];
        unsigned key[4];
#pragma unroll
        for (int t = 0; t < 4; t++) {
            const int w = lane + 32 * t;
            key[t] = 0xFFFFFFFFu;
            if (w < 121) {
                const int wi = w / 11, wj = w - wi * 11;
                const int v = wc[wi * 24 + wj];
                int tok;
                if (v == 0xFFFF) tok = 41;
                else {
                    const int d = v - mid;
                    tok = d > 20 ? 43 : (d < -20 ? 42 : d + 20);
                }
                tb[w] = (uint8_t)tok;
                const int id = wl[wi * 24 + wj];
                if (id >= 0) key[t] = ((unsigned)(abs(wi - 5) + abs(wj - 5)) << 16) | (unsigned)id;
            }
        }
        int count = 0;
#pragma unroll 1
        for (int k = 0; k < 13; k++) {
            const unsigned lm = min(min(key[0], key[1]), min(key[2], key[3]));
            const unsigned m = __reduce_min_sync(0xffffffffu, lm);
            if (m == 0xFFFFFFFFu) break;
#pragma unroll
            for (int t = 0; t < 4; t++)
                if (key[t] == m) key[t] = 0xFFFFFFFFu;
            if (lane == 0) sel[warp][k] = (int)(m & 0xFFFF);
            count++;
        }
        __syncwarp();
        for (int t = lane; t < 135; t += 32) {
            uint8_t tok = 66;
            const int slot = t / 10, f = t - slot * 10;
            if (slot < count) {
                const int j = e * s.N + sel[warp][slot];
                if (f < 4) {
                    const short2 q = (f < 2) ? s.pos[j] : s.goal[j];
                    int d = ((f & 1) ? q.y - p.y : q.x - p.x);
                    if (f >= 2) d = max(-20, min(20, d));
                    else if (d < -20 || d > 20) atomicExch(s.vocab_err, 1);
                    tok = (uint8_t)(d + 20);
                } else if (f < 9) tok = s.hist[(size_t)j * 8 + (f - 4)];
                else tok = (uint8_t)(50 + s.nextb[j]);
            }
            tb[121 + t] = tok;
        }
        __syncwarp();
        reinterpret_cast<uint2 *>(s.tokens + (size_t)idx * 256)[lane] = reinterpret_cast<const uint2 *>(tb)[lane];
        __syncwarp();
    }
}

// ------------------------------------------------------------------------------------------------
// Philox4x32-10 (counter-based; one independent stream per (seed, env, agent, step))
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1,
                                              uint32_t (&out)[4])
{
#pragma unroll
    for (int r = 0; r < 10; r++) {
        const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
        const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
        const uint32_t n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
        c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}
__device__ __forceinline__ float exp1_from_bits(uint32_t b)
{   // u in (0,1], q = -log(u) ~ Exp(1)   (torch.multinomial draws q with exponential_(1))
    const float u = ((float)(b >> 8) + 1.0f) * (1.0f / 16777216.0f);
    return -logf(u);
}

// ------------------------------------------------------------------------------------------------
// sample_step_kernel: one block per env.
//   (1) GPT.act tail (model.py:249-259): softmax over logits[0:5]; greedy argmax, or
//       multinomial == argmax(p / q), q ~ Exp(1) from Philox (mode 1) or supplied (mode 2).
//   (2) POGEMA soft step (App. C.3) in its order-independent fixed-point form:
//       W = waits  U obstacle targets  U swaps  U (movers into one cell, all but the lowest index)
//           U followers of an occupant in W (iterated);   everyone outside W moves at once.
//   (3) episode counters (App. C.4-C.5).
// mode: 0 greedy, 1 philox, 2 supplied q, 3 = actions already in s.act (host supplied).
// ------------------------------------------------------------------------------------------------
// avg_agents_density (pogema's AgentsDensityWrapper; experiment_setup/create_env.py:36-40 wraps every eval env with it;
// SURVEY App. C.5 -- recollection of an un-vendored dependency, parity unpinned): for one observation, the mean over agents of
//   (agents inside the agent's 11x11 FOV, itself included) / (traversable cells of that FOV).
// One warp per agent, a lane per 4 window cells.  Called by all threads of the block; returns the env's mean in thread 0.
__device__ __forceinline__ float fov_density(const EnvState &s, int e, int n, float *scratch /* [8] shared */)
{
    const int cells = s.H * s.P;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
    const int16_t *loc = s.loc + (size_t)e * cells;
    const uint8_t *ob = s.obst + (size_t)e * cells;
    float acc = 0.f;
    for (int i = warp; i < n; i += nw) {
        const short2 p = s.pos[e * s.N + i];
        int ag = 0, fr = 0;
#pragma unroll
        for (int t = 0; t < 4; t++) {
            const int w = lane + 32 * t;
            if (w < 121) {
                const int wi = w / 11, wj = w - wi * 11;
                const int c = (p.x - 5 + wi) * s.P + (p.y - 5 + wj);
                ag += loc[c] >= 0 ? 1 : 0;
                fr += ob[c] == 0 ? 1 : 0;
            }
        }
        ag = __reduce_add_sync(0xffffffffu, ag);
        fr = __reduce_add_sync(0xffffffffu, fr);
        acc += (float)ag / (float)max(fr, 1);
    }
    __syncthreads();
    if (lane == 0) scratch[warp] = acc;
    __syncthreads();
    float tot = 0.f;
    if (threadIdx.x == 0)
        for (int w = 0; w < nw; w++) tot += scratch[w];
    return tot / (float)n;
}

struct StepArgs {
    int mode;
    int do_step;
    const float *q;          // mode 2: [E*N][5]
    unsigned long long seed;
    const int32_t *act_override;  // actions to EXECUTE (history keeps s.act = the sampled ones), or null
    int env_offset;          // global env id of slot 0 (multi-GPU sharding: streams follow the env, not the rank)
    int max_episode_steps;   // 0 = unlimited
};

__global__ void __launch_bounds__(256) sample_step_kernel(EnvState s, StepArgs a)
{
    extern __shared__ __align__(16) uint8_t sm[];
    const int e = blockIdx.x;
    const int n = s.nag[e];
    if (n == 0 || !s.active[e]) return;
    const int cells = s.H * s.P;
    int *tgt = reinterpret_cast<int *>(sm);                   // [N]
    uint8_t *wait = reinterpret_cast<uint8_t *>(tgt + s.N);   // [N]
    int16_t *loc = s.loc + (size_t)e * cells;
    const uint8_t *ob = s.obst + (size_t)e * cells;
    const int t_now = s.steps[e];
    const bool frozen = s.done[e] != 0;

    if (a.mode != 3) {
        for (int i = threadIdx.x; i < n; i += blockDim.x) {
            const int idx = e * s.N + i;
            const float *l = s.logits + (size_t)idx * 8;
            const float l0 = l[0], l1 = l[1], l2 = l[2], l3 = l[3], l4 = l[4];
            // torch.multinomial raises on inf / nan probabilities (model.py:257); here the flag also tells the engine that
            // the max-free attention softmax left its range (engine.cu: safe_softmax)
            if (!(isfinite(l0) && isfinite(l1) && isfinite(l2) && isfinite(l3) && isfinite(l4))) atomicExch(s.vocab_err + 1, 1);
            const float mx = fmaxf(fmaxf(fmaxf(l0, l1), fmaxf(l2, l3)), l4);
            float p[5] = {expf(l0 - mx), expf(l1 - mx), expf(l2 - mx), expf(l3 - mx), expf(l4 - mx)};
            const float sum = (((p[0] + p[1]) + p[2]) + p[3]) + p[4];
            float q[5] = {1.f, 1.f, 1.f, 1.f, 1.f};
            if (a.mode == 1) {
                uint32_t r0[4], r1[4];
                const uint32_t ge = (uint32_t)(a.env_offset + e);
                philox4x32_10(ge, (uint32_t)i, (uint32_t)t_now, 0u, (uint32_t)a.seed, (uint32_t)(a.seed >> 32), r0);
                philox4x32_10(ge, (uint32_t)i, (uint32_t)t_now, 1u, (uint32_t)a.seed, (uint32_t)(a.seed >> 32), r1);
                q[0] = exp1_from_bits(r0[0]); q[1] = exp1_from_bits(r0[1]); q[2] = exp1_from_bits(r0[2]);
                q[3] = exp1_from_bits(r0[3]); q[4] = exp1_from_bits(r1[0]);
            } else if (a.mode == 2) {
#pragma unroll
                for (int k = 0; k < 5; k++) q[k] = a.q[(size_t)idx * 5 + k];
            }
            int best = 0;
            float bv = (p[0] / sum) / q[0];
#pragma unroll
            for (int k = 1; k < 5; k++) {
                const float v = (p[k] / sum) / q[k];
                if (v > bv) { bv = v; best = k; }
            }
            s.act[idx] = best;
        }
    }
    if (!a.do_step || frozen) return;
    __syncthreads();
    __shared__ float dens_scratch[8];
    if (t_now == 0) {   // the observation env.reset() returned
        const float d0 = fov_density(s, e, n, dens_scratch);
        if (threadIdx.x == 0) { s.density_sum[e] = d0; s.density_n[e] = 1; }
    }

    // ---- soft collision step
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const int idx = e * s.N + i;
        int act = a.act_override ? a.act_override[idx] : s.act[idx];
        if (act < 0 || act > 4) act = 0;
        const short2 p = s.pos[idx];
        const int t = (p.x + c_moves[act][0]) * s.P + (p.y + c_moves[act][1]);
        tgt[i] = t;
        wait[i] = (act == 0 || ob[t] != 0) ? 1 : 0;       // (a) waits and obstacle targets
    }
    __syncthreads();
    for (int i = threadIdx.x; i < n; i += blockDim.x) {   // (b) swaps
        if (wait[i]) continue;
        const int j = loc[tgt[i]];
        const short2 p = s.pos[e * s.N + i];
        if (j >= 0 && j != i && tgt[j] == p.x * s.P + p.y) wait[i] = 2;  // mark; applied after the sweep
    }
    __syncthreads();
    for (int i = threadIdx.x; i < n; i += blockDim.x)
        if (wait[i] == 2) wait[i] = 1;
    __syncthreads();
    for (int i = threadIdx.x; i < n; i += blockDim.x) {    // (c) lowest-index mover keeps the claim (n <= 512: plain scan)
        if (wait[i]) continue;
        bool lose = false;
        for (int j = 0; j < i; j++) lose |= (wait[j] != 1 && tgt[j] == tgt[i]);   // 0 or 3: a mover after (a),(b)
        if (lose) wait[i] = 3;                              // applied after the sweep: losers must still block higher ids
    }
    __syncthreads();
    for (int i = threadIdx.x; i < n; i += blockDim.x)
        if (wait[i] == 3) wait[i] = 1;
    // (d) followers of a blocked occupant, to the fixed point
    while (true) {
        __syncthreads();
        int changed = 0;
        for (int i = threadIdx.x; i < n; i += blockDim.x) {
            if (wait[i]) continue;
            const int k = loc[tgt[i]];
            if (k >= 0 && wait[k]) { wait[i] = 1; changed = 1; }
        }
        if (!__syncthreads_or(changed)) break;
    }
    // apply all surviving moves at once
    for (int i = threadIdx.x; i < n; i += blockDim.x)
        if (!wait[i]) {
            const short2 p = s.pos[e * s.N + i];
            loc[p.x * s.P + p.y] = -1;
        }
    __syncthreads();
    int on_goal = 0;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const int idx = e * s.N + i;
        short2 p = s.pos[idx];
        if (!wait[i]) {
            const int t = tgt[i];
            p = make_short2((short)(t / s.P), (short)(t % s.P));
            s.pos[idx] = p;
            loc[t] = (int16_t)i;
        }
        if (s.large) {
            const short4 b = s.bounds[idx];
            if (p.x - 5 < b.x || p.x + 5 > b.y || p.y - 5 < b.z || p.y + 5 > b.w) s.dirty[idx] = 1;
        }
        const short2 g = s.goal[idx];
        const bool og = (p.x == g.x && p.y == g.y);
        if (og) { if (s.arrive[idx] < 0) s.arrive[idx] = t_now + 1; }
        else s.arrive[idx] = -1;
        on_goal += og ? 1 : 0;
    }
    __shared__ int s_on;
    if (threadIdx.x == 0) s_on = 0;
    __syncthreads();
    if (on_goal) atomicAdd(&s_on, on_goal);
    __syncthreads();
    {   // the observation this step returns
        const float d1 = fov_density(s, e, n, dens_scratch);
        if (threadIdx.x == 0) { s.density_sum[e] += d1; s.density_n[e] += 1; }
    }
    if (threadIdx.x == 0) {
        s.steps[e] = t_now + 1;
        s.agent_steps[e] += (unsigned long long)n;
        if (s_on == n || (a.max_episode_steps > 0 && t_now + 1 >= a.max_episode_steps)) s.done[e] = 1;
    }
}

// per-slot metrics (App. C.5): [ep_length, CSR, ISR, SoC, makespan, on_goal_now, agent_steps, n_agents, avg_agents_density, density samples]
#define MG_METRIC_COLS_DEV 10
__global__ void metrics_kernel(EnvState s, double *out)
{
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= s.E) return;
    const int n = s.nag[e];
    double *o = out + (size_t)e * MG_METRIC_COLS_DEV;
    if (n == 0) { for (int k = 0; k < MG_METRIC_COLS_DEV; k++) o[k] = 0.0; return; }
    const int T = s.steps[e];
    int on = 0, soc = 0, mk = 0;
    for (int i = 0; i < n; i++) {
        const int a = s.arrive[e * s.N + i];
        const int cost = a >= 0 ? a : T;
        on += a >= 0 ? 1 : 0;
        soc += cost;
        mk = max(mk, cost);
    }
    o[0] = T; o[1] = on == n ? 1.0 : 0.0; o[2] = (double)on / n; o[3] = soc; o[4] = mk; o[5] = on;
    o[6] = (double)s.agent_steps[e]; o[7] = n;
    const int dn = s.density_n[e];
    o[8] = dn > 0 ? (double)s.density_sum[e] / dn : 0.0;
    o[9] = dn;
}

}  // namespace mg
