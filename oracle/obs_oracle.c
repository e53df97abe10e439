/*
 * obs_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Plain-C restatement of the reference's runtime observation generator /
 * tokenizer (mapf_gpt/observation_generator.{h,cpp}).  It exists only so that
 * tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg can check the
 * CUDA path; nothing under mapf_gpt_b200/ may link or call it.
 *
 * Parity status: PINNED.  tests/test_oracle_obs.py checks this file against
 *   (a) the reference's only known-answer scenario, int main() at
 *       observation_generator.cpp:530-544 (sha256 of the token row recorded in
 *       SURVEY.md section 4),
 *   (b) golden token vectors produced by the compiled reference
 *       (oracle/_ref, built by oracle/Makefile from /root/reference) and
 *       committed under tests/golden/, and
 *   (c) when oracle/_ref is present, live differential runs.
 *
 * Every function cites the reference lines it follows.  Data structures are
 * flat arrays (the reference uses vector<vector<>>, deque<string>, maps); the
 * arithmetic, visiting order and quirks are the reference's.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define OG_INF 65535u /* std::numeric_limits<uint16_t>::max(), cpp:82 etc. */

typedef struct {
    int gx, gy;                 /* goal the field was computed for            */
    int left, right, top, bottom; /* inclusive window bounds (h:67-82)        */
    int rows, cols;             /* right-left+1, bottom-top+1                 */
    uint16_t *c2g;              /* rows*cols                                  */
} og_partial;

typedef struct {
    int px, py, gx, gy;
    uint8_t hist[16];           /* token ids 44..49, oldest first (deque)     */
    uint8_t next_bits;          /* up<<3 | down<<2 | left<<1 | right          */
} og_agent;

typedef struct og {
    int H, W;
    int limit, num_agents, npa, ctx, obs_r, agents_r, gs;
    int32_t *grid;              /* H*W, 0 free                                */
    int32_t *loc;               /* H*W agent id or -1 (agents_locations)      */
    int n;
    og_agent *ag;
    og_partial *part;
    /* precompute_cost2go state (cpp:43-132) */
    int K;                      /* number of distinct precomputed cells       */
    int32_t *cell_idx;          /* H*W -> index in [0,K) or -1                */
    uint16_t *pre;              /* K*K                                        */
    /* scratch */
    int32_t *queue;             /* 2*H*W ints                                 */
    uint16_t *cm;               /* H*W cost matrix scratch                    */
    uint16_t *gcm;              /* H*W goal-block cost matrix scratch         */
} og_t;

static const int MOVES4[4][2] = {{-1, 0}, {1, 0}, {0, -1}, {0, 1}}; /* cpp:7,86,415 */

/* plain BFS over the whole grid from (sx,sy); cpp:86-107 */
static void bfs_full(const og_t *g, int sx, int sy, uint16_t *cost, int32_t *q)
{
    const int H = g->H, W = g->W;
    for (int i = 0; i < H * W; i++) cost[i] = OG_INF;
    int head = 0, tail = 0;
    q[tail++] = sx * W + sy;
    cost[sx * W + sy] = 0;
    while (head < tail) {
        int c = q[head++];
        int ci = c / W, cj = c % W;
        for (int m = 0; m < 4; m++) {
            int ni = ci + MOVES4[m][0], nj = cj + MOVES4[m][1];
            if (ni >= 0 && nj >= 0 && ni < H && nj < W) {
                if (g->grid[ni * W + nj] == 0 && cost[ni * W + nj] == OG_INF) {
                    cost[ni * W + nj] = (uint16_t)(cost[c] + 1);
                    q[tail++] = ni * W + nj;
                }
            }
        }
    }
}

/* precompute_cost2go, cpp:43-132 (the optional precomputed_cost2go.bin cache of
 * cpp:62-80,114-131 is an I/O side effect and is not restated). */
static void precompute(og_t *g)
{
    const int H = g->H, W = g->W, gs = g->gs;
    g->cell_idx = (int32_t *)malloc(sizeof(int32_t) * H * W);
    for (int i = 0; i < H * W; i++) g->cell_idx[i] = -1;
    int K = 0;
    for (int i = 0; i < H; i += gs)
        for (int j = 0; j < W; j++)
            if (g->grid[i * W + j] == 0 && g->cell_idx[i * W + j] < 0) g->cell_idx[i * W + j] = K++;
    for (int i = 0; i < H; i++)
        for (int j = 0; j < W; j += gs)
            if (g->grid[i * W + j] == 0 && g->cell_idx[i * W + j] < 0) g->cell_idx[i * W + j] = K++;
    g->K = K;
    g->pre = (uint16_t *)malloc(sizeof(uint16_t) * (size_t)(K ? K : 1) * (size_t)(K ? K : 1));
    uint16_t *cost = (uint16_t *)malloc(sizeof(uint16_t) * H * W);
    for (int c = 0; c < H * W; c++) {
        int a = g->cell_idx[c];
        if (a < 0) continue;
        bfs_full(g, c / W, c % W, cost, g->queue);
        for (int t = 0; t < H * W; t++)
            if (g->cell_idx[t] >= 0) g->pre[(size_t)a * K + g->cell_idx[t]] = cost[t];
    }
    free(cost);
}

og_t *og_create(const int32_t *grid, int H, int W, int limit, int num_agents, int npa,
                int ctx, int obs_r, int agents_r, int grid_step)
{
    og_t *g = (og_t *)calloc(1, sizeof(og_t));
    g->H = H; g->W = W;
    g->limit = limit; g->num_agents = num_agents; g->npa = npa; g->ctx = ctx;
    g->obs_r = obs_r; g->agents_r = agents_r; g->gs = grid_step;
    g->grid = (int32_t *)malloc(sizeof(int32_t) * H * W);
    memcpy(g->grid, grid, sizeof(int32_t) * H * W);
    g->loc = (int32_t *)malloc(sizeof(int32_t) * H * W);
    for (int i = 0; i < H * W; i++) g->loc[i] = -1;           /* h:116 */
    g->queue = (int32_t *)malloc(sizeof(int32_t) * 2 * H * W + 64);
    g->cm = (uint16_t *)malloc(sizeof(uint16_t) * H * W);
    g->gcm = (uint16_t *)malloc(sizeof(uint16_t) * H * W);
    /* mark_components (cpp:4-41) computes labels nobody reads: skipped. */
    precompute(g);                                            /* h:118 */
    return g;
}

static void free_agents(og_t *g)
{
    if (g->part)
        for (int i = 0; i < g->n; i++) free(g->part[i].c2g);
    free(g->part); free(g->ag);
    g->part = NULL; g->ag = NULL; g->n = 0;
}

void og_destroy(og_t *g)
{
    if (!g) return;
    free_agents(g);
    free(g->grid); free(g->loc); free(g->cell_idx); free(g->pre);
    free(g->queue); free(g->cm); free(g->gcm);
    free(g);
}

static int imin(int a, int b) { return a < b ? a : b; }
static int imax(int a, int b) { return a > b ? a : b; }

/* get_goal_border_and_cost2go, cpp:134-176: BFS restricted to the goal's
 * grid_step block; writes g->gcm (whole-grid sized, INF outside the block). */
static void goal_block_bfs(og_t *g, int gx, int gy, int *L, int *R, int *T, int *B)
{
    const int H = g->H, W = g->W, gs = g->gs;
    int left = gx / gs * gs, right = imin(left + gs, H - 1);
    int top = gy / gs * gs, bottom = imin(top + gs, W - 1);
    uint16_t *cost = g->gcm;
    int32_t *q = g->queue;
    for (int i = 0; i < H * W; i++) cost[i] = OG_INF;
    int head = 0, tail = 0;
    q[tail++] = gx * W + gy;
    cost[gx * W + gy] = 0;
    while (head < tail) {
        int c = q[head++];
        int ci = c / W, cj = c % W;
        for (int m = 0; m < 4; m++) {
            int ni = ci + MOVES4[m][0], nj = cj + MOVES4[m][1];
            if (ni >= left && nj >= top && ni <= right && nj <= bottom)
                if (g->grid[ni * W + nj] == 0 && cost[ni * W + nj] == OG_INF) {
                    cost[ni * W + nj] = (uint16_t)(cost[c] + 1);
                    q[tail++] = ni * W + nj;
                }
        }
    }
    *L = left; *R = right; *T = top; *B = bottom;
}

typedef struct { int cost, x, y; } og_seed;

static int seed_cmp(const void *a, const void *b)
{   /* std::greater<> on pair<int,pair<int,int>> => ascending (cost,x,y), cpp:223 */
    const og_seed *p = (const og_seed *)a, *q = (const og_seed *)b;
    if (p->cost != q->cost) return p->cost < q->cost ? -1 : 1;
    if (p->x != q->x) return p->x < q->x ? -1 : 1;
    if (p->y != q->y) return p->y < q->y ? -1 : 1;
    return 0;
}

/* compute_cost2go_partial, cpp:200-286 */
static void compute_partial(og_t *g, int a)
{
    const int H = g->H, W = g->W, gs = g->gs, r = g->obs_r;
    const int gx = g->ag[a].gx, gy = g->ag[a].gy, px = g->ag[a].px, py = g->ag[a].py;
    int left = imax(px - r, 0) / gs * gs, right = imin(left + 2 * gs, H - 1);
    int top = imax(py - r, 0) / gs * gs, bottom = imin(top + 2 * gs, W - 1);
    og_partial *p = &g->part[a];
    free(p->c2g);
    p->gx = gx; p->gy = gy;
    p->left = left; p->right = right; p->top = top; p->bottom = bottom;

    int gL, gR, gT, gB;
    goal_block_bfs(g, gx, gy, &gL, &gR, &gT, &gB);

    if (H <= gs && W <= gs) {                                /* cpp:214-220 */
        p->rows = H; p->cols = W;                             /* whole cost matrix */
        p->c2g = (uint16_t *)malloc(sizeof(uint16_t) * H * W);
        memcpy(p->c2g, g->gcm, sizeof(uint16_t) * H * W);
        return;
    }

    /* goal-block border cells, cpp:141-152 */
    int ngb = 0;
    int32_t *gb = (int32_t *)malloc(sizeof(int32_t) * 4 * (gs + 2) + 64);
    for (int i = gL; i <= gR; i++) {
        gb[ngb++] = i * W + gT;
        if (gT + gs < W) gb[ngb++] = i * W + gB;
    }
    for (int j = gT; j <= gB; j++) {
        gb[ngb++] = gL * W + j;
        if (gL + gs < H) gb[ngb++] = gR * W + j;
    }
    /* window border cells, get_cells_on_border cpp:178-198 (note: unclipped
     * right/bottom = left+2gs / top+2gs here, and '<' loops) */
    int wr = left + 2 * gs, wb = top + 2 * gs;
    int npb = 0;
    int32_t *pb = (int32_t *)malloc(sizeof(int32_t) * 4 * (2 * gs + 2) + 64);
    for (int i = left; i < imin(wr, H); i++) {
        pb[npb++] = i * W + top;
        if (wb < W) pb[npb++] = i * W + wb;
    }
    for (int j = top; j < imin(wb, W); j++) {
        pb[npb++] = left * W + j;
        if (wr < H) pb[npb++] = wr * W + j;
    }
    /* seeds, cpp:224-245 */
    og_seed *seeds = (og_seed *)malloc(sizeof(og_seed) * (npb + 1));
    int ns = 0;
    for (int k = 0; k < npb; k++) {
        int cell = pb[k];
        if (g->grid[cell] != 0) continue;
        uint16_t min_cost = OG_INF;
        for (int t = 0; t < ngb; t++) {
            int gc = gb[t];
            if (g->grid[gc] != 0 || g->gcm[gc] == OG_INF) continue;
            int new_cost = (int)g->gcm[gc] + (int)g->pre[(size_t)g->cell_idx[gc] * g->K + g->cell_idx[cell]];
            if ((int)min_cost > new_cost) min_cost = (uint16_t)new_cost;
        }
        if (min_cost != OG_INF) { seeds[ns].cost = min_cost; seeds[ns].x = cell / W; seeds[ns].y = cell % W; ns++; }
    }
    uint16_t *cm = g->cm;
    for (int i = 0; i < H * W; i++) cm[i] = OG_INF;
    if (gx >= left && gx <= right && gy >= top && gy <= bottom) {
        seeds[ns].cost = 0; seeds[ns].x = gx; seeds[ns].y = gy; ns++;
        cm[gx * W + gy] = 0;
    }
    qsort(seeds, ns, sizeof(og_seed), seed_cmp);
    /* bucketed multi-source BFS, cpp:246-279.  The fringe is a FIFO; seeds may
     * be pushed more than once-per-cell, so size it generously. */
    int cap = H * W + ns + 8;
    int32_t *fr = (int32_t *)malloc(sizeof(int32_t) * cap);
    int head = 0, tail = 0, sp = 0;
    if (ns > 0) { /* cpp:247 dereferences pq.top() unconditionally; ns==0 is UB there */
        fr[tail++] = seeds[0].x * W + seeds[0].y;
        cm[seeds[0].x * W + seeds[0].y] = (uint16_t)seeds[0].cost;
        sp = 1;
    }
    while (head < tail) {
        int cur = fr[head++];
        int cc = cm[cur];
        while (sp < ns && seeds[sp].cost == cc) {
            fr[tail++] = seeds[sp].x * W + seeds[sp].y;
            cm[seeds[sp].x * W + seeds[sp].y] = (uint16_t)cc;   /* overwrite, no visited check */
            sp++;
        }
        int ci = cur / W, cj = cur % W;
        for (int m = 0; m < 4; m++) { /* the reference's 5th move (0,0) is a no-op */
            int ni = ci + MOVES4[m][0], nj = cj + MOVES4[m][1];
            if (ni >= left && ni <= right && nj >= top && nj <= bottom &&
                g->grid[ni * W + nj] == 0 && cm[ni * W + nj] == OG_INF) {
                cm[ni * W + nj] = (uint16_t)(cc + 1);
                fr[tail++] = ni * W + nj;
            }
        }
        if (head == tail && sp < ns) {
            fr[tail++] = seeds[sp].x * W + seeds[sp].y;
            cm[seeds[sp].x * W + seeds[sp].y] = (uint16_t)seeds[sp].cost;
            sp++;
        }
    }
    p->rows = right - left + 1; p->cols = bottom - top + 1;
    p->c2g = (uint16_t *)malloc(sizeof(uint16_t) * p->rows * p->cols);
    for (int i = left; i <= right; i++)
        memcpy(p->c2g + (size_t)(i - left) * p->cols, cm + i * W + top, sizeof(uint16_t) * p->cols);
    free(fr); free(seeds); free(pb); free(gb);
}

/* get_distance, cpp:313-319 (exclusive upper bounds) */
static int get_distance(const og_t *g, int a, int x, int y)
{
    const og_partial *p = &g->part[a];
    if (x < p->left || x >= p->right || y < p->top || y >= p->bottom) return -1;
    return p->c2g[(x - p->left) * p->cols + (y - p->top)];
}

/* update_next_action, cpp:412-430 */
static void update_next_action(og_t *g, int a)
{
    og_agent *A = &g->ag[a];
    int cur = get_distance(g, a, A->px, A->py);
    uint8_t bits = 0;
    for (int m = 0; m < 4; m++) {
        int nb = get_distance(g, a, A->px + MOVES4[m][0], A->py + MOVES4[m][1]);
        bits = (uint8_t)((bits << 1) | ((nb >= 0 && cur > nb) ? 1 : 0));
    }
    A->next_bits = bits;
}

/* create_agents, cpp:391-410 */
void og_create_agents(og_t *g, const int32_t *pos, const int32_t *goal, int n)
{
    free_agents(g);
    g->n = n;
    g->ag = (og_agent *)calloc(n, sizeof(og_agent));
    g->part = (og_partial *)calloc(n, sizeof(og_partial));
    for (int i = 0; i < n; i++) {
        og_agent *A = &g->ag[i];
        A->px = pos[2 * i]; A->py = pos[2 * i + 1];
        A->gx = goal[2 * i]; A->gy = goal[2 * i + 1];
        for (int j = 0; j < g->npa; j++) A->hist[j] = 44;      /* "n" */
        compute_partial(g, i);
        update_next_action(g, i);
    }
}

/* update_agents, cpp:432-485 */
void og_update_agents(og_t *g, const int32_t *pos, const int32_t *goal, const int32_t *actions, int n)
{
    const int W = g->W, r = g->obs_r;
    (void)n;
    for (int i = 0; i < g->n; i++) g->loc[g->ag[i].px * W + g->ag[i].py] = -1;
    uint8_t *need = (uint8_t *)calloc(g->n ? g->n : 1, 1);
    for (int i = 0; i < g->n; i++) {
        og_agent *A = &g->ag[i];
        g->loc[pos[2 * i] * W + pos[2 * i + 1]] = i;
        A->px = pos[2 * i]; A->py = pos[2 * i + 1];
        int a = actions[i];
        uint8_t tok = (a >= 0 && a <= 4) ? (uint8_t)(45 + a) : 44; /* w u d l r / n */
        for (int j = 0; j + 1 < g->npa; j++) A->hist[j] = A->hist[j + 1];
        if (g->npa > 0) A->hist[g->npa - 1] = tok;
        if (A->gx != goal[2 * i] || A->gy != goal[2 * i + 1]) {
            A->gx = goal[2 * i]; A->gy = goal[2 * i + 1];
            need[i] = 1;
        } else {
            const og_partial *p = &g->part[i];
            if (A->px - r < p->left || A->px + r > p->right || A->py - r < p->top || A->py + r > p->bottom)
                need[i] = 1;
        }
    }
    for (int i = 0; i < g->n; i++) if (need[i]) compute_partial(g, i);
    for (int i = 0; i < g->n; i++) update_next_action(g, i);
    free(need);
}

static int tok_int(int v, int limit)
{   /* Encoder::Encoder int_vocab, cpp:323-341: -L..L -> 0..2L, -4L, -2L, +2L follow */
    if (v >= -limit && v <= limit) return v + limit;
    if (v == -4 * limit) return 2 * limit + 1;
    if (v == -2 * limit) return 2 * limit + 2;
    if (v == 2 * limit) return 2 * limit + 3;
    return -1; /* int_vocab.at() would throw */
}

static int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }

typedef struct { int d, id; } og_cand;
static int cand_cmp(const void *a, const void *b)
{
    const og_cand *p = (const og_cand *)a, *q = (const og_cand *)b;
    if (p->d != q->d) return p->d < q->d ? -1 : 1;
    return p->id < q->id ? -1 : (p->id > q->id ? 1 : 0);
}

/* generate_observations, cpp:516-528 = generate_cost2go_obs (cpp:288-311) +
 * get_agents_info (cpp:487-514) + Encoder::encode (cpp:352-389).
 * Returns 0, or -1 if some value falls outside the vocabulary. */
int og_generate_observations(og_t *g, int32_t *out)
{
    const int W = g->W, r = g->obs_r, ar = g->agents_r, L = g->limit;
    const int base_act = 2 * L + 4;            /* 44 for L=20 */
    const int base_next = base_act + 6;        /* 50 */
    const int pad_tok = base_next + 16;        /* 66 */
    const int slot = 5 + g->npa;
    og_cand *cand = (og_cand *)malloc(sizeof(og_cand) * (2 * ar + 1) * (2 * ar + 1));
    int rc = 0;
    for (int a = 0; a < g->n; a++) {
        int32_t *o = out + (size_t)a * 256;
        const og_agent *A = &g->ag[a];
        const og_partial *p = &g->part[a];
        int k = 0;
        /* cost2go window */
        int x = A->px - p->left - r, y = A->py - p->top - r;
        int mid = p->c2g[(x + r) * p->cols + (y + r)];
        for (int i = 0; i <= 2 * r; i++)
            for (int j = 0; j <= 2 * r; j++) {
                int v = p->c2g[(x + i) * p->cols + (y + j)];
                if (v != (int)OG_INF) {
                    v -= mid;
                    v = v > L ? 2 * L : (v < -L ? -2 * L : v);
                } else v = -4 * L;
                int t = tok_int(v, L);
                if (t < 0) rc = -1;
                o[k++] = t;
            }
        /* neighbours */
        int nc = 0;
        for (int i = -ar; i <= ar; i++)
            for (int j = -ar; j <= ar; j++) {
                int id = g->loc[(A->px + i) * W + (A->py + j)];
                if (id >= 0) {
                    cand[nc].id = id;
                    cand[nc].d = abs(g->ag[id].px - A->px) + abs(g->ag[id].py - A->py);
                    nc++;
                }
            }
        qsort(cand, nc, sizeof(og_cand), cand_cmp);
        int take = nc < g->num_agents ? nc : g->num_agents;
        for (int c = 0; c < take; c++) {
            const og_agent *Bg = &g->ag[cand[c].id];
            int t0 = tok_int(Bg->px - A->px, L), t1 = tok_int(Bg->py - A->py, L);
            if (t0 < 0 || t1 < 0) rc = -1;
            o[k++] = t0; o[k++] = t1;
            o[k++] = tok_int(clampi(Bg->gx - A->px, -L, L), L);
            o[k++] = tok_int(clampi(Bg->gy - A->py, -L, L), L);
            for (int h = 0; h < g->npa; h++) o[k++] = Bg->hist[h] - 44 + base_act;
            o[k++] = base_next + Bg->next_bits;
        }
        for (int c = take * slot; c < g->num_agents * slot; c++) o[k++] = pad_tok;
        while (k < 256) o[k++] = pad_tok;                      /* cpp:386-387 */
    }
    free(cand);
    return rc;
}

/* test helper: copy agent a's field window (for BFS-kernel parity tests) */
int og_get_partial(const og_t *g, int a, int32_t *bounds4, uint16_t *dst, int cap)
{
    const og_partial *p = &g->part[a];
    bounds4[0] = p->left; bounds4[1] = p->right; bounds4[2] = p->top; bounds4[3] = p->bottom;
    int n = p->rows * p->cols;
    if (dst && cap >= n) memcpy(dst, p->c2g, sizeof(uint16_t) * n);
    return n;
}
