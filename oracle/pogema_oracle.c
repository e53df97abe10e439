/*
 * pogema_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * CPU restatement of the POGEMA grid step the reference drives through
 * pogema_toolbox.run_episode (example.py:41-50,65; eval_configs/<set>/<set>.yaml:1-8:
 * collision_system "soft", on_target "nothing", obs_radius 5).
 *
 * PARITY UNPINNED.  The algorithm lives in the third-party package `pogema`
 * (pulled in by pogema-toolbox, pyproject.toml:18: git branch "pogema-2.0", no
 * version or commit pinned; docker/requirements.txt:15 names branch
 * "toolbox-for-pogema-2.0").  It is neither vendored under /root/reference nor
 * installable here (no network), and the reference holds no test or golden
 * vector at this boundary.  This file restates the published algorithm of
 * pogema's `Pogema.move_agents` (soft branch) + `_revert_action` as summarised
 * in SURVEY.md App. C.3, anchored on the in-repo evidence for the move table
 * (dataset/tokenizer/generate_observations.py:10-17) and (row, col) indexing
 * (dataset/lacam/inference.py:142,149-151).  It is written procedurally (claim
 * lists filled in agent-index order, descending-index conflict sweep,
 * recursive reverts) so it can be diffed against the upstream source the moment
 * that is reachable; the CUDA kernel uses the order-independent fixed-point
 * form and is tested against this file.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

static const int PG_MOVES[5][2] = {{0, 0}, {-1, 0}, {1, 0}, {0, -1}, {0, 1}};

typedef struct {
    int W;
    int32_t *cnt;      /* claimants per cell            */
    int32_t *list;     /* cell -> claimant list (cap per cell = 6, plus spill) */
    const int32_t *pos;
    int32_t *act;
    int n;
} pg_ctx;

#define PG_CAP 8  /* own cell + 4 movers is the maximum; keep slack */

static void pg_append(pg_ctx *c, int cell, int agent)
{
    c->list[cell * PG_CAP + c->cnt[cell]] = agent;
    c->cnt[cell]++;
}

static void pg_remove(pg_ctx *c, int cell, int agent)
{
    int m = c->cnt[cell];
    for (int k = 0; k < m; k++)
        if (c->list[cell * PG_CAP + k] == agent) {
            for (int t = k; t + 1 < m; t++) c->list[cell * PG_CAP + t] = c->list[cell * PG_CAP + t + 1];
            c->cnt[cell] = m - 1;
            return;
        }
}

/* pogema `_revert_action(agent_idx, used_cells, cell, actions)` */
static void pg_revert(pg_ctx *c, int agent, int cell)
{
    c->act[agent] = 0;
    pg_remove(c, cell, agent);
    int own = c->pos[2 * agent] * c->W + c->pos[2 * agent + 1];
    if (c->cnt[own] > 0) {
        pg_append(c, own, agent);
        int first = c->list[own * PG_CAP + 0];
        pg_revert(c, first, own);
    } else {
        pg_append(c, own, agent);
    }
}

/*
 * One `soft` step for one env.
 *   obst   H*W, non-zero = obstacle (padded grid)
 *   pos    n*2 (row, col), updated in place
 *   action n values; anything outside 0..4 is treated as wait
 *   moved  optional n flags: 1 if the agent's move was applied
 */
void pg_step_soft(const uint8_t *obst, int H, int W, int n, int32_t *pos, const int32_t *action, int32_t *moved)
{
    pg_ctx c;
    c.W = W; c.n = n; c.pos = pos;
    c.cnt = (int32_t *)calloc((size_t)H * W, sizeof(int32_t));
    c.list = (int32_t *)malloc(sizeof(int32_t) * (size_t)H * W * PG_CAP);
    c.act = (int32_t *)malloc(sizeof(int32_t) * (n ? n : 1));
    int32_t *tgt = (int32_t *)malloc(sizeof(int32_t) * (n ? n : 1));
    for (int i = 0; i < n; i++) c.act[i] = (action[i] >= 0 && action[i] <= 4) ? action[i] : 0;

    /* phase 1: claims in agent-index order */
    for (int i = 0; i < n; i++) {
        int x = pos[2 * i] + PG_MOVES[c.act[i]][0], y = pos[2 * i + 1] + PG_MOVES[c.act[i]][1];
        tgt[i] = x * W + y;
        pg_append(&c, tgt[i], i);
    }
    /* phase 2: edge (swap) conflicts -- i moves a->b while j moves b->a */
    int32_t *occ = (int32_t *)malloc(sizeof(int32_t) * (size_t)H * W);
    for (int k = 0; k < H * W; k++) occ[k] = -1;
    for (int i = 0; i < n; i++) occ[pos[2 * i] * W + pos[2 * i + 1]] = i;
    uint8_t *swap = (uint8_t *)calloc(n ? n : 1, 1);
    for (int i = 0; i < n; i++) {
        if (c.act[i] == 0) continue;
        int j = occ[tgt[i]];
        int own = pos[2 * i] * W + pos[2 * i + 1];
        if (j >= 0 && j != i && c.act[j] != 0 && tgt[j] == own) swap[i] = 1;
    }
    for (int i = 0; i < n; i++)
        if (swap[i]) {
            int own = pos[2 * i] * W + pos[2 * i + 1];
            pg_remove(&c, tgt[i], i);
            pg_append(&c, own, i);
            c.act[i] = 0;
            tgt[i] = own;
        }
    /* phase 3: vertex conflicts and obstacles, descending index, cascading */
    for (int i = n - 1; i >= 0; i--) {
        int x = pos[2 * i] + PG_MOVES[c.act[i]][0], y = pos[2 * i + 1] + PG_MOVES[c.act[i]][1];
        int cell = x * W + y;
        if (c.cnt[cell] > 1 || obst[cell] != 0) pg_revert(&c, i, cell);
    }
    /* phase 4: apply surviving moves simultaneously */
    for (int i = 0; i < n; i++) {
        pos[2 * i] += PG_MOVES[c.act[i]][0];
        pos[2 * i + 1] += PG_MOVES[c.act[i]][1];
        if (moved) moved[i] = c.act[i] != 0;
    }
    free(swap); free(occ); free(tgt); free(c.act); free(c.list); free(c.cnt);
}
