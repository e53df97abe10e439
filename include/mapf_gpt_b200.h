/*
 * mapf_gpt_b200.h -- C ABI of the B200 rollout engine (libmapf_gpt_b200.so).
 *
 * Drop-in boundary for the hot path of CognitiveAISystems/MAPF-GPT:
 *   POGEMA grid step -> per-agent FOV observation/tokenizer -> GPT forward -> action.
 * Plain pointers and sizes only; no torch / pybind types.  All functions return 0 on
 * success and a negative code on failure (mg_last_error() gives the text), except
 * constructors, which return NULL on failure.  The library never falls back to the
 * CPU: without a CUDA device every compute entry point fails with MG_ERR_CUDA.
 *
 * Reference interfaces replaced (paths relative to the reference repo root):
 *   mg_gen_*          <- pybind module `observation_generator`
 *                        (mapf_gpt/observation_generator.cpp:548-563): InputParameters,
 *                        ObservationGenerator(grid, params), create_agents,
 *                        update_agents, generate_observations.
 *   mg_engine_*       <- the same four verbs batched over E environment slots
 *                        (MAPFGPTInference._prepare_inputs, mapf_gpt/inference.py:127-146),
 *                        GPT.act (mapf_gpt/model.py:244-260) behind
 *                        MAPFGPTInference._forward_batch (inference.py:87-101), and the
 *                        POGEMA `soft` step the harness runs between act() calls
 *                        (example.py:41-50,65).
 * INTEGRATION.md shows the ctypes binding a reference maintainer would add.
 *
 * Coordinates are (row, col) in the PADDED grid (POGEMA pads by obs_radius=5,
 * inference.py:130-131).  Actions: 0 wait, 1 up(-1,0), 2 down(+1,0), 3 left(0,-1),
 * 4 right(0,+1); anything else is "none" (observation_generator.cpp:443-462).
 */
#ifndef MAPF_GPT_B200_H
#define MAPF_GPT_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MG_OK 0
#define MG_ERR_ARG (-1)     /* bad argument / unsupported configuration            */
#define MG_ERR_CUDA (-2)    /* CUDA runtime error or no device                      */
#define MG_ERR_STATE (-3)   /* call order (e.g. act before reset / load_model)      */
#define MG_ERR_VOCAB (-4)   /* a value left the token vocabulary (int_vocab.at throws, cpp:357-361) */
#define MG_ERR_NUMERIC (-5) /* a logit is not finite (torch.multinomial raises on inf/nan probabilities, model.py:257) */

#define MG_CONTEXT 256      /* tokens per observation (cpp:386-387, inference.py:22) */
#define MG_VOCAB 67         /* App. A of SURVEY.md                                   */
#define MG_METRIC_COLS 10   /* doubles per slot written by mg_engine_get_metrics     */

/* InputParameters, observation_generator.h:22-40 (field order of the pybind ctor) */
typedef struct mg_params {
    int32_t cost2go_value_limit;   /* 20 */
    int32_t num_agents;            /* 13 */
    int32_t num_previous_actions;  /* 5  */
    int32_t context_size;          /* 256 */
    int32_t obs_radius;            /* 5  */
    int32_t agents_radius;         /* 5  */
    int32_t grid_step;             /* 64 */
    int32_t save_cost2go;          /* bool; "precomputed_cost2go.bin" in the working directory (cpp:62-80,114-131): loaded when
                                    * present, written after the computation otherwise.  Only maps wider than 74 padded cells
                                    * have such a table here; a table of the wrong shape is an MG_ERR_ARG, not silently used */
} mg_params;

/* GPTConfig, mapf_gpt/model.py:107-115 (dropout must be 0, bias must be false) */
typedef struct mg_model_config {
    int32_t block_size;  /* 256 */
    int32_t vocab_size;  /* 67  */
    int32_t n_layer;
    int32_t n_head;
    int32_t n_embd;
} mg_model_config;

typedef struct mg_engine mg_engine;
typedef struct mg_gen mg_gen;

/* ---- library ------------------------------------------------------------------- */
int mg_version(void);
const char *mg_last_error(void);
void mg_default_params(mg_params *p);
int mg_device_count(void);

/* ---- engine lifetime ------------------------------------------------------------ */
/* Capacity is fixed at creation: E env slots x N agents on H x W padded grids, 11..512 cells per side.
 * H, W <= 74: the cost-to-go field is one BFS per agent per goal over the whole grid.  Larger grids run the
 * windowed machinery of observation_generator.cpp:43-286 (per-map precompute tables, per-agent partial fields that are
 * recomputed when the goal changes or the FOV leaves the window). */
mg_engine *mg_engine_create(int device, int max_envs, int max_agents, int H, int W, const mg_params *params);
void mg_engine_destroy(mg_engine *e);

/* ---- policy network (GPT, model.py) --------------------------------------------- */
/* `weights` is one host fp32 buffer holding, in this order (shapes as in the checkpoint,
 * SURVEY App. D.3): wte[V,C], wpe[T,C], then per layer ln_1[C], c_attn[3C,C], attn c_proj[C,C],
 * ln_2[C], c_fc[4C,C], mlp c_proj[C,4C]; finally ln_f[C].  lm_head is tied to wte. */
int mg_engine_load_model(mg_engine *e, const mg_model_config *cfg, const float *weights, size_t n_floats);
size_t mg_model_num_floats(const mg_model_config *cfg);

/* ---- environments: ObservationGenerator ctor + create_agents, batched -------------- */
/* obstacles: n_envs*H*W bytes (non-zero = obstacle), pos/goal: n_envs*n_agents*2 int32, all HOST.
 * Resets slots [first_env, first_env+n_envs): uploads the grid, runs the cost-to-go BFS for every
 * agent (cpp:200-286 collapses to a goal BFS for H,W<=74), history <- "n"x5, last action <- -1,
 * episode counters <- 0.  n_agents <= max_agents may differ per call (ragged slots). */
int mg_engine_reset(mg_engine *e, int first_env, int n_envs, int n_agents,
                    const uint8_t *obstacles, const int32_t *pos_xy, const int32_t *goal_xy);
int mg_engine_num_envs(const mg_engine *e);   /* highest reset slot + 1 */
/* MAPFGPTInference.reset_states (inference.py:174-177) forgets every env slot but keeps the model: drops all slots (num_envs
 * becomes 0) while the loaded weights, the workspace and the per-map precompute tables of large maps stay on the device. */
int mg_engine_clear(mg_engine *e);

/* ---- the four verbs of the reference generator, batched over all reset slots -------- */
/* update_agents (cpp:432-485).  Any pointer may be NULL = "keep what the device holds"
 * (positions moved by mg_engine_env_step, goals fixed, actions = the ones last sampled).
 * HOST int32 arrays of num_envs*max_agents(*2) entries, slot-major, ragged slots padded. */
int mg_engine_update_agents(mg_engine *e, const int32_t *pos_xy, const int32_t *goal_xy, const int32_t *actions);
/* generate_observations (cpp:516-528): tokens stay on the device; out (HOST, may be NULL)
 * receives num_envs*max_agents*256 int8 token ids. */
int mg_engine_generate_observations(mg_engine *e, int8_t *out_tokens);

/* ---- GPT.act (model.py:244-260) over the tokens on the device ----------------------- */
/* mode 0: greedy argmax (do_sample=False); mode 1: sample with the engine's counter-based
 * Philox stream keyed by (seed, env, agent, step); mode 2: sample with caller-supplied
 * q ~ Exp(1) (HOST fp32, num_rows*5 used columns laid out [row][5]) so that the draw equals
 * torch.multinomial's argmax(p/q) for the same q.  actions_out / logits_out: HOST, may be NULL
 * (num_envs*max_agents int32 / *5 fp32). */
int mg_engine_act(mg_engine *e, int mode, const float *q_exp, int32_t *actions_out, float *logits_out);
/* act_batch may address a subset of the slots (inference.py:151-172 touches only the slots it is given): slots whose mask
 * byte is 0 are skipped by update / tokenizer / sampling / step until the mask changes.  mask: HOST, num_envs bytes; NULL = all. */
int mg_engine_set_active(mg_engine *e, const uint8_t *mask);
int mg_engine_set_seed(mg_engine *e, uint64_t seed);
/* global id of slot 0 (env-sharded multi-GPU runs: the Philox stream follows the env, not the rank) */
int mg_engine_set_env_offset(mg_engine *e, int first_global_env);
/* truncation horizon (eval_configs max_episode_steps: 128, 256 on movingai); 0 = unlimited */
int mg_engine_set_max_episode_steps(mg_engine *e, int n);

/* forward only, for tests: tokens HOST int8 [n_rows][256] -> logits HOST fp32 [n_rows][5] */
int mg_engine_forward_tokens(mg_engine *e, const int8_t *tokens, int n_rows, float *logits_out);
/* Validation loss of the training objective on dataset rows (train.py:244-258 estimate_loss; model.py:180-183 with targets:
 * cross-entropy over the 67 tied lm_head logits of position 255, ignore_index -1; dataset/fast_data_loader.py:57 puts the
 * ground-truth action there).  tokens: int8 [n_rows][256] as the Arrow shards hold them (generate_dataset.py:188-191),
 * targets: int8 [n_rows] (-1 = ignore -> loss 0); loss_out: float [n_rows]; pred_out: int32 [n_rows] arg-max action (0..4). */
int mg_engine_eval_tokens(mg_engine *e, const int8_t *tokens, const int8_t *targets, int n_rows, float *loss_out, int32_t *pred_out);

/* ---- POGEMA `soft` step on the device -------------------------- */
/* Applies `actions` (HOST, may be NULL = the actions mg_engine_act sampled) with collision
 * resolution, advances episode counters.  pos_out (HOST, may be NULL) gets the new positions. */
int mg_engine_env_step(mg_engine *e, const int32_t *actions, int32_t *pos_out);

/* one full device-resident timestep: update -> tokenize -> forward -> sample -> move.
 * No host<->device traffic.  `n_steps` timesteps are enqueued back to back. */
int mg_engine_rollout(mg_engine *e, int n_steps, int mode);

/* the same through HOST buffers (the e2e path of MAPFGPTInference.act_batch): positions and
 * goals in, actions out, one call per timestep. */
int mg_engine_act_host(mg_engine *e, const int32_t *pos_xy, const int32_t *goal_xy,
                       int mode, const float *q_exp, int32_t *actions_out);

/* ---- state read-back ------------------------------------------------------------------ */
int mg_engine_get_positions(mg_engine *e, int32_t *pos_xy_out);
int mg_engine_get_tokens(mg_engine *e, int8_t *out_tokens);
int mg_engine_get_cost2go(mg_engine *e, int env, int agent, uint16_t *out_hw);   /* grids <= 74 cells per side */
/* Cost2GoPartial (observation_generator.h:67-82): window bounds {left,right,top,bottom} (inclusive) and the rows x cols
 * field of one agent; returns rows*cols (negative on error).  Works for every grid size. */
int mg_engine_get_partial(mg_engine *e, int env, int agent, int32_t *bounds4, uint16_t *out, int cap);
/* per-slot episode metrics, doubles [num_envs][MG_METRIC_COLS]:
 * 0 ep_length, 1 CSR, 2 ISR, 3 SoC, 4 makespan, 5 agents on goal now, 6 agent-steps executed, 7 n_agents,
 * 8 avg_agents_density (pogema AgentsDensityWrapper, experiment_setup/create_env.py:36-40), 9 observations averaged in 8 */
int mg_engine_get_metrics(mg_engine *e, double *out);
int mg_engine_synchronize(mg_engine *e);
/* CUDA-event time of the last mg_engine_rollout / mg_engine_act call in ms, and per-phase
 * times [observe, forward, sample+step] when profiling is on (mg_engine_set_profiling). */
int mg_engine_set_profiling(mg_engine *e, int on);
int mg_engine_last_timing(mg_engine *e, float *total_ms, float *phases_ms3);
/* number of kernels this library launched since creation (bench.py's gpu_launches) */
long long mg_engine_launch_count(const mg_engine *e);
/* stream lanes the forward uses when a timestep has more than one chunk (2: attention of one chunk overlaps the post-attention
 * kernels of the other on the same SMs; per-kernel event timing -- profiling on -- always runs single-lane); DESIGN.md section 4 */
int mg_engine_num_lanes(const mg_engine *e);
/* timing of the dominant kernels: name list is fixed, see DESIGN.md */
int mg_engine_kernel_times(mg_engine *e, float *ms_out, int n);

/* ---- single-env twin of the pybind class (observation_generator.cpp:548-563) ------------ */
mg_gen *mg_gen_create(const int32_t *grid, int H, int W, const mg_params *params);      /* ctor, h:112 */
int mg_gen_create_agents(mg_gen *g, const int32_t *pos_xy, const int32_t *goal_xy, int n);   /* cpp:391 */
int mg_gen_update_agents(mg_gen *g, const int32_t *pos_xy, const int32_t *goal_xy,
                         const int32_t *actions, int n);                                 /* cpp:432 */
int mg_gen_generate_observations(mg_gen *g, int32_t *out /* n x 256 */);                /* cpp:516 */
void mg_gen_destroy(mg_gen *g);

/* ---- kernel-level test hooks (device pointers; used by tests/ only) ---------------------- */
/* C[M,N] (fp32 row-major) = A[M,K] (bf16 row-major) * B[N,K]^T (bf16 row-major) through the
 * production tcgen05 GEMM (operands are re-packed into tile images on the device first). */
int mg_test_gemm(int device, const void *A_bf16, const void *B_bf16, float *C, int M, int N, int K, int variant);
/* softmax(Q K^T / sqrt(hs)) V for n_seq*n_head blocks of 256 tokens through the production
 * attention kernel.  q,k,v,out: bf16 [n_seq][n_head][256][hs] row-major device pointers. */
int mg_test_attention(int device, const void *q, const void *k, const void *v, void *out,
                      int n_seq, int n_head, int hs);
/* the same with the kernel variant chosen explicitly: 0 = max-subtracting softmax, 1 = max-free softmax on pre-scaled q (the
 * engine's default; the hook multiplies q by log2(e)/sqrt(hs) first, as the engine folds that factor into Wq), 2 = the
 * classic one-CTA-per-item kernel */
int mg_test_attention_ex(int device, const void *q, const void *k, const void *v, void *out,
                         int n_seq, int n_head, int hs, int variant);
/* precision mode of engines created afterwards by this process: 0 = bf16 tensor-core path (default), 1 = fp32 CUDA-core
 * verification path (also selected by MAPF_GPT_B200_PRECISION=fp32); returns the previous value */
int mg_set_precision(int mode);

/* UMMA-rate microbenchmark: average ms per launch of the production GEMM kernel (tile width BN) on dummy operands */
int mg_test_gemm_time(int device, int M, int N, int K, int BN, int iters, float *ms);
/* pure UMMA rate: one thread issues iters*4 UMMAs (M128 x N x K16) on smem-resident operands; cycles2 = {issue->done, issue loop} */
int mg_test_umma_rate(int device, int N, int iters, int ctas, long long *cycles2);
/* profiling aid, out[17][128]: clock64() stamps of 4 CTAs of the last fused post-attention / attention launch (rows 0-3),
 * per-CTA attention load-balance data (rows 4-15), phase-cycle accumulators of -DMG_PHASE_PROF builds (row 16) */
int mg_test_timeline(mg_engine *e, int enable, long long *out);

#ifdef __cplusplus
}
#endif
#endif /* MAPF_GPT_B200_H */
