/* ssd_b200 -- C ABI of the B200-native batched SSD grid-world simulator
 * (Cleanup / Harvest) that replaces the env step + observation path of
 * drdh/Homophily-MARL.
 *
 * The reference has no FFI: its seam is the duck-typed PyMARL MultiAgentEnv
 * surface (src/envs/multiagentenv.py:6-75) implemented by MapEnv
 * (src/envs/ssd/map_env.py:874-1022) and reached through
 * src/envs/__init__.py:6-11 REGISTRY.  Every entry point below names the
 * reference method(s) it replaces; homophily_marl_b200/pymarl_env.py is the
 * Python binding a maintainer registers in that REGISTRY.
 *
 * Conventions: plain C types only; all buffers are caller-owned DEVICE
 * pointers unless the name starts with h_ (host, pinned); `stream` is a
 * cudaStream_t passed as void*; every call returns 0 or a negative SSD_ERR_*
 * code and never throws; no internal threads; calls on one handle must be
 * serialised by the caller (except ssd_step_range on disjoint ranges).  There is no CPU fallback: without a CUDA device
 * ssd_create fails with SSD_ERR_CUDA.
 * Launches of at most 2048 envs are chained to the preceding kernel of `stream` by programmatic dependent launch: their
 * CTAs may become resident early, but they read nothing before all earlier work of the stream has completed, so stream
 * order holds exactly as for a plain launch (environment variable SSD_B200_PDL=0 disables it).
 */
#ifndef SSD_B200_H
#define SSD_B200_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SSD_B200_ABI_VERSION 2

#define SSD_OK             0
#define SSD_ERR_INVALID   -1   /* bad argument / unsupported geometry            */
#define SSD_ERR_CUDA      -2   /* CUDA runtime error (see ssd_last_cuda_error)   */
#define SSD_ERR_SPAWN     -3   /* fewer spawn points than agents (map_env.py:783) */
#define SSD_ERR_MAP       -4   /* map is not wall-enclosed / unknown character   */

#define SSD_KIND_CLEANUP 0
#define SSD_KIND_HARVEST 1

#define SSD_MAX_AGENTS 16
#define SSD_MAX_CELLS  2048
#define SSD_MAX_SPAWN  32

/* cell codes stored in the device grid (one byte per cell) */
#define SSD_CELL_EMPTY  0
#define SSD_CELL_WALL   1
#define SSD_CELL_APPLE  2
#define SSD_CELL_WASTE  3
#define SSD_CELL_RIVER  4
#define SSD_CELL_STREAM 5

typedef struct ssd_handle ssd_handle;

/* Constructor arguments == the reference's env_args (config/envs/*.yaml:3-15,
 * cleanup.py:29-54, harvest.py:18-22) after the host resolved `map=` into an
 * ASCII map and probabilities. */
typedef struct ssd_config {
    int32_t kind;                 /* SSD_KIND_*                                        */
    int32_t n_envs;               /* B: env instances resident on this device          */
    int32_t n_agents;             /* num_agents (<= SSD_MAX_AGENTS)                    */
    int32_t height, width;        /* ASCII map shape                                    */
    int32_t view;                 /* view_size V; observations are (2V+1)^2            */
    int32_t episode_limit;
    int32_t fire_cost;            /* agent.py:188-190,239-241 -> 1                      */
    int32_t hit_penalty;          /* agent.py:184-186,246-248 -> 0                      */
    int32_t beam_len;             /* cleanup.py:10-11, harvest.py:11 -> 5               */
    int32_t random_spawn_point;   /* extra_args.random_spawn_point                      */
    int32_t spawn_rotation;       /* extra_args.random_spawn_rotation: 0..3, -1 = random */
    int32_t device;               /* CUDA device ordinal                                */
    int32_t reserved0;
    uint64_t seed;                /* Philox key                                         */
    uint32_t env_gid_base;        /* global id of env 0 (multi-GPU shards keep trajectories) */
    uint32_t n_waste_lut;         /* entries in thr_apple / thr_waste (= #waste points + 1)   */
    const char* ascii_map;        /* [height*width] row-major, reference alphabet       */
    const uint32_t* thr_apple;    /* cleanup: ceil(pA(h) * 2^32) per waste count h (cleanup.py:189-204) */
    const uint32_t* thr_waste;    /* cleanup: ceil(pW(h) * 2^32), 0 when isclose(pW, 0)  */
    uint32_t thr_harvest[4];      /* harvest: ceil(SPAWN_PROB[k] * 2^32) (harvest.py:118) */
    uint8_t color_lut[16][4];     /* RGB0 per colour index: 0-5 cell codes, 6 outside, 6+c agent char c */
} ssd_config;

/* Strides the caller must allocate with (all in elements of the buffer type). */
typedef struct ssd_layout {
    int32_t n_actions;        /* 9 cleanup / 8 harvest                          */
    int32_t n_cells;          /* G = H*W                                        */
    int32_t obs_n;            /* N = 2V+1                                       */
    int32_t grid_stride;      /* bytes per env in `grid` (G rounded up to 16)   */
    int32_t agent_stride;     /* entries per env in `agent` / `ep_ret`          */
    int32_t obs_plane_stride; /* bytes between colour planes (= N * obs_row_stride)   */
    int32_t obs_agent_stride; /* bytes between agents (3 planes rounded to 16)  */
    int32_t obs_env_stride;   /* bytes between envs = n_agents * obs_agent_stride */
    int32_t n_apple_pts, n_waste_pts, n_spawn_pts;
    int32_t obs_row_stride;   /* bytes between pixel rows (N rounded up to 4); pad bytes are 0 */
} ssd_layout;

/* Persistent per-env state (MapEnv.world_map, Agent.pos/orientation, _episode_steps, rewards). */
typedef struct ssd_state {
    uint8_t*  grid;      /* [B][grid_stride]   cell codes, padding bytes stay 0           */
    uint32_t* agent;     /* [B][agent_stride]  row | col<<8 | orientation<<16             */
    int32_t*  ep_ret;    /* [B][agent_stride]  episode return per agent (map_env.py:885-888) */
    int32_t*  t;         /* [B]                _episode_steps                             */
    uint32_t* tick;      /* [B]                Philox step counter (never reset)          */
    uint32_t* counts;    /* [B]                #'A' cells | #'H' cells << 16 of `grid` (the reference recounts the map every
                          *                     step: map_env.py:291-292, cleanup.py:206-212).  Whoever writes `grid` directly
                          *                     must refresh this word; ssd_reset / ssd_step keep it up to date              */
} ssd_state;

/* Outputs of one step == (reward, terminated, info) of MapEnv.step + get_obs/get_state. */
typedef struct ssd_step_out {
    int8_t*   reward;     /* [B][n]                                                     */
    uint8_t*  clean;      /* [B][n]   info["clean_num"]                                 */
    uint16_t* apple_cnt;  /* [B]      info["apple_den"] * H*W                           */
    uint8_t*  done;       /* [B]      terminated                                        */
    uint8_t*  obs;        /* [B][obs_env_stride] u8 RGB planes, rows padded to obs_row_stride (get_obs * 256) or NULL */
    uint8_t*  state_rgb;  /* [B][3][H][W] (get_state * 256) or NULL                     */
} ssd_step_out;

/* Injected, position-indexed random draws (parity harness).  NULL members fall
 * back to Philox4x32-10 keyed (seed; env_gid, tick, stream, index). */
typedef struct ssd_draws {
    const uint32_t* prio;       /* [B][n]     np.random.shuffle of movers (map_env.py:541): ascending (key, index) */
    const uint32_t* u_apple;    /* [B][G]     np.random.rand per apple cell (cleanup.py:172, harvest.py:119)     */
    const uint32_t* u_waste;    /* [B][G]     np.random.rand per waste cell (cleanup.py:183)                      */
    const uint32_t* wkey;       /* [B][G]     random.shuffle(waste_points) (cleanup.py:178): ascending (key, cell) */
    const uint32_t* spawn_key;  /* [B][n][G]  random.shuffle(spawn_points) (map_env.py:777): max (key, cell) wins  */
    const uint8_t*  rot;        /* [B][n]     np.random.randint(4) (map_env.py:789)                               */
} ssd_draws;

int ssd_abi_version(void);
const char* ssd_error_string(int code);
/* CUDA error of the most recent failing call on this thread (cudaError_t as int). */
int ssd_last_cuda_error(void);

/* u32 threshold T with (k / 2^32 < p) <=> (k < T); helper for non-Python hosts. */
uint32_t ssd_prob_to_threshold(double p);

/* CleanupEnv/HarvestEnv.__init__ (cleanup.py:29-105, harvest.py:18-48, map_env.py:116-175). */
int ssd_create(const ssd_config* cfg, ssd_handle** out);
int ssd_destroy(ssd_handle* h);
int ssd_get_layout(const ssd_handle* h, ssd_layout* out);

/* MapEnv.reset (map_env.py:986-993 -> _reset 297-326).  mask: [B] bytes, non-zero = reset this
 * env, NULL = all.  When obs != NULL the first observations are rendered too. */
int ssd_reset(ssd_handle* h, const ssd_state* st, const uint8_t* mask, const ssd_draws* draws,
              uint8_t* obs, void* stream);

/* MapEnv.step (map_env.py:874-915 -> _step 227-295) fused with get_obs (923-945) and,
 * optionally, get_state (950-957).  actions: [B][n] u8, values < n_actions; a larger value makes that agent do
 * nothing this step (the reference raises KeyError, which the Python facade reproduces before calling). */
int ssd_step(ssd_handle* h, const ssd_state* st, const uint8_t* actions, const ssd_draws* draws,
             const ssd_step_out* out, void* stream);

/* ssd_step restricted to the env instances [env_begin, env_begin + env_count).  Every pointer is still the base of the
 * whole-batch buffer.  Independent ranges may be stepped concurrently on different streams (an asynchronous sampler
 * that overlaps the policy of one group with the env step of another); a range's trajectory does not depend on how the
 * batch is split, because draws are keyed by the global env id. */
int ssd_step_range(ssd_handle* h, const ssd_state* st, const uint8_t* actions, const ssd_draws* draws,
                   const ssd_step_out* out, int32_t env_begin, int32_t env_count, void* stream);

/* get_obs / get_state without stepping (map_env.py:923-957).  Either pointer may be NULL. */
int ssd_render(ssd_handle* h, const ssd_state* st, uint8_t* obs, uint8_t* state_rgb, void* stream);

/* Host-buffer variant of ssd_step: copies h_actions -> d_actions, steps, copies every non-NULL
 * member of d_out to the same member of h_out, then synchronises the stream. */
int ssd_step_host(ssd_handle* h, const ssd_state* st, const uint8_t* h_actions, uint8_t* d_actions,
                  const ssd_step_out* d_out, const ssd_step_out* h_out, void* stream);

/* Incentive bookkeeping of the homophily learner (homophily_learner.py:98-115):
 * actions_inc [R][n][n] int64 (0 none, 1 reward, 2 punish; diagonal ignored), reward [R][n] f32
 * -> rewards_for_env, rewards_for_inc [R][n] f32, recv_sign [R][n] f32. */
int ssd_incentive(const int64_t* actions_inc, const float* reward, int64_t rows, int32_t n_agents,
                  float incentive, float cost, float ratio, int32_t max_seq_length,
                  float* rewards_for_env, float* rewards_for_inc, float* recv_sign, void* stream);

/* ---- policy side of the rollout loop ------------------------------------------------- */

/* EpsilonGreedyActionSelector.select_action (src/components/action_selectors.py:44-68) on the device.
 * q [rows][n_actions] f32; avail [rows][n_actions] i32 0/1 or NULL (everything available); picked [rows] i64.
 *   greedy      = first index of max over q with unavailable actions at -inf          (lines 57-58, 67)
 *   pick_random = u_pick[row] < epsilon                                                (lines 62-64)
 *   random      = the floor(u_act[row] * #available)-th available action               (line 66: multinomial over the mask)
 * u_pick / u_act: injected uniforms in [0,1) (both or neither); NULL -> Philox4x32-10 keyed (seed; row, counter).
 * The caller evaluates the epsilon schedule (epsilon_schedules.py) and passes 0 in test mode. */
int ssd_select_actions(const float* q, const int32_t* avail, int64_t rows, int32_t n_actions, float epsilon,
                       const float* u_pick, const float* u_act, uint64_t seed, uint64_t counter,
                       int64_t* picked, void* stream);
/* CUDA error of the most recent failing policy-side call on this thread. */
int ssd_policy_last_cuda_error(void);

/* Fused observation front end of the agent network: HomophilyAgent.rgb_preprocess == conv_to_fc
 * (src/modules/agents/homophily_agent.py:19-27; call site src/controllers/homophily_controller.py:132-136) for rollouts:
 *   u8 obs planes (the env's layout) -> /256 -> Conv2d(3,6,k3,s1)+LeakyReLU -> Flatten -> Linear(6(N-2)^2, 32)+LeakyReLU.
 * conv on the CUDA cores, the Linear contraction on tcgen05 tensor cores (tf32 hi/lo split, fp32 accumulate in TMEM).
 * conv_w [6][3][3][3], conv_b [6], fc_w [32][6(N-2)^2] (nn.Linear.weight, Flatten order), fc_b [32]: HOST pointers, copied.
 * Shapes are the reference's defaults (config/default.yaml: conv_out 6, conv_kernel 3, conv_stride 1, obs_dim_net 32). */
typedef struct ssd_frontend ssd_frontend;
int ssd_frontend_create(int32_t view, const float* conv_w, const float* conv_b, const float* fc_w, const float* fc_b,
                        float negative_slope, int32_t device, ssd_frontend** out);
/* obs: DEVICE u8, `rows` agent views `obs_agent_stride` bytes apart (ssd_layout strides); out: DEVICE f32 [rows][32].
 * One kernel launch on `stream`.  The handle owns scratch for tiles shared between CTAs: one forward at a time per handle
 * (stream order is enough); the result does not depend on timing. */
int ssd_frontend_forward(ssd_frontend* f, const uint8_t* obs, int64_t rows, int32_t obs_agent_stride, int32_t obs_plane_stride,
                         int32_t obs_row_stride, float* out, void* stream);
int ssd_frontend_destroy(ssd_frontend* f);
int ssd_frontend_last_cuda_error(void);

/* Kernels launched through this handle since creation (bench.py's gpu_launches). */
int64_t ssd_launch_count(const ssd_handle* h);

/* Checked build only (-DSSD_BOUNDS_CHECK, libssd_b200_check.so): number of shared-memory index violations seen so
 * far on the current device; -1 in the normal build. */
int64_t ssd_debug_oob_count(void);

#ifdef __cplusplus
}
#endif
#endif /* SSD_B200_H */
