/* TEST INFRASTRUCTURE ONLY -- CPU restatement ("oracle") of the reference SSD
 * env step + observation path of drdh/Homophily-MARL.  Only tests/, the
 * smoke() check and bench.py's cpu_baseline / --impl reference legs may link
 * or call this.  The product path (homophily_marl_b200/csrc) never does.
 *
 * Parity status: PINNED.  The restatement is checked (a) against golden traces
 * generated from the unmodified reference (tests/golden/, generator committed)
 * and (b) live against the reference in the dev container
 * (tests/test_oracle_vs_reference.py, skipped where /root/reference is absent).
 *
 * All file:line citations are relative to the reference root.
 */
#ifndef SSD_ORACLE_H
#define SSD_ORACLE_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SSDO_MAX_CELLS   2048
#define SSDO_MAX_AGENTS  16
#define SSDO_KIND_CLEANUP 0
#define SSDO_KIND_HARVEST 1

/* cell codes */
enum { SSDO_EMPTY = 0, SSDO_WALL = 1, SSDO_APPLE = 2, SSDO_WASTE = 3, SSDO_RIVER = 4, SSDO_STREAM = 5 };
/* orientation index == position in map_env.py:28-31 ORIENTATIONS */
enum { SSDO_LEFT = 0, SSDO_RIGHT = 1, SSDO_UP = 2, SSDO_DOWN = 3 };

typedef struct ssdo_map {
    int kind, H, W, G, n, V, N, episode_limit, full_color;
    int fire_cost, hit_penalty, beam_len, n_actions;
    uint8_t base[SSDO_MAX_CELLS];          /* ascii chars of the base map            */
    uint8_t wall[SSDO_MAX_CELLS];
    int n_apple, n_waste, n_spawn;
    int apple_pts[SSDO_MAX_CELLS];         /* row-major cell indices                 */
    int waste_pts[SSDO_MAX_CELLS];
    int spawn_pts[SSDO_MAX_CELLS];
    double thr_depletion, thr_restoration, p_waste, p_apple, spawn_prob[4];
    uint32_t thr_apple_lut[SSDO_MAX_CELLS + 1]; /* cleanup: ceil(pA(h) * 2^32), h = #waste */
    uint32_t thr_waste_lut[SSDO_MAX_CELLS + 1]; /* cleanup: ceil(pW(h) * 2^32)             */
    uint32_t thr_harvest[4];               /* harvest: ceil(SPAWN_PROB[k] * 2^32)    */
    uint8_t color[16][3];                  /* 0-5 cell codes, 6 void, 6+c agent char c=1..9 */
} ssdo_map;

/* injected, position-indexed draws; any NULL member => Philox4x32-10 */
typedef struct ssdo_draws {
    const uint32_t* prio;      /* [n]     mover priority key (ascending, ties by index) */
    const uint32_t* u_apple;   /* [G]     apple spawn draw per cell                     */
    const uint32_t* u_waste;   /* [G]     waste spawn draw per cell                     */
    const uint32_t* wkey;      /* [G]     waste visiting order key (ascending, ties by cell) */
    const uint32_t* spawn_key; /* [n][G]  reset: agent i takes free spawn point with max (key, cell) */
    const uint8_t*  rot;       /* [n]     reset: orientation index                      */
} ssdo_draws;

typedef struct ssdo_env {          /* one env instance, plain arrays */
    uint8_t  grid[SSDO_MAX_CELLS];
    int32_t  pos[SSDO_MAX_AGENTS];      /* cell index r*W+c */
    uint8_t  orient[SSDO_MAX_AGENTS];
    int32_t  ep_ret[SSDO_MAX_AGENTS];
    int32_t  t;                         /* _episode_steps */
    uint32_t tick;                      /* Philox step counter, never reset */
    int32_t  error;                     /* set if a reference KeyError path would be hit */
} ssdo_env;

void ssdo_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]);

int  ssdo_map_init(ssdo_map* m, int kind, const char* ascii, int H, int W, int n, int V,
                   int episode_limit, int full_color,
                   double thr_depletion, double thr_restoration, double p_waste, double p_apple,
                   const double spawn_prob[4], int fire_cost, int hit_penalty);

void ssdo_reset(const ssdo_map* m, ssdo_env* e, int random_spawn_point, int spawn_rotation /* -1 = random */,
                const ssdo_draws* d, uint64_t seed, uint32_t env_gid);

void ssdo_step(const ssdo_map* m, ssdo_env* e, const uint8_t* actions, const ssdo_draws* d,
               uint64_t seed, uint32_t env_gid,
               int8_t* reward, uint8_t* clean, uint16_t* apple_cnt, uint8_t* done);

void ssdo_render_obs(const ssdo_map* m, const ssdo_env* e, uint8_t* obs /* [n][3][N][N] */);
void ssdo_render_state(const ssdo_map* m, const ssdo_env* e, uint8_t* out /* [3][H][W] */);

/* batch helpers (OpenMP over envs); layouts are dense, see ssd_oracle.c */
void ssdo_batch_reset(const ssdo_map* m, ssdo_env* envs, int B, int random_spawn_point, int spawn_rotation,
                      uint64_t seed, uint32_t env_gid0, int threads);
void ssdo_batch_step(const ssdo_map* m, ssdo_env* envs, int B, const uint8_t* actions /* [B][n] */,
                     uint64_t seed, uint32_t env_gid0,
                     int8_t* reward, uint8_t* clean, uint16_t* apple_cnt, uint8_t* done,
                     uint8_t* obs /* [B][n][3][N][N] or NULL */, int threads);
unsigned long ssdo_sizeof_map(void);
unsigned long ssdo_sizeof_env(void);

#ifdef __cplusplus
}
#endif
#endif
