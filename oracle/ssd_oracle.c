/* TEST INFRASTRUCTURE ONLY -- see ssd_oracle.h.  Plain C restatement of the
 * reference algorithm; sequential, one env at a time, written to mirror the
 * reference's control flow (NOT the CUDA design) so it can be audited line by
 * line.  Citations: src/envs/ssd/{map_env,cleanup,harvest,agent}.py,
 * src/utils/utility_funcs.py of drdh/Homophily-MARL.
 *
 * Build: gcc -O2 -ffp-contract=off -fopenmp -shared -fPIC  (oracle/Makefile)
 */
#include "ssd_oracle.h"
#include <math.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* ------------------------------------------------------------------ Philox */
/* Philox4x32-10 (Salmon et al., SC'11; Random123 reference constants). */
void ssdo_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) {
    uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3];
    uint32_t k0 = key[0], k1 = key[1];
    for (int r = 0; r < 10; ++r) {
        uint64_t p0 = (uint64_t)0xD2511F53u * c0;
        uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
        uint32_t n1 = (uint32_t)p1;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
        uint32_t n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

/* Draw streams (the contract shared with the CUDA path):
 *   key = (seed_lo, seed_hi), counter = (env_gid, tick, stream, j)
 *   stream 0  mover priority : agent i       -> j = i>>2,  word i&3
 *   stream 1  apple draw     : k-th apple pt -> j = k>>2,  word k&3
 *   stream 2  waste          : k-th waste pt -> j = k>>1,  u = word 2(k&1), order key = word 2(k&1)+1
 *   stream 3  spawn key      : idx = i*S+s   -> j = idx>>2, word idx&3
 *   stream 4  spawn rotation : agent i       -> j = i>>2,  word i&3, rot = word>>30          */
static uint32_t philox_word(uint64_t seed, uint32_t gid, uint32_t tick, uint32_t stream, uint32_t idx) {
    uint32_t ctr[4] = { gid, tick, stream, idx >> 2 }, key[2] = { (uint32_t)seed, (uint32_t)(seed >> 32) }, o[4];
    ssdo_philox4x32_10(ctr, key, o);
    return o[idx & 3];
}

/* ------------------------------------------------------------------ tables */
/* map_env.py:20-31, 826-861 dumped as (drow, dcol); SURVEY Appendix A.1.   */
static const int MOVE_D[4][4][2] = {
    /* LEFT  */ { {0, 1}, {0, -1}, {-1, 0}, {1, 0} },
    /* RIGHT */ { {0, -1}, {0, 1}, {1, 0}, {-1, 0} },
    /* UP    */ { {-1, 0}, {1, 0}, {0, -1}, {0, 1} },
    /* DOWN  */ { {1, 0}, {-1, 0}, {0, 1}, {0, -1} } };
static const int FIRE_D[4][2]  = { {-1, 0}, {1, 0}, {0, -1}, {0, 1} };   /* ORIENTATIONS[o]            */
static const int RIGHT_D[4][2] = { {0, -1}, {0, 1}, {1, 0}, {-1, 0} };   /* rotate_right(ORIENTATIONS) */
static const uint8_t CW_T[4]  = { SSDO_UP, SSDO_DOWN, SSDO_RIGHT, SSDO_LEFT };   /* map_env.py:853-861 */
static const uint8_t CCW_T[4] = { SSDO_DOWN, SSDO_UP, SSDO_LEFT, SSDO_RIGHT };   /* map_env.py:844-852 */

static uint32_t prob_to_thr(double p) {
    if (!(p > 0.0)) return 0u;
    double t = ceil(p * 4294967296.0);
    if (t >= 4294967295.0) return 0xFFFFFFFFu;
    return (uint32_t)t;
}

static uint8_t base_to_code(int kind, uint8_t ch) {
    /* map_env.py:817-820 build_walls; cleanup.py:117-125; harvest.py:74-77 */
    if (ch == '@') return SSDO_WALL;
    if (kind == SSDO_KIND_CLEANUP) {
        if (ch == 'H') return SSDO_WASTE;
        if (ch == 'R') return SSDO_RIVER;
        if (ch == 'S') return SSDO_STREAM;
        return SSDO_EMPTY;
    }
    if (ch == 'A') return SSDO_APPLE;
    return SSDO_EMPTY;
}

int ssdo_map_init(ssdo_map* m, int kind, const char* ascii, int H, int W, int n, int V,
                  int episode_limit, int full_color,
                  double thr_depletion, double thr_restoration, double p_waste, double p_apple,
                  const double spawn_prob[4], int fire_cost, int hit_penalty) {
    memset(m, 0, sizeof(*m));
    if (H * W > SSDO_MAX_CELLS || n > SSDO_MAX_AGENTS || n < 1) return -1;
    m->kind = kind; m->H = H; m->W = W; m->G = H * W; m->n = n; m->V = V; m->N = 2 * V + 1;
    m->episode_limit = episode_limit; m->full_color = full_color;
    m->fire_cost = fire_cost; m->hit_penalty = hit_penalty; m->beam_len = 5;   /* cleanup.py:10-11, harvest.py:11 */
    m->n_actions = kind == SSDO_KIND_CLEANUP ? 9 : 8;                          /* agent.py:153-154, 207-209 */
    m->thr_depletion = thr_depletion; m->thr_restoration = thr_restoration;
    m->p_waste = p_waste; m->p_apple = p_apple;
    for (int c = 0; c < m->G; ++c) {
        uint8_t ch = (uint8_t)ascii[c];
        m->base[c] = ch;
        m->wall[c] = ch == '@';
        if (ch == 'P') m->spawn_pts[m->n_spawn++] = c;                          /* map_env.py:143-146 */
        if (kind == SSDO_KIND_CLEANUP) {
            if (ch == 'B') m->apple_pts[m->n_apple++] = c;                      /* cleanup.py:78-90 */
            if (ch == 'H') m->waste_pts[m->n_waste++] = c;
        } else if (ch == 'A') m->apple_pts[m->n_apple++] = c;                   /* harvest.py:30-34 */
    }
    if (m->n_spawn < n) return -2;                                              /* map_env.py:783 */
    /* cleanup.py:189-204 evaluated for every waste count h */
    int P = m->n_waste;
    for (int h = 0; h <= P; ++h) {
        double density = 0.0, pA, pW;
        if (P > 0) density = 1.0 - (double)(P - h) / (double)P;
        if (density >= thr_depletion) { pA = 0.0; pW = 0.0; }
        else {
            pW = p_waste;
            if (density <= thr_restoration) pA = p_apple;
            else pA = (1.0 - (density - thr_restoration) / (thr_depletion - thr_restoration)) * p_apple;
        }
        m->thr_apple_lut[h] = prob_to_thr(pA);
        /* np.isclose(pW, 0) (cleanup.py:177): |pW| <= 1e-8 skips the waste phase */
        m->thr_waste_lut[h] = fabs(pW) <= 1e-8 ? 0u : prob_to_thr(pW);
    }
    for (int k = 0; k < 4; ++k) {
        m->spawn_prob[k] = spawn_prob ? spawn_prob[k] : 0.0;
        m->thr_harvest[k] = prob_to_thr(m->spawn_prob[k]);
    }
    /* colour tables: map_env.py:33-62, cleanup.py:14-17, 92-105, harvest.py:37-48 */
    static const uint8_t AGENT_FULL[10][3] = { {0, 0, 0}, {159, 67, 255}, {2, 81, 154}, {204, 0, 204}, {216, 30, 54},
        {254, 151, 0}, {205, 155, 155}, {99, 99, 255}, {250, 204, 255}, {238, 223, 16} };
    if (full_color) {
        uint8_t c[6][3] = { {0, 0, 0}, {180, 180, 180}, {0, 255, 0}, {99, 156, 194}, {113, 75, 24}, {113, 75, 24} };
        memcpy(m->color, c, sizeof(c));
        for (int a = 1; a <= 9; ++a) memcpy(m->color[6 + a], AGENT_FULL[a], 3);
    } else {
        m->color[SSDO_WALL][2] = 255;
        m->color[SSDO_APPLE][1] = 255;
        if (kind == SSDO_KIND_CLEANUP) m->color[SSDO_WASTE][0] = 255;
        for (int a = 1; a <= 9; ++a) m->color[6 + a][2] = 255;
    }
    return 0;
}

/* agent i is drawn as str(int(agent_id[-1]) + 1) clipped to one char: map_env.py:370 */
static int agent_char(int i) { int v = i % 10 + 1; return v == 10 ? 1 : v; }

static int occupied(const ssdo_map* m, const ssdo_env* e, int cell) {
    for (int j = 0; j < m->n; ++j) if (e->pos[j] == cell) return 1;
    return 0;
}
static int by_pos(const int32_t* pos, int n, int cell) {        /* dict comprehension: last index wins */
    int occ = -1;
    for (int j = 0; j < n; ++j) if (pos[j] == cell) occ = j;
    return occ;
}

/* ------------------------------------------------------------ update_moves */
/* map_env.py:477-661. */
static void update_moves(const ssdo_map* m, ssdo_env* e, const uint8_t* act, const uint32_t* prio) {
    const int n = m->n, W = m->W;
    int prop[SSDO_MAX_AGENTS], mv[SSDO_MAX_AGENTS], in_moves[SSDO_MAX_AGENTS];
    int order[SSDO_MAX_AGENTS], n_movers = 0;
    for (int i = 0; i < n; ++i) {
        int a = act[i];
        in_moves[i] = 0; prop[i] = -1; mv[i] = -1;
        if (a <= 4) {                                             /* map_env.py:502-508 */
            int q = e->pos[i];
            if (a < 4) q += MOVE_D[e->orient[i]][a][0] * W + MOVE_D[e->orient[i]][a][1];
            if (m->wall[q]) q = e->pos[i];                        /* agent.py:111-119 */
            prop[i] = mv[i] = q; in_moves[i] = 1; order[n_movers++] = i;
        } else if (a == 5) e->orient[i] = CW_T[e->orient[i]];     /* map_env.py:509-511 */
        else if (a == 6) e->orient[i] = CCW_T[e->orient[i]];
    }
    if (n_movers == 0) return;                                    /* map_env.py:534 */
    /* np.random.shuffle -> ascending (prio, index); insertion sort */
    for (int a = 1; a < n_movers; ++a) {
        int x = order[a], b = a - 1;
        while (b >= 0 && (prio[order[b]] > prio[x] || (prio[order[b]] == prio[x] && order[b] > x))) { order[b + 1] = order[b]; --b; }
        order[b + 1] = x;
    }
    /* phase 1: contested cells in lexicographic (= cell index) order of the ORIGINAL proposals, 543-609 */
    int cells[SSDO_MAX_AGENTS], nc = 0;
    for (int a = 0; a < n_movers; ++a) {
        int c = prop[order[a]], seen = 0;
        for (int b = 0; b < nc; ++b) if (cells[b] == c) seen = 1;
        if (!seen) cells[nc++] = c;
    }
    for (int a = 1; a < nc; ++a) { int x = cells[a], b = a - 1; while (b >= 0 && cells[b] > x) { cells[b + 1] = cells[b]; --b; } cells[b + 1] = x; }
    for (int ci = 0; ci < nc; ++ci) {
        int cell = cells[ci], cont[SSDO_MAX_AGENTS], k = 0;
        for (int a = 0; a < n_movers; ++a) if (prop[order[a]] == cell) cont[k++] = order[a];
        if (k < 2) continue;
        int free_cell = 1;
        for (int a = 0; a < k; ++a) {
            int i = cont[a];
            if (occupied(m, e, cell)) {                           /* 567 */
                int occ = by_pos(e->pos, n, cell);
                if (occ == i) free_cell = 0;                                              /* (1) 578 */
                else if (!in_moves[occ] || e->pos[occ] == mv[occ]) free_cell = 0;         /* (2) 584-586 */
                else if (mv[occ] == e->pos[i] && cell == e->pos[occ]) free_cell = 0;      /* (3) 590-594 */
            }
        }
        if (free_cell) e->pos[cont[0]] = cell;                    /* 598-601 winner moves now */
        for (int a = 0; a < k; ++a) mv[cont[a]] = e->pos[cont[a]]; /* 604-609 */
    }
    /* phase 2: 612-661 */
    for (;;) {
        int num = 0;
        for (int i = 0; i < n; ++i) num += in_moves[i];
        if (num == 0) break;
        int32_t spos[SSDO_MAX_AGENTS]; int snap[SSDO_MAX_AGENTS], deleted[SSDO_MAX_AGENTS];
        for (int i = 0; i < n; ++i) { spos[i] = e->pos[i]; snap[i] = in_moves[i]; deleted[i] = 0; }
        for (int i = 0; i < n; ++i) {
            if (!snap[i] || deleted[i]) continue;
            int mvi = mv[i];
            if (occupied(m, e, mvi)) {                            /* 621 live positions */
                int occ = by_pos(spos, n, mvi);                   /* snapshot dict, 613 */
                if (occ < 0) { e->error = 1; in_moves[i] = 0; deleted[i] = 1; continue; }  /* reference KeyError */
                int occ_mv = in_moves[occ] ? mv[occ] : e->pos[occ];
                if (occ == i) { in_moves[i] = 0; deleted[i] = 1; }                                   /* (1) 630 */
                else if (!snap[occ] || e->pos[occ] == occ_mv) { in_moves[i] = 0; deleted[i] = 1; }   /* (2) 636-639 */
                else if (mv[occ] == e->pos[i] && mvi == e->pos[occ]) {                               /* (3) 642-648 */
                    in_moves[i] = in_moves[occ] = 0; deleted[i] = deleted[occ] = 1;
                }
            } else { e->pos[i] = mvi; in_moves[i] = 0; deleted[i] = 1; }                             /* 650-653 */
        }
        int left = 0;
        for (int i = 0; i < n; ++i) left += in_moves[i];
        if (left == num) {                                        /* 658-661 cycles move together */
            for (int i = 0; i < n; ++i) if (in_moves[i]) e->pos[i] = mv[i];
            break;
        }
    }
}

/* --------------------------------------------------------------- beams ---- */
/* map_env.py:663-769; cleanup.py:127-144; harvest.py:79-84; agent.py:184-190,239-248 */
static int fire_beam(const ssdo_map* m, ssdo_env* e, int i, int clean, int8_t* reward) {
    const int W = m->W, H = m->H;
    int o = e->orient[i], pr = e->pos[i] / W, pc = e->pos[i] % W;
    int dr = FIRE_D[o][0], dc = FIRE_D[o][1], rr = RIGHT_D[o][0], rc = RIGHT_D[o][1];
    int starts[3][2] = { {pr, pc}, {pr + rr - dr, pc + rc - dc}, {pr - rr - dr, pc - rc - dc} };   /* 728-730 */
    int upd[3], nu = 0;
    for (int s = 0; s < 3; ++s) {
        int r = starts[s][0] + dr, c = starts[s][1] + dc;
        for (int k = 0; k < m->beam_len; ++k) {
            if (r < 0 || r >= H || c < 0 || c >= W) break;        /* 736, 865-872 */
            int cell = r * W + c;
            if (e->grid[cell] == SSDO_WALL) break;                /* 737 */
            if (occupied(m, e, cell)) {                           /* 741 agents absorb beams */
                if (!clean) reward[by_pos(e->pos, m->n, cell)] -= (int8_t)m->hit_penalty;   /* hit('F'), penalty 0 in the reference */
                if (clean && e->grid[cell] == SSDO_WASTE) upd[nu++] = cell;                 /* 745-748 */
                break;
            }
            if (clean && e->grid[cell] == SSDO_WASTE) { upd[nu++] = cell; break; }          /* 752-760 */
            r += dr; c += dc;
        }
    }
    for (int k = 0; k < nu; ++k) e->grid[upd[k]] = SSDO_RIVER;   /* 669-671 after all three rays */
    return nu;
}

/* ------------------------------------------------------------- spawning ---- */
static void spawn_cleanup(const ssdo_map* m, ssdo_env* e, const ssdo_draws* d, uint64_t seed, uint32_t gid) {
    /* cleanup.py:146-149, 165-212 */
    int h = 0;
    for (int c = 0; c < m->G; ++c) h += e->grid[c] == SSDO_WASTE;
    uint32_t tA = m->thr_apple_lut[h], tW = m->thr_waste_lut[h];
    int sp[SSDO_MAX_CELLS], ns = 0;
    for (int k = 0; k < m->n_apple; ++k) {
        int c = m->apple_pts[k];
        if (!occupied(m, e, c) && e->grid[c] != SSDO_APPLE) {
            uint32_t u = d && d->u_apple ? d->u_apple[c] : philox_word(seed, gid, e->tick, 1, (uint32_t)k);
            if (u < tA) sp[ns++] = c;
        }
    }
    int wcell = -1;
    if (tW != 0) {
        /* first success in ascending (wkey, cell) order == min over successes */
        uint32_t best_key = 0;
        for (int k = 0; k < m->n_waste; ++k) {
            int c = m->waste_pts[k];
            if (e->grid[c] == SSDO_WASTE) continue;
            uint32_t u, key;
            if (d && d->u_waste) { u = d->u_waste[c]; key = d->wkey[c]; }
            else { u = philox_word(seed, gid, e->tick, 2, 2u * k); key = philox_word(seed, gid, e->tick, 2, 2u * k + 1); }
            if (u < tW && (wcell < 0 || key < best_key)) { wcell = c; best_key = key; }   /* waste_pts ascending => ties by cell */
        }
    }
    for (int k = 0; k < ns; ++k) e->grid[sp[k]] = SSDO_APPLE;
    if (wcell >= 0) e->grid[wcell] = SSDO_WASTE;
}

static void spawn_harvest(const ssdo_map* m, ssdo_env* e, const ssdo_draws* d, uint64_t seed, uint32_t gid) {
    /* harvest.py:86-122; radius test j^2+k^2 <= 2 == the 3x3 block (line 110) */
    int sp[SSDO_MAX_CELLS], ns = 0;
    for (int k = 0; k < m->n_apple; ++k) {
        int c = m->apple_pts[k], r0 = c / m->W, c0 = c % m->W;
        if (occupied(m, e, c) || e->grid[c] == SSDO_APPLE) continue;
        int cnt = 0;
        for (int j = -2; j <= 2; ++j) for (int q = -2; q <= 2; ++q) if (j * j + q * q <= 2) {
            int r = r0 + j, cc = c0 + q;
            if (r >= 0 && r < m->H && cc >= 0 && cc < m->W && e->grid[r * m->W + cc] == SSDO_APPLE) ++cnt;
        }
        uint32_t u = d && d->u_apple ? d->u_apple[c] : philox_word(seed, gid, e->tick, 1, (uint32_t)k);
        if (u < m->thr_harvest[cnt < 3 ? cnt : 3]) sp[ns++] = c;
    }
    for (int k = 0; k < ns; ++k) e->grid[sp[k]] = SSDO_APPLE;
}

static void custom_map_update(const ssdo_map* m, ssdo_env* e, const ssdo_draws* d, uint64_t seed, uint32_t gid) {
    if (m->kind == SSDO_KIND_CLEANUP) spawn_cleanup(m, e, d, seed, gid); else spawn_harvest(m, e, d, seed, gid);
}

/* ----------------------------------------------------------------- step ---- */
/* map_env.py:874-915 -> _step 227-295 */
void ssdo_step(const ssdo_map* m, ssdo_env* e, const uint8_t* actions, const ssdo_draws* d,
               uint64_t seed, uint32_t env_gid,
               int8_t* reward, uint8_t* clean, uint16_t* apple_cnt, uint8_t* done) {
    const int n = m->n;
    uint32_t prio[SSDO_MAX_AGENTS];
    for (int i = 0; i < n; ++i) {
        prio[i] = d && d->prio ? d->prio[i] : philox_word(seed, env_gid, e->tick, 0, (uint32_t)i);
        reward[i] = 0; clean[i] = 0;
    }
    update_moves(m, e, actions, prio);                            /* 251 */
    for (int i = 0; i < n; ++i)                                   /* 253-256 consume, index order */
        if (e->grid[e->pos[i]] == SSDO_APPLE) { reward[i] += 1; e->grid[e->pos[i]] = SSDO_EMPTY; }
    for (int i = 0; i < n; ++i) {                                 /* 259-260, 663-673 */
        if (actions[i] == 7) {                                    /* FIRE: agent.py:188-190, 239-241 */
            reward[i] -= (int8_t)m->fire_cost;
            fire_beam(m, e, i, 0, reward);
        } else if (actions[i] == 8 && m->kind == SSDO_KIND_CLEANUP) {
            clean[i] = (uint8_t)fire_beam(m, e, i, 1, reward);    /* 672-673 */
        }
    }
    custom_map_update(m, e, d, seed, env_gid);                    /* 263 */
    int apples = 0;                                               /* 291-292 counted on map WITH agents */
    for (int c = 0; c < m->G; ++c) apples += e->grid[c] == SSDO_APPLE && !occupied(m, e, c);
    *apple_cnt = (uint16_t)apples;
    for (int i = 0; i < n; ++i) e->ep_ret[i] += reward[i];        /* 885-888 */
    e->t += 1; e->tick += 1;
    *done = e->t >= m->episode_limit;                             /* 890-894; get_done() is always False */
}

/* ---------------------------------------------------------------- reset ---- */
/* map_env.py:297-326, 681-685, 771-793, 986-993; cleanup.py:117-125,151-163; harvest.py:60-77 */
void ssdo_reset(const ssdo_map* m, ssdo_env* e, int random_spawn_point, int spawn_rotation,
                const ssdo_draws* d, uint64_t seed, uint32_t env_gid) {
    const int n = m->n, S = m->n_spawn;
    for (int i = 0; i < n; ++i) e->pos[i] = -1;
    for (int i = 0; i < n; ++i) {
        int best = -1; uint32_t best_key = 0;
        for (int s = 0; s < S; ++s) {
            int c = m->spawn_pts[s], taken = 0;
            for (int j = 0; j < i; ++j) taken |= e->pos[j] == c;
            if (taken) continue;
            uint32_t key = 0;                                     /* fixed order: last free point in row-major order */
            if (random_spawn_point)
                key = d && d->spawn_key ? d->spawn_key[i * m->G + c] : philox_word(seed, env_gid, e->tick, 3, (uint32_t)(i * S + s));
            if (best < 0 || key >= best_key) { best = c; best_key = key; }   /* spawn_pts ascending => max (key, cell) */
        }
        e->pos[i] = best;
        if (spawn_rotation >= 0) e->orient[i] = (uint8_t)spawn_rotation;     /* 786-793 */
        else e->orient[i] = d && d->rot ? d->rot[i] : (uint8_t)(philox_word(seed, env_gid, e->tick, 4, (uint32_t)i) >> 30);
        e->ep_ret[i] = 0;
    }
    for (int c = 0; c < m->G; ++c) e->grid[c] = base_to_code(m->kind, m->base[c]);
    /* custom_map_update() at reset (map_env.py:313): with draws keyed by this tick */
    custom_map_update(m, e, d, seed, env_gid);
    e->t = 0; e->tick += 1; e->error = 0;
}

/* --------------------------------------------------------------- render ---- */
/* map_env.py:360-379, 418-446, 795-815, 923-957; agent.py:82-84; utility_funcs.py:58-116 */
static void overlay(const ssdo_map* m, const ssdo_env* e, uint8_t* idx /* [G] colour index */) {
    for (int c = 0; c < m->G; ++c) idx[c] = e->grid[c];
    for (int i = 0; i < m->n; ++i) idx[e->pos[i]] = (uint8_t)(6 + agent_char(i));   /* later index overwrites */
}

void ssdo_render_state(const ssdo_map* m, const ssdo_env* e, uint8_t* out) {
    uint8_t idx[SSDO_MAX_CELLS];
    overlay(m, e, idx);
    for (int ch = 0; ch < 3; ++ch) for (int c = 0; c < m->G; ++c) out[ch * m->G + c] = m->color[idx[c]][ch];
}

void ssdo_render_obs(const ssdo_map* m, const ssdo_env* e, uint8_t* obs) {
    const int N = m->N, V = m->V, W = m->W, H = m->H;
    uint8_t idx[SSDO_MAX_CELLS];
    overlay(m, e, idx);
    for (int i = 0; i < m->n; ++i) {
        int pr = e->pos[i] / W, pc = e->pos[i] % W;
        for (int y = 0; y < N; ++y) for (int x = 0; x < N; ++x) {
            int a, b;                                            /* view[a][b] index before rotation */
            switch (e->orient[i]) {                              /* np.rot90 k = 1, 3, 0, 2 (map_env.py:806-813) */
                case SSDO_LEFT:  a = x;         b = N - 1 - y; break;
                case SSDO_RIGHT: a = N - 1 - x; b = y;         break;
                case SSDO_UP:    a = y;         b = x;         break;
                default:         a = N - 1 - y; b = N - 1 - x; break;
            }
            int r = pr - V + a, c = pc - V + b;                  /* utility_funcs.py:76-90, zero padded */
            int ci = (r >= 0 && r < H && c >= 0 && c < W) ? idx[r * W + c] : 6;
            for (int ch = 0; ch < 3; ++ch) obs[((i * 3 + ch) * N + y) * N + x] = m->color[ci][ch];
        }
    }
}

/* ---------------------------------------------------------------- batch ---- */
void ssdo_batch_reset(const ssdo_map* m, ssdo_env* envs, int B, int random_spawn_point, int spawn_rotation,
                      uint64_t seed, uint32_t env_gid0, int threads) {
    (void)threads;
#pragma omp parallel for num_threads(threads) schedule(static)
    for (int b = 0; b < B; ++b) ssdo_reset(m, &envs[b], random_spawn_point, spawn_rotation, 0, seed, env_gid0 + (uint32_t)b);
}

void ssdo_batch_step(const ssdo_map* m, ssdo_env* envs, int B, const uint8_t* actions,
                     uint64_t seed, uint32_t env_gid0,
                     int8_t* reward, uint8_t* clean, uint16_t* apple_cnt, uint8_t* done,
                     uint8_t* obs, int threads) {
    (void)threads;
    const int n = m->n;
    const long obs_env = (long)n * 3 * m->N * m->N;
#pragma omp parallel for num_threads(threads) schedule(static)
    for (int b = 0; b < B; ++b) {
        ssdo_step(m, &envs[b], actions + (long)b * n, 0, seed, env_gid0 + (uint32_t)b,
                  reward + (long)b * n, clean + (long)b * n, apple_cnt + b, done + b);
        if (obs) ssdo_render_obs(m, &envs[b], obs + b * obs_env);
    }
}

unsigned long ssdo_sizeof_map(void) { return sizeof(ssdo_map); }
unsigned long ssdo_sizeof_env(void) { return sizeof(ssdo_env); }
