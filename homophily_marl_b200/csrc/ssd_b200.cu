// ssd_b200.cu -- sm_100a kernels + C ABI of the batched SSD grid-world simulator.
//
// Design: one warp owns one env instance for a whole step.
//   * lanes 0..n-1 ARE the agents during move/rotate, conflict resolution, consume and beams:
//     collisions, swaps, chains and cycles are resolved with __ballot/__shfl/redux (bit-filter pre-tests
//     with REDUX.OR, exact shuffle loops only on a possible hit), keeping the reference's sequential
//     phase structure (map_env.py:477-661);
//   * lanes are spawn candidates during apple/waste spawning (4 apple points or 2 waste
//     points per Philox4x32-10 call);
//   * lanes are output pixel ROWS during the egocentric render.  Small views (N <= 16, Cleanup) are
//     gathered straight from the staged byte grid; large views (Harvest, N = 31) read one run of N
//     nibbles from a zero-padded nibble-packed index map (or its transpose) held in shared memory.
//     Either way a lane colours 4 pixels per PRMT through an 8-entry register LUT and writes the
//     finished row with one 256-bit (st.global.v8.b32, SASS STG.E.ENL2.256) or 128-bit store per plane.
// The env's grid is staged in shared memory for the whole step.  No tensor cores: nothing here is a
// contraction.  HBM traffic per env-step is the algorithmic 2G + n(3N^2+11)+3 bytes (+ row padding),
// ~98% of it observation WRITES.
//
// Reference citations are relative to drdh/Homophily-MARL.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>
#include <new>
#include <mutex>

#include "ssd_b200.h"

namespace {

#ifndef SSD_WARPS
#define SSD_WARPS 4
#endif
#ifndef SSD_MIN_BLOCKS
#define SSD_MIN_BLOCKS 8
#endif
constexpr int kWarps = SSD_WARPS;         // env instances per CTA
constexpr uint16_t kNoPoint = 0xFFFF;
constexpr int kMaxDevices = 64;

enum { MODE_STEP = 0, MODE_RESET = 1, MODE_RENDER = 2 };

// -DSSD_BOUNDS_CHECK builds the checked variant (libssd_b200_check.so): every data-dependent shared-memory index is
// tested against its tile and violations are counted (compute-sanitizer is closed on the GPU pool).
#ifdef SSD_BOUNDS_CHECK
__device__ unsigned long long g_oob_count = 0;
#define SSD_CHECK(cond) do { if (!(cond)) atomicAdd(&g_oob_count, 1ull); } while (0)
#else
#define SSD_CHECK(cond) do { } while (0)
#endif

// Static per-map tables, resident in global memory (read through the L1 read-only path).
struct MapDev {
    uint8_t  base_grid[SSD_MAX_CELLS];        // reset grid in cell codes
    uint16_t apple_pts[SSD_MAX_CELLS + 4];    // row-major, padded to x4 with kNoPoint
    uint16_t waste_pts[SSD_MAX_CELLS + 4];    // padded to x2
    uint16_t spawn_pts[SSD_MAX_SPAWN];
    uint32_t thr_apple[SSD_MAX_CELLS + 1];
    uint32_t thr_waste[SSD_MAX_CELLS + 1];
};

struct KParams {
    int kind, B, n, H, W, G, V, N;
    int env0, env1;                           // this launch covers env instances [env0, env1) of the B resident ones
    int pdl;                                  // launched with the programmatic-serialization attribute (set per launch)
    int GS, NA, RP, PS, AS, ES;               // strides (ssd_layout)
    int pitchM, pitchT, off_map[4], PMS;      // nibble maps M / MT: row pitches, byte offset by orientation (0,1 -> MT; 2,3 -> M), total bytes
    int direct, padF, padB;                   // direct-gather render (small views): slack bytes before / after the staged grid
    int LPn, nw8M, nw8T;                      // left pad in nibbles (V rounded up to 8), words per map row holding cells
    uint32_t maskM8, maskT8;                  // valid-nibble mask of the last word of a map row
    int agents_uniform;                       // all agent colours equal -> no per-agent repaint
    int episode_limit, fire_cost, hit_penalty, beam_len, n_actions;
    int n_apple, n_waste, n_spawn, n_apple4, n_waste2;
    int random_spawn, spawn_rot;
    int base_apples, base_waste;              // 'A' / 'H' cells of the reset grid (cell counts are carried in ssd_state.counts)
    int smem_per_warp, off_pmap;
    uint32_t invN20, invW20, invMW20, invMTW20;   // floor(x / d) == (x * inv) >> 20 on the used ranges
    uint32_t seed_lo, seed_hi, gid_base;
    uint32_t thr_harvest[4];
    uint32_t lut[16];
    uint32_t lut8[6];                         // 8-entry colour LUT per plane (lo, hi words): R, G, B
    const MapDev* map;
    uint8_t* grid; uint32_t* agent; int32_t* ep_ret; int32_t* t; uint32_t* tick; uint32_t* counts;
    const uint8_t* actions; const uint8_t* mask;
    int8_t* reward; uint8_t* clean; uint16_t* apple_cnt; uint8_t* done; uint8_t* obs; uint8_t* state_rgb;
    const uint32_t* d_prio; const uint32_t* d_uapple; const uint32_t* d_uwaste; const uint32_t* d_wkey;
    const uint32_t* d_spawnkey; const uint8_t* d_rot;
};


// ------------------------------------------------------------------ geometry
// Everything that follows from (kind, H, W, V).  Computed by ONE constexpr function used by the host
// (ssd_create) and, for the reference's shipped maps, at compile time by the kernels (GeoS), so that index
// arithmetic folds into immediates; any other geometry runs the same code on runtime values (GeoD).
struct GeoVals {
    int kind, H, W, V, G, GS, N, RP, PS, AS, LPn, nw8M, nw8T, pitchM, pitchT, off0, off1, off2, off3, PMS;
    int direct, padF, padB;                   // small views (RP <= 16) are gathered straight from the staged grid: no maps, but
                                              // V rows of slack before / after the grid so that column runs need no bounds test
    uint32_t maskM8, maskT8;
};
__host__ __device__ constexpr int cround_up(int v, int m) { return (v + m - 1) / m * m; }
__host__ __device__ constexpr GeoVals make_geo(int kind, int H, int W, int V) {
    GeoVals g{};
    g.kind = kind; g.H = H; g.W = W; g.V = V; g.G = H * W; g.GS = cround_up(H * W, 16);
    g.N = 2 * V + 1; g.RP = cround_up(g.N, 4); g.PS = g.N * g.RP; g.AS = cround_up(3 * g.PS, 16);
    // nibble-packed padded maps; word pitch kept odd so that lanes gathering consecutive rows hit distinct banks
    g.LPn = cround_up(V, 8); g.nw8M = (W + 7) / 8; g.nw8T = (H + 7) / 8;
    g.maskM8 = (W & 7) ? (1u << (4 * (W & 7))) - 1u : 0xffffffffu;
    g.maskT8 = (H & 7) ? (1u << (4 * (H & 7))) - 1u : 0xffffffffu;
    int wm = (g.LPn + W + V + 7) / 8; if (wm < g.LPn / 8 + g.nw8M) wm = g.LPn / 8 + g.nw8M; if ((wm & 1) == 0) ++wm;
    int wt = (g.LPn + H + V + 7) / 8; if (wt < g.LPn / 8 + g.nw8T) wt = g.LPn / 8 + g.nw8T; if ((wt & 1) == 0) ++wt;
    g.pitchM = 4 * wm; g.pitchT = 4 * wt;
    const int szM = cround_up((H + 2 * V) * g.pitchM, 16) + 16, szT = cround_up((W + 2 * V) * g.pitchT, 16) + 16;
    g.off0 = 16; g.off1 = 16; g.off2 = 16 + szT; g.off3 = 16 + szT;        // by orientation: LEFT/RIGHT -> MT, UP/DOWN -> M
    g.PMS = szT + szM + 32;
    g.direct = g.RP <= 16;
    g.padF = g.direct ? cround_up(V * W + 48, 16) : 16;      // + 32 bytes at the very front for the (class, rank) -> lane table
    g.padB = g.direct ? cround_up(V * W + 32, 16) : 32;
    if (g.direct) g.PMS = 0;
    return g;
}

template <int KIND_, int H_, int W_, int V_>
struct GeoS {                                                 // compile-time geometry
    static constexpr GeoVals v = make_geo(KIND_, H_, W_, V_);
    __device__ __forceinline__ explicit GeoS(const KParams&) {}
    __device__ __forceinline__ int kind() const { return v.kind; }
    __device__ __forceinline__ int H() const { return v.H; }
    __device__ __forceinline__ int W() const { return v.W; }
    __device__ __forceinline__ int V() const { return v.V; }
    __device__ __forceinline__ int G() const { return v.G; }
    __device__ __forceinline__ int GS() const { return v.GS; }
    __device__ __forceinline__ int N() const { return v.N; }
    __device__ __forceinline__ int RP() const { return v.RP; }
    __device__ __forceinline__ int PS() const { return v.PS; }
    __device__ __forceinline__ int AS() const { return v.AS; }
    __device__ __forceinline__ int LPn() const { return v.LPn; }
    __device__ __forceinline__ int nw8M() const { return v.nw8M; }
    __device__ __forceinline__ int nw8T() const { return v.nw8T; }
    __device__ __forceinline__ int pitchM() const { return v.pitchM; }
    __device__ __forceinline__ int pitchT() const { return v.pitchT; }
    __device__ __forceinline__ int PMS() const { return v.PMS; }
    __device__ __forceinline__ bool direct() const { return v.direct != 0; }
    __device__ __forceinline__ int padF() const { return v.padF; }
    __device__ __forceinline__ int padB() const { return v.padB; }
    __device__ __forceinline__ uint32_t maskM8() const { return v.maskM8; }
    __device__ __forceinline__ int off_map(int o) const { return o == 0 ? v.off0 : (o == 1 ? v.off1 : (o == 2 ? v.off2 : v.off3)); }
    __device__ __forceinline__ int divW(int x) const { return (int)((unsigned)x / (unsigned)v.W); }
    __device__ __forceinline__ int divN(int x) const { return (int)((unsigned)x / (unsigned)v.N); }
    __device__ __forceinline__ int divMW(int x) const { return (int)((unsigned)x / (unsigned)v.nw8M); }
};

struct GeoD {                                                 // runtime geometry (any wall-enclosed map)
    const KParams& p;
    __device__ __forceinline__ explicit GeoD(const KParams& p_) : p(p_) {}
    __device__ __forceinline__ int kind() const { return p.kind; }
    __device__ __forceinline__ int H() const { return p.H; }
    __device__ __forceinline__ int W() const { return p.W; }
    __device__ __forceinline__ int V() const { return p.V; }
    __device__ __forceinline__ int G() const { return p.G; }
    __device__ __forceinline__ int GS() const { return p.GS; }
    __device__ __forceinline__ int N() const { return p.N; }
    __device__ __forceinline__ int RP() const { return p.RP; }
    __device__ __forceinline__ int PS() const { return p.PS; }
    __device__ __forceinline__ int AS() const { return p.AS; }
    __device__ __forceinline__ int LPn() const { return p.LPn; }
    __device__ __forceinline__ int nw8M() const { return p.nw8M; }
    __device__ __forceinline__ int nw8T() const { return p.nw8T; }
    __device__ __forceinline__ int pitchM() const { return p.pitchM; }
    __device__ __forceinline__ int pitchT() const { return p.pitchT; }
    __device__ __forceinline__ int PMS() const { return p.PMS; }
    __device__ __forceinline__ bool direct() const { return p.direct != 0; }
    __device__ __forceinline__ int padF() const { return p.padF; }
    __device__ __forceinline__ int padB() const { return p.padB; }
    __device__ __forceinline__ uint32_t maskM8() const { return p.maskM8; }
    __device__ __forceinline__ int off_map(int o) const { return p.off_map[o]; }
    __device__ __forceinline__ int divW(int x) const { return (int)(((uint32_t)x * p.invW20) >> 20); }
    __device__ __forceinline__ int divN(int x) const { return (int)(((uint32_t)x * p.invN20) >> 20); }
    __device__ __forceinline__ int divMW(int x) const { return (int)(((uint32_t)x * p.invMW20) >> 20); }
};

// ------------------------------------------------------------------ Philox4x32-10
__device__ __forceinline__ uint4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                               uint32_t k0, uint32_t k1) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
        const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
        const uint32_t n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
        c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    return make_uint4(c0, c1, c2, c3);
}
__device__ __forceinline__ uint32_t pick(const uint4& v, int i) {
    return i == 0 ? v.x : (i == 1 ? v.y : (i == 2 ? v.z : v.w));
}
// A sub-warp of LPE lanes owns one env instance.  The shipped configuration is LPE = 32 (one env per warp).  LPE = 16 (two
// envs per warp, so that the agent-lane phases issue one instruction for two envs; -DSSD_ENABLE_LPE16) passes the parity
// suite but measured 7-20 % slower on every workload: the lane-parallel render phases take twice the trips per warp.
// Every collective is restricted to the sub-warp's member mask, so the halves of a warp may diverge freely.
template <int LPE>
struct SubWarp {
    static constexpr int kLanes = LPE, kEnvs = 32 / LPE;
    int lane, sub, base;                                     // lane within the sub-warp, sub-warp index, first warp lane
    unsigned mask;
    __device__ __forceinline__ explicit SubWarp(int warp_lane)
        : lane(warp_lane % LPE), sub(warp_lane / LPE), base(warp_lane / LPE * LPE),
          mask(LPE == 32 ? 0xffffffffu : (0xffffu << (warp_lane / LPE * LPE))) {}
    __device__ __forceinline__ unsigned ballot(bool pred) const { return __ballot_sync(mask, pred) >> base; }
    template <typename T> __device__ __forceinline__ T shfl(T v, int src) const { return __shfl_sync(mask, v, base + src); }
    template <typename T> __device__ __forceinline__ unsigned match(T v) const { return __match_any_sync(mask, v) >> base; }
    template <typename T> __device__ __forceinline__ T rmin(T v) const { return __reduce_min_sync(mask, v); }
    template <typename T> __device__ __forceinline__ T rmax(T v) const { return __reduce_max_sync(mask, v); }
    template <typename T> __device__ __forceinline__ T radd(T v) const { return __reduce_add_sync(mask, v); }
    __device__ __forceinline__ unsigned ror(unsigned v) const { return __reduce_or_sync(mask, v); }
    __device__ __forceinline__ void sync() const { __syncwarp(mask); }
};

// lanes j < n whose value equals this lane's: what __match_any_sync returns for the agent lanes, from n pipelined shuffles
// (MATCH.ANY costs 3-4 % of a Cleanup step per use, profiles/r2_notes.md)
template <class SW>
__device__ __forceinline__ unsigned equal_lanes(const SW& w, int v, int n) {
    unsigned m = 0;
    for (int j = 0; j < n; ++j) m |= (unsigned)(w.shfl(v, j) == v) << j;
    return m;
}
// Cheap sufficient test for "the values of the lanes in `who` are pairwise distinct": two 32-bucket bit filters, OR-reduced over
// the warp (REDUX.OR); if either filter shows one bit per participating lane there is no duplicate.  False alarms (hash
// collisions) only send the caller to the exact shuffle loop.
template <class SW>
__device__ __forceinline__ bool surely_distinct(const SW& w, int v, bool in, unsigned who) {
    const int cnt = __popc(who);
    const unsigned f0 = w.ror(in ? 1u << (v & 31) : 0u);
    if (__popc(f0) == cnt) return true;
    const unsigned f1 = w.ror(in ? 1u << ((v >> 5) & 31) : 0u);
    return __popc(f1) == cnt;
}

// sub-warp-strided loop with the first trip peeled (trip counts here are almost always 0 or 1)
template <int STRIDE, typename F>
__device__ __forceinline__ void warp_for(int n, int lane, F f) {
    int i = lane;
    if (i < n) f(i);
    for (i += STRIDE; i < n; i += STRIDE) f(i);
}
// draw streams: 0 mover priority, 1 apple, 2 waste (u, order key), 3 spawn key, 4 spawn rotation
__device__ __forceinline__ uint32_t philox_word(const KParams& p, uint32_t gid, uint32_t tick, uint32_t stream, uint32_t idx) {
    const uint4 r = philox4x32_10(gid, tick, stream, idx >> 2, p.seed_lo, p.seed_hi);
    return pick(r, idx & 3);
}

// agent i is drawn as the first char of str(i % 10 + 1)  (map_env.py:370 into a <U1 array)
__device__ __forceinline__ int agent_colour_index(int i) {
    const int v = i % 10 + 1;
    return 6 + (v == 10 ? 1 : v);
}

// ------------------------------------------------------------------ update_moves (map_env.py:477-661)
// lanes < n hold one agent each: pos (cell index), ori, act.  Non-agent lanes carry unique
// negative positions so they never match a cell.
template <class SW, class GEO>
__device__ __forceinline__ void update_moves(const SW& w, const GEO& g, const KParams& p, const uint8_t* __restrict__ sg, int lane, bool is_agent,
                                             int act, int& pos, int& ori, int env, uint32_t gid, uint32_t tick) {
    // turns take effect immediately (map_env.py:509-511, 843-861)
    if (act == 5) ori = (0x0132 >> (4 * ori)) & 3;           // CW : LEFT->UP, RIGHT->DOWN, UP->RIGHT, DOWN->LEFT
    else if (act == 6) ori = (0x1023 >> (4 * ori)) & 3;      // CCW: LEFT->DOWN, RIGHT->UP, UP->LEFT, DOWN->RIGHT
    const bool mover = is_agent && act <= 4;
    int prop = pos;
    if (mover && act < 4) {
        // (drow, dcol) of action a under orientation o, 2-bit fields holding value+1 (map_env.py:826-835)
        //   LEFT : (0,+1) (0,-1) (-1,0) (+1,0)   RIGHT: (0,-1) (0,+1) (+1,0) (-1,0)
        //   UP   : (-1,0) (+1,0) (0,-1) (0,+1)   DOWN : (+1,0) (-1,0) (0,+1) (0,-1)
        constexpr uint32_t kDr = (1u) | (1u << 2) | (0u << 4) | (2u << 6) | (1u << 8) | (1u << 10) | (2u << 12) | (0u << 14) |
                                 (0u << 16) | (2u << 18) | (1u << 20) | (1u << 22) | (2u << 24) | (0u << 26) | (1u << 28) | (1u << 30);
        constexpr uint32_t kDc = (2u) | (0u << 2) | (1u << 4) | (1u << 6) | (0u << 8) | (2u << 10) | (1u << 12) | (1u << 14) |
                                 (1u << 16) | (1u << 18) | (0u << 20) | (2u << 22) | (1u << 24) | (1u << 26) | (2u << 28) | (0u << 30);
        const int f = 2 * (ori * 4 + act);
        const int dr = (int)((kDr >> f) & 3u) - 1, dc = (int)((kDc >> f) & 3u) - 1;
        const int q = pos + dr * g.W() + dc;
        SSD_CHECK(q >= 0 && q < g.G());
        prop = sg[q] == SSD_CELL_WALL ? pos : q;             // agent.py:111-119
    }
    unsigned in_moves = w.ballot(mover);
    if (in_moves == 0) return;                                // map_env.py:534
    int mv = prop;                                            // live agent_moves[i]

    // ---- phase 1: contested cells in lexicographic order of the ORIGINAL proposals (543-609)
    unsigned pending = 0;                                     // movers whose proposal is shared with another mover
    if (!surely_distinct(w, prop, mover, in_moves)) {
        const unsigned grp = equal_lanes(w, mover ? prop : -1000 - lane, p.n);
        pending = w.ballot(mover && __popc(grp) >= 2);
    }
    if (pending) {
        uint32_t prio = 0;
        if (mover) prio = p.d_prio ? p.d_prio[(size_t)env * p.n + lane] : philox_word(p, gid, tick, 0, (uint32_t)lane);
        while (pending) {
            const int cell = w.rmin(((pending >> lane) & 1u) ? prop : 0x7fffffff);
            const bool in_cont = mover && prop == cell;
            const unsigned cont = w.ballot(in_cont);
            const unsigned occm = w.ballot(pos == cell);          // live positions (567)
            bool free_cell = true;
            if (occm) {
                const int occ = 31 - __clz(occm);                              // dict build: last index wins (516)
                const int mv_occ = w.shfl(mv, occ);
                const bool occ_moves = (in_moves >> occ) & 1u;
                const bool c1 = (cont >> occ) & 1u;                            // (1) 578
                const bool c2 = !occ_moves || mv_occ == cell;                  // (2) 584-586
                const unsigned swp = w.ballot(in_cont && mv_occ == pos);   // (3) 590-594
                free_cell = !(c1 || c2 || swp != 0);
            }
            if (free_cell) {                                                   // winner = first in shuffled order (598-601)
                const uint32_t mk = w.rmin(in_cont ? prio : 0xffffffffu);
                const unsigned eq = w.ballot(in_cont && prio == mk);
                if (lane == __ffs(eq) - 1) pos = cell;
            }
            if (in_cont) mv = pos;                                             // 604-609
            pending &= ~cont;
        }
    }
    // ---- phase 2: iterate until every move is made or dropped (612-661)
    // Fast path: if no mover's target is occupied by ANOTHER agent, the sequential walk below moves every mover (targets are
    // unique after phase 1 or the mover's own cell, and nobody enters a cell that is somebody's target), so all move at once.
    // (A mover whose target is its own cell never changes anything, whoever else stands there.)  The occupancy test is first
    // made against a 32-bucket bit filter of all agents' cells; only a possible hit pays for the exact shuffle loop.
    {
        const bool mine = (in_moves >> lane) & 1u;
        const unsigned cells = w.ror(is_agent ? 1u << (pos & 31) : 0u);
        bool blocked = mine && mv != pos && ((cells >> (mv & 31)) & 1u);
        if (w.ballot(blocked)) {
            blocked = false;
            for (int j = 0; j < p.n; ++j) blocked |= (w.shfl(pos, j) == mv) && (j != lane);
            blocked = blocked && mine && mv != pos;
        }
        if (w.ballot(blocked) == 0) {
            if (mine) pos = mv;
            return;
        }
    }
    while (in_moves) {
        const int spos = pos;                                  // snapshot dict of this pass (613)
        const unsigned snap = in_moves;
        unsigned todo = snap;
        while (todo) {
            const int i = __ffs(todo) - 1;
            todo &= todo - 1;
            if (!((in_moves >> i) & 1u)) continue;             // deleted earlier in this pass (619-620)
            const int mvi = w.shfl(mv, i);
            const int posi = w.shfl(pos, i);
            const unsigned live = w.ballot(pos == mvi);            // 621
            if (!live) {                                                       // 650-653
                if (lane == i) pos = mvi;
                in_moves &= ~(1u << i);
                continue;
            }
            const unsigned sm = w.ballot(spos == mvi);
            if (!sm) { in_moves &= ~(1u << i); continue; }                     // reference KeyError; unreachable
            const int occ = 31 - __clz(sm);
            const int pos_occ = w.shfl(pos, occ);
            const int mv_occ = w.shfl(mv, occ);
            const int occ_mv = ((in_moves >> occ) & 1u) ? mv_occ : pos_occ;
            if (occ == i) in_moves &= ~(1u << i);                                              // (1) 630
            else if (!((snap >> occ) & 1u) || pos_occ == occ_mv) in_moves &= ~(1u << i);       // (2) 636-639
            else if (mv_occ == posi && mvi == pos_occ) in_moves &= ~((1u << i) | (1u << occ)); // (3) 642-648
        }
        if (in_moves == snap) {                                // nobody could move: rotate cycles together (658-661)
            if ((in_moves >> lane) & 1u) pos = mv;
            break;
        }
    }
}

// Agent occupancy is kept IN the staged grid: bit 7 of a cell byte is set while an agent stands on it
// (the dict `agent_by_pos` / `[r, c] in self.agent_pos` tests of the reference become one bit test).
constexpr uint8_t kOcc = 0x80;

// ------------------------------------------------------------------ beams (map_env.py:663-769)
template <class SW, class GEO>
__device__ __forceinline__ void beams(const SW& w, const GEO& g, const KParams& p, uint8_t* sg, int lane, bool is_agent,
                                      int act, int pos, int ori, int& reward, int& clean_num, int& waste) {
    const bool fire = is_agent && act == 7;
    const bool clean = is_agent && act == 8 && g.kind() == SSD_KIND_CLEANUP;
    if (fire) reward -= p.fire_cost;                          // agent.py:188-190, 239-241
    unsigned need = w.ballot(clean || (fire && p.hit_penalty != 0));
    while (need) {                                            // agent index order, map effects applied per agent (669-671)
        const int i = __ffs(need) - 1;
        need &= need - 1;
        const int posi = w.shfl(pos, i), orii = w.shfl(ori, i);
        const bool is_clean = w.shfl(act, i) == 8;
        // firing direction ORIENTATIONS[o] and its right-hand rotation (map_env.py:28-31, 840-841)
        const int dr = orii == 0 ? -1 : (orii == 1 ? 1 : 0), dc = orii == 2 ? -1 : (orii == 3 ? 1 : 0);
        const int rr = orii == 2 ? 1 : (orii == 3 ? -1 : 0), rc = orii == 0 ? -1 : (orii == 1 ? 1 : 0);
        const int d = dr * g.W() + dc, rs = rr * g.W() + rc;
        int upd = -1, hit = -1;
        if (lane < 3) {                                       // the three parallel rays (728-730)
            int q = posi + (lane == 0 ? d : (lane == 1 ? rs : -rs));
            for (int k = 0; k < p.beam_len; ++k) {            // maps are wall-enclosed: a ray cannot leave the map
                SSD_CHECK(q >= 0 && q < g.G());
                const int v = sg[q], code = v & 0x7f;
                if (code == SSD_CELL_WALL) break;             // 737
                if (v & kOcc) {                               // agents absorb beams (741-749)
                    if (!is_clean) hit = q;
                    if (is_clean && code == SSD_CELL_WASTE) upd = q;
                    break;
                }
                if (is_clean && code == SSD_CELL_WASTE) { upd = q; break; }   // 752-760
                q += d;
            }
        }
        w.sync();
        if (upd >= 0) sg[upd] = (uint8_t)(SSD_CELL_RIVER | (sg[upd] & kOcc));
        w.sync();
        const unsigned um = w.ballot(upd >= 0);
        if (lane == i && is_clean) clean_num = __popc(um);    // 672-673
        waste -= __popc(um);                                  // only CLEAN produces updates: 'H' -> 'R'
        if (p.hit_penalty != 0) {                             // hit(): the LAST index standing on the cell (agent.py:184-186)
#pragma unroll
            for (int s = 0; s < 3; ++s) {
                const int hq = w.shfl(hit, s);
                const unsigned on = w.ballot(hq >= 0 && pos == hq);
                if (on && lane == 31 - __clz(on)) reward -= p.hit_penalty;
            }
        }
    }
}

// ------------------------------------------------------------------ spawning (cleanup.py:165-204, harvest.py:92-122)
// `apples` / `waste` are the env's running cell counts (ssd_state.counts): the reference recounts the map every step
// (map_env.py:291-292, cleanup.py:206-212); here consume / clean / spawn keep them up to date, which is the same number.
template <class SW, class GEO>
__device__ __forceinline__ void spawn(const SW& w, const GEO& g, const KParams& p, uint8_t* sg, int lane, int env, uint32_t gid, uint32_t tick,
                                      int& apples, int& waste) {
    const MapDev* __restrict__ m = p.map;
    uint32_t tA = 1, tW = 0;
    if (g.kind() == SSD_KIND_CLEANUP) {
        SSD_CHECK(waste >= 0 && waste <= p.n_waste);
        tA = __ldg(&m->thr_apple[waste]);
        tW = __ldg(&m->thr_waste[waste]);
    }
    // apples: 4 candidate points per lane per Philox call; decisions use the pre-spawn grid.  In Cleanup a decision reads
    // only its own cell (and apple points and waste points are different cells), so the apple is written at once; Harvest's
    // 3x3 neighbour count needs every decision to see the pre-spawn grid, so its apples are applied in a second pass.
    unsigned long long decided = 0;
    int n_new = 0;
    if (tA != 0) {
        int it = 0;
        for (int j = lane; j < p.n_apple4; j += SW::kLanes, ++it) {
            const ushort4 pts = __ldg(reinterpret_cast<const ushort4*>(m->apple_pts) + j);
            const uint16_t c4[4] = { pts.x, pts.y, pts.z, pts.w };
            uint4 r = make_uint4(0, 0, 0, 0);
            if (!p.d_uapple) r = philox4x32_10(gid, tick, 1u, (uint32_t)j, p.seed_lo, p.seed_hi);
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const int c = c4[q];
                if (c == kNoPoint) continue;
                const int v = sg[c];
                if ((v & kOcc) || v == SSD_CELL_APPLE) continue;               // cleanup.py:171, harvest.py:105
                uint32_t thr = tA;
                if (g.kind() == SSD_KIND_HARVEST) {
                    int cnt = 0;                                               // j^2+k^2 <= 2: the 3x3 block (harvest.py:107-116)
#pragma unroll
                    SSD_CHECK(c - g.W() - 1 >= 0 && c + g.W() + 1 < g.G());
#pragma unroll
                    for (int a = -1; a <= 1; ++a)
#pragma unroll
                        for (int b = -1; b <= 1; ++b) cnt += sg[c + a * g.W() + b] == SSD_CELL_APPLE;
                    thr = p.thr_harvest[cnt < 3 ? cnt : 3];
                }
                const uint32_t u = p.d_uapple ? p.d_uapple[(size_t)env * g.G() + c] : pick(r, q);
                if (u < thr) {
                    if (g.kind() == SSD_KIND_CLEANUP) { sg[c] = SSD_CELL_APPLE; ++n_new; }
                    else decided |= 1ull << (it * 4 + q);
                }
            }
        }
    }
    // waste: visit eligible waste points in ascending (order key, cell); the first success spawns one 'H'
    int wcell = -1;
    if (tW != 0) {
        uint32_t bk = 0xffffffffu, bc = 0xffffffffu;
        for (int j = lane; j < p.n_waste2; j += SW::kLanes) {
            const ushort2 pts = __ldg(reinterpret_cast<const ushort2*>(m->waste_pts) + j);
            uint4 r = make_uint4(0, 0, 0, 0);
            if (!p.d_uwaste) r = philox4x32_10(gid, tick, 2u, (uint32_t)j, p.seed_lo, p.seed_hi);
#pragma unroll
            for (int q = 0; q < 2; ++q) {
                const uint32_t c = q ? pts.y : pts.x;
                if (c == kNoPoint || (sg[c] & 0x7f) == SSD_CELL_WASTE) continue;   // cleanup.py:182
                const uint32_t u = p.d_uwaste ? p.d_uwaste[(size_t)env * g.G() + c] : (q ? r.z : r.x);
                const uint32_t key = p.d_uwaste ? p.d_wkey[(size_t)env * g.G() + c] : (q ? r.w : r.y);
                if (u < tW && (key < bk || (key == bk && c < bc))) { bk = key; bc = c; }
            }
        }
        const bool have = bc != 0xffffffffu;
        if (w.ballot(have)) {
            const uint32_t mk = w.rmin(have ? bk : 0xffffffffu);
            wcell = (int)w.rmin((have && bk == mk) ? bc : 0xffffffffu);
        }
    }
    w.sync();                                             // every decision read the pre-spawn grid
    if (g.kind() == SSD_KIND_HARVEST && decided) {
        int it = 0;
        for (int j = lane; j < p.n_apple4; j += SW::kLanes, ++it) {
            const ushort4 pts = __ldg(reinterpret_cast<const ushort4*>(m->apple_pts) + j);
            const uint16_t c4[4] = { pts.x, pts.y, pts.z, pts.w };
#pragma unroll
            for (int q = 0; q < 4; ++q) if ((decided >> (it * 4 + q)) & 1ull) sg[c4[q]] = SSD_CELL_APPLE;
        }
    }
    if (wcell >= 0 && lane == 0) sg[wcell] = (uint8_t)(SSD_CELL_WASTE | (sg[wcell] & kOcc));
    w.sync();
    if (g.kind() == SSD_KIND_HARVEST) n_new = __popcll(decided);
    if (w.ballot(n_new != 0)) apples += w.radd(n_new);
    waste += wcell >= 0;
}

// ------------------------------------------------------------------ render (map_env.py:360-379, 418-446, 795-815, 923-957)
#ifndef SSD_ST_HINT
#define SSD_ST_HINT ""                                        // cache operator of the observation stores (see profiles/r1_notes.md)
#endif
__device__ __forceinline__ void st_row32(uint8_t* dst, const uint32_t* w) {     // one full 32-byte sector per lane
    asm volatile("st.global" SSD_ST_HINT ".v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};"
                 :: "l"(dst), "r"(w[0]), "r"(w[1]), "r"(w[2]), "r"(w[3]), "r"(w[4]), "r"(w[5]), "r"(w[6]), "r"(w[7]) : "memory");
}
__device__ __forceinline__ void st_row16(uint8_t* dst, const uint32_t* w) {
    asm volatile("st.global" SSD_ST_HINT ".v4.b32 [%0], {%1,%2,%3,%4};" :: "l"(dst), "r"(w[0]), "r"(w[1]), "r"(w[2]), "r"(w[3]) : "memory");
}
__device__ __forceinline__ uint32_t prmt(uint32_t a, uint32_t b, uint32_t sel) {   // raw PRMT: no selector masking
    uint32_t d;
    asm("prmt.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(sel));
    return d;
}

// Row gather from NIBBLE-PACKED padded index maps.
// Two zero-padded copies of the map live in shared memory as 4-bit colour indices (8 cells per word):
//   M [padded row][col]   and   MT [padded col][row]
// so that for every orientation an output row of the rotated egocentric window (np.rot90 k=1,3,0,2 for
// LEFT, RIGHT, UP, DOWN; map_env.py:806-813) is ONE run of N nibbles:
//   UP    out[y][x] = M [r-V+y][c-V+x]  ascending       DOWN  out[y][x] = M [r+V-y][c+V-x]  descending
//   LEFT  out[y][x] = MT[c+V-y][r-V+x]  ascending       RIGHT out[y][x] = MT[c-V+y][r+V-x]  descending
// A lane owns one (agent, y) output row: it loads the <=5 words holding the run (nibble-reversing them for a
// descending run: one PRMT byte swap + one shift/merge), funnel-shifts them to pixel 0 (SHF), and every 16 bits
// of the result ARE the PRMT selector that looks 4 pixels up in the 8-entry colour LUT held in two registers
// per plane.  The finished row leaves the SM as one 32-byte (st.global.v8.b32 -> STG.E.ENL2.256) or 16-byte
// store per plane.  Indices: 0-5 cell codes, 6 outside, 7 agent.
__device__ __forceinline__ uint32_t nibble_reverse(uint32_t x) {
    const uint32_t y = prmt(x, 0u, 0x0123);
    return ((y << 4) & 0xf0f0f0f0u) | ((y >> 4) & 0x0f0f0f0fu);
}

template <int OCT_T>
struct RowUnit {                                              // one (agent, y) output row in flight
    uint32_t L[OCT_T > 0 ? OCT_T + 1 : 1];
    const uint32_t* w;
    uint8_t* dst;
    uint32_t sh, rev;
    bool valid;
};

template <int OCT_T, class SW, class GEO>
__device__ __forceinline__ void fetch_row(const SW& w, const GEO& g, const KParams& p, const uint8_t* pmap, uint8_t* gobs, int lane, int it, int units,
                                          uint32_t abase_sh, int astep, RowUnit<OCT_T>& U) {
    const int u = it * SW::kLanes + lane;
    U.valid = u < units;
    const int al = U.valid ? g.divN(u) : 0;
    const int y = u - al * g.N();
    const uint32_t ab = w.shfl(abase_sh, al);
    const int as = w.shfl(astep, al);
    U.w = reinterpret_cast<const uint32_t*>(pmap + (ab & 0xffffu) + (U.valid ? y * as : 0));
    U.sh = (ab >> 16) & 31u;
    U.rev = (ab >> 24) & 1u;
    U.dst = gobs + al * g.AS() + y * g.RP();
    {   // the <= OCT+1 words of the run, ascending or descending, must lie inside this warp's map tile
        const int nwords = (OCT_T > 0 ? OCT_T : ((g.RP() >> 2) + 1) >> 1) + 1;
        const long off = reinterpret_cast<const uint8_t*>(U.w) - pmap;
        SSD_CHECK(off >= (U.rev ? 4L * (nwords - 1) : 0L) && off + (U.rev ? 4L : 4L * nwords) <= g.PMS());
    }
    if constexpr (OCT_T > 0) {
        if (U.rev) {                                          // descending run: words walk down, nibbles are reversed
#pragma unroll
            for (int k = 0; k <= OCT_T; ++k) U.L[k] = nibble_reverse(*(U.w - k));
        } else {
#pragma unroll
            for (int k = 0; k <= OCT_T; ++k) U.L[k] = U.w[k]; // always in bounds: invalid lanes read agent 0's row 0
        }
    }
}

template <int OCT_T, class GEO>
__device__ __forceinline__ void emit_row(const GEO& g, const KParams& p, const RowUnit<OCT_T>& U, uint32_t lastmask) {
    if (!U.valid) return;
    const uint32_t sh = U.sh;
    if constexpr (OCT_T > 0) {
        uint32_t v[2 * OCT_T];
#pragma unroll
        for (int k = 0; k < OCT_T; ++k) { v[2 * k] = __funnelshift_r(U.L[k], U.L[k + 1], sh); v[2 * k + 1] = v[2 * k] >> 16; }
#pragma unroll
        for (int pl = 0; pl < 3; ++pl) {                      // one colour plane at a time: 2*OCT PRMTs, then one wide store
            const uint32_t t0 = p.lut8[2 * pl], t1 = p.lut8[2 * pl + 1];
            uint32_t o[2 * OCT_T];
#pragma unroll
            for (int k = 0; k < 2 * OCT_T; ++k) o[k] = prmt(t0, t1, v[k]);
            o[2 * OCT_T - 1] &= lastmask;
            if (OCT_T == 4) st_row32(U.dst + pl * g.PS(), o);
            else st_row16(U.dst + pl * g.PS(), o);
        }
    } else {
        const int WR = g.RP() >> 2, OCT = (WR + 1) >> 1;
        uint32_t prev = U.rev ? nibble_reverse(U.w[0]) : U.w[0];
        for (int k = 0; k < OCT; ++k) {
            const uint32_t cur = U.rev ? nibble_reverse(*(U.w - (k + 1))) : U.w[k + 1];
            uint32_t v = __funnelshift_r(prev, cur, sh);
            prev = cur;
            for (int hlf = 0; hlf < 2; ++hlf, v >>= 16) {
                const int wd = 2 * k + hlf;
                if (wd >= WR) break;
                const uint32_t m = wd == WR - 1 ? lastmask : 0xffffffffu;
#pragma unroll
                for (int pl = 0; pl < 3; ++pl)
                    *reinterpret_cast<uint32_t*>(U.dst + pl * g.PS() + 4 * wd) = prmt(p.lut8[2 * pl], p.lut8[2 * pl + 1], v) & m;
            }
        }
    }
}

template <int OCT_T, class SW, class GEO>
__device__ __forceinline__ void gather_rows(const SW& w, const GEO& g, const KParams& p, const uint8_t* pmap, uint8_t* gobs, int lane,
                                            uint32_t abase_sh, int astep) {
    const int WR = g.RP() >> 2;
    const uint32_t lastmask = 0xffffffffu >> (8 * (4 * WR - g.N()));
    const int units = p.n * g.N(), iters = (units + SW::kLanes - 1) / SW::kLanes;
    RowUnit<OCT_T> cur, nxt;
    fetch_row<OCT_T>(w, g, p, pmap, gobs, lane, 0, units, abase_sh, astep, cur);
    for (int it = 0; it < iters; ++it) {                      // software pipeline: row it+1 is fetched while row it is emitted
        if (it + 1 < iters) fetch_row<OCT_T>(w, g, p, pmap, gobs, lane, it + 1, units, abase_sh, astep, nxt);
        emit_row<OCT_T>(g, p, cur, lastmask);
        cur = nxt;
    }
}

// Nibble-packed padded maps, built one aligned 32-bit word (8 cells) at a time.  Map cells start at nibble
// LPn (V rounded up to 8) of a padded row, so only the last word of a row needs its tail set to "outside".
template <class SW, class GEO>
__device__ __forceinline__ void build_rowmap(const SW& w, const GEO& g, const KParams& p, const uint8_t* sg, uint8_t* map, int lane) {
    const uint32_t* sgw = reinterpret_cast<const uint32_t*>(sg);
    const int total = g.H() * g.nw8M();
#pragma unroll 4
    for (int i = lane; i < total; i += SW::kLanes) {
        const int r = g.divMW(i), j = i - r * g.nw8M();
        const int sb = r * g.W() + 8 * j, wi = sb >> 2;       // first source byte of this word (grid row r, col 8j)
        const uint32_t sel = 0x3210u + 0x1111u * (sb & 3);
        SSD_CHECK(wi >= 0 && 4 * (wi + 3) <= g.GS() + g.PMS());          // the last words may run into the (ignored) map tile
        const uint32_t a0 = sgw[wi], a1 = sgw[wi + 1], a2 = sgw[wi + 2];
        const uint32_t x0 = prmt(a0, a1, sel), x1 = prmt(a1, a2, sel);
        uint32_t w = prmt(x0 | (x0 >> 4), x1 | (x1 >> 4), 0x6420);            // 8 bytes -> 8 nibbles
        if (j == g.nw8M() - 1) w = (w & g.maskM8()) | (0x66666666u & ~g.maskM8());
        SSD_CHECK((r + g.V()) * g.pitchM() + (g.LPn() >> 1) + 4 * j + 4 <= (g.H() + 2 * g.V()) * g.pitchM());
        *reinterpret_cast<uint32_t*>(map + (r + g.V()) * g.pitchM() + (g.LPn() >> 1) + 4 * j) = w;
    }
}
// Transposed map from M: one lane transposes one 8 x 8 block of nibbles (8 words in, 8 words out) with three butterfly
// stages (swap 1-, 2-, 4-nibble sub-blocks across the diagonal).  Rows below the map come from M's first padding row and
// columns right of the map from M's tail nibbles, both "outside", which is exactly what MT's own padding holds.
template <class SW, class GEO>
__device__ __forceinline__ void build_colmap(const SW& w, const GEO& g, const KParams& p, const uint8_t* M, uint8_t* MT, int lane) {
    const int total = g.nw8T() * g.nw8M();
    for (int blk = lane; blk < total; blk += SW::kLanes) {
        const int i = g.divMW(blk), j = blk - i * g.nw8M();   // rows 8i.., columns 8j..
        uint32_t a[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const int r = min(8 * i + k, g.H());              // H = first bottom padding row of M
            SSD_CHECK((r + g.V()) * g.pitchM() + (g.LPn() >> 1) + 4 * j + 4 <= (g.H() + 2 * g.V()) * g.pitchM());
            a[k] = *reinterpret_cast<const uint32_t*>(M + (r + g.V()) * g.pitchM() + (g.LPn() >> 1) + 4 * j);
        }
#pragma unroll
        for (int k = 0; k < 8; k += 2) {                      // 1-nibble sub-blocks
            const uint32_t t = ((a[k] >> 4) ^ a[k + 1]) & 0x0f0f0f0fu;
            a[k + 1] ^= t; a[k] ^= t << 4;
        }
#pragma unroll
        for (int k = 0; k < 8; ++k) if (!(k & 2)) {           // 2-nibble sub-blocks
            const uint32_t t = ((a[k] >> 8) ^ a[k + 2]) & 0x00ff00ffu;
            a[k + 2] ^= t; a[k] ^= t << 8;
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) {                         // 4-nibble sub-blocks
            const uint32_t t = ((a[k] >> 16) ^ a[k + 4]) & 0x0000ffffu;
            a[k + 4] ^= t; a[k] ^= t << 16;
        }
#pragma unroll
        for (int x = 0; x < 8; ++x) {
            const int c = 8 * j + x;
            if (c < g.W()) {
                SSD_CHECK((c + g.V()) * g.pitchT() + (g.LPn() >> 1) + 4 * i + 4 <= (g.W() + 2 * g.V()) * g.pitchT());
                *reinterpret_cast<uint32_t*>(MT + (c + g.V()) * g.pitchT() + (g.LPn() >> 1) + 4 * i) = a[x];
            }
        }
    }
}

// "outside the map" everywhere (utility_funcs.py:58-116 without np.pad); issued while the state loads are in flight
template <class SW, class GEO>
__device__ __forceinline__ void fill_outside(const SW& w, const GEO& g, const KParams& p, uint8_t* pmap, int lane) {
    const uint4 v6 = make_uint4(0x66666666u, 0x66666666u, 0x66666666u, 0x66666666u);
    uint4* q = reinterpret_cast<uint4*>(pmap);
    const int n16 = g.PMS() >> 4;
#pragma unroll 8
    for (int i = lane; i < n16; i += SW::kLanes) q[i] = v6;           // trip count is a compile-time constant for the shipped maps
}

// Direct gather for small views (N <= 16, i.e. every Cleanup configuration): a lane owns one (agent, y) output row and
// reads its N cells straight from the staged byte grid -- one unaligned 16-byte run for UP / DOWN (5 aligned words + funnel
// shifts), N byte loads down a column for LEFT / RIGHT -- packs them to nibbles, replaces the cells outside the map by
// index 6, reverses the run for DOWN / RIGHT, and colours 4 pixels per PRMT as the map path does.  No padded maps are built.
// The grid tile has V rows of slack on both sides (zeroed at kernel start), so no load needs a bounds test; agents were
// written into the staged grid as index 7 by the caller.
template <bool COLS, class SW, class GEO>
__device__ __forceinline__ void gather_class(const SW& w, const GEO& g, const KParams& p, const uint8_t* sg, const uint8_t* rank2lane,
                                             uint8_t* gobs, int lane, uint32_t apack, int n_in_class) {
    const int N = g.N(), V = g.V(), W = g.W(), units = n_in_class * N, WR = g.RP() >> 2;
    const uint32_t lastmask = 0xffffffffu >> (8 * (4 * WR - N));
    for (int u0 = 0; u0 < units; u0 += SW::kLanes) {
        const int u = u0 + lane;
        const bool valid = u < units;
        const int rank = valid ? g.divN(u) : 0;
        const int y = u - rank * N;
        const int al = rank2lane[rank];
        const uint32_t ap = w.shfl(apack, al);
        if (!valid) continue;
        const int ar = (int)(ap & 0xffu), ac = (int)((ap >> 8) & 0xffu), o = (int)(ap >> 16);
        uint32_t b0, b1, b2, b3;                              // element i of the run (ascending map coordinate) in byte i
        int lo, hi;                                           // elements inside the map: [lo, hi)
        if (!COLS) {                                          // UP / DOWN: a run along map row R
            const int R = o == 2 ? ar - V + y : ar + V - y;
            const bool rv = (unsigned)R < (unsigned)g.H();
            lo = rv ? max(0, V - ac) : 0;
            hi = rv ? min(N, W + V - ac) : 0;
            const int start = (rv ? R : 0) * W + ac - V;      // >= -V: inside the front slack
            SSD_CHECK(start >= -g.padF() + 32 && start + 20 <= g.GS() + g.padB());
            const uint32_t* wp = reinterpret_cast<const uint32_t*>(sg + (start & ~3));
            const uint32_t sh = (uint32_t)(start & 3) * 8u;
            const uint32_t a0 = wp[0], a1 = wp[1], a2 = wp[2], a3 = wp[3], a4 = wp[4];
            b0 = __funnelshift_r(a0, a1, sh); b1 = __funnelshift_r(a1, a2, sh);
            b2 = __funnelshift_r(a2, a3, sh); b3 = __funnelshift_r(a3, a4, sh);
        } else {                                              // LEFT / RIGHT: a run down map column C
            const int C = o == 0 ? ac + V - y : ac - V + y;
            const bool cv = (unsigned)C < (unsigned)W;
            lo = cv ? max(0, V - ar) : 0;
            hi = cv ? min(N, g.H() + V - ar) : 0;
            const uint8_t* col = sg + (ar - V) * W + (cv ? C : 0);     // rows ar-V .. ar+V: at most V rows into either slack
            SSD_CHECK((ar - V) * W >= -g.padF() + 32 && (ar + V) * W + W <= g.GS() + g.padB());
            uint32_t e[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) e[i] = i < N ? (uint32_t)col[i * W] : 0u;
            b0 = e[0] | (e[1] << 8) | (e[2] << 16) | (e[3] << 24);
            b1 = e[4] | (e[5] << 8) | (e[6] << 16) | (e[7] << 24);
            b2 = e[8] | (e[9] << 8) | (e[10] << 16) | (e[11] << 24);
            b3 = e[12] | (e[13] << 8) | (e[14] << 16) | (e[15] << 24);
        }
        // bytes -> nibbles (every byte is a colour index < 8), cells outside the map -> 6
        uint32_t n0 = prmt(b0 | (b0 >> 4), b1 | (b1 >> 4), 0x6420);
        uint32_t n1 = prmt(b2 | (b2 >> 4), b3 | (b3 >> 4), 0x6420);
        const unsigned long long vm = (hi >= 16 ? ~0ull : ((1ull << (4 * hi)) - 1ull)) & ~((1ull << (4 * lo)) - 1ull);
        const uint32_t m0 = (uint32_t)vm, m1 = (uint32_t)(vm >> 32);
        n0 = (n0 & m0) | (0x66666666u & ~m0);
        n1 = (n1 & m1) | (0x66666666u & ~m1);
        if (o & 1) {                                          // RIGHT / DOWN: out[x] = element N-1-x (np.rot90 k=3 / k=2)
            const uint32_t r1 = nibble_reverse(n0), r0w = nibble_reverse(n1);       // nibble j of (r1:r0w) = element 15-j
            const unsigned long long rr = (((unsigned long long)r1 << 32) | r0w) >> (4 * (16 - N));   // up to 60 bits for tiny views
            n0 = (uint32_t)rr;
            n1 = (uint32_t)(rr >> 32);
        }
        const uint32_t v[4] = { n0, n0 >> 16, n1, n1 >> 16 };
        uint8_t* dst = gobs + al * g.AS() + y * g.RP();
#pragma unroll
        for (int pl = 0; pl < 3; ++pl) {
            const uint32_t t0 = p.lut8[2 * pl], t1 = p.lut8[2 * pl + 1];
            uint32_t o4[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) o4[k] = prmt(t0, t1, v[k]);
            if (WR == 4) { o4[3] &= lastmask; st_row16(dst + pl * g.PS(), o4); }
            else {
#pragma unroll
                for (int k = 0; k < 4; ++k)
                    if (k < WR) *reinterpret_cast<uint32_t*>(dst + pl * g.PS() + 4 * k) = k == WR - 1 ? (o4[k] & lastmask) : o4[k];
            }
        }
    }
}

// The rows of UP / DOWN agents and of LEFT / RIGHT agents are gathered in two separate passes, so that a trip of the warp runs
// ONE of the two load schemes instead of both under divergence.  `tbl` (32 bytes at the very start of the tile's front slack,
// never touched by a gather load) maps (class, rank within class) -> agent lane.
template <class SW, class GEO>
__device__ __forceinline__ void gather_direct(const SW& w, const GEO& g, const KParams& p, const uint8_t* sg, uint8_t* tbl, uint8_t* gobs, int lane,
                                              bool is_agent, int r0, int c0, int ori) {
    const uint32_t apack = (uint32_t)r0 | ((uint32_t)c0 << 8) | ((uint32_t)ori << 16);
    const unsigned m_rows = w.ballot(is_agent && ori >= 2), m_cols = w.ballot(is_agent && ori < 2);
    if (is_agent) {
        const bool cols = ori < 2;
        tbl[(cols ? 16 : 0) + __popc((cols ? m_cols : m_rows) & ((1u << lane) - 1u))] = (uint8_t)lane;
    }
    w.sync();
    if (m_rows) gather_class<false>(w, g, p, sg, tbl, gobs, lane, apack, __popc(m_rows));
    if (m_cols) gather_class<true>(w, g, p, sg, tbl + 16, gobs, lane, apack, __popc(m_cols));
}

template <class SW, class GEO>
__device__ __forceinline__ void render(const SW& w, const GEO& g, const KParams& p, const uint8_t* sg, uint8_t* pmap, const uint32_t* lut_s,
                                       int lane, bool is_agent, int pos, int ori, int env, unsigned same) {
    const bool top = is_agent && lane == 31 - __clz(same);   // later index overwrites (map_env.py:370); same = lanes on this lane's cell
    int r0 = 0, c0 = 0;
    if (is_agent) { r0 = g.divW(pos); c0 = pos - r0 * g.W(); }

    if (p.state_rgb) {                                        // get_state: unrotated full map (map_env.py:950-957)
        uint8_t* out = p.state_rgb + (size_t)env * 3 * g.G();
        for (int c = lane; c < g.G(); c += SW::kLanes) {
            const uint32_t rgb = lut_s[sg[c]];
            out[c] = (uint8_t)rgb; out[g.G() + c] = (uint8_t)(rgb >> 8); out[2 * g.G() + c] = (uint8_t)(rgb >> 16);
        }
        w.sync();                                         // orders the agent overlay after the cell colours
        if (top) {
            const uint32_t rgb = lut_s[agent_colour_index(lane)];
            out[pos] = (uint8_t)rgb; out[g.G() + pos] = (uint8_t)(rgb >> 8); out[2 * g.G() + pos] = (uint8_t)(rgb >> 16);
        }
    }
    if (!p.obs) return;
    uint8_t* gobs = p.obs + (size_t)env * (p.n * g.AS());

    if (g.direct()) {
        gather_direct(w, g, p, sg, const_cast<uint8_t*>(sg) - g.padF(), gobs, lane, is_agent, r0, c0, ori);
    } else {
        // the maps were pre-filled with "outside the map" at kernel start (fill_outside); now the cells (agents are already in
        // the staged grid as index 7)
        // per agent: which map, first word of its window row 0, row step, funnel shift, direction   (o: 0 LEFT 1 RIGHT 2 UP 3 DOWN)
        const bool colmap = ori < 2, rev = (ori == 1) || (ori == 3);
        const int rowc = colmap ? c0 : r0;                        // coordinate that selects the map row
        const int runc = colmap ? r0 : c0;                        // coordinate along the run
        const int pitch = colmap ? g.pitchT() : g.pitchM();
        const bool down = (ori == 0) || (ori == 3);               // window row y walks towards smaller map rows
        const int s = g.LPn() + runc + (rev ? g.V() : -g.V());    // first pixel of the run (highest nibble when descending)
        const uint32_t abase_sh = (uint32_t)(g.off_map(ori) + (rowc + (down ? 2 * g.V() : 0)) * pitch + ((s >> 3) << 2))
                                  | ((uint32_t)((rev ? 7 - (s & 7) : (s & 7)) * 4) << 16) | ((uint32_t)rev << 24);
        const int astep = down ? -pitch : pitch;
        const bool needT = w.ballot(is_agent && ori < 2) != 0;
        uint8_t* MT = pmap + g.off_map(0);
        uint8_t* M = pmap + g.off_map(2);
        build_rowmap(w, g, p, sg, M, lane);
        w.sync();
        if (needT) {                                              // LEFT / RIGHT views read the transpose (agents included)
            build_colmap(w, g, p, M, MT, lane);
            w.sync();
        }
        if (g.RP() == 32) gather_rows<4>(w, g, p, pmap, gobs, lane, abase_sh, astep);
        else gather_rows<0>(w, g, p, pmap, gobs, lane, abase_sh, astep);
    }
    const int tail = g.AS() - 3 * g.PS();                         // pad bytes per agent block (0 for the shipped views)
    if (tail) for (int i = lane; i < p.n * (tail >> 2); i += SW::kLanes)
        *reinterpret_cast<uint32_t*>(gobs + (i / (tail >> 2)) * g.AS() + 3 * g.PS() + 4 * (i % (tail >> 2))) = 0u;
    if (!p.agents_uniform) {
        // full-colour scheme: every visible agent is re-painted with its own colour (after the row stores)
        w.sync();
        const uint32_t apack = (uint32_t)r0 | ((uint32_t)c0 << 10) | ((uint32_t)ori << 20);
        const int pairs = p.n * p.n, iters = (pairs + SW::kLanes - 1) / SW::kLanes;
        for (int it = 0; it < iters; ++it) {
            const int q = it * SW::kLanes + lane;
            const bool valid = q < pairs;
            const int al = valid ? q / p.n : 0, j = valid ? q - al * p.n : 0;
            const uint32_t api = w.shfl(apack, al), apj = w.shfl(apack, j);
            const bool topj = w.shfl((int)top, j);
            if (!valid || !topj) continue;
            const int a = (int)(apj & 1023) - (int)(api & 1023) + g.V(), b = (int)((apj >> 10) & 1023) - (int)((api >> 10) & 1023) + g.V();
            if (a < 0 || a >= g.N() || b < 0 || b >= g.N()) continue;
            const int o = api >> 20;
            int y, x;
            if (o == 2) { y = a; x = b; } else if (o == 0) { x = a; y = g.N() - 1 - b; }
            else if (o == 3) { y = g.N() - 1 - a; x = g.N() - 1 - b; } else { x = g.N() - 1 - a; y = b; }
            const uint32_t rgb = lut_s[agent_colour_index(j)];
            uint8_t* d = gobs + al * g.AS() + y * g.RP() + x;
            d[0] = (uint8_t)rgb; d[g.PS()] = (uint8_t)(rgb >> 8); d[2 * g.PS()] = (uint8_t)(rgb >> 16);
        }
    }
}

// ------------------------------------------------------------------ the kernel
template <int MODE, class GEO, int LPE>
__global__ void __launch_bounds__(kWarps * 32, SSD_MIN_BLOCKS) ssd_kernel(const __grid_constant__ KParams p) {
    using SW = SubWarp<LPE>;
    const GEO g(p);
    extern __shared__ __align__(16) uint8_t smem[];
    __shared__ uint32_t lut_all[kWarps * 2][16];              // colour LUT, one private copy per (sub-)warp: no CTA barrier anywhere
    const int warp = threadIdx.x >> 5;
    const SW w(threadIdx.x & 31);
    const int lane = w.lane;                                  // lane within the sub-warp that owns an env
    uint32_t* lut_s = lut_all[warp * SW::kEnvs + w.sub];
    // per-env tile: [front slack][grid GS][back slack][nibble maps (map path only)]
    uint8_t* tile = smem + ((size_t)warp * SW::kEnvs + w.sub) * (g.padF() + g.GS() + g.padB() + g.PMS());
    uint8_t* sg = tile + g.padF();
    uint8_t* pmap = sg + g.GS() + g.padB();
    const bool is_agent = lane < p.n;
    // One env instance per warp and per launch.  (A persistent loop over env instances was measured and dropped: warps that
    // start together stay phase-locked -- everybody resolves moves, then everybody stores observations -- which turns a
    // 65536-env launch into 14 back-to-back single-wave steps: Harvest 204 us instead of 167 us.  Letting the block scheduler
    // start a fresh CTA whenever one retires decorrelates the phases; profiles/r2_notes.md.)
    const int env = p.env0 + (blockIdx.x * kWarps + warp) * SW::kEnvs + w.sub;
    if (env >= p.env1) return;
    // Prologue that touches nothing an earlier launch writes (colour LUT, slack / "outside" fill of the staging tile).  Under
    // programmatic dependent launch (small launches, p.pdl) it runs first, while the previous launch of the stream is still
    // storing observations, and the state loads follow the grid dependency; otherwise it runs under the state loads.
    auto prologue = [&]() {
        if (lane < 16) lut_s[lane] = p.lut[lane];             // (made visible by w.sync below)
        if (p.obs) {
            if (g.direct()) {                                 // zero the slack around the grid (read, then masked, by the gather)
                for (int i = lane; i < (g.padF() >> 4); i += SW::kLanes) reinterpret_cast<uint4*>(tile)[i] = make_uint4(0, 0, 0, 0);
                for (int i = lane; i < (g.padB() >> 4); i += SW::kLanes) reinterpret_cast<uint4*>(sg + g.GS())[i] = make_uint4(0, 0, 0, 0);
            } else fill_outside(w, g, p, pmap, lane);         // "outside the map"
        }
    };
    if (p.pdl) {
        prologue();
        asm volatile("griddepcontrol.wait;" ::: "memory");    // the previous launch of the stream has completed, its writes are visible
    }
    {
        if (MODE == MODE_RESET && p.mask && __ldcg(p.mask + env) == 0) return;
        const uint32_t gid = p.gid_base + (uint32_t)env;
        // every global load of the step is issued up front.  Mutable state is read past L1 (ld.global.cg): each entry is read once
        // per step anyway, and a launch that became resident before its predecessors finished (programmatic dependent launch,
        // concurrent env ranges sharing a cache line at a range boundary) must never see a line an earlier CTA left in this SM's L1.
        const uint32_t tick = __ldcg(p.tick + env);
        const int t_prev = MODE == MODE_STEP ? __ldcg(p.t + env) : 0;
        const uint32_t cnt = MODE == MODE_STEP ? __ldcg(p.counts + env) : 0u;
        const int act = (MODE == MODE_STEP && is_agent) ? (int)__ldcg(p.actions + (size_t)env * p.n + lane) : 255;
        int pos = -1 - lane, ori = 0, ep_ret = 0;
        uint32_t a_rec = 0;
        if (MODE != MODE_RESET && is_agent) {
            a_rec = __ldcg(p.agent + (size_t)env * p.NA + lane);
            ep_ret = __ldcg(p.ep_ret + (size_t)env * p.NA + lane);
        }
        const int n16 = g.GS() >> 4;
        {
            const uint4* src = MODE == MODE_RESET ? reinterpret_cast<const uint4*>(p.map->base_grid)
                                                  : reinterpret_cast<const uint4*>(p.grid + (size_t)env * g.GS());
            uint4 g0 = make_uint4(0, 0, 0, 0);
            if (lane < n16) g0 = __ldcg(src + lane);
            if (!p.pdl) prologue();                           // while the state loads are in flight
            if (lane < n16) reinterpret_cast<uint4*>(sg)[lane] = g0;
            for (int i = lane + SW::kLanes; i < n16; i += SW::kLanes) reinterpret_cast<uint4*>(sg)[i] = __ldcg(src + i);
        }
        if (MODE != MODE_RESET && is_agent) {
            pos = (int)(a_rec & 0xff) * g.W() + (int)((a_rec >> 8) & 0xff);
            ori = (int)((a_rec >> 16) & 3);
        }
        w.sync();

        int apples = (int)(cnt & 0xffffu), waste = (int)(cnt >> 16);
        // lanes standing on this lane's cell: MATCH.ANY is slow (3-4 % of the step each in the profile), positions do not change
        // after update_moves, so it is evaluated once and shared by consume, the occupancy strip and the render
        unsigned same = 1u << lane;                            // reset: distinct spawn points
        if (MODE == MODE_RENDER) same = equal_lanes(w, pos, p.n);
        if (MODE == MODE_STEP) {
            int reward = 0, clean_num = 0;
            update_moves(w, g, p, sg, lane, is_agent, act, pos, ori, env, gid, tick);                 // map_env.py:251
            // consume in index order: the lowest index on a cell eats the apple (253-256); mark occupancy
            same = surely_distinct(w, pos, is_agent, w.ballot(is_agent)) ? (1u << lane) : equal_lanes(w, pos, p.n);
            const int here = is_agent ? sg[pos] : 0;
            const bool first = is_agent && lane == __ffs(same) - 1;
            const unsigned ate = w.ballot(first && here == SSD_CELL_APPLE);
            w.sync();
            if (first) {
                if (here == SSD_CELL_APPLE) { reward += 1; sg[pos] = kOcc | SSD_CELL_EMPTY; }
                else sg[pos] = (uint8_t)(kOcc | here);
            }
            apples -= __popc(ate);
            w.sync();
            beams(w, g, p, sg, lane, is_agent, act, pos, ori, reward, clean_num, waste);              // 259-260
            spawn(w, g, p, sg, lane, env, gid, tick, apples, waste);                                  // 263, 291-292
            const int t = t_prev + 1;
            if (is_agent) {
                p.reward[(size_t)env * p.n + lane] = (int8_t)reward;
                p.clean[(size_t)env * p.n + lane] = (uint8_t)clean_num;
                ep_ret += reward;                                                               // 885-888
            }
            if (lane == 0) {
                p.apple_cnt[env] = (uint16_t)apples;
                p.done[env] = t >= p.episode_limit;                                             // 890-894
                p.t[env] = t;
                p.tick[env] = tick + 1;
                p.counts[env] = (uint32_t)apples | ((uint32_t)waste << 16);
            }
        } else if (MODE == MODE_RESET) {
            // setup_agents: agent i takes the free spawn point with the largest (key, cell) (map_env.py:771-784)
            const bool is_sp = lane < p.n_spawn;
            const int mycell = is_sp ? (int)__ldg(&p.map->spawn_pts[lane]) : -1;
            bool taken = false;
            for (int i = 0; i < p.n; ++i) {
                uint32_t key = 0;
                if (p.random_spawn && is_sp)
                    key = p.d_spawnkey ? p.d_spawnkey[((size_t)env * p.n + i) * g.G() + mycell]
                                       : philox_word(p, gid, tick, 3u, (uint32_t)(i * p.n_spawn + lane));
                const bool cand = is_sp && !taken;
                const uint32_t mk = w.rmax(cand ? key : 0u);
                const unsigned eq = w.ballot(cand && key == mk);
                const int win = 31 - __clz(eq);
                const int cell = w.shfl(mycell, win);
                if (lane == win) taken = true;
                if (lane == i) pos = cell;
            }
            if (is_agent) {                                                                     // spawn_rotation (786-793)
                if (p.spawn_rot >= 0) ori = p.spawn_rot;
                else ori = p.d_rot ? (int)(p.d_rot[(size_t)env * p.n + lane] & 3) : (int)(philox_word(p, gid, tick, 4u, (uint32_t)lane) >> 30);
            }
            if (is_agent) sg[pos] |= kOcc;                         // distinct spawn points: no two lanes share a cell
            w.sync();
            apples = p.base_apples; waste = p.base_waste;
            spawn(w, g, p, sg, lane, env, gid, tick, apples, waste);                                  // custom_map_update (313)
            ep_ret = 0;
            if (lane == 0) { p.t[env] = 0; p.tick[env] = tick + 1; p.counts[env] = (uint32_t)apples | ((uint32_t)waste << 16); }
        }

        // the next launch of this stream may become resident now (its prologue overlaps this launch's stores; it still waits
        // for this whole grid before it reads any state)
        if (p.pdl) asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
        if (MODE != MODE_RENDER) {                                 // strip the occupancy bits, write the state back
            if (is_agent && lane == __ffs(same) - 1) sg[pos] &= 0x7f;
            w.sync();
            uint4* dst = reinterpret_cast<uint4*>(p.grid + (size_t)env * g.GS());
            warp_for<SW::kLanes>(g.GS() >> 4, lane, [&](int i) { dst[i] = reinterpret_cast<const uint4*>(sg)[i]; });
            if (is_agent) {
                const int r = g.divW(pos);
                p.agent[(size_t)env * p.NA + lane] = (uint32_t)r | ((uint32_t)(pos - r * g.W()) << 8) | ((uint32_t)ori << 16);
                p.ep_ret[(size_t)env * p.NA + lane] = ep_ret;
            }
        }
        if (p.obs || p.state_rgb) {
            // agent overlay: every occupied cell shows an agent (index 7; which agent only matters for the full-colour repaint and
            // the state image, both handled in render).  Same value from every lane on a shared cell: no race.
            w.sync();                                          // the write-back above has read the staged grid
            if (is_agent) sg[pos] = 7;
            w.sync();
            render(w, g, p, sg, pmap, lut_s, lane, is_agent, pos, ori, env, same);
        }
    }
}

// ------------------------------------------------------------------ incentive bookkeeping (homophily_learner.py:98-115)
__global__ void incentive_kernel(const long long* __restrict__ a_inc, const float* __restrict__ reward, long long rows, int n,
                                 float incentive, float cost, float ratio, float T, int recip,
                                 float* __restrict__ r_env, float* __restrict__ r_inc, float* __restrict__ sgn) {
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= rows * n) return;
    const long long row = idx / n;
    const int k = (int)(idx - row * n);
    const long long* a = a_inc + row * n * n;
    int give = 0, rp = 0, rn = 0;
    for (int j = 0; j < n; ++j) {
        if (j == k) continue;                                  // inc_mask_actions = 1 - eye
        give += a[k * n + j] != 0;
        const long long v = a[j * n + k];
        rp += v == 1; rn += v == 2;
    }
    const float rv = (float)(rp - rn), r = reward[idx];
    const float e = __fadd_rn(r, __fmul_rn(__fmul_rn(rv, ratio), incentive));
    const float c = __fsub_rn(r, __fmul_rn(__fmul_rn((float)give, cost), incentive));
    if (recip) { const float inv = __fdiv_rn(1.0f, T); r_env[idx] = __fmul_rn(e, inv); r_inc[idx] = __fmul_rn(c, inv); }
    else { r_env[idx] = __fdiv_rn(e, T); r_inc[idx] = __fdiv_rn(c, T); }
    if (sgn) sgn[idx] = rv > 0.f ? 1.f : (rv < 0.f ? -1.f : 0.f);
}

thread_local int g_last_cuda_error = 0;

inline int cuda_fail(cudaError_t e) { g_last_cuda_error = (int)e; return SSD_ERR_CUDA; }
#define SSD_CUDA(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) return cuda_fail(e_); } while (0)

inline int round_up(int v, int m) { return (v + m - 1) / m * m; }

// Makes the handle's device current for the duration of a call and restores the caller's device afterwards.
struct DeviceGuard {
    int prev = -1;
    bool switched = false;
    cudaError_t err = cudaSuccess;
    explicit DeviceGuard(int dev) {
        err = cudaGetDevice(&prev);
        if (err == cudaSuccess && prev != dev) { err = cudaSetDevice(dev); switched = err == cudaSuccess; }
    }
    ~DeviceGuard() { if (switched) cudaSetDevice(prev); }
};

uint32_t magic20(int d, int max_x) {
    // smallest m with floor(x*m / 2^20) == floor(x / d) for all 0 <= x <= max_x (verified exhaustively)
    uint32_t m = (uint32_t)(((1u << 20) + d - 1) / d);
    for (int x = 0; x <= max_x; ++x)
        if ((int)(((uint64_t)x * m) >> 20) != x / d) return 0;
    return m;
}

}  // namespace

struct ssd_handle {
    KParams kp;
    MapDev* d_map;
    int device;
    size_t smem_bytes;
    int64_t launches;
    int force_generic;                                        // SSD_B200_GENERIC=1: always use the runtime-geometry kernels
    int pdl;                                                  // programmatic dependent launch for small launches (SSD_B200_PDL=0 turns it off)
    int lanes_per_env;                                        // 16: two envs per warp (needs <= 16 spawn points); 32: one
};

static int fill_common(const ssd_handle* h, const ssd_state* st, const ssd_draws* d, KParams& k) {
    if (!h || !st || !st->grid || !st->agent || !st->ep_ret || !st->t || !st->tick || !st->counts) return SSD_ERR_INVALID;
    k = h->kp;
    k.env0 = 0; k.env1 = k.B;
    k.grid = st->grid; k.agent = st->agent; k.ep_ret = st->ep_ret; k.t = st->t; k.tick = st->tick; k.counts = st->counts;
    if (d) {
        if ((d->u_waste == nullptr) != (d->wkey == nullptr)) return SSD_ERR_INVALID;
        k.d_prio = d->prio; k.d_uapple = d->u_apple; k.d_uwaste = d->u_waste; k.d_wkey = d->wkey;
        k.d_spawnkey = d->spawn_key; k.d_rot = d->rot;
    }
    return SSD_OK;
}

template <int MODE, class GEO, int LPE>
static int launch_lpe(ssd_handle* h, const KParams& k, void* stream) {
    DeviceGuard guard(h->device);
    SSD_CUDA(guard.err);
    {   // cudaFuncAttributeMaxDynamicSharedMemorySize belongs to the kernel FUNCTION (per device), not to a handle: all
        // runtime-geometry handles share one instantiation, so the limit is only ever raised (high-water mark).
        static std::mutex mu;
        static int high_water[kMaxDevices] = {};
        std::lock_guard<std::mutex> lock(mu);
        const int dev = h->device;
        if (dev < 0 || dev >= kMaxDevices) return SSD_ERR_INVALID;
        if ((int)h->smem_bytes > high_water[dev]) {
            SSD_CUDA(cudaFuncSetAttribute(ssd_kernel<MODE, GEO, LPE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->smem_bytes));
            high_water[dev] = (int)h->smem_bytes;
        }
    }
    const int envs_per_cta = kWarps * (32 / LPE);
    int grid = (k.env1 - k.env0 + envs_per_cta - 1) / envs_per_cta;
    if (grid <= 0) return SSD_OK;

    if (k.pdl) {                                               // programmatic dependent launch (see launch_geo)
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3((unsigned)grid); cfg.blockDim = dim3(kWarps * 32); cfg.dynamicSmemBytes = h->smem_bytes;
        cfg.stream = (cudaStream_t)stream;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[0].val.programmaticStreamSerializationAllowed = 1;
        cfg.attrs = attr; cfg.numAttrs = 1;
        SSD_CUDA(cudaLaunchKernelEx(&cfg, ssd_kernel<MODE, GEO, LPE>, k));
    } else {
        ssd_kernel<MODE, GEO, LPE><<<grid, kWarps * 32, h->smem_bytes, (cudaStream_t)stream>>>(k);
    }
    ++h->launches;
    SSD_CUDA(cudaGetLastError());
    return SSD_OK;
}

// Launches of at most a few waves are bound by the latency of their own chain (launch -> state loads -> logic -> stores), so
// consecutive launches of a stream are chained programmatically: the next one becomes resident and runs its prologue while this
// one stores observations (cleanup5, 4 096 envs in 8 ranges: 8.7 -> 7.2 us per step; harvest5: 12.1 -> 10.2 us).  Large
// launches gain nothing and pay for the prologue no longer running under the state loads, so they keep the plain launch.
constexpr int kPdlMaxEnvs = 2048;

template <int MODE, class GEO>
static int launch_geo(ssd_handle* h, const KParams& k_in, void* stream) {
    KParams k = k_in;
    k.pdl = h->pdl && (k.env1 - k.env0) <= kPdlMaxEnvs;
#ifdef SSD_ENABLE_LPE16                                        // experiment: two envs per warp (measured slower, profiles/r1_notes.md)
    if (h->lanes_per_env == 16) return launch_lpe<MODE, GEO, 16>(h, k, stream);
#endif
    return launch_lpe<MODE, GEO, 32>(h, k, stream);
}

// The reference's shipped geometries (constants.py maps x yaml view sizes) get compile-time index arithmetic.
template <int MODE>
static int launch(ssd_handle* h, const KParams& k, void* stream) {
    if (!h->force_generic) {
        if (k.kind == SSD_KIND_HARVEST && k.H == 9 && k.W == 38 && k.V == 15) return launch_geo<MODE, GeoS<SSD_KIND_HARVEST, 9, 38, 15>>(h, k, stream);
        if (k.kind == SSD_KIND_HARVEST && k.H == 9 && k.W == 38 && k.V == 7) return launch_geo<MODE, GeoS<SSD_KIND_HARVEST, 9, 38, 7>>(h, k, stream);
        if (k.kind == SSD_KIND_CLEANUP && k.H == 25 && k.W == 18 && k.V == 7) return launch_geo<MODE, GeoS<SSD_KIND_CLEANUP, 25, 18, 7>>(h, k, stream);
        if (k.kind == SSD_KIND_CLEANUP && k.H == 48 && k.W == 18 && k.V == 7) return launch_geo<MODE, GeoS<SSD_KIND_CLEANUP, 48, 18, 7>>(h, k, stream);
        if (k.kind == SSD_KIND_CLEANUP && k.H == 10 && k.W == 10 && k.V == 7) return launch_geo<MODE, GeoS<SSD_KIND_CLEANUP, 10, 10, 7>>(h, k, stream);
    }
    return launch_geo<MODE, GeoD>(h, k, stream);
}

extern "C" {

int ssd_abi_version(void) { return SSD_B200_ABI_VERSION; }

const char* ssd_error_string(int code) {
    switch (code) {
        case SSD_OK: return "ok";
        case SSD_ERR_INVALID: return "invalid argument or unsupported geometry";
        case SSD_ERR_CUDA: return "CUDA runtime error";
        case SSD_ERR_SPAWN: return "There are not enough spawn points! Check your map?";
        case SSD_ERR_MAP: return "map must be wall-enclosed and use the reference alphabet";
        default: return "unknown error";
    }
}

int ssd_last_cuda_error(void) { return g_last_cuda_error; }

uint32_t ssd_prob_to_threshold(double p) {
    if (!(p > 0.0)) return 0u;
    const double t = ceil(p * 4294967296.0);
    return t >= 4294967295.0 ? 0xFFFFFFFFu : (uint32_t)t;
}

int ssd_create(const ssd_config* cfg, ssd_handle** out) {
    if (!cfg || !out || !cfg->ascii_map) return SSD_ERR_INVALID;
    *out = nullptr;
    const int H = cfg->height, W = cfg->width, G = H * W, n = cfg->n_agents, V = cfg->view;
    if (cfg->kind != SSD_KIND_CLEANUP && cfg->kind != SSD_KIND_HARVEST) return SSD_ERR_INVALID;
    if (H < 3 || W < 3 || H > 255 || W > 255 || G > SSD_MAX_CELLS) return SSD_ERR_INVALID;
    if (n < 1 || n > SSD_MAX_AGENTS || V < 1 || V > 31 || cfg->n_envs < 1) return SSD_ERR_INVALID;
    if (cfg->beam_len < 0 || cfg->episode_limit < 1 || cfg->spawn_rotation > 3) return SSD_ERR_INVALID;
    if (abs(cfg->fire_cost) + n * abs(cfg->hit_penalty) + 1 > 127) return SSD_ERR_INVALID;   // reward is int8

    MapDev* hm = new (std::nothrow) MapDev;
    if (!hm) return SSD_ERR_INVALID;
    memset(hm, 0, sizeof(MapDev));
    int na = 0, nw = 0, ns = 0;
    for (int c = 0; c < G; ++c) {
        const char ch = cfg->ascii_map[c];
        const int r = c / W, col = c % W;
        const bool border = r == 0 || col == 0 || r == H - 1 || col == W - 1;
        if (border && ch != '@') { delete hm; return SSD_ERR_MAP; }
        uint8_t code = SSD_CELL_EMPTY;
        switch (ch) {
            case '@': code = SSD_CELL_WALL; break;
            case ' ': break;
            case 'P': if (ns < SSD_MAX_SPAWN) hm->spawn_pts[ns] = (uint16_t)c; ++ns; break;   // map_env.py:143-146
            case 'A': if (cfg->kind == SSD_KIND_HARVEST) { code = SSD_CELL_APPLE; hm->apple_pts[na++] = (uint16_t)c; } break;
            case 'B': if (cfg->kind == SSD_KIND_CLEANUP) hm->apple_pts[na++] = (uint16_t)c; break;   // cleanup.py:81-82
            case 'H': if (cfg->kind == SSD_KIND_CLEANUP) { code = SSD_CELL_WASTE; hm->waste_pts[nw++] = (uint16_t)c; } break;
            case 'R': if (cfg->kind == SSD_KIND_CLEANUP) code = SSD_CELL_RIVER; break;
            case 'S': if (cfg->kind == SSD_KIND_CLEANUP) code = SSD_CELL_STREAM; break;
            default: delete hm; return SSD_ERR_MAP;
        }
        hm->base_grid[c] = code;
    }
    if (ns < n) { delete hm; return SSD_ERR_SPAWN; }
    if (ns > SSD_MAX_SPAWN) { delete hm; return SSD_ERR_INVALID; }
    for (int k = na; k < round_up(na, 4) + 4 && k < SSD_MAX_CELLS + 4; ++k) hm->apple_pts[k] = kNoPoint;
    for (int k = nw; k < round_up(nw, 2) + 2 && k < SSD_MAX_CELLS + 4; ++k) hm->waste_pts[k] = kNoPoint;
    if (cfg->kind == SSD_KIND_CLEANUP) {
        if (cfg->n_waste_lut != (uint32_t)nw + 1 || !cfg->thr_apple || !cfg->thr_waste) { delete hm; return SSD_ERR_INVALID; }
        memcpy(hm->thr_apple, cfg->thr_apple, sizeof(uint32_t) * (nw + 1));
        memcpy(hm->thr_waste, cfg->thr_waste, sizeof(uint32_t) * (nw + 1));
    }

    ssd_handle* h = new (std::nothrow) ssd_handle;
    if (!h) { delete hm; return SSD_ERR_INVALID; }
    memset(h, 0, sizeof(*h));
    KParams& k = h->kp;
    const GeoVals gv = make_geo(cfg->kind, H, W, V);
    k.kind = cfg->kind; k.B = cfg->n_envs; k.n = n; k.H = H; k.W = W; k.G = G; k.V = V; k.N = gv.N;
    k.GS = gv.GS; k.NA = round_up(n, 4);
    k.RP = gv.RP; k.PS = gv.PS; k.AS = gv.AS; k.ES = n * k.AS;
    k.LPn = gv.LPn; k.nw8M = gv.nw8M; k.nw8T = gv.nw8T; k.maskM8 = gv.maskM8; k.maskT8 = gv.maskT8;
    k.pitchM = gv.pitchM; k.pitchT = gv.pitchT; k.PMS = gv.PMS;
    k.direct = gv.direct; k.padF = gv.padF; k.padB = gv.padB;
    k.off_map[0] = gv.off0; k.off_map[1] = gv.off1; k.off_map[2] = gv.off2; k.off_map[3] = gv.off3;
    k.episode_limit = cfg->episode_limit; k.fire_cost = cfg->fire_cost; k.hit_penalty = cfg->hit_penalty;
    k.beam_len = cfg->beam_len; k.n_actions = cfg->kind == SSD_KIND_CLEANUP ? 9 : 8;
    k.n_apple = na; k.n_waste = nw; k.n_spawn = ns; k.n_apple4 = (na + 3) / 4; k.n_waste2 = (nw + 1) / 2;
    k.base_apples = 0; k.base_waste = 0;
    for (int c = 0; c < G; ++c) { k.base_apples += hm->base_grid[c] == SSD_CELL_APPLE; k.base_waste += hm->base_grid[c] == SSD_CELL_WASTE; }
    k.random_spawn = cfg->random_spawn_point != 0; k.spawn_rot = cfg->spawn_rotation < 0 ? -1 : cfg->spawn_rotation;
    k.invN20 = magic20(k.N, SSD_MAX_AGENTS * k.N + 64); k.invW20 = magic20(W, G + 1);
    k.seed_lo = (uint32_t)cfg->seed; k.seed_hi = (uint32_t)(cfg->seed >> 32); k.gid_base = cfg->env_gid_base;
    for (int i = 0; i < 4; ++i) k.thr_harvest[i] = cfg->thr_harvest[i];
    for (int i = 0; i < 16; ++i)
        k.lut[i] = (uint32_t)cfg->color_lut[i][0] | ((uint32_t)cfg->color_lut[i][1] << 8) | ((uint32_t)cfg->color_lut[i][2] << 16);
    k.agents_uniform = 1;
    for (int i = 8; i < 16; ++i) if (k.lut[i] != k.lut[7]) k.agents_uniform = 0;
    for (int ch = 0; ch < 3; ++ch) {                          // 8-entry per-plane LUT: indices 0-6 + 7 = agent
        uint32_t lo = 0, hi = 0;
        for (int i = 0; i < 4; ++i) { lo |= (uint32_t)cfg->color_lut[i][ch] << (8 * i); hi |= (uint32_t)cfg->color_lut[4 + i][ch] << (8 * i); }
        k.lut8[2 * ch] = lo; k.lut8[2 * ch + 1] = hi;
    }
    if (k.invN20 == 0 || k.invW20 == 0 || k.n_apple4 > 32 * 16) { delete hm; delete h; return SSD_ERR_INVALID; }

    k.invMW20 = magic20(k.nw8M, H * k.nw8M + 32); k.invMTW20 = magic20(k.nw8T, W * k.nw8T + 32);
    if (k.invMW20 == 0 || k.invMTW20 == 0 || k.PMS >= 60000) { delete hm; delete h; return SSD_ERR_INVALID; }
    k.off_pmap = k.padF + k.GS + k.padB;
    k.smem_per_warp = k.off_pmap + k.PMS;
    // two envs per warp need the spawn points (reset) and the per-lane apple decision bits (64) to fit 16 lanes
    h->lanes_per_env = 32;
#ifdef SSD_ENABLE_LPE16
    if (ns <= 16 && k.n_apple4 <= 16 * 16) h->lanes_per_env = 16;
    { const char* e = getenv("SSD_B200_LPE"); if (e && atoi(e) == 32) h->lanes_per_env = 32; }
#endif
    h->smem_bytes = (size_t)kWarps * (32 / h->lanes_per_env) * k.smem_per_warp;
    { const char* e = getenv("SSD_B200_GENERIC"); h->force_generic = e && e[0] == '1'; }
    { const char* e = getenv("SSD_B200_PDL"); h->pdl = !(e && e[0] == '0'); }
    h->device = cfg->device;

    DeviceGuard guard(cfg->device);
    cudaError_t e = guard.err;
    if (e == cudaSuccess) {                                   // the launch needs smem_bytes dynamic + the static LUT
        int optin = 0;
        e = cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, cfg->device);
        if (e == cudaSuccess && h->smem_bytes + 256 > (size_t)optin) { delete hm; delete h; return SSD_ERR_INVALID; }
    }
    if (e == cudaSuccess) e = cudaMalloc(&h->d_map, sizeof(MapDev));
    if (e == cudaSuccess) e = cudaMemcpy(h->d_map, hm, sizeof(MapDev), cudaMemcpyHostToDevice);
    delete hm;
    if (e != cudaSuccess) {
        if (h->d_map) cudaFree(h->d_map);
        delete h;
        return cuda_fail(e);
    }
    k.map = h->d_map;
    *out = h;
    return SSD_OK;
}

int ssd_destroy(ssd_handle* h) {
    if (!h) return SSD_ERR_INVALID;
    DeviceGuard guard(h->device);
    cudaFree(h->d_map);
    delete h;
    return SSD_OK;
}

int ssd_get_layout(const ssd_handle* h, ssd_layout* o) {
    if (!h || !o) return SSD_ERR_INVALID;
    const KParams& k = h->kp;
    o->n_actions = k.n_actions; o->n_cells = k.G; o->obs_n = k.N; o->grid_stride = k.GS; o->agent_stride = k.NA;
    o->obs_plane_stride = k.PS; o->obs_agent_stride = k.AS; o->obs_env_stride = k.ES;
    o->n_apple_pts = k.n_apple; o->n_waste_pts = k.n_waste; o->n_spawn_pts = k.n_spawn; o->obs_row_stride = k.RP;
    return SSD_OK;
}

int ssd_reset(ssd_handle* h, const ssd_state* st, const uint8_t* mask, const ssd_draws* draws, uint8_t* obs, void* stream) {
    KParams k;
    const int rc = fill_common(h, st, draws, k);
    if (rc) return rc;
    k.mask = mask; k.obs = obs;
    return launch<MODE_RESET>(h, k, stream);
}

int ssd_step(ssd_handle* h, const ssd_state* st, const uint8_t* actions, const ssd_draws* draws,
             const ssd_step_out* out, void* stream) {
    KParams k;
    const int rc = fill_common(h, st, draws, k);
    if (rc) return rc;
    if (!actions || !out || !out->reward || !out->clean || !out->apple_cnt || !out->done) return SSD_ERR_INVALID;
    k.actions = actions; k.reward = out->reward; k.clean = out->clean; k.apple_cnt = out->apple_cnt; k.done = out->done;
    k.obs = out->obs; k.state_rgb = out->state_rgb;
    return launch<MODE_STEP>(h, k, stream);
}

int ssd_step_range(ssd_handle* h, const ssd_state* st, const uint8_t* actions, const ssd_draws* draws,
                   const ssd_step_out* out, int32_t env_begin, int32_t env_count, void* stream) {
    KParams k;
    const int rc = fill_common(h, st, draws, k);
    if (rc) return rc;
    if (!actions || !out || !out->reward || !out->clean || !out->apple_cnt || !out->done) return SSD_ERR_INVALID;
    if (env_begin < 0 || env_count < 0 || (int64_t)env_begin + env_count > k.B) return SSD_ERR_INVALID;
    k.env0 = env_begin; k.env1 = env_begin + env_count;
    k.actions = actions; k.reward = out->reward; k.clean = out->clean; k.apple_cnt = out->apple_cnt; k.done = out->done;
    k.obs = out->obs; k.state_rgb = out->state_rgb;
    return launch<MODE_STEP>(h, k, stream);
}

int ssd_render(ssd_handle* h, const ssd_state* st, uint8_t* obs, uint8_t* state_rgb, void* stream) {
    KParams k;
    const int rc = fill_common(h, st, nullptr, k);
    if (rc) return rc;
    if (!obs && !state_rgb) return SSD_ERR_INVALID;
    k.obs = obs; k.state_rgb = state_rgb;
    return launch<MODE_RENDER>(h, k, stream);
}

int ssd_step_host(ssd_handle* h, const ssd_state* st, const uint8_t* h_actions, uint8_t* d_actions,
                  const ssd_step_out* d_out, const ssd_step_out* h_out, void* stream) {
    if (!h || !h_actions || !d_actions || !d_out || !h_out) return SSD_ERR_INVALID;
    DeviceGuard guard(h->device);
    SSD_CUDA(guard.err);
    const KParams& k = h->kp;
    cudaStream_t s = (cudaStream_t)stream;
    const size_t B = (size_t)k.B;
    SSD_CUDA(cudaMemcpyAsync(d_actions, h_actions, B * k.n, cudaMemcpyHostToDevice, s));
    const int rc = ssd_step(h, st, d_actions, nullptr, d_out, stream);
    if (rc) return rc;
    if (h_out->reward) SSD_CUDA(cudaMemcpyAsync(h_out->reward, d_out->reward, B * k.n, cudaMemcpyDeviceToHost, s));
    if (h_out->clean) SSD_CUDA(cudaMemcpyAsync(h_out->clean, d_out->clean, B * k.n, cudaMemcpyDeviceToHost, s));
    if (h_out->apple_cnt) SSD_CUDA(cudaMemcpyAsync(h_out->apple_cnt, d_out->apple_cnt, B * 2, cudaMemcpyDeviceToHost, s));
    if (h_out->done) SSD_CUDA(cudaMemcpyAsync(h_out->done, d_out->done, B, cudaMemcpyDeviceToHost, s));
    if (h_out->obs && d_out->obs) SSD_CUDA(cudaMemcpyAsync(h_out->obs, d_out->obs, B * k.ES, cudaMemcpyDeviceToHost, s));
    if (h_out->state_rgb && d_out->state_rgb)
        SSD_CUDA(cudaMemcpyAsync(h_out->state_rgb, d_out->state_rgb, B * 3 * k.G, cudaMemcpyDeviceToHost, s));
    SSD_CUDA(cudaStreamSynchronize(s));
    return SSD_OK;
}

int ssd_incentive(const int64_t* actions_inc, const float* reward, int64_t rows, int32_t n_agents,
                  float incentive, float cost, float ratio, int32_t max_seq_length,
                  float* rewards_for_env, float* rewards_for_inc, float* recv_sign, void* stream) {
    if (!actions_inc || !reward || !rewards_for_env || !rewards_for_inc || rows < 0 || n_agents < 1 || max_seq_length == 0)
        return SSD_ERR_INVALID;
    if (rows == 0) return SSD_OK;
    // max_seq_length > 0: true division (torch CPU); < 0: multiply by 1/|T| (torch CUDA scalar-divisor path)
    const int recip = max_seq_length < 0;
    const float T = (float)abs(max_seq_length);
    const long long total = (long long)rows * n_agents;
    const int threads = 256;
    incentive_kernel<<<(unsigned)((total + threads - 1) / threads), threads, 0, (cudaStream_t)stream>>>(
        reinterpret_cast<const long long*>(actions_inc), reward, rows, n_agents, incentive, cost, ratio, T, recip,
        rewards_for_env, rewards_for_inc, recv_sign);
    SSD_CUDA(cudaGetLastError());
    return SSD_OK;
}

int64_t ssd_launch_count(const ssd_handle* h) { return h ? h->launches : -1; }

int64_t ssd_debug_oob_count(void) {
#ifdef SSD_BOUNDS_CHECK
    unsigned long long v = 0;
    if (cudaDeviceSynchronize() != cudaSuccess || cudaMemcpyFromSymbol(&v, g_oob_count, sizeof(v)) != cudaSuccess) return -2;
    return (int64_t)v;
#else
    return -1;
#endif
}

}  // extern "C"
