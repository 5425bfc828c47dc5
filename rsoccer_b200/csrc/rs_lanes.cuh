// rs_lanes.cuh -- ONE LANE PER BODY: the scene of a match is spread over L = 2^k >= R + 1
// consecutive lanes of a warp (lane b of the group: b = 0 ball, 1..R robots, the rest idle),
// a warp advances 32 / L matches (sm_100a).
//
//   world                      R    L   matches / warp
//   SSL 1 v 1                  2    4        8
//   VSS 3 v 3                  6    8        4
//   SSL 1 v 6                  7    8        4
//   VSS 5 v 5                 10   16        2
//   SSL 11 v 11               22   32        1      (one warp per match)
//
// Why (profiles/r1_steady_final.txt): with one lane per MATCH the
// benchmark size (65 536 matches) is only 2 048 warps of 128 registers -- 3.5 warps per SM
// sub-partition, issue slots 48 % used, fixed-latency and instruction-fetch stalls on top,
// and 4 096-match worlds leave most SMs empty.  One lane per BODY gives L x the warps at
// ~40 registers (>= 10 resident warps per scheduler), a sub-step body of a few hundred
// SASS instructions that lives in the L0 instruction cache, and no per-body unrolling.
//
// Per sub-step every lane drives / integrates / wall-tests its own body; the all-pairs scan
// is a ring of warp shuffles (offsets d = 1 .. L/2: each unordered pair is tested exactly
// once, by the lane of its lower ring end); contacts are rare (~7 % of the matches per
// sub-step), so the warp leaves the fast path only when a ballot finds one and then
// resolves the contacts of a match sequentially in lexicographic pair order (Gauss-Seidel,
// same order and arithmetic as the oracle), all lanes of the group
// computing the impulse and the two owners applying it.
//
// Replaces the arithmetic inside `robosim.VSS.step` / `robosim.SSL.step`
// (rsoccer_gym/Simulators/rsim.py:102, :155).
#pragma once
#include "rs_tasks.cuh"

#define RS_FULL_MASK 0xffffffffu

struct LaneBody { float x, y, vx, vy, th, om; };
struct LaneDrive { float tf, tl, tw, kick; bool drib; };

template <int L>
struct LaneGroup {
    static_assert(L == 2 || L == 4 || L == 8 || L == 16 || L == 32, "lanes per match");
    static constexpr int MPW = 32 / L;                   // matches per warp
    static constexpr int ND = L / 2;                     // ring offsets of the pair scan
    static constexpr uint32_t BITS = L == 32 ? 0xffffffffu : ((1u << (L & 31)) - 1u);
    int base;                                            // first lane of this group in the warp
    uint32_t mask;                                       // member mask of this group
    __device__ __forceinline__ LaneGroup() {
        base = (threadIdx.x & 31) & ~(L - 1);
        mask = BITS << base;
    }
    // value held by body j of this group; warp-uniform control flow
    __device__ __forceinline__ float get(float v, int j) const { return __shfl_sync(RS_FULL_MASK, v, base + j); }
    // same, inside control flow that is uniform only within the group
    __device__ __forceinline__ float gget(float v, int j) const { return __shfl_sync(mask, v, base + j); }
    __device__ __forceinline__ uint32_t gget(uint32_t v, int j) const { return __shfl_sync(mask, v, base + j); }
    // bits of this group in a warp ballot
    __device__ __forceinline__ uint32_t bits(uint32_t ballot) const { return (ballot >> base) & BITS; }
};

__device__ __forceinline__ float dist2(float dx, float dy) { return dx * dx + dy * dy; }

// One contact (i < j, body indices) of this group; every lane of the group computes it from
// shuffled copies, lanes i and j keep their side.  Positions are those of the phase start
// (corrections accumulate in cx, cy); velocities are live (sequential impulses).
template <int KIND, int L, class PP>
__device__ __forceinline__ void lanes_resolve(const PP &P, const LaneGroup<L> &g, const int b,
                                              const int i, const int j, LaneBody &s, float &cx, float &cy) {
    const float xi = g.gget(s.x, i), yi = g.gget(s.y, i), xj = g.gget(s.x, j), yj = g.gget(s.y, j);
    float vxi = g.gget(s.vx, i), vyi = g.gget(s.vy, i), vxj = g.gget(s.vx, j), vyj = g.gget(s.vy, j);
    float cxi = 0.0f, cyi = 0.0f, cxj = 0.0f, cyj = 0.0f;
    bool any = false;
    if (i == 0) {
        const float omj = g.gget(s.om, j);
        const float thj = KIND == RS_KIND_SSL ? g.gget(s.th, j) : 0.0f;
        ball_robot<KIND>(P, xi, yi, vxi, vyi, xj, yj, thj, vxj, vyj, omj, cxi, cyi, cxj, cyj, any);
    } else {
        robot_robot(P, xi, yi, vxi, vyi, xj, yj, vxj, vyj, cxi, cyi, cxj, cyj, any);
    }
    if (b == i) { s.vx = vxi; s.vy = vyi; cx += cxi; cy += cyi; }
    if (b == j) { s.vx = vxj; s.vy = vyj; cx += cxj; cy += cyj; }
}

// one control step (RS_SUBSTEPS sub-steps) of the match this lane's group holds.
//   b: body index of this lane (0 ball, 1..R robots, > R idle); idle lanes carry zeros.
// PP = DevParams (run-time constants) or VssF0 (the benchmark world: immediates, VSS only)
template <int KIND, int L, class PP>
__device__ __forceinline__ void lanes_physics_step(const PP &P, const int R, const int b,
                                                   LaneBody &s, const LaneDrive &d) {
    constexpr int ND = LaneGroup<L>::ND;
    const LaneGroup<L> g;
    const bool is_ball = b == 0, is_robot = b >= 1 && b <= R;
    const float h = P.h;
    const float rad = is_ball ? P.ball_r : P.rbt_r, ew = is_ball ? P.e_ball_wall : P.e_rbt_wall;
    // ring scan: offset k pairs body b with body (b + k) mod L; offset L/2 would see every
    // pair twice, the upper half of the ring skips it.  thr = squared contact distance of
    // the pair, negative when the pair does not exist.
    float thr[ND];
#pragma unroll
    for (int k = 1; k <= ND; ++k) {
        const int p = (b + k) & (L - 1);
        const bool exists = b <= R && p <= R && (k < ND || b < ND);
        thr[k - 1] = !exists ? -1.0f : ((b == 0 || p == 0) ? P.rs_br2 : P.rs_rr2);
    }

    bool kicked = false;
    if constexpr (KIND == RS_KIND_SSL) {
        // kick: once per control step; robots in row order overwrite the ball velocity, so
        // the highest kicking row that touches the ball wins
        const bool wants = is_robot && d.kick > 0.0f;
        if (__any_sync(RS_FULL_MASK, wants)) {
            const float bx = g.get(s.x, 0), by = g.get(s.y, 0);
            float sn, cs;
            __sincosf(s.th, &sn, &cs);
            kicked = wants && touching(P, s.x, s.y, cs, sn, bx, by);
            const uint32_t km = g.bits(__ballot_sync(RS_FULL_MASK, kicked));
            const int jk = km ? 31 - __clz((int)km) : 0;
            const float kvx = g.get(cs * d.kick, jk), kvy = g.get(sn * d.kick, jk);
            if (km && is_ball) { s.vx = kvx; s.vy = kvy; }
        }
    }

#pragma unroll 1
    for (int sub = 0; sub < RS_SUBSTEPS; ++sub) {
        float sn, cs;
        __sincosf(s.th, &sn, &cs);
        // (a) drive: velocity in the robot frame moves toward the target, traction limited
        float vf = cs * s.vx + sn * s.vy, vl = -sn * s.vx + cs * s.vy;
        if (KIND == RS_KIND_VSS) {
            // v + clamp(t - v, -a, a) == clamp(t, v - a, v + a): two independent adds, then min / max (a shorter chain)
            vf = fmaxf(fminf(d.tf, vf + P.acc_fwd_h), vf - P.acc_fwd_h);
            vl = fmaxf(fminf(d.tl, vl + P.acc_lat_h), vl - P.acc_lat_h);
        } else {
            const float df = d.tf - vf, dl = d.tl - vl;
            const float n2 = df * df + dl * dl;
            float sc = 1.0f;
            if (n2 > P.acc_fwd_h * P.acc_fwd_h) sc = P.acc_fwd_h * rsqrtf(n2);
            vf += df * sc; vl += dl * sc;
        }
        const float om_n = fmaxf(fminf(d.tw, s.om + P.acc_ang_h), s.om - P.acc_ang_h);
        const float dvx = cs * vf - sn * vl, dvy = sn * vf + cs * vl;
        // (b) dribbler latch: first robot in row order with the dribbler on, not kicking,
        // and the ball in its kicker box holds the ball for this sub-step
        int holder = -1; float hx = 0.0f, hy = 0.0f;
        if constexpr (KIND == RS_KIND_SSL) {
            const bool cand = is_robot && d.drib && !kicked;
            if (__any_sync(RS_FULL_MASK, cand)) {
                const float dx = g.get(s.x, 0) - s.x, dy = g.get(s.y, 0) - s.y;
                const float lx = cs * dx + sn * dy, ly = -sn * dx + cs * dy;
                const uint32_t hm = g.bits(__ballot_sync(RS_FULL_MASK, cand && touching_local(P, lx, ly)));
                const int hj = hm ? __ffs((int)hm) - 1 : 0;
                hx = g.get(lx, hj); hy = g.get(ly, hj);
                if (hm) holder = hj;
            }
        }
        // (c) ball rolling friction (Coulomb, exact stop) / robots take the driven velocity
        {
            const float sp2 = s.vx * s.vx + s.vy * s.vy;
            const float fr = fmaxf(1.0f - P.ball_decel_h * rsqrtf(sp2 + 1e-12f), 0.0f);
            const float sc = holder < 0 ? fr : 1.0f;
            s.vx = is_ball ? s.vx * sc : dvx; s.vy = is_ball ? s.vy * sc : dvy;
            s.om = is_ball ? 0.0f : om_n;
        }
        // (d) integrate (semi-implicit Euler); a held ball is slaved to its holder's new pose
        {
            const float hh = (is_ball && holder >= 0) ? 0.0f : h;
            s.x += s.vx * hh; s.y += s.vy * hh;
            s.th = wrap_pi(s.th + s.om * h);
        }
        if constexpr (KIND == RS_KIND_SSL) {
            if (__any_sync(RS_FULL_MASK, holder >= 0)) {
                float s2, c2;
                __sincosf(s.th, &s2, &c2);
                const float ox = c2 * hx - s2 * hy, oy = s2 * hx + c2 * hy;
                const int hj = holder >= 0 ? holder : 0;
                const float nbx = g.get(s.x + ox, hj), nby = g.get(s.y + oy, hj);
                const float nvx = g.get(s.vx - s.om * oy, hj), nvy = g.get(s.vy + s.om * ox, hj);
                if (is_ball && holder >= 0) { s.x = nbx; s.y = nby; s.vx = nvx; s.vy = nvy; }
            }
        }
        // (e) pairs: ring scan on the positions at phase start
        uint32_t hits = 0;
#pragma unroll
        for (int k = 1; k <= ND; ++k) {
            const float px = __shfl_sync(RS_FULL_MASK, s.x, b + k, L), py = __shfl_sync(RS_FULL_MASK, s.y, b + k, L);
            if (dist2(px - s.x, py - s.y) < thr[k - 1]) hits |= 1u << k;
        }
        if (__any_sync(RS_FULL_MASK, hits != 0u)) {
            // row of the upper-triangular contact matrix owned by this body: partners j > b.
            // A pair found at offset k by lane q belongs to row min(q, (q + k) mod L).
            uint32_t up = 0;
#pragma unroll
            for (int k = 1; k <= ND; ++k) {
                const int p = (b + k) & (L - 1), q = (b - k) & (L - 1);
                const uint32_t qh = __shfl_sync(RS_FULL_MASK, hits, b + L - k, L);
                if (((hits >> k) & 1u) && p > b) up |= 1u << p;
                if (((qh >> k) & 1u) && q > b) up |= 1u << q;
            }
            uint32_t rows = g.bits(__ballot_sync(RS_FULL_MASK, up != 0u));
            if (rows) {                                   // uniform within the group from here
                float cx = 0.0f, cy = 0.0f;
                while (rows) {
                    const int i = __ffs((int)rows) - 1;
                    rows &= rows - 1;
                    uint32_t row = g.gget(up, i);
                    while (row) {
                        const int j = __ffs((int)row) - 1;
                        row &= row - 1;
                        lanes_resolve<KIND, L>(P, g, b, i, j, s, cx, cy);
                    }
                }
                s.x += cx; s.y += cy;
            }
            __syncwarp();
        }
        // (f) walls (the lean forms of the lane-per-match kernels, same decisions and arithmetic)
        if constexpr (KIND == RS_KIND_VSS) vss_walls(P, rad, ew, s.x, s.y, s.vx, s.vy);
        else ssl_walls(P, rad, ew, s.x, s.y, s.vx, s.vy);
    }
}

// ---------------------------------------------------------------- lane <-> HBM
// body / ang rows of this lane's body for match e
template <int L>
__device__ __forceinline__ void lanes_load(const StatePtrs &S, const int R, const int b, const int e, LaneBody &s) {
    s.x = s.y = s.vx = s.vy = s.th = s.om = 0.0f;
    if (b <= R) {
        const float4 q = __ldcg(S.body + (size_t)b * S.np + e);      // around L1, as in load_scene (rs_device.cuh)
        s.x = q.x; s.y = q.y; s.vx = q.z; s.vy = q.w;
    }
    if (b >= 1 && b <= R) {
        const float2 a = __ldcg(S.ang + (size_t)(b - 1) * S.np + e);
        s.th = a.x; s.om = a.y;
    }
}
template <int L>
__device__ __forceinline__ void lanes_store(const StatePtrs &S, const int R, const int b, const int e, const LaneBody &s) {
    if (b <= R) S.body[(size_t)b * S.np + e] = make_float4(s.x, s.y, s.vx, s.vy);
    if (b >= 1 && b <= R) S.ang[(size_t)(b - 1) * S.np + e] = make_float2(s.th, s.om);
}

// The per-match task scalars are rows of ONE array of 32-bit words with pitch Np:
// word 0 = previous ball potential, word 1 = step counter | has_prev << 24, words 2.. =
// reward_shaping_total accumulators (RS_ARR_PREV, RS_ARR_STEPS, RS_ARR_INFO are laid out
// back to back by rs_create).  Lane b of a group owns words b, b + L, b + 2L, ...
#define RS_AUX_PREV 0
#define RS_AUX_STEPS 1
#define RS_AUX_INFO 2


// ---------------------------------------------------------------- warp-cooperative auto-reset
// `reset` is uniform within a group of L lanes.  An ending match used to make all L lanes of its
// group run the scalar placement (dependent Philox calls, rejection sampling on a scene in local
// memory): 1.9 of the 10.1 us of SSLStaticDefenders-v0 at 4 096 matches, where ~15 matches end
// per step.  Now, per ending match: the whole warp draws the first 128 words of its placement
// stream (one Philox block per lane), the group leader places from shared memory with the placed
// robots in statically indexed registers and publishes (ball x y, robots x y theta), the group
// picks its bodies up.  buf: 128 + MPW x (2 + 3 R) words of shared memory per warp.
template <int TASK, int R, int L>
__device__ __forceinline__ void lanes_reset_place(const DevParams &P, uint32_t *buf, const bool reset, const bool valid,
                                                  const int b, const bool is_robot, const uint32_t gid,
                                                  const uint64_t seed, const uint32_t t_now, LaneBody &s) {
    constexpr int RES = 2 + 3 * R;
    const int lane = threadIdx.x & 31;
    unsigned need = __ballot_sync(0xffffffffu, reset && valid && b == 0);
    if (!need) return;                                  // warp-uniform
    float *const res = reinterpret_cast<float *>(buf + 128);
    const uint2 key = make_uint2((uint32_t)seed, (uint32_t)(seed >> 32));
    while (need) {
        const int src = __ffs((int)need) - 1;
        need &= need - 1;
        const uint32_t env_src = __shfl_sync(0xffffffffu, gid, src);
        warp_placement_words(buf, 0xffffffffu, seed, env_src, t_now);
        if (lane == src) {
            Scene<R> tmp;
            const PlaceStream g{buf, 32, key, env_src, t_now, 0};
            place_from_stream<TASK, R>(P, g, tmp);
            float *o = res + (src / L) * RES;
            o[0] = tmp.bx; o[1] = tmp.by;
#pragma unroll
            for (int r = 0; r < R; ++r) { o[2 + 3 * r] = tmp.x[r]; o[3 + 3 * r] = tmp.y[r]; o[4 + 3 * r] = tmp.th[r]; }
        }
    }
    __syncwarp();
    if (reset && valid) {
        const float *o = res + (lane / L) * RES;
        if (b == 0) { s.x = o[0]; s.y = o[1]; }
        else if (is_robot) { s.x = o[2 + 3 * (b - 1)]; s.y = o[3 + 3 * (b - 1)]; s.th = o[4 + 3 * (b - 1)]; }
    }
}
