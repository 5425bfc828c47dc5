// rs_capi.cu -- kernels + the C ABI declared in include/rsoccer_b200.h (sm_100a only).
//
// Build (see __graft_entry__.build):
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -shared -Xcompiler -fPIC
//        -o rsoccer_b200/librsoccer_b200.so rsoccer_b200/csrc/rs_capi.cu
#include <atomic>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>

#include "../../include/rsoccer_b200.h"
#include "rs_tasks.cuh"
#include "rs_lanes.cuh"

// ============================================================================ kernels

__global__ void k_set_ctr(uint32_t *ctr, int n_ctr, uint32_t t) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n_ctr) ctr[i] = t;
}

// resident CTAs per SM the headline kernel is compiled for: 448 threads = 14 warps per SM is what
// 65 536 matches need (6.9 CTAs of 64), and it leaves ptxas 144 registers per thread
#ifndef RS_VSS_THREADS
#define RS_VSS_THREADS 448
#endif
#define RS_VSS_MINB(BS) ((RS_VSS_THREADS / (BS)) > 0 ? (RS_VSS_THREADS / (BS)) : 1)
#ifndef RS_VSS_DENSE_THREADS
#define RS_VSS_DENSE_THREADS 704
#endif

struct VssStepArgs {
    const float2 *actions;   // [N]
    const float *normals;    // [N][2(R-1)] or null
    float *obs;              // [N][NOBS]
    float *reward;           // [N]
    uint8_t *done, *trunc;   // [N]
    float *cmds_out;         // [N][R][2] or null
    int auto_reset, max_steps;
    uint64_t seed;
    uint32_t *ctr;           // world step counter t (Philox counter word 1), one copy per RS_CTR_GROUP matches
    uint32_t env_offset;
    // step-to-step overlap (rs_device.cuh, tile_acquire): one flag word per 32-match tile + one error
    // word behind them, or null.  chain = 0: grid-wide wait first (griddepcontrol.wait), 1: per-tile wait
    // for the state, grid-wide wait before the first read of a caller buffer (actions, normals),
    // 2 / 3: per-tile wait only (the caller vouches for its buffers, RS_OPT_STEP_OVERLAP = 2 / 3)
    uint32_t *flags;
    int chain;
};

// VSSEnv.step for BS matches per CTA, one lane per match.  ONE launch = commands (agent +
// OU noise), 5 physics sub-steps, reward/done/truncation, info accumulators, masked
// auto-reset and the observation tile (leaves through a TMA bulk store).
//
// DENSE: the same source compiled for 11 resident CTAs per SM (80 registers instead of 114, no spills, same
// instruction count).  When several worlds are stepped round-robin on one stream (RS_OPT_STEP_OVERLAP = 3)
// consecutive launches are independent and as many CTAs of the NEXT step as registers allow should be
// resident (65 536 matches, 8 worlds: 10.7 -> 9.2 us per step).  A launch behind a true dependency (every step
// waits for the previous grid, or one world stepped again and again) wants its 1 024 CTAs spread evenly, 7 per
// SM.  With DENSE and programmatic launches up to 5 CTAs per SM of the NEXT step become resident early and
// wait; the ~280 that did not fit then land, it seems, on whichever SMs finish first, 7 at a time, so that a
// few SMs run 12 CTAs of the step and the others 5 (15.8 -> 19.5 us serialised, 12.8 -> 16.6 us chained on one world; with
// RS_PDL=0 the dense build runs 15.6 us and alone under ncu all 148 SMs are busy, profiles/r2_logs/
// dense_pdl.log), so the caller chooses.
template <int NB, int NY, int BS, int F0 /* 0: run-time physics constants.  1: VssF0's, immediates instead of
          constant-bank loads.  2: VssF0P, the same with the packed fp32x2 instruction forms (rs_device.cuh) */,
          bool DENSE = false>
__global__ void __launch_bounds__(BS, DENSE ? RS_VSS_DENSE_THREADS / BS : RS_VSS_MINB(BS))
k_vss_env_step(const __grid_constant__ DevParams P, const StatePtrs S, const VssStepArgs A) {
    constexpr int R = NB + NY, NZ = 2 * (R - 1), NOBS = 4 + 7 * NB + 5 * NY;
    constexpr bool PK = F0 == 2 && VssF0P::packed;
    // one region per WARP: its contact scratch during the physics (2 x (R + 1) rows of 32 float4
    // columns), then its 32 observation rows.  Nothing in it is ever touched by another warp.
    constexpr int SCRATCH = 2 * (R + 1) * 32 * 4, WF = 32 * NOBS > SCRATCH ? 32 * NOBS : SCRATCH;
    __shared__ __align__(128) float smem[BS / 32][WF];
    float *const wtile = smem[threadIdx.x >> 5];
    float4 *const cq = reinterpret_cast<float4 *>(wtile) + (threadIdx.x & 31), *const cp0 = cq + (R + 1) * 32;
    const int tid = threadIdx.x;
    const int e0 = blockIdx.x * BS;
    const int e = e0 + tid;
    const int w0 = e0 + (tid & ~31);                       // first env of this warp
    const int wrows = min(32, S.n - w0);
    const unsigned live = __ballot_sync(0xffffffffu, e < S.n);
    const StepTile tile_lock = step_begin<BS>(A.ctr, A.flags, A.chain, w0);
    if (e < S.n) {
        // ---- every global load of the step is issued first: the step counter ahead of the state,
        // so that it is not queued behind 100 KB of requests per SM and Philox can start at once
        // (with step overlap it arrived with the tile lock)
        const uint32_t t_now = tile_lock.lock ? tile_lock.t : step_counter_read<RS_CTR_GROUP>(A.ctr, e, live);
        Scene<R> s;
        load_scene<R>(P, S, e, s);
        const int st = __ldcg(S.steps + e);
        float prev = __ldcg(S.prev + e);
        // reward_shaping_total is never loaded: the six accumulators are zeroed by a store at the
        // first step of an episode and updated by fire-and-forget reductions (RED.ADD.F32, one
        // add per word and step: same rounding as load-add-store) -- move / ball_grad / energy every
        // step, goal_score / goals_blue / goals_yellow only when a goal is scored.  36 of the 640
        // bytes a step used to move per env.
        float2 ou[R - 1];
#pragma unroll
        for (int r = 1; r < R; ++r) ou[r - 1] = __ldcg(S.ou + (size_t)(r - 1) * S.np + e);
        float2 act = make_float2(0.0f, 0.0f);
        if (A.chain != 1) act = __ldcg(A.actions + e);      // around L1, like the state: read once, and never stale

        // ---- ... and the OU noise (Philox + Box-Muller, ~15 % of the instructions, needs
        // only the env id and the step counter) is computed while they are in flight
        float z[NZ];
        if (A.normals) {
            if (A.chain == 1) pdl_wait();
#pragma unroll
            for (int k = 0; k < NZ; ++k) z[k] = __ldcg(A.normals + (size_t)e * NZ + k);
        } else {
            // Philox stream (global env id, t, OU): Box-Muller on consecutive u32 pairs
            const uint2 key = make_uint2((uint32_t)A.seed, (uint32_t)(A.seed >> 32));
#pragma unroll
            for (int j = 0; j < (NZ + 3) / 4; ++j) {
                const uint4 u = philox4x32_10(make_uint4(A.env_offset + (uint32_t)e, t_now, RS_STREAM_OU, j), key);
                float sn, cs;
                float rr = sqrtf(-2.0f * __logf(u01(u.x)));
                __sincosf(2.0f * RS_PI_F * (u01(u.y) - 0.5f), &sn, &cs);   // angle - pi: flip signs
                if (4 * j < NZ) z[4 * j] = -rr * cs;
                if (4 * j + 1 < NZ) z[4 * j + 1] = -rr * sn;
                rr = sqrtf(-2.0f * __logf(u01(u.z)));
                __sincosf(2.0f * RS_PI_F * (u01(u.w) - 0.5f), &sn, &cs);
                if (4 * j + 2 < NZ) z[4 * j + 2] = -rr * cs;
                if (4 * j + 3 < NZ) z[4 * j + 3] = -rr * sn;
            }
        }

        // caller buffers are read only after the predecessor grid has completed (chain 1: the state
        // loads and the noise above already ran under its tail)
        if (A.chain == 1) { pdl_wait(); act = __ldcg(A.actions + e); }

        // ---- _get_commands, vss_gym.py:119-142
        Drive<R> d;
        d.drib = 0;
        float wl0, wr0;
        vss_action_to_wheels<PK>(P, act.x, act.y, wl0, wr0);
        vss_target<PK>(P, wl0, wr0, d.tf[0], d.tw[0]);
        d.tl[0] = 0.0f; d.kick[0] = 0.0f;
        if (A.cmds_out) { A.cmds_out[(size_t)e * R * 2] = wl0; A.cmds_out[(size_t)e * R * 2 + 1] = wr0; }
#pragma unroll
        for (int r = 1; r < R; ++r) {
            // Utils/Utils.py:14-21 OU sample (mu = 0, sigma = 0.5, theta = 0.17)
            float2 o = ou[r - 1];
            // x + theta (0 - x) dt + sigma sqrt(dt) z for both wheels at once
            o = v_fma<PK>(o, bc2(-(float)RS_OU_THETA * P.dt), o);
            o = v_fma<PK>(make_float2(z[2 * (r - 1)], z[2 * (r - 1) + 1]), bc2((float)RS_OU_SIGMA * P.sqrt_dt), o);
            S.ou[(size_t)(r - 1) * S.np + e] = o;
            float wl, wr;
            vss_action_to_wheels<PK>(P, o.x, o.y, wl, wr);
            vss_target<PK>(P, wl, wr, d.tf[r], d.tw[r]);
            d.tl[r] = 0.0f; d.kick[r] = 0.0f;
            if (A.cmds_out) { A.cmds_out[((size_t)e * R + r) * 2] = wl; A.cmds_out[((size_t)e * R + r) * 2 + 1] = wr; }
        }

        // ---- rsim.send_commands + get_frame, vss_gym_base.py:77-82
        if constexpr (F0 == 2) physics_step<RS_KIND_VSS, R>(VssF0P{}, s, d, live, cq, cp0, 32);
        else if constexpr (F0 == 1) physics_step<RS_KIND_VSS, R>(VssF0{}, s, d, live, cq, cp0, 32);
        else physics_step<RS_KIND_VSS, R>(P, s, d, live, cq, cp0, 32);

        // ---- the task words (loaded at the top) are first needed here, after the physics
        int steps = st & 0xFFFFFF;
        bool has_prev = (st >> 24) & 1;
        if (steps == 0) {
#pragma unroll
            for (int i = 0; i < RS_VSS_INFO; ++i) S.info[(size_t)i * S.np + e] = 0.0f;
        }
        steps += 1;                                                 // vss_gym_base.py:73

        // ---- _calculate_reward_and_done, vss_gym.py:144-192
        float rew; bool goal = false;
        if (s.bx > P.half_len) {
            atomicAdd(&S.info[e], 1.0f); atomicAdd(&S.info[(size_t)4 * S.np + e], 1.0f); rew = 10.0f; goal = true;
        } else if (s.bx < -P.half_len) {
            atomicAdd(&S.info[e], -1.0f); atomicAdd(&S.info[(size_t)5 * S.np + e], 1.0f); rew = -10.0f; goal = true;
        } else {
            const float length_cm = 2.0f * P.half_len * 100.0f, hl = P.half_len + P.goal_depth;
            const float dx_d = (hl + s.bx) * 100.0f, dx_a = (hl - s.bx) * 100.0f, dy = s.by * 100.0f;
            const float pot = ((-sqrtf(dx_a * dx_a + 2.0f * dy * dy) + sqrtf(dx_d * dx_d + 2.0f * dy * dy)) / length_cm - 1.0f) * 0.5f;
            float grad = 0.0f;
            if (has_prev) grad = clampf((pot - prev) * 3.0f / P.dt, -5.0f, 5.0f);
            prev = pot; has_prev = true;
            const float rx = s.bx - s.x[0], ry = s.by - s.y[0];
            const float rinv = rsqrtf(rx * rx + ry * ry);
            const float move = clampf((rx * rinv * s.vx[0] + ry * rinv * s.vy[0]) * (1.0f / 0.4f), -5.0f, 5.0f);
            const float energy = -(fabsf(wl0) + fabsf(wr0));
            rew = 0.2f * move + 0.8f * grad + 2e-4f * energy;
            atomicAdd(&S.info[(size_t)1 * S.np + e], 0.2f * move);
            atomicAdd(&S.info[(size_t)2 * S.np + e], 0.8f * grad);
            atomicAdd(&S.info[(size_t)3 * S.np + e], 2e-4f * energy);
        }
        const bool tr = steps >= A.max_steps;                       // TimeLimit, __init__.py:4
        A.reward[e] = rew; A.done[e] = goal ? 1 : 0; A.trunc[e] = tr ? 1 : 0;

        if (A.auto_reset) {
            // rare (one match in ~6 000 per step) but on the critical path of the kernel: the
            // whole warp draws the first 4 x wrows words of the ending match's placement stream
            // at once, the ending lane then places from shared memory
            unsigned need = __ballot_sync(live, goal || tr);
            while (need) {
                const int src = __ffs((int)need) - 1;
                need &= need - 1;
                const uint2 key = make_uint2((uint32_t)A.seed, (uint32_t)(A.seed >> 32));
                const uint32_t env_src = A.env_offset + (uint32_t)(w0 + src);
                warp_placement_words(reinterpret_cast<uint32_t *>(wtile), live, A.seed, env_src, t_now);
                if ((tid & 31) == src) {
                    PlaceStream g{reinterpret_cast<const uint32_t *>(wtile), wrows, key, env_src, t_now, 0};
                    vss_place_stream<R>(P, g, s);
#pragma unroll
                    for (int r = 1; r < R; ++r) S.ou[(size_t)(r - 1) * S.np + e] = make_float2(0.0f, 0.0f);
                    steps = 0; has_prev = false; prev = 0.0f;
                }
            }
        }
        store_scene<R>(P, S, e, s);
        S.steps[e] = steps | ((has_prev ? 1 : 0) << 24);
        S.prev[e] = prev;
        __syncwarp(live);      // the rows below overlay the other lanes' contact scratch
        vss_obs<NB, NY, PK>(P, s, wtile + (tid & 31) * NOBS);
        step_counter_bump<RS_CTR_GROUP>(A.ctr, e, t_now, tile_lock.lock != nullptr);
    }
    // the rows of a warp are one contiguous span of global memory: every lane of the warp (live
    // or not) ships 16-byte pieces of it, coalesced -- half the L2 write sectors of row-per-lane
    // stores, and unlike a TMA bulk store nothing to wait for before the warp exits
    __syncwarp();
    {
        const float4 *src = reinterpret_cast<const float4 *>(wtile);
        float4 *dst = reinterpret_cast<float4 *>(A.obs + (size_t)w0 * NOBS);
        const int total = wrows * (NOBS / 4);
#pragma unroll
        for (int i = 0; i < NOBS / 4; ++i) { const int k = i * 32 + (tid & 31); if (k < total) dst[k] = src[k]; }
    }
    step_end(tile_lock);
}


// VSSEnv.step, one lane per BODY (rs_lanes.cuh): 8 lanes per 3 v 3 match, BS / 8 matches per
// CTA.  Same launch contract as k_vss_env_step.  The eight task words of a match (previous
// potential, step counter, six reward_shaping_total accumulators) are one word per lane:
// one load and one store instruction for all of them.
template <int BS, bool F0 = false /* physics constants are VssF0's immediates */>
__global__ void __launch_bounds__(BS)
k_vss_env_step_lanes(const __grid_constant__ DevParams P, const StatePtrs S, const VssStepArgs A) {
    constexpr int L = 8, NB = 3, NY = 3, R = NB + NY, NZ = 2 * (R - 1), NOBS = 4 + 7 * NB + 5 * NY;
    constexpr int EPB = BS / L, MPW = 32 / L;
    static_assert(RS_AUX_INFO + RS_VSS_INFO == L, "one task word per lane");
    __shared__ __align__(128) float tile[EPB * NOBS];
    __shared__ __align__(16) uint32_t pbuf[BS / 32][128 + MPW * (2 + 3 * R)];      // auto-reset placement (rs_lanes.cuh)
    const int tid = threadIdx.x;
    const int b = tid & (L - 1);                               // body of this lane
    const int el = tid / L;                                    // match within the CTA
    const int e = blockIdx.x * EPB + el;
    const bool valid = e < S.n;
    const int ec = valid ? e : S.n - 1;                        // dead groups shadow the last match (no stores)
    const int w0 = blockIdx.x * EPB + (tid >> 5) * MPW;        // first match of this warp
    const int wrows = min(MPW, S.n - w0);
    const LaneGroup<L> g;
    const bool is_robot = b >= 1 && b <= R, is_ou = b >= 2 && b <= R;
    const int p = is_ou ? b - 2 : 0;                           // OU process of this lane (blue 0 is the agent)
    const uint32_t gid = A.env_offset + (uint32_t)ec;

    const StepTile tile_lock = step_begin<BS>(A.ctr, A.flags, A.chain, w0);
    // ---- loads first: own body, own task word, own action source
    LaneBody s;
    lanes_load<L>(S, R, b, ec, s);
    const uint32_t aux = __ldcg(S.aux + (size_t)b * S.np + ec);
    float2 a = make_float2(0.0f, 0.0f);
    if (b == 1) a = __ldcg(A.actions + ec);
    if (is_ou) a = __ldcg(S.ou + (size_t)p * S.np + ec);
    const uint32_t t_now = tile_lock.lock ? tile_lock.t : step_counter_read<RS_CTR_GROUP * L>(A.ctr, ec);

    // ---- OU noise under the load latency: normals (2p, 2p + 1) are Box-Muller of the
    // u32 pair (p & 1) of Philox call p / 2 of the (global env id, t, OU) stream
    float z0 = 0.0f, z1 = 0.0f;
    if (A.normals) {
        if (is_ou) { z0 = __ldcg(A.normals + (size_t)ec * NZ + 2 * p); z1 = __ldcg(A.normals + (size_t)ec * NZ + 2 * p + 1); }
    } else {
        const uint2 key = make_uint2((uint32_t)A.seed, (uint32_t)(A.seed >> 32));
        const uint4 u = philox4x32_10(make_uint4(gid, t_now, RS_STREAM_OU, (uint32_t)(p >> 1)), key);
        const uint32_t ua = (p & 1) ? u.z : u.x, ub = (p & 1) ? u.w : u.y;
        float sn, cs;
        const float rr = sqrtf(-2.0f * __logf(u01(ua)));
        __sincosf(2.0f * RS_PI_F * (u01(ub) - 0.5f), &sn, &cs);   // angle - pi: flip signs
        z0 = -rr * cs; z1 = -rr * sn;
    }

    // ---- _get_commands, vss_gym.py:119-142 (Utils/Utils.py:14-21 OU sample)
    if (is_ou) {
        a.x = a.x + (float)RS_OU_THETA * (0.0f - a.x) * P.dt + (float)RS_OU_SIGMA * P.sqrt_dt * z0;
        a.y = a.y + (float)RS_OU_THETA * (0.0f - a.y) * P.dt + (float)RS_OU_SIGMA * P.sqrt_dt * z1;
    }
    float wl, wr;
    vss_action_to_wheels(P, a.x, a.y, wl, wr);
    LaneDrive d;
    vss_target(P, wl, wr, d.tf, d.tw);
    d.tl = 0.0f; d.kick = 0.0f; d.drib = false;
    if (A.cmds_out && valid && is_robot)
        reinterpret_cast<float2 *>(A.cmds_out)[(size_t)e * R + (b - 1)] = make_float2(wl, wr);

    // ---- rsim.send_commands + get_frame, vss_gym_base.py:77-82
    if constexpr (F0) lanes_physics_step<RS_KIND_VSS, L>(VssF0{}, R, b, s, d);
    else lanes_physics_step<RS_KIND_VSS, L>(P, R, b, s, d);

    // ---- _calculate_reward_and_done, vss_gym.py:144-192: every lane evaluates it on the
    // ball (lane 0) and blue 0 (lane 1) and keeps the update of its own task word
    const float bx = g.get(s.x, 0), by = g.get(s.y, 0);
    const float r0x = g.get(s.x, 1), r0y = g.get(s.y, 1), r0vx = g.get(s.vx, 1), r0vy = g.get(s.vy, 1);
    const float wl0 = g.get(wl, 1), wr0 = g.get(wr, 1);
    const float prev = g.get(__uint_as_float(aux), RS_AUX_PREV);
    const uint32_t stw = __float_as_uint(g.get(__uint_as_float(aux), RS_AUX_STEPS));
    int steps = (int)(stw & 0xFFFFFFu);
    bool has_prev = (stw >> 24) & 1u;
    const bool fresh = steps == 0;
    steps += 1;                                                 // vss_gym_base.py:73
    float rew, prev_n = prev;
    float dI[RS_VSS_INFO] = {0.0f, 0.0f, 0.0f, 0.0f, 0.0f, 0.0f};
    bool goal = false;
    if (bx > P.half_len) { dI[0] = 1.0f; dI[4] = 1.0f; rew = 10.0f; goal = true; }
    else if (bx < -P.half_len) { dI[0] = -1.0f; dI[5] = 1.0f; rew = -10.0f; goal = true; }
    else {
        const float length_cm = 2.0f * P.half_len * 100.0f, hl = P.half_len + P.goal_depth;
        const float dx_d = (hl + bx) * 100.0f, dx_a = (hl - bx) * 100.0f, dy = by * 100.0f;
        const float pot = ((-sqrtf(dx_a * dx_a + 2.0f * dy * dy) + sqrtf(dx_d * dx_d + 2.0f * dy * dy)) / length_cm - 1.0f) * 0.5f;
        float grad = 0.0f;
        if (has_prev) grad = clampf((pot - prev) * 3.0f / P.dt, -5.0f, 5.0f);
        prev_n = pot; has_prev = true;
        const float rx = bx - r0x, ry = by - r0y;
        const float rinv = rsqrtf(rx * rx + ry * ry);
        const float move = clampf((rx * rinv * r0vx + ry * rinv * r0vy) * (1.0f / 0.4f), -5.0f, 5.0f);
        const float energy = -(fabsf(wl0) + fabsf(wr0));
        rew = 0.2f * move + 0.8f * grad + 2e-4f * energy;
        dI[1] = 0.2f * move; dI[2] = 0.8f * grad; dI[3] = 2e-4f * energy;
    }
    const bool tr = steps >= A.max_steps;                       // TimeLimit, __init__.py:4
    const bool reset = A.auto_reset && (goal || tr);
    if (valid && b == 0) { A.reward[e] = rew; A.done[e] = goal ? 1 : 0; A.trunc[e] = tr ? 1 : 0; }
    {
        float delta = 0.0f;
#pragma unroll
        for (int i = 0; i < RS_VSS_INFO; ++i) if (b == RS_AUX_INFO + i) delta = dI[i];
        uint32_t out = __float_as_uint((fresh ? 0.0f : __uint_as_float(aux)) + delta);
        if (b == RS_AUX_PREV) out = __float_as_uint(reset ? 0.0f : prev_n);
        if (b == RS_AUX_STEPS) out = reset ? 0u : ((uint32_t)steps | ((has_prev ? 1u : 0u) << 24));
        if (valid) S.aux[(size_t)b * S.np + e] = out;
    }
    lanes_reset_place<RS_TASK_VSS, R, L>(P, pbuf[tid >> 5], reset, valid, b, is_robot, gid, A.seed, t_now, s);
    if (reset) { s.vx = 0.0f; s.vy = 0.0f; s.om = 0.0f; a = make_float2(0.0f, 0.0f); }
    __syncwarp();
    // ---- stores
    if (valid) {
        lanes_store<L>(S, R, b, e, s);
        if (is_ou) S.ou[(size_t)p * S.np + e] = a;
    }
    // ---- observation row (vss_gym.py:93-117): ball [0, 4), blue 7 each, yellow 5 each
    {
        float *o = tile + el * NOBS;
        const float nx = nrm(s.x, P.inv_max_pos), ny = nrm(s.y, P.inv_max_pos);
        const float nvx = nrm(s.vx, P.inv_max_v), nvy = nrm(s.vy, P.inv_max_v), nw = nrm(s.om, P.inv_max_w_rad);
        if (b == 0) { o[0] = nx; o[1] = ny; o[2] = nvx; o[3] = nvy; }
        else if (b <= NB) {
            float sn, cs;
            __sincosf(s.th, &sn, &cs);
            float *q = o + 4 + 7 * (b - 1);
            q[0] = nx; q[1] = ny; q[2] = sn; q[3] = cs; q[4] = nvx; q[5] = nvy; q[6] = nw;
        } else if (b <= R) {
            float *q = o + 4 + 7 * NB + 5 * (b - NB - 1);
            q[0] = nx; q[1] = ny; q[2] = nvx; q[3] = nvy; q[4] = nw;
        }
    }
    if (valid) step_counter_bump<RS_CTR_GROUP * L>(A.ctr, e, t_now, tile_lock.lock != nullptr);
    warp_tile_store(A.obs + (size_t)w0 * NOBS, tile + (tid >> 5) * MPW * NOBS, wrows, NOBS, tile_lock.lock == nullptr);
    step_end(tile_lock);
}

struct SslStepArgs {
    const float *actions;    // [N][5]
    float *obs, *reward;
    uint8_t *done, *trunc;
    float *cmds_out;         // [N][R][8] or null
    int auto_reset, max_steps;
    uint64_t seed;
    uint32_t *ctr;
    uint32_t env_offset;
    uint32_t *flags;         // step-to-step overlap, as in VssStepArgs (chain 0 or 2 here)
    int chain;
};

// SSLHWStaticDefendersEnv.step / SSLContestedPossessionEnv.step
template <int TASK, int NB, int NY, int BS>
__global__ void __launch_bounds__(BS)
k_ssl_env_step(const __grid_constant__ DevParams P, const StatePtrs S, const SslStepArgs A) {
    constexpr int R = NB + NY, NOBS = 4 + 8 * NB + 2 * NY;
    // one region per WARP, as in k_vss_env_step: contact scratch during the physics (worlds of >= 3 robots:
    // per-lane resolve through shared memory, contacts_via_smem_ssl), then placement words, then its 32 obs rows
    constexpr bool SCR = R >= 3;
    constexpr int SCRATCH = SCR ? 2 * (R + 1) * 32 * 4 : 0, WF = 32 * NOBS > SCRATCH ? 32 * NOBS : SCRATCH;
    __shared__ __align__(128) float smem[BS / 32][WF];
    float *const wtile = smem[threadIdx.x >> 5];
    float4 *const cq = reinterpret_cast<float4 *>(wtile) + (threadIdx.x & 31);
    const int tid = threadIdx.x;
    const int e0 = blockIdx.x * BS;
    const int e = e0 + tid;
    const int w0 = e0 + (tid & ~31);                       // first env of this warp
    const int wrows = min(32, S.n - w0);
    const unsigned live = __ballot_sync(0xffffffffu, e < S.n);
    const StepTile tile_lock = step_begin<BS>(A.ctr, A.flags, A.chain, w0);
    if (e < S.n) {
        Scene<R> s;
        load_scene<R>(P, S, e, s);
        const uint32_t t_now = tile_lock.lock ? tile_lock.t : step_counter_read<RS_CTR_GROUP>(A.ctr, e, live);
        const int st = __ldcg(S.steps + e);
        // reward_shaping_total is never loaded (as in k_vss_env_step): zeroed by a store at the first step of an
        // episode, updated by fire-and-forget reductions (one RED.ADD.F32 per word = the rounding of load-add-store);
        // 72 bytes per env-step less, and no second exposed load latency behind the step word
        // ---- _get_commands + convert_actions, static_defenders.py:114-148
        float a[RS_SSL_ACT];
#pragma unroll
        for (int i = 0; i < RS_SSL_ACT; ++i) a[i] = __ldcg(A.actions + (size_t)e * RS_SSL_ACT + i);
        const float max_v = 2.5f, max_w = 10.0f, kick_speed = 5.0f;
        float cmd[8];
        {
            float sn, cs;
            __sincosf(s.th[0], &sn, &cs);
            const float vx = a[0] * max_v, vy = a[1] * max_v;
            const float lx = vx * cs + vy * sn, ly = -vx * sn + vy * cs;
            const float vn2 = lx * lx + ly * ly;
            const float c = vn2 < max_v * max_v ? 1.0f : max_v * rsqrtf(vn2);
            cmd[0] = 0.0f; cmd[1] = lx * c; cmd[2] = ly * c; cmd[3] = a[2] * max_w; cmd[4] = 0.0f;
            cmd[5] = a[3] > 0.0f ? kick_speed : 0.0f; cmd[6] = 0.0f; cmd[7] = a[4] > 0.0f ? 1.0f : 0.0f;
        }
        Drive<R> d;
        bool drib0;
        ssl_target(P, cmd, d.tf[0], d.tl[0], d.tw[0], d.kick[0], drib0);
        d.drib = drib0 ? 1u : 0u;
#pragma unroll
        for (int r = 1; r < R; ++r) { d.tf[r] = 0.0f; d.tl[r] = 0.0f; d.tw[r] = 0.0f; d.kick[r] = 0.0f; }
        if (A.cmds_out) {
            for (int i = 0; i < 8; ++i) A.cmds_out[(size_t)e * R * 8 + i] = cmd[i];
            for (int i = 8; i < R * 8; ++i) A.cmds_out[(size_t)e * R * 8 + i] = 0.0f;
        }
        const float lbx = s.bx, lby = s.by, lrx = s.x[0], lry = s.y[0];   // last_frame

        if constexpr (SCR) physics_step<RS_KIND_SSL, R>(P, s, d, live, cq, cq + (R + 1) * 32, 32);
        else physics_step<RS_KIND_SSL, R>(P, s, d, live);

        // ---- the task words (loaded at the top) are first needed here, after the physics
        int steps = st & 0xFFFFFF;
        float *const info = S.info + e;                        // word i of this match: info[i * np]
        const size_t np = (size_t)S.np;
        if (steps == 0) {
#pragma unroll
            for (int i = 0; i < RS_SSL_INFO; ++i) info[i * np] = 0.0f;
        }
        steps += 1;
        // ---- _calculate_reward_and_done, static_defenders.py:150-212 / contested_possession.py:136-208
        float rew = 0.0f; bool dn = false;
        if (TASK == RS_TASK_SSL_CONTESTED_POSSESSION) {
            int cnt = 0;
#pragma unroll
            for (int r = NB; r < R; ++r)
                if (fabsf(s.vx[r]) > 0.1f || fabsf(s.vy[r]) > 0.1f) { ++cnt; dn = true; }
            if (cnt) atomicAdd(info + 8 * np, (float)cnt);
        }
        const float hl = P.half_len, hw = P.half_wid;
        if (s.x[0] < -0.2f || fabsf(s.y[0]) > hw) { dn = true; atomicAdd(info + 4 * np, 1.0f); }
        else if (s.x[0] > hl - P.pen_len && fabsf(s.y[0]) < P.half_pen_wid) { dn = true; atomicAdd(info + 1 * np, 1.0f); }
        else if (s.bx < 0.0f || fabsf(s.by) > hw) { dn = true; atomicAdd(info + 2 * np, 1.0f); }
        else if (s.bx > hl) {
            dn = true;
            if (fabsf(s.by) < P.half_goal_wid) { rew = 5.0f; atomicAdd(info, 1.0f); } else { atomicAdd(info + 3 * np, 1.0f); }
        } else {
            const float ball_dist_scale = sqrtf(4.0f * hw * hw + hl * hl);
            const float ball_grad_scale = sqrtf(hw * hw + hl * hl) * 0.25f;
            const float energy_scale = 160.0f * 4.0f * (TASK == RS_TASK_SSL_STATIC_DEFENDERS ? 1000.0f : 1200.0f);
            const float ld = sqrtf((lrx - lbx) * (lrx - lbx) + (lry - lby) * (lry - lby));
            const float nd = sqrtf((s.x[0] - s.bx) * (s.x[0] - s.bx) + (s.y[0] - s.by) * (s.y[0] - s.by));
            const float bd = clampf(ld - nd, -1.0f, 1.0f) / ball_dist_scale;
            const float lg = sqrtf((hl - lbx) * (hl - lbx) + lby * lby);
            const float ng = sqrtf((hl - s.bx) * (hl - s.bx) + s.by * s.by);
            const float bg = clampf(lg - ng, -1.0f, 1.0f) / ball_grad_scale;
            float sn, cs;
            __sincosf(s.th[0], &sn, &cs);
            const float vf = cs * s.vx[0] + sn * s.vy[0], vl = -sn * s.vx[0] + cs * s.vy[0];
            float en = 0.0f;
#pragma unroll
            for (int i = 0; i < 4; ++i) en += fabsf((P.J[i][0] * vf + P.J[i][1] * vl + P.J[i][2] * s.om[0]) * P.inv_rw);
            const float er = -en / energy_scale;
            atomicAdd(info + 5 * np, bd); atomicAdd(info + 6 * np, bg); atomicAdd(info + 7 * np, er);
            rew = bd + bg + er;
        }
        const bool tr = steps >= A.max_steps;
        A.reward[e] = rew; A.done[e] = dn ? 1 : 0; A.trunc[e] = tr ? 1 : 0;
        if (TASK == RS_TASK_SSL_CONTESTED_POSSESSION) {
            // two draws, no rejection loop: the scalar placement is cheaper than a warp-wide one (7.19 vs 7.34 us)
            if (A.auto_reset && (dn || tr)) {
                const PlaceStream g{nullptr, 0, make_uint2((uint32_t)A.seed, (uint32_t)(A.seed >> 32)),
                                    A.env_offset + (uint32_t)e, t_now, 0};
                place_from_stream<TASK, R>(P, g, s);
                steps = 0;
            }
        } else if (A.auto_reset) {
            // warp-cooperative placement of the ending matches (see k_vss_env_step); the word buffer is
            // this warp's part of the observation tile, not yet written
            uint32_t *const wbuf = reinterpret_cast<uint32_t *>(wtile);
            static_assert(32 * NOBS >= 128, "128 placement words fit the warp's observation rows");
            __syncwarp(live);      // the words overlay the other lanes' contact scratch
            unsigned need = __ballot_sync(live, dn || tr);
            while (need) {
                const int src = __ffs((int)need) - 1;
                need &= need - 1;
                const uint32_t env_src = A.env_offset + (uint32_t)(w0 + src);
                warp_placement_words(wbuf, live, A.seed, env_src, t_now);
                if ((tid & 31) == src) {
                    const PlaceStream g{wbuf, wrows, make_uint2((uint32_t)A.seed, (uint32_t)(A.seed >> 32)), env_src, t_now, 0};
                    place_from_stream<TASK, R>(P, g, s);
                    steps = 0;
                }
            }
        }
        store_scene<R>(P, S, e, s);
        S.steps[e] = steps;
        __syncwarp(live);          // the rows below overlay the scratch / the placement words an ending lane may still read
        ssl_obs<NB, NY>(P, s, wtile + (tid & 31) * NOBS);
        step_counter_bump<RS_CTR_GROUP>(A.ctr, e, t_now, tile_lock.lock != nullptr);
    }
    warp_tile_store(A.obs + (size_t)w0 * NOBS, wtile, wrows, NOBS, tile_lock.lock == nullptr);
    step_end(tile_lock);
}

// SSLHWDribblingEnv.step (TASK 3: 1 blue + 4 yellow, dribbling.py) and SSLPassEnduranceEnv.step
// (TASK 4: 2 blue, pass_endurance.py), one lane per match.  The per-episode counter of the
// task (dribbling: checkpoints_count; pass endurance: stopped_steps) lives in the `prev` word;
// info[0..1] of pass endurance = reward_shaping_total {reversed_dist, ball_grad}
// (pass_endurance.py:113-114).  pass_endurance.py never increments holding_steps (:56, :92,
// :121), so its `> 15` test never fires and is not restated.
template <int TASK, int NB, int NY>
__device__ __forceinline__ void ssl_hw_env_step_match(const DevParams &P, const StatePtrs &S, const SslStepArgs &A,
                                                      const int e, const unsigned live, float *const wtile, const StepTile &tile_lock) {
    constexpr int R = NB + NY;
    constexpr bool DRIB = TASK == RS_TASK_SSL_DRIBBLING;
    constexpr int NACT = DRIB ? RS_DRIB_ACT : RS_PASS_ACT, NOBS = DRIB ? RS_DRIB_OBS : RS_PASS_OBS;
    float4 *const cq = reinterpret_cast<float4 *>(wtile) + (threadIdx.x & 31);
    Scene<R> s;
    load_scene<R>(P, S, e, s);
    int steps = __ldcg(S.steps + e) & 0xFFFFFF;
    float counter = steps == 0 ? 0.0f : __ldcg(S.prev + e);
    float info0 = steps == 0 ? 0.0f : __ldcg(S.info + e), info1 = steps == 0 ? 0.0f : __ldcg(S.info + (size_t)S.np + e);
    steps += 1;
    const uint32_t t_now = tile_lock.lock ? tile_lock.t : step_counter_read<RS_CTR_GROUP>(A.ctr, e, live);
    float a[NACT];
#pragma unroll
    for (int i = 0; i < NACT; ++i) a[i] = __ldcg(A.actions + (size_t)e * NACT + i);
    const float max_v = 2.5f, max_w = 10.0f, max_kick_x = 5.0f;
    float cmd[R][8];
#pragma unroll
    for (int r = 0; r < R; ++r)
#pragma unroll
        for (int i = 0; i < 8; ++i) cmd[r][i] = 0.0f;
    if (DRIB) {                                   // dribbling.py:107-133
        float sn, cs;
        __sincosf(s.th[0], &sn, &cs);
        const float vx = a[0] * max_v, vy = a[1] * max_v;
        const float lx = vx * cs + vy * sn, ly = -vx * sn + vy * cs;
        const float vn2 = lx * lx + ly * ly;
        const float c = vn2 < max_v * max_v ? 1.0f : max_v * rsqrtf(vn2);
        cmd[0][1] = lx * c; cmd[0][2] = ly * c; cmd[0][3] = a[2] * max_w; cmd[0][7] = a[NACT - 1] > 0.0f ? 1.0f : 0.0f;
    } else {                                      // pass_endurance.py:100-124
        const float a1 = fabsf(a[1]) > 0.5f ? a[1] : 0.0f;
        cmd[0][3] = a[0] * max_w; cmd[0][5] = a1 * max_kick_x; cmd[0][7] = a[2] > 0.0f ? 1.0f : 0.0f;
        cmd[R > 1 ? 1 : 0][7] = 1.0f;             // the receiver: dribbler on, everything else 0
    }
    Drive<R> d;
    d.drib = 0;
#pragma unroll
    for (int r = 0; r < R; ++r) {
        if (r < NB) {
            bool drib;
            ssl_target(P, cmd[r], d.tf[r], d.tl[r], d.tw[r], d.kick[r], drib);
            if (drib) d.drib |= 1u << r;
        } else { d.tf[r] = 0.0f; d.tl[r] = 0.0f; d.tw[r] = 0.0f; d.kick[r] = 0.0f; }
    }
    if (A.cmds_out) {
#pragma unroll
        for (int r = 0; r < R; ++r)
#pragma unroll
            for (int i = 0; i < 8; ++i) A.cmds_out[((size_t)e * R + r) * 8 + i] = cmd[r][i];
    }
    const float lbx = s.bx, lby = s.by;           // last_frame.ball

    if constexpr (R >= 3) physics_step<RS_KIND_SSL, R>(P, s, d, live, cq, cq + (R + 1) * 32, 32);
    else physics_step<RS_KIND_SSL, R>(P, s, d, live);

    // ssl_gym_base.py:83-85: the observation is taken BEFORE the reward updates the counter.  The row is staged
    // in the warp's shared-memory tile (over the contact scratch) and leaves with the warp copy of the kernel.
    __syncwarp(live);
    float *o = wtile + (threadIdx.x & 31) * NOBS;
    ssl_hw_obs<TASK>(P, s, counter, o);
    float rew = 0.0f; bool dn = false;
    if (DRIB) {                                   // dribbling.py:135-185
        const float n0 = -0.5f, n1 = -1.0f, n2 = -1.5f, n3 = -2.0f, fm = 1.0f;
        int cc = (int)counter;
#pragma unroll
        for (int r = NB; r < R; ++r) if (fabsf(s.vx[r]) > 0.05f || fabsf(s.vy[r]) > 0.05f) dn = true;
        if (s.x[0] < n3 - fm || s.x[0] > fm || fabsf(s.y[0]) > fm) dn = true;
        else if (cc == 0) {
            if (s.bx < n0 && s.bx > n1 && lby >= 0.0f && s.by < 0.0f) { rew = 1.0f; cc += 1; }
        } else if (cc == 1) {
            if (s.bx < n1 && s.bx > n2 && lby < 0.0f && s.by >= 0.0f) { rew = 1.0f; cc += 1; }
        } else if (cc % 2 == 0) {
            if (s.bx < n2 && s.bx > n3) {
                if (lby >= 0.0f && s.by < 0.0f) { rew = 1.0f; cc += 1; if (cc == 7) dn = true; }
                else if (lby < 0.0f && s.by >= 0.0f) dn = true;
            }
        } else {
            if (s.bx > n3 - fm && s.bx < n3 && lby < 0.0f && s.by >= 0.0f) { rew = 1.0f; cc += 1; }
        }
        counter = (float)cc;
    } else {                                      // pass_endurance.py:126-150, 183-233
        constexpr int RC = R > 1 ? 1 : 0;
        const float ball_grad_scale = sqrtf(P.half_wid * P.half_wid + P.half_len * P.half_len) * 0.25f;
        const float ld = sqrtf((lbx - s.x[RC]) * (lbx - s.x[RC]) + (lby - s.y[RC]) * (lby - s.y[RC]));
        const float nd = sqrtf((s.bx - s.x[RC]) * (s.bx - s.x[RC]) + (s.by - s.y[RC]) * (s.by - s.y[RC]));
        float sn, cs;
        __sincosf(s.th[RC], &sn, &cs);
        if (touching(P, s.x[RC], s.y[RC], cs, sn, s.bx, s.by)) { rew += 1.0f; dn = true; }
        else {
            const float g = clampf(ld - nd, -1.0f, 1.0f) / ball_grad_scale;
            rew = g; info1 += g;
        }
        // __wrong_ball: centimetre-truncated box between shooter and receiver, 20-step stall counter
        const int ibx = (int)(s.bx * 100.0f), iby = (int)(s.by * 100.0f);
        const int isx = (int)(s.x[0] * 100.0f), isy = (int)(s.y[0] * 100.0f);
        const int irx = (int)(s.x[RC] * 100.0f), iry = (int)(s.y[RC] * 100.0f);
        const bool inside = min(isx, irx) <= ibx && ibx <= max(isx, irx) && min(isy, iry) <= iby && iby <= max(isy, iry);
        int stopped = (int)counter;
        stopped = fabsf(ld - nd) < 0.01f ? stopped + 1 : 0;
        counter = (float)stopped;
        if (stopped > 20 || !inside) { rew -= 1.0f; dn = true; }
        if (dn) {
            const float dr = sqrtf((s.x[RC] - s.x[0]) * (s.x[RC] - s.x[0]) + (s.y[RC] - s.y[0]) * (s.y[RC] - s.y[0]));
            info0 = (dr - nd) / dr;
        }
    }
    const bool tr = steps >= A.max_steps;
    A.reward[e] = rew; A.done[e] = dn ? 1 : 0; A.trunc[e] = tr ? 1 : 0;
    S.info[e] = info0; S.info[(size_t)S.np + e] = info1;
    if (A.auto_reset && (dn || tr)) {
        Scene<0> tmp;
        task_place<TASK, 0>(P, Rng(A.seed, A.env_offset + (uint32_t)e, t_now, RS_STREAM_AUTORESET), tmp);
        s.bx = tmp.bx; s.by = tmp.by; s.bvx = 0.0f; s.bvy = 0.0f;
#pragma unroll
        for (int r = 0; r < R; ++r) {
            s.x[r] = tmp.x[r]; s.y[r] = tmp.y[r]; s.th[r] = tmp.th[r];
            s.vx[r] = 0.0f; s.vy[r] = 0.0f; s.om[r] = 0.0f;
        }
        steps = 0; counter = 0.0f;
        ssl_hw_obs<TASK>(P, s, 0.0f, o);
    }
    store_scene<R>(P, S, e, s);
    S.steps[e] = steps;
    S.prev[e] = counter;
    step_counter_bump<RS_CTR_GROUP>(A.ctr, e, t_now, tile_lock.lock != nullptr);
}
template <int TASK, int NB, int NY, int BS>
__global__ void __launch_bounds__(BS)
k_ssl_hw_env_step(const __grid_constant__ DevParams P, const StatePtrs S, const SslStepArgs A) {
    constexpr int R = NB + NY, NOBS = TASK == RS_TASK_SSL_DRIBBLING ? RS_DRIB_OBS : RS_PASS_OBS;
    constexpr int SCRATCH = R >= 3 ? 2 * (R + 1) * 32 * 4 : 0, WF = 32 * NOBS > SCRATCH ? 32 * NOBS : SCRATCH;
    __shared__ __align__(128) float smem[BS / 32][WF];
    float *const wtile = smem[threadIdx.x >> 5];
    const int e = blockIdx.x * BS + threadIdx.x;
    const int w0 = blockIdx.x * BS + (threadIdx.x & ~31);
    const unsigned live = __ballot_sync(0xffffffffu, e < S.n);
    const StepTile tile_lock = step_begin<BS>(A.ctr, A.flags, A.chain, w0);
    if (e < S.n) ssl_hw_env_step_match<TASK, NB, NY>(P, S, A, e, live, wtile, tile_lock);
    warp_tile_store(A.obs + (size_t)w0 * NOBS, wtile, min(32, S.n - w0), NOBS, tile_lock.lock == nullptr);
    step_end(tile_lock);
}


// SSLHWStaticDefendersEnv.step / SSLContestedPossessionEnv.step, one lane per BODY
// (rs_lanes.cuh): L lanes per match (8 for 1 v 6, 4 for 1 v 1).  Lane b owns task words
// b, b + L, ... of the 2 + RS_SSL_INFO words of its match.
template <int TASK, int NB, int NY, int L, int BS>
__global__ void __launch_bounds__(BS)
k_ssl_env_step_lanes(const __grid_constant__ DevParams P, const StatePtrs S, const SslStepArgs A) {
    constexpr int R = NB + NY, NOBS = 4 + 8 * NB + 2 * NY, EPB = BS / L, MPW = 32 / L;
    constexpr int NW = RS_AUX_INFO + RS_SSL_INFO, KW = (NW + L - 1) / L;
    static_assert(NB == 1 && R + 1 <= L, "the benchmarked SSL tasks have one agent, blue 0");
    __shared__ __align__(128) float tile[EPB * NOBS];
    __shared__ __align__(16) uint32_t pbuf[BS / 32][128 + MPW * (2 + 3 * R)];      // auto-reset placement (rs_lanes.cuh)
    const int tid = threadIdx.x;
    const int b = tid & (L - 1);
    const int el = tid / L;
    const int e = blockIdx.x * EPB + el;
    const bool valid = e < S.n;
    const int ec = valid ? e : S.n - 1;
    const int w0 = blockIdx.x * EPB + (tid >> 5) * MPW;
    const int wrows = min(MPW, S.n - w0);
    const LaneGroup<L> g;
    const bool is_robot = b >= 1 && b <= R;

    const StepTile tile_lock = step_begin<BS>(A.ctr, A.flags, A.chain, w0);
    LaneBody s;
    lanes_load<L>(S, R, b, ec, s);
    uint32_t aux[KW];
#pragma unroll
    for (int k = 0; k < KW; ++k) {
        const int w = b + k * L;                               // (word 0, the VSS ball potential, is unused here)
        aux[k] = (w < NW && w != RS_AUX_PREV) ? __ldcg(S.aux + (size_t)w * S.np + ec) : 0u;
    }
    float a[RS_SSL_ACT];
#pragma unroll
    for (int i = 0; i < RS_SSL_ACT; ++i) a[i] = b == 1 ? __ldcg(A.actions + (size_t)ec * RS_SSL_ACT + i) : 0.0f;
    const uint32_t t_now = tile_lock.lock ? tile_lock.t : step_counter_read<RS_CTR_GROUP * L>(A.ctr, ec);

    // ---- _get_commands + convert_actions, static_defenders.py:114-148 (blue 0; the other
    // robots get all-zero rows, rsim.py:129-130)
    const float max_v = 2.5f, max_w = 10.0f, kick_speed = 5.0f;
    float cmd[8];
    {
        float sn, cs;
        __sincosf(s.th, &sn, &cs);
        const float vx = a[0] * max_v, vy = a[1] * max_v;
        const float lx = vx * cs + vy * sn, ly = -vx * sn + vy * cs;
        const float vn2 = lx * lx + ly * ly;
        const float c = vn2 < max_v * max_v ? 1.0f : max_v * rsqrtf(vn2);
        cmd[0] = 0.0f; cmd[1] = lx * c; cmd[2] = ly * c; cmd[3] = a[2] * max_w; cmd[4] = 0.0f;
        cmd[5] = a[3] > 0.0f ? kick_speed : 0.0f; cmd[6] = 0.0f; cmd[7] = a[4] > 0.0f ? 1.0f : 0.0f;
    }
    LaneDrive d;
    ssl_target(P, cmd, d.tf, d.tl, d.tw, d.kick, d.drib);
    if (b != 1) { d.tf = 0.0f; d.tl = 0.0f; d.tw = 0.0f; d.kick = 0.0f; d.drib = false; }
    if (A.cmds_out && valid && is_robot) {
        float *o = A.cmds_out + ((size_t)e * R + (b - 1)) * 8;
#pragma unroll
        for (int i = 0; i < 8; ++i) o[i] = b == 1 ? cmd[i] : 0.0f;
    }
    const float ox = s.x, oy = s.y;                              // last_frame (own body)

    lanes_physics_step<RS_KIND_SSL, L>(P, R, b, s, d);

    // ---- _calculate_reward_and_done, static_defenders.py:150-212 / contested_possession.py:136-208
    const float bx = g.get(s.x, 0), by = g.get(s.y, 0), lbx = g.get(ox, 0), lby = g.get(oy, 0);
    const float rx = g.get(s.x, 1), ry = g.get(s.y, 1), lrx = g.get(ox, 1), lry = g.get(oy, 1);
    const float rvx = g.get(s.vx, 1), rvy = g.get(s.vy, 1), rth = g.get(s.th, 1), rom = g.get(s.om, 1);
    int steps = (int)(__float_as_uint(g.get(__uint_as_float(aux[0]), RS_AUX_STEPS)) & 0xFFFFFFu);
    const bool fresh = steps == 0;
    steps += 1;
    float rew = 0.0f; bool dn = false;
    float dI[RS_SSL_INFO];
#pragma unroll
    for (int i = 0; i < RS_SSL_INFO; ++i) dI[i] = 0.0f;
    if (TASK == RS_TASK_SSL_CONTESTED_POSSESSION) {
        const bool moved = b > NB && b <= R && (fabsf(s.vx) > 0.1f || fabsf(s.vy) > 0.1f);
        const int cnt = __popc(g.bits(__ballot_sync(RS_FULL_MASK, moved)));
        if (cnt) { dI[8] = (float)cnt; dn = true; }
    }
    const float hl = P.half_len, hw = P.half_wid;
    if (rx < -0.2f || fabsf(ry) > hw) { dn = true; dI[4] = 1.0f; }
    else if (rx > hl - P.pen_len && fabsf(ry) < P.half_pen_wid) { dn = true; dI[1] = 1.0f; }
    else if (bx < 0.0f || fabsf(by) > hw) { dn = true; dI[2] = 1.0f; }
    else if (bx > hl) {
        dn = true;
        if (fabsf(by) < P.half_goal_wid) { rew = 5.0f; dI[0] = 1.0f; } else { dI[3] = 1.0f; }
    } else {
        const float ball_dist_scale = sqrtf(4.0f * hw * hw + hl * hl);
        const float ball_grad_scale = sqrtf(hw * hw + hl * hl) * 0.25f;
        const float energy_scale = 160.0f * 4.0f * (TASK == RS_TASK_SSL_STATIC_DEFENDERS ? 1000.0f : 1200.0f);
        const float ld = sqrtf((lrx - lbx) * (lrx - lbx) + (lry - lby) * (lry - lby));
        const float nd = sqrtf((rx - bx) * (rx - bx) + (ry - by) * (ry - by));
        const float bd = clampf(ld - nd, -1.0f, 1.0f) / ball_dist_scale;
        const float lg = sqrtf((hl - lbx) * (hl - lbx) + lby * lby);
        const float ng = sqrtf((hl - bx) * (hl - bx) + by * by);
        const float bg = clampf(lg - ng, -1.0f, 1.0f) / ball_grad_scale;
        float sn, cs;
        __sincosf(rth, &sn, &cs);
        const float vf = cs * rvx + sn * rvy, vl = -sn * rvx + cs * rvy;
        float en = 0.0f;
#pragma unroll
        for (int i = 0; i < 4; ++i) en += fabsf((P.J[i][0] * vf + P.J[i][1] * vl + P.J[i][2] * rom) * P.inv_rw);
        const float er = -en / energy_scale;
        dI[5] = bd; dI[6] = bg; dI[7] = er;
        rew = bd + bg + er;
    }
    const bool tr = steps >= A.max_steps;
    const bool reset = A.auto_reset && (dn || tr);
    if (valid && b == 0) { A.reward[e] = rew; A.done[e] = dn ? 1 : 0; A.trunc[e] = tr ? 1 : 0; }
#pragma unroll
    for (int k = 0; k < KW; ++k) {
        const int w = b + k * L;
        float delta = 0.0f;
#pragma unroll
        for (int i = 0; i < RS_SSL_INFO; ++i) if (w == RS_AUX_INFO + i) delta = dI[i];
        uint32_t out = __float_as_uint((fresh ? 0.0f : __uint_as_float(aux[k])) + delta);
        if (w == RS_AUX_STEPS) out = reset ? 0u : (uint32_t)steps;
        if (valid && w < NW && w != RS_AUX_PREV) S.aux[(size_t)w * S.np + e] = out;
    }
    lanes_reset_place<TASK, R, L>(P, pbuf[tid >> 5], reset, valid, b, is_robot, A.env_offset + (uint32_t)ec, A.seed, t_now, s);
    if (reset) { s.vx = 0.0f; s.vy = 0.0f; s.om = 0.0f; }
    __syncwarp();
    if (valid) lanes_store<L>(S, R, b, e, s);
    // ---- observation row (static_defenders.py:90-112): ball 4, blue 8 (with infrared), yellow x y
    {
        const float nbx = g.get(s.x, 0), nby = g.get(s.y, 0);       // ball after a possible reset
        const float inv_v = 1.0f / 2.5f, inv_w = RS_DEG_F / 10.0f;  // quirks preserved, SURVEY A.1
        float *o = tile + el * NOBS;
        const float nx = nrm(s.x, P.inv_max_pos), ny = nrm(s.y, P.inv_max_pos);
        if (b == 0) { o[0] = nx; o[1] = ny; o[2] = nrm(s.vx, inv_v); o[3] = nrm(s.vy, inv_v); }
        else if (b <= NB) {
            float sn, cs;
            __sincosf(s.th, &sn, &cs);
            float *q = o + 4 + 8 * (b - 1);
            q[0] = nx; q[1] = ny; q[2] = sn; q[3] = cs;
            q[4] = nrm(s.vx, inv_v); q[5] = nrm(s.vy, inv_v); q[6] = nrm(s.om, inv_w);
            q[7] = touching(P, s.x, s.y, cs, sn, nbx, nby) ? 1.0f : 0.0f;
        } else if (b <= R) {
            float *q = o + 4 + 8 * NB + 2 * (b - NB - 1);
            q[0] = nx; q[1] = ny;
        }
    }
    if (valid) step_counter_bump<RS_CTR_GROUP * L>(A.ctr, e, t_now, tile_lock.lock != nullptr);
    warp_tile_store(A.obs + (size_t)w0 * NOBS, tile + (tid >> 5) * MPW * NOBS, wrows, NOBS, tile_lock.lock == nullptr);
    step_end(tile_lock);
}

// simulator.step(cmds), one lane per BODY: any (kind, R <= 31) with L = 2^k >= R + 1
template <int KIND, int L, int BS, bool F0 = false /* VSS 3 v 3, field 0, 25 ms: VssF0's immediates */>
__global__ void __launch_bounds__(BS)
k_step_lanes(const __grid_constant__ DevParams P, const StatePtrs S, const float *__restrict__ cmds) {
    const int R = P.n_robots;
    const int tid = threadIdx.x;
    const int b = tid & (L - 1);
    const int e = blockIdx.x * (BS / L) + tid / L;
    const bool valid = e < S.n;
    const int ec = valid ? e : S.n - 1;
    const bool is_robot = b >= 1 && b <= R;
    pdl_wait();
    pdl_release();
    LaneBody s;
    lanes_load<L>(S, R, b, ec, s);
    LaneDrive d;
    d.tf = 0.0f; d.tl = 0.0f; d.tw = 0.0f; d.kick = 0.0f; d.drib = false;
    if (is_robot) {
        if (KIND == RS_KIND_VSS) {
            const float2 c = reinterpret_cast<const float2 *>(cmds)[(size_t)ec * R + (b - 1)];
            vss_target(P, c.x, c.y, d.tf, d.tw);
        } else {
            const float4 *c4 = reinterpret_cast<const float4 *>(cmds) + ((size_t)ec * R + (b - 1)) * 2;
            const float4 lo = c4[0], hi = c4[1];
            const float cmd[8] = {lo.x, lo.y, lo.z, lo.w, hi.x, hi.y, hi.z, hi.w};
            ssl_target(P, cmd, d.tf, d.tl, d.tw, d.kick, d.drib);
        }
    }
    __syncwarp();
    if constexpr (F0) lanes_physics_step<KIND, L>(VssF0{}, R, b, s, d);
    else lanes_physics_step<KIND, L>(P, R, b, s, d);
    if (valid) lanes_store<L>(S, R, b, e, s);
}

// simulator.step(cmds): physics only, any (kind, R).  RT > 0: register resident scene.
// F0 (VSS 3 v 3 on field 0 at 25 ms only): 1 = compile-time physics constants (VssF0), 2 = the same with the packed
// fp32x2 forms (VssF0P), as in k_vss_env_step.
template <int KIND, int RT, int BS, int F0 = 0>
__global__ void __launch_bounds__(BS)
k_step(const __grid_constant__ DevParams P, const StatePtrs S, const float *__restrict__ cmds) {
    const int e = blockIdx.x * BS + threadIdx.x;
    const unsigned live = __ballot_sync(0xffffffffu, e < S.n);
    pdl_wait();
    pdl_release();
    if (e >= S.n) return;
    const int R = RT > 0 ? RT : P.n_robots;
    Scene<RT> s;
    load_scene<RT>(P, S, e, s);
    Drive<RT> d;
    d.drib = 0;
    if (KIND == RS_KIND_VSS) {
        const float2 *c2 = reinterpret_cast<const float2 *>(cmds) + (size_t)e * R;
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const float2 c = c2[r];
            vss_target<F0 == 2 && VssF0P::packed>(P, c.x, c.y, d.tf[r], d.tw[r]);
            d.tl[r] = 0.0f; d.kick[r] = 0.0f;
        }
    } else {
        const float4 *c4 = reinterpret_cast<const float4 *>(cmds) + (size_t)e * R * 2;
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const float4 lo = c4[2 * r], hi = c4[2 * r + 1];
            const float cmd[8] = {lo.x, lo.y, lo.z, lo.w, hi.x, hi.y, hi.z, hi.w};
            bool drib;
            ssl_target(P, cmd, d.tf[r], d.tl[r], d.tw[r], d.kick[r], drib);
            if (drib) d.drib |= 1u << r;
        }
    }
    if constexpr ((KIND == RS_KIND_VSS && RT >= 1 && RT <= 7) || (KIND == RS_KIND_SSL && RT >= 3 && RT <= 7)) {
        // per-lane contact resolve through shared memory (rs_device.cuh): one region per warp
        __shared__ __align__(16) float4 scratch[BS / 32][2 * (RT + 1) * 32];
        float4 *const cq = scratch[threadIdx.x >> 5] + (threadIdx.x & 31);
        if constexpr (F0 == 2) physics_step<KIND, RT>(VssF0P{}, s, d, live, cq, cq + (RT + 1) * 32, 32);
        else if constexpr (F0 == 1) physics_step<KIND, RT>(VssF0{}, s, d, live, cq, cq + (RT + 1) * 32, 32);
        else physics_step<KIND, RT>(P, s, d, live, cq, cq + (RT + 1) * 32, 32);
    } else {
        physics_step<KIND, RT>(P, s, d, live);
    }
    store_scene<RT>(P, S, e, s);
}

// simulator.get_state(): Entities/Frame.py:20-47 / :55-93 rows, degrees on the wire
__global__ void k_get_state(const DevParams P, const StatePtrs S, float *__restrict__ out) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= S.n) return;
    const int R = P.n_robots;
    const int K = P.kind == RS_KIND_VSS ? RS_STATE_VSS_ROBOT : RS_STATE_SSL_ROBOT;
    float *o = out + (size_t)e * (RS_STATE_BALL + K * R);
    const float4 b = S.body[e];
    o[0] = b.x; o[1] = b.y; o[2] = P.ball_r; o[3] = b.z; o[4] = b.w;
    for (int r = 0; r < R; ++r) {
        const float4 q = S.body[(size_t)(r + 1) * S.np + e];
        const float2 a = S.ang[(size_t)r * S.np + e];
        float *p = o + RS_STATE_BALL + K * r;
        p[0] = q.x; p[1] = q.y; p[2] = a.x * RS_DEG_F; p[3] = q.z; p[4] = q.w; p[5] = a.y * RS_DEG_F;
        if (P.kind == RS_KIND_SSL) {
            float sn, cs;
            __sincosf(a.x, &sn, &cs);
            p[6] = touching(P, q.x, q.y, cs, sn, b.x, b.y) ? 1.0f : 0.0f;
            const float vf = cs * q.z + sn * q.w, vl = -sn * q.z + cs * q.w;
            for (int i = 0; i < 4; ++i) p[7 + i] = (P.J[i][0] * vf + P.J[i][1] * vl + P.J[i][2] * a.y) * P.inv_rw;
        }
    }
}

__device__ __forceinline__ void clear_task(const DevParams &P, const StatePtrs &S, int e) {
    for (int r = 0; r + 1 < P.n_robots; ++r) S.ou[(size_t)r * S.np + e] = make_float2(0.0f, 0.0f);
    S.prev[e] = 0.0f; S.steps[e] = 0;
}

// simulator.reset(ball, blue, yellow): rsim.py:36-38, 52-75
__global__ void k_reset(const DevParams P, const StatePtrs S, const float *__restrict__ ball,
                        const float *__restrict__ blue, const float *__restrict__ yellow,
                        const uint8_t *__restrict__ mask) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= S.n) return;
    if (mask && !mask[e]) return;
    const float4 b = reinterpret_cast<const float4 *>(ball)[e];
    S.body[e] = b;
    for (int r = 0; r < P.n_robots; ++r) {
        const float *src = r < P.n_blue ? blue + ((size_t)e * P.n_blue + r) * 3
                                        : yellow + ((size_t)e * P.n_yellow + (r - P.n_blue)) * 3;
        float th = remainderf(src[2], 360.0f) * (1.0f / RS_DEG_F);
        if (th <= -RS_PI_F) th += 2.0f * RS_PI_F;
        S.body[(size_t)(r + 1) * S.np + e] = make_float4(src[0], src[1], 0.0f, 0.0f);
        S.ang[(size_t)r * S.np + e] = make_float2(th, 0.0f);
    }
    clear_task(P, S, e);
}

// rsim.py:19-24 dummy initial poses after bind
__global__ void k_init(const DevParams P, const StatePtrs S) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= S.n) return;
    S.body[e] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
    for (int r = 0; r < P.n_robots; ++r) {
        const float x = r < P.n_blue ? -0.2f * (float)(r + 1) : 0.2f * (float)(r - P.n_blue + 1);
        S.body[(size_t)(r + 1) * S.np + e] = make_float4(x, 0.0f, 0.0f, 0.0f);
        S.ang[(size_t)r * S.np + e] = make_float2(0.0f, 0.0f);
    }
}

__global__ void k_set_raw(const DevParams P, const StatePtrs S, const float *__restrict__ in) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= S.n) return;
    const float *r = in + (size_t)e * (4 + 6 * P.n_robots);
    S.body[e] = make_float4(r[0], r[1], r[2], r[3]);
    for (int k = 0; k < P.n_robots; ++k) {
        const float *q = r + 4 + 6 * k;
        S.body[(size_t)(k + 1) * S.np + e] = make_float4(q[0], q[1], q[3], q[4]);
        S.ang[(size_t)k * S.np + e] = make_float2(q[2], q[5]);
    }
}
__global__ void k_get_raw(const DevParams P, const StatePtrs S, float *__restrict__ out) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= S.n) return;
    float *r = out + (size_t)e * (4 + 6 * P.n_robots);
    const float4 b = S.body[e];
    r[0] = b.x; r[1] = b.y; r[2] = b.z; r[3] = b.w;
    for (int k = 0; k < P.n_robots; ++k) {
        const float4 q = S.body[(size_t)(k + 1) * S.np + e];
        const float2 a = S.ang[(size_t)k * S.np + e];
        float *o = r + 4 + 6 * k;
        o[0] = q.x; o[1] = q.y; o[2] = a.x; o[3] = q.z; o[4] = q.w; o[5] = a.y;
    }
}

// env.reset(): initial frame on device + first observation
template <int TASK>
__global__ void k_task_reset(const DevParams P, const StatePtrs S, const uint8_t *__restrict__ mask,
                             float *__restrict__ obs, int obs_dim, uint64_t seed,
                             const uint32_t *__restrict__ ctr, uint32_t env_offset) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= S.n) return;
    if (mask && !mask[e]) return;
    const uint32_t t = ctr[e / RS_CTR_GROUP] & RS_T_MASK;
    Scene<0> s;
    task_place<TASK, 0>(P, Rng(seed, env_offset + (uint32_t)e, t, RS_STREAM_RESET), s);
    store_scene<0>(P, S, e, s);
    clear_task(P, S, e);
    if (!obs) return;
    float *o = obs + (size_t)e * obs_dim;
    int k = 0;
    const int NB = P.n_blue, R = P.n_robots;
    if (TASK == RS_TASK_VSS) {
        o[k++] = nrm(s.bx, P.inv_max_pos); o[k++] = nrm(s.by, P.inv_max_pos);
        o[k++] = nrm(s.bvx, P.inv_max_v); o[k++] = nrm(s.bvy, P.inv_max_v);
        for (int r = 0; r < R; ++r) {
            o[k++] = nrm(s.x[r], P.inv_max_pos); o[k++] = nrm(s.y[r], P.inv_max_pos);
            if (r < NB) { float sn, cs; __sincosf(s.th[r], &sn, &cs); o[k++] = sn; o[k++] = cs; }
            o[k++] = nrm(s.vx[r], P.inv_max_v); o[k++] = nrm(s.vy[r], P.inv_max_v);
            o[k++] = nrm(s.om[r], P.inv_max_w_rad);
        }
    } else if (TASK == RS_TASK_SSL_DRIBBLING || TASK == RS_TASK_SSL_PASS_ENDURANCE) {
        ssl_hw_obs<TASK>(P, s, 0.0f, o);
    } else {
        const float inv_v = 1.0f / 2.5f, inv_w = RS_DEG_F / 10.0f;
        o[k++] = nrm(s.bx, P.inv_max_pos); o[k++] = nrm(s.by, P.inv_max_pos);
        o[k++] = nrm(s.bvx, inv_v); o[k++] = nrm(s.bvy, inv_v);
        for (int r = 0; r < NB; ++r) {
            float sn, cs;
            __sincosf(s.th[r], &sn, &cs);
            o[k++] = nrm(s.x[r], P.inv_max_pos); o[k++] = nrm(s.y[r], P.inv_max_pos);
            o[k++] = sn; o[k++] = cs;
            o[k++] = nrm(s.vx[r], inv_v); o[k++] = nrm(s.vy[r], inv_v); o[k++] = nrm(s.om[r], inv_w);
            o[k++] = touching(P, s.x[r], s.y[r], cs, sn, s.bx, s.by) ? 1.0f : 0.0f;
        }
        for (int r = NB; r < R; ++r) { o[k++] = nrm(s.x[r], P.inv_max_pos); o[k++] = nrm(s.y[r], P.inv_max_pos); }
    }
}

// launch floor probe (rs_debug_empty_step): the grid, CTA size, parameter block and launch attributes of a step
// kernel, no work.  chain == 0: waits for its predecessor like a serialised step; else only triggers its dependents.
__global__ void __launch_bounds__(64)
k_empty_step(const __grid_constant__ DevParams P, const StatePtrs S, const SslStepArgs A) {
    if (A.chain == 0) pdl_wait();
    pdl_release();
    if (S.n < 0) A.reward[0] = P.dt;          // never true: keeps the parameters alive
}

// ============================================================================ host side

static thread_local std::string g_err;
static int fail(int code, const std::string &msg) { g_err = msg; return code; }
#define CUDA_TRY(x)                                                                        \
    do {                                                                                   \
        cudaError_t _e = (x);                                                              \
        if (_e != cudaSuccess)                                                             \
            return fail(RS_E_CUDA, std::string(#x) + ": " + cudaGetErrorString(_e));       \
    } while (0)

struct rs_world {
    rs_params p;
    DevParams dp;
    int n, np, device;
    uint64_t seed;
    int64_t env_offset;
    uint64_t t;              // host mirror of the device step counter (exact unless launches were replayed from a graph)
    bool t_dirty;            // rs_set_t changed the host value; the next step launch writes it to the device
    uint32_t *d_ctr;         // device: t, one copy per RS_CTR_GROUP matches (library-owned)
    int n_ctr;
    mutable std::atomic<uint64_t> launches;   // counted by const getters too: no const_cast, no torn counts
    void *state;
    int64_t off[RS_ARR_COUNT];
    size_t state_bytes;
    int f0;                  // 1: the physics constants equal VssF0 (kernels with immediates); RS_NO_PRESET=1 forces 0
    int block;               // CTA size of the step kernels
    int per_match;           // 1: one lane per MATCH kernels (rs_device.cuh); 0: one lane per BODY (rs_lanes.cuh); -1: by world size
    int lane_block;          // CTA size of the lane-per-body kernels
    int packed;              // 1: packed fp32x2 instruction forms in the VssF0 kernels; 0: scalar forms; -1: by world size
    bool pdl;                // step kernels are launched with programmatic stream serialization (RS_PDL=0 turns it off)
    int host_copy_actions;   // RS_OPT_HOST_COPY_ACTIONS: 1 stage pinned host actions with a copy, 0 the kernel reads them in place, -1 by row size
    // step-to-step overlap (RS_OPT_STEP_OVERLAP; rs_device.cuh, tile_acquire)
    int overlap;             // 0 off, 1 state through tile flags + grid wait before caller buffers, 2 tile flags only, 3 = 2 + dense CTAs
    uint32_t *d_flags;       // error counter of the step-overlap protocol (library-owned; the tile locks live in d_ctr)
    bool chain_ok;           // the previous launch that touched the state was a flag-protocol step ...
    cudaStream_t chain_stream;   // ... on this stream ...
    int chain_family;        // ... of this kernel family (= tile mapping; a different family starts with a grid-wide wait)
    // scratch for the *_host entry points (library owned)
    float *s_actions, *s_obs, *s_reward;
    uint8_t *s_done, *s_trunc;
    int s_act_dim, s_obs_dim;
    const void *h_act_seen;  // last host action pointer looked up with cudaPointerGetAttributes ...
    const float *h_act_dev;  // ... and its device alias (null: pageable, staged with a copy)
    cudaEvent_t host_done;   // split-phase host step: recorded behind the D2H copies of rs_*_env_step_host_begin
    bool host_pending;       // ... and not waited for yet
};

// Every entry point that launches or allocates runs on the world's own device, whatever device is
// current in the calling thread, and leaves the caller's current device as it found it.
struct DeviceGuard {
    int prev;
    bool ok;
    explicit DeviceGuard(int dev) : prev(-1), ok(true) {
        int cur = -1;
        if (cudaGetDevice(&cur) != cudaSuccess) { ok = false; return; }
        if (cur != dev) { ok = cudaSetDevice(dev) == cudaSuccess; if (ok) prev = cur; }
    }
    ~DeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
};
#define ON_DEVICE(w, name)                                                                  \
    DeviceGuard _dg((w)->device);                                                           \
    if (!_dg.ok) return fail(RS_E_CUDA, name ": cannot select the world's CUDA device")

static void fill_dev_params(const rs_params &p, DevParams &d) {
    memset(&d, 0, sizeof(d));
    d.kind = p.kind; d.n_blue = p.n_blue; d.n_yellow = p.n_yellow; d.n_robots = p.n_robots;
    d.dt = (float)p.dt; d.h = (float)p.h;
    d.half_len = (float)(p.length / 2); d.half_wid = (float)(p.width / 2);
    d.goal_depth = (float)p.goal_depth; d.pen_len = (float)p.penalty_length;
    d.half_pen_wid = (float)(p.penalty_width / 2); d.half_goal_wid = (float)(p.goal_width / 2);
    d.ball_r = (float)p.ball_radius; d.rbt_r = (float)p.rbt_radius;
    d.rw = (float)p.rbt_wheel_radius; d.inv_rw = (float)(1.0 / p.rbt_wheel_radius);
    d.wmax = (float)p.wheel_max_rad_s;
    d.x_out = (float)p.x_out; d.y_out = (float)p.y_out; d.x_near = (float)p.x_near;
    d.n_box = p.n_box;
    for (int k = 0; k < RS_MAX_BOXES; ++k) for (int i = 0; i < 4; ++i) d.box[k][i] = (float)p.box[k][i];
    const double wb = 1.0 / p.ball_mass, wr = 1.0 / p.rbt_mass;
    d.wb = (float)wb; d.wr = (float)wr; d.inv_wsum = (float)(1.0 / (wb + wr));
    d.fb = (float)(wb / (wb + wr)); d.fr = (float)(wr / (wb + wr));
    d.e_ball_wall = (float)p.e_ball_wall; d.e_rbt_wall = (float)p.e_rbt_wall;
    d.e_ball_rbt = (float)p.e_ball_rbt; d.e_rbt_rbt = (float)p.e_rbt_rbt;
    d.mu_ball_rbt = (float)p.mu_ball_rbt;
    d.ball_decel_h = (float)(p.ball_decel * p.h);
    const double rs_br = p.rbt_radius + p.ball_radius, rs_rr = 2.0 * p.rbt_radius;
    d.rs_br = (float)rs_br; d.rs_br2 = (float)(rs_br * rs_br);
    d.rs_rr = (float)rs_rr; d.rs_rr2 = (float)(rs_rr * rs_rr);
    d.inv_2b = p.half_track > 0 ? (float)(1.0 / (2.0 * p.half_track)) : 0.0f;
    d.acc_fwd_h = (float)(p.acc_fwd * p.h); d.acc_lat_h = (float)(p.acc_lat * p.h);
    d.acc_ang_h = (float)(p.acc_ang * p.h);
    for (int i = 0; i < 4; ++i) for (int a = 0; a < 3; ++a) { d.J[i][a] = (float)p.omni_J[i][a]; d.Jp[a][i] = (float)p.omni_Jpinv[a][i]; }
    d.dk = (float)p.rbt_distance_center_kicker; d.kick_centre = (float)p.kick_centre;
    d.kick_reach = (float)p.kick_reach; d.kick_hw = (float)p.kick_half_width;
    d.mouth_hc = (float)p.mouth_half_chord; d.kick_max = (float)p.kick_speed_max;
    // vss_gym_base.py:52-58
    const double PI = 3.14159265358979323846;
    const double max_pos = p.width / 2 > p.length / 2 + p.penalty_length ? p.width / 2 : p.length / 2 + p.penalty_length;
    const double max_v = (p.rbt_motor_max_rpm / 60.0) * 2.0 * PI * p.rbt_wheel_radius;
    d.inv_max_pos = (float)(1.0 / max_pos); d.max_v = (float)max_v; d.inv_max_v = (float)(1.0 / max_v);
    d.inv_max_w_rad = (float)(0.04 / max_v);       // omega[rad/s] * DEG / (rad2deg(max_v / 0.04))
    d.sqrt_dt = (float)sqrt(p.dt);
}

// Do the compile-time constants of VssF0 (rs_device.cuh) equal this world's run-time block?
// Field by field, bit for bit: the kernels built on VssF0 are used only then.
static bool matches_vss_f0(const DevParams &d) {
    if (d.kind != RS_KIND_VSS || d.n_robots != VssF0::n_robots) return false;
    const float a[] = {d.h, d.ball_r, d.rbt_r, d.x_out, d.y_out, d.box[0][0], d.box[0][1], d.wb, d.wr, d.inv_wsum, d.fb, d.fr,
                       d.e_ball_wall, d.e_rbt_wall, d.e_ball_rbt, d.e_rbt_rbt, d.mu_ball_rbt, d.ball_decel_h,
                       d.rs_br, d.rs_br2, d.rs_rr, d.rs_rr2, d.acc_fwd_h, d.acc_lat_h, d.acc_ang_h};
    const float b[] = {VssF0::h, VssF0::ball_r, VssF0::rbt_r, VssF0::x_out, VssF0::y_out, VssF0::wall_lx, VssF0::wall_ly,
                       VssF0::wb, VssF0::wr, VssF0::inv_wsum, VssF0::fb, VssF0::fr,
                       VssF0::e_ball_wall, VssF0::e_rbt_wall, VssF0::e_ball_rbt, VssF0::e_rbt_rbt, VssF0::mu_ball_rbt,
                       VssF0::ball_decel_h, VssF0::rs_br, VssF0::rs_br2, VssF0::rs_rr, VssF0::rs_rr2,
                       VssF0::acc_fwd_h, VssF0::acc_lat_h, VssF0::acc_ang_h};
    static_assert(sizeof(a) == sizeof(b), "one run-time field per VssF0 constant");
    return memcmp(a, b, sizeof(a)) == 0;
}

static StatePtrs state_ptrs(const rs_world *w) {
    StatePtrs S;
    char *b = (char *)w->state;
    S.body = (float4 *)(b + w->off[RS_ARR_BODY]);
    S.ang = (float2 *)(b + w->off[RS_ARR_ANG]);
    S.ou = (float2 *)(b + w->off[RS_ARR_OU]);
    S.prev = (float *)(b + w->off[RS_ARR_PREV]);
    S.steps = (int *)(b + w->off[RS_ARR_STEPS]);
    S.info = (float *)(b + w->off[RS_ARR_INFO]);
    S.aux = (uint32_t *)(b + w->off[RS_ARR_PREV]);     // PREV, STEPS, INFO are back to back (rs_create)
    S.n = w->n; S.np = w->np;
    return S;
}


// Which mapping steps this world?  One lane per MATCH issues the fewest instructions but
// needs n / 32 warps of 128 registers to fill 592 SM sub-partitions; one lane per BODY
// issues ~2.6x the instructions in 8x the warps.  Measured on B200 with
// tools/step_timing.py, us per step, lane per body vs lane per match (round 1c: per-lane
// contact resolve through shared memory, warp-cooperative auto-reset, immediates):
//   VSS-v0 3 v 3        7.5 vs  8.0 @ 4 096     8.8 vs  8.2 @ 8 192    12.3 vs  8.6 @ 16 384    20.5 vs 12.1 @ 32 768   36.0 vs 18.3 @ 65 536
//   SSL 1 v 6 task      9.8 vs 13.3 @ 4 096    14.9 vs 14.2 @ 16 384    37.6 vs 30.6 @ 65 536
//   SSL 1 v 1 task      8.3 vs  6.5 @ 4 096    10.3 vs  7.0 @ 16 384    24.0 vs 11.9 @ 65 536
//   rs_step SSL 1 v 6 (all seven robots driven)  8.6 vs 24.8 @ 4 096    36.1 vs 57.4 @ 65 536
// Worlds with more than 7 robots have no register-resident lane-per-match kernel (its
// generic variant keeps the scene in local memory), so they always go lane per body.
//
// RS_OPT_STEP_OVERLAP = 3 (worlds stepped round-robin: consecutive launches overlap freely, the GPU is
// filled by several launches together) is a throughput regime: issued instructions decide, so every task
// world with a register-resident kernel runs one lane per match -- us per step, lane per body vs lane per
// match in rotation: VSS-v0 @ 4 096 2.63 vs 1.69, SSL 1 v 6 @ 4 096 2.40 vs 2.01, SSL 1 v 1 @ 16 384 5.31 vs
// 2.00 (profiles/r2_overlap.txt).
static bool use_lane_per_body(const rs_world *w, bool task_kernel) {
    if (w->per_match >= 0) return w->per_match == 0;
    const int R = w->p.n_robots;
    if (R > 7) return true;
    if (R <= 2) return false;
    if (task_kernel && w->overlap == 3) return false;
    if (w->p.kind == RS_KIND_VSS) return task_kernel ? w->n < 6000 : w->n < 20000;
    return task_kernel ? w->n < 14000 : true;
}

// Packed fp32x2 forms (VssF0P, rs_device.cuh) or scalar forms in the VssF0 lane-per-match kernels?  Same
// results bit for bit (tests/test_gpu_api.py); the packed forms save issue slots, which only large worlds are
// short of, and the pair table adds loads to the contact chain, so it is a matter of world size: 16 384 matches 8.21 (scalar) vs 8.30 us (packed), 24 576
// matches 9.90 vs 9.50 us, 65 536 matches 17.05 vs 15.71 us (profiles/r1g_packed.txt).  The whole world decides, not
// the chunk a launch covers.
#ifndef RS_PACKED_MIN_MATCHES
#define RS_PACKED_MIN_MATCHES 20480
#endif
static bool use_packed(const rs_world *w) {
    if (w->packed >= 0) return w->packed != 0;
    // RS_OPT_STEP_OVERLAP = 3 is the throughput regime whatever the world size (several worlds fill the GPU together:
    // issue slots decide), and only the packed kernel has the dense build
    if (w->overlap == 3) return true;
    return w->n >= RS_PACKED_MIN_MATCHES;
}

// Launch of a step kernel, by default with programmatic stream serialization (PDL): the
// kernel's pre-wait part overlaps the tail of its predecessor in the stream (rs_device.cuh).
// The first failing launch of a call is kept in the world-independent, per-thread g_launch_err
// and turned into RS_E_CUDA by the entry point (LAUNCH_CHECK).
static thread_local cudaError_t g_launch_err = cudaSuccess;
template <typename... KArgs, typename... Args>
static void launch_step_kernel(const rs_world *w, void (*kernel)(KArgs...), int grid, int block, cudaStream_t st, Args... args) {
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3((unsigned)grid); cfg.blockDim = dim3((unsigned)block); cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at; cfg.numAttrs = w->pdl ? 1u : 0u;
    const cudaError_t e = cudaLaunchKernelEx(&cfg, kernel, args...);
    if (e != cudaSuccess && g_launch_err == cudaSuccess) g_launch_err = e;
    w->launches++;
}
#define LAUNCH_CHECK(name)                                                                  \
    do {                                                                                    \
        cudaError_t _e = g_launch_err;                                                      \
        g_launch_err = cudaSuccess;                                                         \
        if (_e == cudaSuccess) _e = cudaGetLastError();                                     \
        if (_e != cudaSuccess) return fail(RS_E_CUDA, std::string(name ": kernel launch failed: ") + cudaGetErrorString(_e)); \
    } while (0)

template <int KIND, int RT>
static void launch_step(rs_world *w, const float *cmds, cudaStream_t st) {
    const int g64 = (w->n + 63) / 64, g128 = (w->n + 127) / 128;
    if constexpr (KIND == RS_KIND_VSS && RT == VssF0::n_robots) {
        if (w->f0) {            // rs_create compared the constants bit for bit (matches_vss_f0)
            if (use_packed(w)) launch_step_kernel(w, k_step<KIND, RT, 64, 2>, g64, 64, st, w->dp, state_ptrs(w), cmds);
            else launch_step_kernel(w, k_step<KIND, RT, 64, 1>, g64, 64, st, w->dp, state_ptrs(w), cmds);
            return;
        }
    }
    if (w->block == 128) launch_step_kernel(w, k_step<KIND, RT, 128>, g128, 128, st, w->dp, state_ptrs(w), cmds);
    else launch_step_kernel(w, k_step<KIND, RT, 64>, g64, 64, st, w->dp, state_ptrs(w), cmds);
}

// lane-per-body launches: grid = matches / (BS / L)
template <int KIND, int L>
static void launch_step_lanes_l(rs_world *w, const float *cmds, cudaStream_t st) {
    const StatePtrs S = state_ptrs(w);
    if (w->lane_block == 256) launch_step_kernel(w, k_step_lanes<KIND, L, 256>, (w->n + 256 / L - 1) / (256 / L), 256, st, w->dp, S, cmds);
    else if (w->lane_block == 64) launch_step_kernel(w, k_step_lanes<KIND, L, 64>, (w->n + 64 / L - 1) / (64 / L), 64, st, w->dp, S, cmds);
    else if (KIND == RS_KIND_VSS && L == 8 && w->f0) launch_step_kernel(w, k_step_lanes<RS_KIND_VSS, 8, 128, true>, (w->n + 15) / 16, 128, st, w->dp, S, cmds);
    else launch_step_kernel(w, k_step_lanes<KIND, L, 128>, (w->n + 128 / L - 1) / (128 / L), 128, st, w->dp, S, cmds);
}
template <int KIND>
static void launch_step_lanes(rs_world *w, const float *cmds, cudaStream_t st) {
    const int bodies = w->p.n_robots + 1;
    if (bodies <= 2) launch_step_lanes_l<KIND, 2>(w, cmds, st);
    else if (bodies <= 4) launch_step_lanes_l<KIND, 4>(w, cmds, st);
    else if (bodies <= 8) launch_step_lanes_l<KIND, 8>(w, cmds, st);
    else if (bodies <= 16) launch_step_lanes_l<KIND, 16>(w, cmds, st);
    else launch_step_lanes_l<KIND, 32>(w, cmds, st);
}

extern "C" {

int rs_version(void) { return 200; }
const char *rs_last_error(void) { return g_err.c_str(); }

int rs_create(int kind, int field_type, int n_blue, int n_yellow, int time_step_ms, int n_envs,
              int device, uint64_t seed, int64_t env_offset, rs_world **out) {
    if (!out) return fail(RS_E_INVALID, "rs_create: out is null");
    *out = nullptr;
    if (n_envs < 1) return fail(RS_E_INVALID, "rs_create: n_envs must be >= 1");
    if (env_offset < 0 || (uint64_t)env_offset + (uint64_t)n_envs > 0xFFFFFFFFull)
        return fail(RS_E_INVALID, "rs_create: global env ids must fit 32 bits");
    rs_params p;
    if (rs_params_fill(&p, kind, field_type, n_blue, n_yellow, time_step_ms) != 0)
        return fail(RS_E_INVALID, "rs_create: unknown world (kind/field_type) or bad robot counts");
    int count = 0;
    cudaError_t ce = cudaGetDeviceCount(&count);
    if (ce != cudaSuccess || count == 0)
        return fail(RS_E_CUDA, std::string("rs_create: no CUDA device (there is no CPU fallback): ") +
                                   cudaGetErrorString(ce));
    if (device < 0) CUDA_TRY(cudaGetDevice(&device));      // "the current device", resolved once, here
    if (device >= count) return fail(RS_E_INVALID, "rs_create: no such CUDA device");
    DeviceGuard dg(device);                                // the caller's current device is left alone
    if (!dg.ok) return fail(RS_E_CUDA, "rs_create: cannot select the CUDA device");
    rs_world *w = new rs_world();
    w->p = p; fill_dev_params(p, w->dp);
    w->n = n_envs; w->np = (n_envs + 127) / 128 * 128; w->device = device;
    w->seed = seed; w->env_offset = env_offset; w->t = 0;
    const int R = p.n_robots;
    const size_t np = (size_t)w->np;
    size_t o = 0;
    w->off[RS_ARR_BODY] = (int64_t)o; o += 16 * (size_t)(R + 1) * np;
    w->off[RS_ARR_ANG] = (int64_t)o; o += 8 * (size_t)R * np;
    w->off[RS_ARR_OU] = (int64_t)o; o += 8 * (size_t)(R > 1 ? R - 1 : 1) * np;
    w->off[RS_ARR_PREV] = (int64_t)o; o += 4 * np;
    w->off[RS_ARR_STEPS] = (int64_t)o; o += 4 * np;
    w->off[RS_ARR_INFO] = (int64_t)o; o += 4 * (size_t)RS_SSL_INFO * np;
    w->state_bytes = o;
    // the environment is read here, once per handle, never on a step
    w->block = 64;
    w->per_match = -1; w->lane_block = 128;
    if (const char *ls = getenv("RS_PER_MATCH")) w->per_match = atoi(ls) != 0;
    w->packed = -1;
    if (const char *ls = getenv("RS_PACKED")) w->packed = atoi(ls) != 0;
    w->pdl = true;
    if (const char *ls = getenv("RS_PDL")) w->pdl = atoi(ls) != 0;
    w->host_copy_actions = -1;
    if (const char *ls = getenv("RS_HOST_COPY_ACTIONS")) w->host_copy_actions = atoi(ls) != 0 ? 1 : 0;
    if (const char *bs = getenv("RS_LANE_BLOCK")) { const int b = atoi(bs); if (b == 64 || b == 128 || b == 256) w->lane_block = b; }
    w->f0 = matches_vss_f0(w->dp) ? 1 : 0;
    if (const char *nv = getenv("RS_NO_PRESET")) { if (atoi(nv) == 1) w->f0 = 0; }
    if (const char *bs = getenv("RS_BLOCK")) { const int b = atoi(bs); if (b == 32 || b == 64 || b == 128 || b == 256) w->block = b; }
    w->overlap = 0;
    if (const char *ov = getenv("RS_STEP_OVERLAP")) { const int v = atoi(ov); if (v >= 0 && v <= 3) w->overlap = v; }
    w->n_ctr = w->np / RS_CTR_GROUP;
    const size_t flag_bytes = 4 * sizeof(uint32_t);
    if (cudaMalloc(&w->d_ctr, w->n_ctr * sizeof(uint32_t)) != cudaSuccess || cudaMemset(w->d_ctr, 0, w->n_ctr * sizeof(uint32_t)) != cudaSuccess ||
        cudaMalloc(&w->d_flags, flag_bytes) != cudaSuccess || cudaMemset(w->d_flags, 0, flag_bytes) != cudaSuccess) {
        cudaFree(w->d_ctr); cudaFree(w->d_flags);
        cudaGetLastError();
        delete w;
        return fail(RS_E_CUDA, "rs_create: cudaMalloc of the step counter / tile flags failed");
    }
    *out = w;
    return RS_OK;
}

int rs_destroy(rs_world *w) {
    if (!w) return RS_OK;
    {
        DeviceGuard dg(w->device);
        cudaFree(w->s_actions); cudaFree(w->s_obs);      // s_reward / s_done / s_trunc live inside s_obs's allocation
        if (w->host_done) cudaEventDestroy(w->host_done);
        cudaFree(w->d_ctr); cudaFree(w->d_flags);
    }
    delete w;
    return RS_OK;
}

size_t rs_state_bytes(const rs_world *w) { return w ? w->state_bytes : 0; }

int rs_layout(const rs_world *w, int64_t *out_offsets, int64_t *out_np) {
    if (!w || !out_offsets || !out_np) return fail(RS_E_INVALID, "rs_layout: null argument");
    for (int i = 0; i < RS_ARR_COUNT; ++i) out_offsets[i] = w->off[i];
    *out_np = w->np;
    return RS_OK;
}

int rs_bind_state(rs_world *w, void *d_state, void *stream) {
    if (!w || !d_state) return fail(RS_E_INVALID, "rs_bind_state: null argument");
    if ((uintptr_t)d_state & 255u) return fail(RS_E_INVALID, "rs_bind_state: buffer must be 256-byte aligned");
    ON_DEVICE(w, "rs_bind_state");
    cudaStream_t st = (cudaStream_t)stream;
    w->state = d_state;
    w->chain_ok = false;
    CUDA_TRY(cudaMemsetAsync(d_state, 0, w->state_bytes, st));
    k_init<<<(w->n + 127) / 128, 128, 0, st>>>(w->dp, state_ptrs(w));
    w->launches++;
    CUDA_TRY(cudaGetLastError());
    return RS_OK;
}

int rs_field_params(const rs_world *w, double out[17]) {
    if (!w || !out) return fail(RS_E_INVALID, "rs_field_params: null argument");
    rs_params_field(&w->p, out);
    return RS_OK;
}

#define NEED_STATE(w, name)                                                                 \
    if (!(w)) return fail(RS_E_INVALID, name ": null world");                               \
    if (!(w)->state) return fail(RS_E_STATE, name ": no state buffer bound (rs_bind_state)"); \
    ON_DEVICE(w, name)

int rs_reset(rs_world *w, const float *d_ball, const float *d_blue, const float *d_yellow,
             const uint8_t *d_mask, void *stream) {
    NEED_STATE(w, "rs_reset");
    if (!d_ball || (w->p.n_blue && !d_blue) || (w->p.n_yellow && !d_yellow))
        return fail(RS_E_INVALID, "rs_reset: null placement array");
    w->chain_ok = false;
    k_reset<<<(w->n + 127) / 128, 128, 0, (cudaStream_t)stream>>>(w->dp, state_ptrs(w), d_ball, d_blue, d_yellow, d_mask);
    w->launches++;
    CUDA_TRY(cudaGetLastError());
    return RS_OK;
}

// robosim.step: pure physics, no random numbers -- the Philox step counter is not touched
int rs_step(rs_world *w, const float *d_cmds, void *stream) {
    NEED_STATE(w, "rs_step");
    if (!d_cmds) return fail(RS_E_INVALID, "rs_step: null commands");
    cudaStream_t st = (cudaStream_t)stream;
    const int R = w->p.n_robots;
    w->chain_ok = false;
    if (use_lane_per_body(w, false)) {
        if (w->p.kind == RS_KIND_VSS) launch_step_lanes<RS_KIND_VSS>(w, d_cmds, st);
        else launch_step_lanes<RS_KIND_SSL>(w, d_cmds, st);
    } else if (w->p.kind == RS_KIND_VSS) {
        if (R == 6) launch_step<RS_KIND_VSS, 6>(w, d_cmds, st);
        else if (R == 10) launch_step<RS_KIND_VSS, 10>(w, d_cmds, st);
        else if (R == 3) launch_step<RS_KIND_VSS, 3>(w, d_cmds, st);
        else if (R == 2) launch_step<RS_KIND_VSS, 2>(w, d_cmds, st);
        else launch_step<RS_KIND_VSS, 0>(w, d_cmds, st);
    } else {
        if (R == 7) launch_step<RS_KIND_SSL, 7>(w, d_cmds, st);
        else if (R == 2) launch_step<RS_KIND_SSL, 2>(w, d_cmds, st);
        else if (R == 1) launch_step<RS_KIND_SSL, 1>(w, d_cmds, st);
        else launch_step<RS_KIND_SSL, 0>(w, d_cmds, st);
    }
    LAUNCH_CHECK("rs_step");
    return RS_OK;
}

int rs_get_state(const rs_world *w, float *d_out, void *stream) {
    NEED_STATE(w, "rs_get_state");
    if (!d_out) return fail(RS_E_INVALID, "rs_get_state: null output");
    k_get_state<<<(w->n + 127) / 128, 128, 0, (cudaStream_t)stream>>>(w->dp, state_ptrs(w), d_out);
    w->launches++;
    CUDA_TRY(cudaGetLastError());
    return RS_OK;
}

int rs_set_raw(rs_world *w, const float *d_in, void *stream) {
    NEED_STATE(w, "rs_set_raw");
    if (!d_in) return fail(RS_E_INVALID, "rs_set_raw: null input");
    w->chain_ok = false;
    k_set_raw<<<(w->n + 127) / 128, 128, 0, (cudaStream_t)stream>>>(w->dp, state_ptrs(w), d_in);
    w->launches++;
    CUDA_TRY(cudaGetLastError());
    return RS_OK;
}
int rs_get_raw(const rs_world *w, float *d_out, void *stream) {
    NEED_STATE(w, "rs_get_raw");
    if (!d_out) return fail(RS_E_INVALID, "rs_get_raw: null output");
    k_get_raw<<<(w->n + 127) / 128, 128, 0, (cudaStream_t)stream>>>(w->dp, state_ptrs(w), d_out);
    w->launches++;
    CUDA_TRY(cudaGetLastError());
    return RS_OK;
}

// The step counter t (Philox counter word 1, 32 bits on the device) is DEVICE-authoritative:
// every task step advances it on the device (so captured graphs replay with fresh noise) and
// the host keeps a mirror that is exact as long as no launch was replayed from a graph.  Only
// rs_set_t writes host -> device (at the next step launch), only rs_sync_t reads device -> host.
uint64_t rs_get_t(const rs_world *w) { return w ? w->t : 0; }
int rs_set_t(rs_world *w, uint64_t t) {
    if (!w) return fail(RS_E_INVALID, "rs_set_t: null world");
    if (t > 0x7FFFFFFFull) return fail(RS_E_INVALID, "rs_set_t: the step counter has 31 bits (bit 31 of its device copies is the tile lock)");
    w->t = t; w->t_dirty = true;
    return RS_OK;
}
// applies a pending rs_set_t.  Never inside a stream capture: the write would be baked into the
// graph and every replay would rewind the counter (and repeat the noise).
static int push_t(rs_world *w, cudaStream_t st) {
    if (!w->t_dirty) return RS_OK;
    cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
    if (cudaStreamIsCapturing(st, &cs) != cudaSuccess) { cudaGetLastError(); cs = cudaStreamCaptureStatusNone; }
    if (cs != cudaStreamCaptureStatusNone)
        return fail(RS_E_STATE, "rs_set_t is pending and the stream is capturing: run one step (or rs_sync_t) outside the capture first");
    k_set_ctr<<<(w->n_ctr + 255) / 256, 256, 0, st>>>(w->d_ctr, w->n_ctr, (uint32_t)w->t);
    w->launches++; w->t_dirty = false; w->chain_ok = false;
    return RS_OK;
}
int rs_sync_t(rs_world *w, void *stream) {
    if (!w) return fail(RS_E_INVALID, "rs_sync_t: null world");
    ON_DEVICE(w, "rs_sync_t");
    cudaStream_t st = (cudaStream_t)stream;
    if (w->t_dirty) {                  // the host value is the newer one: write it, then both agree
        const int rc = push_t(w, st);
        if (rc) return rc;
        CUDA_TRY(cudaStreamSynchronize(st));
        return RS_OK;
    }
    uint32_t t = 0;
    CUDA_TRY(cudaMemcpyAsync(&t, w->d_ctr, sizeof(t), cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    w->t = t & 0x7FFFFFFFu;
    return RS_OK;
}
uint64_t rs_launch_count(const rs_world *w) { return w ? w->launches.load() : 0; }

int rs_debug_empty_step(rs_world *w, int chain, void *stream) {
    NEED_STATE(w, "rs_debug_empty_step");
    SslStepArgs A;
    memset(&A, 0, sizeof(A));
    A.chain = chain;
    launch_step_kernel(w, k_empty_step, (w->n + 63) / 64, 64, (cudaStream_t)stream, w->dp, state_ptrs(w), A);
    LAUNCH_CHECK("rs_debug_empty_step");
    return RS_OK;
}

int rs_kernel_flags(const rs_world *w) {
    if (!w) return 0;
    return (use_lane_per_body(w, true) ? 1 : 0) | (use_lane_per_body(w, false) ? 2 : 0) | (w->f0 ? 4 : 0) |
           (w->f0 && !use_lane_per_body(w, true) && use_packed(w) ? 8 : 0);
}

int rs_set_option(rs_world *w, int option, int64_t value) {
    if (!w) return fail(RS_E_INVALID, "rs_set_option: null world");
    switch (option) {
        case RS_OPT_STEP_OVERLAP:
            if (value < 0 || value > 3) return fail(RS_E_INVALID, "rs_set_option: RS_OPT_STEP_OVERLAP takes 0, 1, 2 or 3");
            w->overlap = (int)value; w->chain_ok = false;     // also: "the state was written behind the library's back"
            return RS_OK;
        case RS_OPT_PDL:
            w->pdl = value != 0; w->chain_ok = false;
            return RS_OK;
        case RS_OPT_HOST_COPY_ACTIONS:
            if (value < -1 || value > 1) return fail(RS_E_INVALID, "rs_set_option: RS_OPT_HOST_COPY_ACTIONS takes -1, 0 or 1");
            w->host_copy_actions = (int)value; w->h_act_seen = nullptr;
            return RS_OK;
        default:
            return fail(RS_E_INVALID, "rs_set_option: unknown or read-only option");
    }
}
int rs_get_option(const rs_world *w, int option, int64_t *value, void *stream) {
    if (!w || !value) return fail(RS_E_INVALID, "rs_get_option: null argument");
    switch (option) {
        case RS_OPT_STEP_OVERLAP: *value = w->overlap; return RS_OK;
        case RS_OPT_PDL: *value = w->pdl ? 1 : 0; return RS_OK;
        case RS_OPT_HOST_COPY_ACTIONS: *value = w->host_copy_actions; return RS_OK;
        case RS_OPT_OVERLAP_ERRORS: {
            ON_DEVICE(w, "rs_get_option");
            uint32_t e = 0;
            CUDA_TRY(cudaMemcpyAsync(&e, w->d_flags, sizeof(e), cudaMemcpyDeviceToHost, (cudaStream_t)stream));
            CUDA_TRY(cudaStreamSynchronize((cudaStream_t)stream));
            *value = e;
            return RS_OK;
        }
        default:
            return fail(RS_E_INVALID, "rs_get_option: unknown option");
    }
}

static bool task_matches(const rs_world *w, int task) {
    const rs_params &p = w->p;
    if (task == RS_TASK_VSS_V0) return p.kind == RS_KIND_VSS && p.n_blue == 3 && p.n_yellow == 3;
    if (task == RS_TASK_SSL_STATIC_DEFENDERS_V0) return p.kind == RS_KIND_SSL && p.n_blue == 1 && p.n_yellow == 6;
    if (task == RS_TASK_SSL_CONTESTED_POSSESSION_V0) return p.kind == RS_KIND_SSL && p.n_blue == 1 && p.n_yellow == 1;
    if (task == RS_TASK_SSL_DRIBBLING_V0) return p.kind == RS_KIND_SSL && p.n_blue == 1 && p.n_yellow == 4;
    if (task == RS_TASK_SSL_PASS_ENDURANCE_V0) return p.kind == RS_KIND_SSL && p.n_blue == 2 && p.n_yellow == 0;
    return false;
}

int rs_task_obs_dim(const rs_world *w, int task) {
    if (!w) return fail(RS_E_INVALID, "rs_task_obs_dim: null world");
    if (task == RS_TASK_VSS_V0) return 4 + 7 * w->p.n_blue + 5 * w->p.n_yellow;
    if (task == RS_TASK_SSL_STATIC_DEFENDERS_V0 || task == RS_TASK_SSL_CONTESTED_POSSESSION_V0)
        return 4 + 8 * w->p.n_blue + 2 * w->p.n_yellow;
    if (task == RS_TASK_SSL_DRIBBLING_V0) return 5 + 8 * w->p.n_blue + 2 * w->p.n_yellow;     // dribbling.py:53
    if (task == RS_TASK_SSL_PASS_ENDURANCE_V0) return 4 + 6 * w->p.n_blue;                     // pass_endurance.py:55
    return fail(RS_E_INVALID, "rs_task_obs_dim: unknown task");
}

int rs_task_act_dim(int task) {
    switch (task) {
        case RS_TASK_VSS_V0: return RS_VSS_ACT;
        case RS_TASK_SSL_STATIC_DEFENDERS_V0: case RS_TASK_SSL_CONTESTED_POSSESSION_V0: return RS_SSL_ACT;
        case RS_TASK_SSL_DRIBBLING_V0: return RS_DRIB_ACT;
        case RS_TASK_SSL_PASS_ENDURANCE_V0: return RS_PASS_ACT;
        default: return fail(RS_E_INVALID, "rs_task_act_dim: unknown task");
    }
}

int rs_task_reset(rs_world *w, int task, const uint8_t *d_mask, float *d_obs, void *stream) {
    NEED_STATE(w, "rs_task_reset");
    if (!task_matches(w, task)) return fail(RS_E_UNSUPPORTED, "rs_task_reset: task does not match this world");
    cudaStream_t st = (cudaStream_t)stream;
    const int g = (w->n + 127) / 128, od = rs_task_obs_dim(w, task);
    const uint32_t off = (uint32_t)w->env_offset;
    const int rc = push_t(w, st);
    if (rc) return rc;
    w->chain_ok = false;
    const uint32_t *t = w->d_ctr;
    if (task == RS_TASK_VSS_V0) k_task_reset<RS_TASK_VSS><<<g, 128, 0, st>>>(w->dp, state_ptrs(w), d_mask, d_obs, od, w->seed, t, off);
    else if (task == RS_TASK_SSL_STATIC_DEFENDERS_V0) k_task_reset<RS_TASK_SSL_STATIC_DEFENDERS><<<g, 128, 0, st>>>(w->dp, state_ptrs(w), d_mask, d_obs, od, w->seed, t, off);
    else if (task == RS_TASK_SSL_DRIBBLING_V0) k_task_reset<RS_TASK_SSL_DRIBBLING><<<g, 128, 0, st>>>(w->dp, state_ptrs(w), d_mask, d_obs, od, w->seed, t, off);
    else if (task == RS_TASK_SSL_PASS_ENDURANCE_V0) k_task_reset<RS_TASK_SSL_PASS_ENDURANCE><<<g, 128, 0, st>>>(w->dp, state_ptrs(w), d_mask, d_obs, od, w->seed, t, off);
    else k_task_reset<RS_TASK_SSL_CONTESTED_POSSESSION><<<g, 128, 0, st>>>(w->dp, state_ptrs(w), d_mask, d_obs, od, w->seed, t, off);
    w->launches++;
    CUDA_TRY(cudaGetLastError());
    return RS_OK;
}

// Step-to-step overlap, host side: a launch takes part in the tile-flag protocol whenever the option is on,
// and may skip the grid-wide wait (chain != 0) when the launch before it -- same world, same stream, same
// kernel family, i.e. same tile mapping -- took part too.  Mode 1 (grid-wide wait before the caller's buffers
// are read) exists in the lane-per-match VSS-v0 kernel only; every other kernel treats it as 0.
static void chain_setup(rs_world *w, int family, bool has_mode1, cudaStream_t st, uint32_t *&flags, int &chain) {
    flags = nullptr; chain = 0;
    if (w->overlap) {
        flags = w->d_flags;
        if (w->pdl && w->chain_ok && w->chain_stream == st && w->chain_family == family && (w->overlap != 1 || has_mode1))
            chain = w->overlap;
    }
    w->chain_ok = w->overlap != 0; w->chain_stream = st; w->chain_family = family;
}

// One launch of VSSEnv.step over the S.n matches S / A point at.
static void launch_vss(rs_world *w, VssStepArgs &A, const StatePtrs &S, cudaStream_t st) {
    const int n = S.n;
    if (use_lane_per_body(w, true)) {
        chain_setup(w, 100 + w->lane_block, false, st, A.flags, A.chain);
        if (w->lane_block == 256) launch_step_kernel(w, k_vss_env_step_lanes<256>, (n + 31) / 32, 256, st, w->dp, S, A);
        else if (w->lane_block == 64) launch_step_kernel(w, k_vss_env_step_lanes<64>, (n + 7) / 8, 64, st, w->dp, S, A);
        else if (w->f0) launch_step_kernel(w, k_vss_env_step_lanes<128, true>, (n + 15) / 16, 128, st, w->dp, S, A);
        else launch_step_kernel(w, k_vss_env_step_lanes<128>, (n + 15) / 16, 128, st, w->dp, S, A);
        return;
    }
    chain_setup(w, 1, true, st, A.flags, A.chain);          // every lane-per-match build has the same tiles: 32 matches
    if (w->f0 && use_packed(w)) switch (w->block) {
        case 32: launch_step_kernel(w, k_vss_env_step<3, 3, 32, 2>, (n + 31) / 32, 32, st, w->dp, S, A); break;
        case 128: launch_step_kernel(w, k_vss_env_step<3, 3, 128, 2>, (n + 127) / 128, 128, st, w->dp, S, A); break;
        default:
            if (A.chain == 3) launch_step_kernel(w, k_vss_env_step<3, 3, 64, 2, true>, (n + 63) / 64, 64, st, w->dp, S, A);
            else launch_step_kernel(w, k_vss_env_step<3, 3, 64, 2>, (n + 63) / 64, 64, st, w->dp, S, A);
            break;
    } else if (w->f0) {
        launch_step_kernel(w, k_vss_env_step<3, 3, 64, 1>, (n + 63) / 64, 64, st, w->dp, S, A);
    } else {
        launch_step_kernel(w, k_vss_env_step<3, 3, 64, 0>, (n + 63) / 64, 64, st, w->dp, S, A);
    }
}

int rs_vss_env_step(rs_world *w, const float *d_actions, const float *d_normals, int auto_reset,
                    int max_steps, float *d_obs, float *d_reward, uint8_t *d_done,
                    uint8_t *d_trunc, float *d_cmds_out, void *stream) {
    NEED_STATE(w, "rs_vss_env_step");
    if (!task_matches(w, RS_TASK_VSS_V0))
        return fail(RS_E_UNSUPPORTED, "rs_vss_env_step: world is not VSS 3v3");
    if (!d_actions || !d_obs || !d_reward || !d_done || !d_trunc)
        return fail(RS_E_INVALID, "rs_vss_env_step: null argument");
    if (max_steps < 1 || max_steps > 0xFFFFFF) return fail(RS_E_INVALID, "rs_vss_env_step: bad max_steps");
    if (((uintptr_t)d_obs & 15u) || ((uintptr_t)d_actions & 7u)) return fail(RS_E_INVALID, "rs_vss_env_step: obs must be 16-byte and actions 8-byte aligned");
    VssStepArgs A;
    A.actions = reinterpret_cast<const float2 *>(d_actions); A.normals = d_normals;
    A.obs = d_obs; A.reward = d_reward; A.done = d_done; A.trunc = d_trunc; A.cmds_out = d_cmds_out;
    A.auto_reset = auto_reset; A.max_steps = max_steps;
    cudaStream_t st = (cudaStream_t)stream;
    const int rc = push_t(w, st);
    if (rc) return rc;
    A.seed = w->seed; A.ctr = w->d_ctr; A.env_offset = (uint32_t)w->env_offset;
    launch_vss(w, A, state_ptrs(w), st);
    w->t = (w->t + 1) & 0x7FFFFFFFull;
    LAUNCH_CHECK("rs_vss_env_step");
    return RS_OK;
}

int rs_ssl_env_step(rs_world *w, int task, const float *d_actions, int auto_reset, int max_steps,
                    float *d_obs, float *d_reward, uint8_t *d_done, uint8_t *d_trunc,
                    float *d_cmds_out, void *stream) {
    NEED_STATE(w, "rs_ssl_env_step");
    if (task < RS_TASK_SSL_STATIC_DEFENDERS_V0 || task > RS_TASK_SSL_PASS_ENDURANCE_V0 || !task_matches(w, task))
        return fail(RS_E_UNSUPPORTED, "rs_ssl_env_step: task does not match this world");
    if (!d_actions || !d_obs || !d_reward || !d_done || !d_trunc)
        return fail(RS_E_INVALID, "rs_ssl_env_step: null argument");
    if (max_steps < 1 || max_steps > 0xFFFFFF) return fail(RS_E_INVALID, "rs_ssl_env_step: bad max_steps");
    SslStepArgs A;
    A.actions = d_actions; A.obs = d_obs; A.reward = d_reward; A.done = d_done; A.trunc = d_trunc;
    A.cmds_out = d_cmds_out; A.auto_reset = auto_reset; A.max_steps = max_steps;
    cudaStream_t st = (cudaStream_t)stream;
    const int rc = push_t(w, st);
    if (rc) return rc;
    A.seed = w->seed; A.ctr = w->d_ctr; A.env_offset = (uint32_t)w->env_offset;
    const StatePtrs S = state_ptrs(w);
    const int g64 = (w->n + 63) / 64, g128 = (w->n + 127) / 128;
    // kernel family = tile mapping: lane per match (32 matches per warp) or lane per body (by CTA size: the
    // lanes per match are fixed by the world)
    chain_setup(w, (task == RS_TASK_SSL_DRIBBLING_V0 || task == RS_TASK_SSL_PASS_ENDURANCE_V0 || !use_lane_per_body(w, true))
                       ? 1 : 100 + w->lane_block, false, st, A.flags, A.chain);
    if (task == RS_TASK_SSL_DRIBBLING_V0) {
        launch_step_kernel(w, k_ssl_hw_env_step<RS_TASK_SSL_DRIBBLING, 1, 4, 64>, g64, 64, st, w->dp, S, A);
    } else if (task == RS_TASK_SSL_PASS_ENDURANCE_V0) {
        launch_step_kernel(w, k_ssl_hw_env_step<RS_TASK_SSL_PASS_ENDURANCE, 2, 0, 64>, g64, 64, st, w->dp, S, A);
    } else if (use_lane_per_body(w, true)) {
        // lane per body: 8 lanes per 1 v 6 match, 4 per 1 v 1 match
        const int bs = w->lane_block;
        if (task == RS_TASK_SSL_STATIC_DEFENDERS_V0) {
            if (bs == 256) launch_step_kernel(w, k_ssl_env_step_lanes<RS_TASK_SSL_STATIC_DEFENDERS, 1, 6, 8, 256>, (w->n + 31) / 32, 256, st, w->dp, S, A);
            else if (bs == 64) launch_step_kernel(w, k_ssl_env_step_lanes<RS_TASK_SSL_STATIC_DEFENDERS, 1, 6, 8, 64>, (w->n + 7) / 8, 64, st, w->dp, S, A);
            else launch_step_kernel(w, k_ssl_env_step_lanes<RS_TASK_SSL_STATIC_DEFENDERS, 1, 6, 8, 128>, (w->n + 15) / 16, 128, st, w->dp, S, A);
        } else {
            if (bs == 256) launch_step_kernel(w, k_ssl_env_step_lanes<RS_TASK_SSL_CONTESTED_POSSESSION, 1, 1, 4, 256>, (w->n + 63) / 64, 256, st, w->dp, S, A);
            else if (bs == 64) launch_step_kernel(w, k_ssl_env_step_lanes<RS_TASK_SSL_CONTESTED_POSSESSION, 1, 1, 4, 64>, (w->n + 15) / 16, 64, st, w->dp, S, A);
            else launch_step_kernel(w, k_ssl_env_step_lanes<RS_TASK_SSL_CONTESTED_POSSESSION, 1, 1, 4, 128>, (w->n + 31) / 32, 128, st, w->dp, S, A);
        }
    } else if (task == RS_TASK_SSL_STATIC_DEFENDERS_V0) {
        if (w->block == 128) launch_step_kernel(w, k_ssl_env_step<RS_TASK_SSL_STATIC_DEFENDERS, 1, 6, 128>, g128, 128, st, w->dp, S, A);
        else launch_step_kernel(w, k_ssl_env_step<RS_TASK_SSL_STATIC_DEFENDERS, 1, 6, 64>, g64, 64, st, w->dp, S, A);
    } else {
        if (w->block == 128) launch_step_kernel(w, k_ssl_env_step<RS_TASK_SSL_CONTESTED_POSSESSION, 1, 1, 128>, g128, 128, st, w->dp, S, A);
        else launch_step_kernel(w, k_ssl_env_step<RS_TASK_SSL_CONTESTED_POSSESSION, 1, 1, 64>, g64, 64, st, w->dp, S, A);
    }
    w->t = (w->t + 1) & 0x7FFFFFFFull;
    LAUNCH_CHECK("rs_ssl_env_step");
    return RS_OK;
}

// Device staging of the *_host entry points.  The four outputs are ONE allocation laid out
// [obs | reward | done | trunc]: a caller whose host buffers are laid out the same way (one
// pinned block, BatchedWorld.alloc_host_outputs) gets a single device-to-host copy instead of
// four (each copy costs ~8 us of fixed latency next to a 20 us kernel).
static int ensure_scratch(rs_world *w, int act_dim, int obs_dim) {
    if (w->s_actions && w->s_act_dim == act_dim && w->s_obs_dim == obs_dim) return RS_OK;
    cudaFree(w->s_actions); cudaFree(w->s_obs);
    w->s_actions = w->s_obs = w->s_reward = nullptr; w->s_done = w->s_trunc = nullptr;
    const size_t n = (size_t)w->n;
    CUDA_TRY(cudaMalloc(&w->s_actions, sizeof(float) * n * act_dim));
    CUDA_TRY(cudaMalloc(&w->s_obs, sizeof(float) * n * obs_dim + sizeof(float) * n + 2 * n));
    w->s_reward = w->s_obs + n * obs_dim;
    w->s_done = reinterpret_cast<uint8_t *>(w->s_reward + n);
    w->s_trunc = w->s_done + n;
    w->s_act_dim = act_dim; w->s_obs_dim = obs_dim;
    return RS_OK;
}

// The actions of a host-buffer step as the kernel will read them.  Pinned host memory is mapped
// into the device address space (UVA), so a kernel can load the actions straight over PCIe while its
// state loads are in flight, which saves the separate H2D copy and its latency -- IF the rows are read
// coalesced.  VSS-v0 (one float2 per match, 256 contiguous bytes per warp): 248 -> 237 us per 65 536-match
// step, 44.6 -> 42.7 us at 4 096.  The SSL tasks read five scalars at a 20-byte stride, i.e. every line
// crosses the link five times: in place 54.9 / 79.9 / 282 us against 41.6 / 59.3 / 190 us with the copy
// (SSLStaticDefenders-v0 at 4 096, SSLContestedPossession-v0 at 16 384, SSLStaticDefenders-v0 at 65 536;
// profiles/r2_e2e_probe.txt).  So by default (-1) only 8-byte rows are read in place; RS_OPT_HOST_COPY_ACTIONS
// forces either.  Pageable memory is always staged with a copy.
// The pointer lookup is cached per handle: a rollout loop passes the same buffer every step.
static const float *host_actions_on_device(rs_world *w, const float *h_actions, size_t bytes, cudaStream_t st) {
    const bool copy = w->host_copy_actions >= 0 ? w->host_copy_actions != 0 : bytes != 8u * (size_t)w->n;
    if (h_actions != w->h_act_seen) {
        cudaPointerAttributes at;
        w->h_act_seen = h_actions; w->h_act_dev = nullptr;
        if (!copy && cudaPointerGetAttributes(&at, h_actions) == cudaSuccess &&
            at.type == cudaMemoryTypeHost && at.devicePointer)
            w->h_act_dev = static_cast<const float *>(at.devicePointer);
        cudaGetLastError();
    }
    if (w->h_act_dev) return w->h_act_dev;
    if (cudaMemcpyAsync(w->s_actions, h_actions, bytes, cudaMemcpyHostToDevice, st) != cudaSuccess) {
        fail(RS_E_CUDA, std::string("host step: H2D copy of the actions failed: ") + cudaGetErrorString(cudaGetLastError()));
        return nullptr;
    }
    return w->s_actions;
}

// D2H of the four outputs: one copy when the caller's buffers are the packed block, else four
static int host_copy_out(rs_world *w, int obs_dim, float *h_obs, float *h_reward, uint8_t *h_done,
                         uint8_t *h_trunc, cudaStream_t st) {
    const size_t n = (size_t)w->n, ob = sizeof(float) * n * obs_dim;
    if (reinterpret_cast<char *>(h_reward) == reinterpret_cast<char *>(h_obs) + ob &&
        h_done == reinterpret_cast<uint8_t *>(h_reward + n) && h_trunc == h_done + n) {
        CUDA_TRY(cudaMemcpyAsync(h_obs, w->s_obs, ob + sizeof(float) * n + 2 * n, cudaMemcpyDeviceToHost, st));
    } else {
        CUDA_TRY(cudaMemcpyAsync(h_obs, w->s_obs, ob, cudaMemcpyDeviceToHost, st));
        CUDA_TRY(cudaMemcpyAsync(h_reward, w->s_reward, sizeof(float) * n, cudaMemcpyDeviceToHost, st));
        CUDA_TRY(cudaMemcpyAsync(h_done, w->s_done, n, cudaMemcpyDeviceToHost, st));
        CUDA_TRY(cudaMemcpyAsync(h_trunc, w->s_trunc, n, cudaMemcpyDeviceToHost, st));
    }
    return RS_OK;
}
// blocking form: the call returns when the outputs have landed.  split-phase form (begin = true): an event is
// recorded behind the copies and rs_host_step_wait blocks on it -- the caller overlaps this world's transfer with
// the step of ANOTHER world (double-buffered env groups) or with its own host work.
static int host_epilogue(rs_world *w, int obs_dim, float *h_obs, float *h_reward, uint8_t *h_done,
                         uint8_t *h_trunc, cudaStream_t st, bool begin) {
    const int rc = host_copy_out(w, obs_dim, h_obs, h_reward, h_done, h_trunc, st);
    if (rc) return rc;
    if (!begin) { CUDA_TRY(cudaStreamSynchronize(st)); return RS_OK; }
    if (!w->host_done) CUDA_TRY(cudaEventCreateWithFlags(&w->host_done, cudaEventDisableTiming));
    CUDA_TRY(cudaEventRecord(w->host_done, st));
    w->host_pending = true;
    return RS_OK;
}

// Host-buffer VSSEnv.step: the actions (read in place when pinned, else copied), the fused step,
// ONE D2H of the packed outputs, a stream synchronize.  The D2H of the observations (10.5 MB at 65 536 matches, ~200 us at the
// 54 GB/s a pinned copy reaches here) is 80 % of the call; two alternatives were measured on
// B200 and are slower than this plain form (251 us per step): four sub-range launches pipelined
// against four smaller copies on a second stream (284 us: the smaller copies lose more than
// the 17 us kernel hides), the kernel storing straight into the mapped host buffers (267 us), and (round 1g)
// the packed copy as two halves on two streams / copy engines (239 vs 234 us: one link, one more launch).
// Round 2 re-measured the pipelined form with UNEQUAL sub-ranges (1/4 + 3/4, 1/8 + 3/8 + 1/2: a small first launch
// so that the link starts early, early chunks copied on a second stream behind events): 240 / 244 us against 237 us
// -- each cross-stream event edge costs the few microseconds the earlier start gains -- and 54 vs 43 us at 4 096
// matches (profiles/r2_e2e_probe.txt).  The call runs at 0.86 of the plain copy of its outputs; what is left is one
// launch from idle, a 13 us kernel and the copy's start-up.
static int vss_env_step_host(rs_world *w, const float *h_actions, int auto_reset, int max_steps,
                             float *h_obs, float *h_reward, uint8_t *h_done, uint8_t *h_trunc,
                             void *stream, bool begin) {
    NEED_STATE(w, "rs_vss_env_step_host");
    if (w->host_pending) return fail(RS_E_STATE, "rs_vss_env_step_host: a split-phase host step of this world is pending (rs_host_step_wait first)");
    if (!h_actions || !h_obs || !h_reward || !h_done || !h_trunc)
        return fail(RS_E_INVALID, "rs_vss_env_step_host: null argument");
    const int od = rs_task_obs_dim(w, RS_TASK_VSS_V0);
    int rc = ensure_scratch(w, RS_VSS_ACT, od);
    if (rc) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    const float *d_act = host_actions_on_device(w, h_actions, sizeof(float) * (size_t)w->n * RS_VSS_ACT, st);
    if (!d_act) return RS_E_CUDA;
    // a host step ends with a synchronize: its launch starts on an idle GPU and gains nothing from the tile protocol,
    // so it is launched as if RS_OPT_STEP_OVERLAP were 0 (mode 3 would pick the dense build, 80 registers for nothing
    // here, and the lane-per-match kernels for small worlds)
    const int ov = w->overlap;
    w->overlap = 0;
    rc = rs_vss_env_step(w, d_act, nullptr, auto_reset, max_steps, w->s_obs, w->s_reward, w->s_done, w->s_trunc, nullptr, stream);
    w->overlap = ov;
    if (rc) return rc;
    return host_epilogue(w, od, h_obs, h_reward, h_done, h_trunc, st, begin);
}
int rs_vss_env_step_host(rs_world *w, const float *h_actions, int auto_reset, int max_steps,
                         float *h_obs, float *h_reward, uint8_t *h_done, uint8_t *h_trunc, void *stream) {
    return vss_env_step_host(w, h_actions, auto_reset, max_steps, h_obs, h_reward, h_done, h_trunc, stream, false);
}
int rs_vss_env_step_host_begin(rs_world *w, const float *h_actions, int auto_reset, int max_steps,
                               float *h_obs, float *h_reward, uint8_t *h_done, uint8_t *h_trunc, void *stream) {
    return vss_env_step_host(w, h_actions, auto_reset, max_steps, h_obs, h_reward, h_done, h_trunc, stream, true);
}

static int ssl_env_step_host(rs_world *w, int task, const float *h_actions, int auto_reset,
                             int max_steps, float *h_obs, float *h_reward, uint8_t *h_done,
                             uint8_t *h_trunc, void *stream, bool begin) {
    NEED_STATE(w, "rs_ssl_env_step_host");
    if (w->host_pending) return fail(RS_E_STATE, "rs_ssl_env_step_host: a split-phase host step of this world is pending (rs_host_step_wait first)");
    if (!h_actions || !h_obs || !h_reward || !h_done || !h_trunc)
        return fail(RS_E_INVALID, "rs_ssl_env_step_host: null argument");
    const int od = rs_task_obs_dim(w, task);
    if (od < 0) return od;
    const int ad = rs_task_act_dim(task);
    if (ad < 0) return ad;
    int rc = ensure_scratch(w, ad, od);
    if (rc) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    const float *d_act = host_actions_on_device(w, h_actions, sizeof(float) * (size_t)w->n * ad, st);
    if (!d_act) return RS_E_CUDA;
    const int ov = w->overlap;            // as in rs_vss_env_step_host
    w->overlap = 0;
    rc = rs_ssl_env_step(w, task, d_act, auto_reset, max_steps, w->s_obs, w->s_reward, w->s_done, w->s_trunc, nullptr, stream);
    w->overlap = ov;
    if (rc) return rc;
    return host_epilogue(w, od, h_obs, h_reward, h_done, h_trunc, st, begin);
}
int rs_ssl_env_step_host(rs_world *w, int task, const float *h_actions, int auto_reset, int max_steps,
                         float *h_obs, float *h_reward, uint8_t *h_done, uint8_t *h_trunc, void *stream) {
    return ssl_env_step_host(w, task, h_actions, auto_reset, max_steps, h_obs, h_reward, h_done, h_trunc, stream, false);
}
int rs_ssl_env_step_host_begin(rs_world *w, int task, const float *h_actions, int auto_reset, int max_steps,
                               float *h_obs, float *h_reward, uint8_t *h_done, uint8_t *h_trunc, void *stream) {
    return ssl_env_step_host(w, task, h_actions, auto_reset, max_steps, h_obs, h_reward, h_done, h_trunc, stream, true);
}
int rs_host_step_wait(rs_world *w) {
    if (!w) return fail(RS_E_INVALID, "rs_host_step_wait: null world");
    if (!w->host_pending) return RS_OK;
    ON_DEVICE(w, "rs_host_step_wait");
    w->host_pending = false;
    CUDA_TRY(cudaEventSynchronize(w->host_done));
    return RS_OK;
}

}  // extern "C"
