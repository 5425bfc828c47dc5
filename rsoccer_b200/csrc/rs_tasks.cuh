// rs_tasks.cuh -- device restatement of the task logic of the benchmarked reference envs
// (commands, observation, reward, done, initial placement), fused into the step kernels.
//   VSS-v0                      rsoccer_gym/vss/env_vss/vss_gym.py
//   SSLStaticDefenders-v0       rsoccer_gym/ssl/ssl_hw_challenge/static_defenders.py
//   SSLContestedPossession-v0   rsoccer_gym/ssl/ssl_hw_challenge/contested_possession.py
#pragma once
#include "rs_device.cuh"

// vss_gym.py:235-254 _actions_to_v_wheels: action in [-1,1] -> wheel rad/s, clip, 0.05 m/s deadzone
template <bool PK = false>
__device__ __forceinline__ void vss_action_to_wheels(const DevParams &P, float a0, float a1,
                                                     float &wl, float &wr) {
    const float2 v = v_mul<PK>(make_float2(a0, a1), bc2(P.max_v));
    float l = clampf(v.x, -P.max_v, P.max_v), r = clampf(v.y, -P.max_v, P.max_v);
    if (fabsf(l) < (float)RS_VSS_DEADZONE) l = 0.0f;      // -dz < l < dz
    if (fabsf(r) < (float)RS_VSS_DEADZONE) r = 0.0f;
    const float2 w = v_mul<PK>(make_float2(l, r), bc2(P.inv_rw));
    wl = w.x; wr = w.y;
}

__device__ __forceinline__ float nrm(float v, float inv) {
    return clampf(v * inv, -(float)RS_NORM_BOUNDS, (float)RS_NORM_BOUNDS);
}

// two components with one scale, then the clamps
template <bool PK>
__device__ __forceinline__ void nrm2(const float a, const float b, const float inv, float &oa, float &ob) {
    const float2 t = v_mul<PK>(make_float2(a, b), bc2(inv));
    oa = clampf(t.x, -(float)RS_NORM_BOUNDS, (float)RS_NORM_BOUNDS); ob = clampf(t.y, -(float)RS_NORM_BOUNDS, (float)RS_NORM_BOUNDS);
}

// vss_gym.py:93-117 _frame_to_observations; o = row of 4 + 7 NB + 5 NY floats
template <int NB, int NY, bool PK = false>
__device__ __forceinline__ void vss_obs(const DevParams &P, const Scene<NB + NY> &s, float *o) {
    constexpr int NOBS = 4 + 7 * NB + 5 * NY;
    float v[NOBS];
    int k = 0;
    nrm2<PK>(s.bx, s.by, P.inv_max_pos, v[k], v[k + 1]); nrm2<PK>(s.bvx, s.bvy, P.inv_max_v, v[k + 2], v[k + 3]); k += 4;
#pragma unroll
    for (int r = 0; r < NB; ++r) {
        float sn, cs;
        __sincosf(s.th[r], &sn, &cs);
        nrm2<PK>(s.x[r], s.y[r], P.inv_max_pos, v[k], v[k + 1]);
        v[k + 2] = sn; v[k + 3] = cs;
        nrm2<PK>(s.vx[r], s.vy[r], P.inv_max_v, v[k + 4], v[k + 5]);
        v[k + 6] = nrm(s.om[r], P.inv_max_w_rad);
        k += 7;
    }
#pragma unroll
    for (int r = NB; r < NB + NY; ++r) {
        nrm2<PK>(s.x[r], s.y[r], P.inv_max_pos, v[k], v[k + 1]);
        nrm2<PK>(s.vx[r], s.vy[r], P.inv_max_v, v[k + 2], v[k + 3]);
        v[k + 4] = nrm(s.om[r], P.inv_max_w_rad);
        k += 5;
    }
    if ((NOBS & 3) == 0 && ((reinterpret_cast<uintptr_t>(o) & 15u) == 0u)) {
        float4 *o4 = reinterpret_cast<float4 *>(o);
#pragma unroll
        for (int i = 0; i < NOBS / 4; ++i) o4[i] = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
    } else {
#pragma unroll
        for (int i = 0; i < NOBS; ++i) o[i] = v[i];
    }
}

// static_defenders.py:90-112 / contested_possession.py:78-105; row of 4 + 8 NB + 2 NY floats.
// The reference divides v by the overridden max_v = 2.5 and v_theta [deg/s] by max_w = 10
// (static_defenders.py:76-77) -- preserved.
template <int NB, int NY>
__device__ __forceinline__ void ssl_obs(const DevParams &P, const Scene<NB + NY> &s, float *o) {
    int k = 0;
    const float inv_v = 1.0f / 2.5f, inv_w = RS_DEG_F / 10.0f;
    o[k++] = nrm(s.bx, P.inv_max_pos); o[k++] = nrm(s.by, P.inv_max_pos);
    o[k++] = nrm(s.bvx, inv_v); o[k++] = nrm(s.bvy, inv_v);
#pragma unroll
    for (int r = 0; r < NB; ++r) {
        float sn, cs;
        __sincosf(s.th[r], &sn, &cs);
        o[k++] = nrm(s.x[r], P.inv_max_pos); o[k++] = nrm(s.y[r], P.inv_max_pos);
        o[k++] = sn; o[k++] = cs;
        o[k++] = nrm(s.vx[r], inv_v); o[k++] = nrm(s.vy[r], inv_v);
        o[k++] = nrm(s.om[r], inv_w);
        o[k++] = touching(P, s.x[r], s.y[r], cs, sn, s.bx, s.by) ? 1.0f : 0.0f;
    }
#pragma unroll
    for (int r = NB; r < NB + NY; ++r) { o[k++] = nrm(s.x[r], P.inv_max_pos); o[k++] = nrm(s.y[r], P.inv_max_pos); }
}

// ---- initial placement ----
// vss_gym.py:194-233 (min_dist 0.1, exact all-pairs distance instead of the KD-tree)
template <int RT>
__device__ __noinline__ void vss_place(const DevParams &P, Rng g, Scene<RT> &s) {
    const int R = RT > 0 ? RT : P.n_robots;
    const float hl = P.half_len, hw = P.half_wid;
    s.bx = g.uniform(-hl + 0.1f, hl - 0.1f); s.by = g.uniform(-hw + 0.1f, hw - 0.1f);
    s.bvx = 0.0f; s.bvy = 0.0f;
    for (int r = 0; r < R; ++r) {
        float x = 0.0f, y = 0.0f;
        for (int tries = 0; tries < 64; ++tries) {
            x = g.uniform(-hl + 0.1f, hl - 0.1f); y = g.uniform(-hw + 0.1f, hw - 0.1f);
            float dx = x - s.bx, dy = y - s.by;
            bool ok = !(dx * dx + dy * dy < 0.01f);
            for (int k = 0; k < r; ++k) {
                dx = x - s.x[k]; dy = y - s.y[k];
                if (dx * dx + dy * dy < 0.01f) ok = false;
            }
            if (ok) break;
        }
        s.x[r] = x; s.y[r] = y; s.vx[r] = 0.0f; s.vy[r] = 0.0f; s.om[r] = 0.0f;
        s.th[r] = wrap_pi(g.uniform(0.0f, 360.0f) * (1.0f / RS_DEG_F));
    }
}
// The same placement as vss_place, fed from a warp-generated block of the SAME Philox stream
// and with the placed robots in (statically indexed) registers.  An episode end is rare per
// lane but the kernel ends with its slowest warp, and ~10 matches of 65 536 end every step:
// the scalar routine above (six dependent Philox calls, scene in local memory) was 1.1 us of
// every 20 us step (profiles/r1c_decomposition.txt).  `buf` holds words 0 .. 4 navail - 1 of
// the stream (counter word 3 = j, component = i & 3, exactly what Rng::next() walks through);
// anything beyond is computed on the spot.
__device__ __noinline__ uint32_t placement_word(uint2 key, uint32_t env, uint32_t t, int i) {
    const uint4 u = philox4x32_10(make_uint4(env, t, RS_STREAM_AUTORESET, (uint32_t)(i >> 2)), key);
    const int c = i & 3;
    return c == 0 ? u.x : c == 1 ? u.y : c == 2 ? u.z : u.w;
}
struct PlaceStream {
    const uint32_t *buf; int navail; uint2 key; uint32_t env, t; int i;
    __device__ __forceinline__ float uniform(float a, float b) {
        const uint32_t v = (i >> 2) < navail ? buf[i] : placement_word(key, env, t, i);
        ++i;
        return a + (b - a) * u01(v);
    }
};
template <int R, class PP>
__device__ __forceinline__ void vss_place_stream(const PP &P, PlaceStream g, Scene<R> &s) {
    const float hl = P.half_len, hw = P.half_wid;
    s.bx = g.uniform(-hl + 0.1f, hl - 0.1f); s.by = g.uniform(-hw + 0.1f, hw - 0.1f);
    s.bvx = 0.0f; s.bvy = 0.0f;
#pragma unroll
    for (int r = 0; r < R; ++r) {
        float x = 0.0f, y = 0.0f;
#pragma unroll 1
        for (int tries = 0; tries < 64; ++tries) {
            x = g.uniform(-hl + 0.1f, hl - 0.1f); y = g.uniform(-hw + 0.1f, hw - 0.1f);
            float dx = x - s.bx, dy = y - s.by;
            bool ok = !(dx * dx + dy * dy < 0.01f);
#pragma unroll
            for (int k = 0; k < r; ++k) {
                dx = x - s.x[k]; dy = y - s.y[k];
                if (dx * dx + dy * dy < 0.01f) ok = false;
            }
            if (ok) break;
        }
        s.x[r] = x; s.y[r] = y; s.vx[r] = 0.0f; s.vy[r] = 0.0f; s.om[r] = 0.0f;
        s.th[r] = wrap_pi(g.uniform(0.0f, 360.0f) * (1.0f / RS_DEG_F));
    }
}
// static_defenders.py:214-254 fed from a PlaceStream (see vss_place_stream): same draws, same results
template <int R, class PP>
__device__ __forceinline__ void ssl_sd_place_stream(const PP &P, PlaceStream g, Scene<R> &s) {
    const float hl = P.half_len, hw = P.half_wid;
    s.x[0] = 0.0f; s.y[0] = 0.0f; s.th[0] = 0.0f; s.vx[0] = 0.0f; s.vy[0] = 0.0f; s.om[0] = 0.0f;
#pragma unroll 1
    for (int tries = 0; tries < 64; ++tries) {
        s.bx = g.uniform(0.2f, hl - 0.1f); s.by = g.uniform(-hw + 0.1f, hw - 0.1f);
        if (!(s.bx > hl - P.pen_len && fabsf(s.by) < P.half_pen_wid)) break;
    }
    s.bvx = 0.0f; s.bvy = 0.0f;
#pragma unroll
    for (int r = 1; r < R; ++r) {
        float x = 0.0f, y = 0.0f;
#pragma unroll 1
        for (int tries = 0; tries < 64; ++tries) {
            x = g.uniform(0.2f, hl - 0.1f); y = g.uniform(-hw + 0.1f, hw - 0.1f);
            float dx = x - s.bx, dy = y - s.by;
            bool ok = !(dx * dx + dy * dy < 0.04f);
#pragma unroll
            for (int k = 0; k < r; ++k) {
                dx = x - s.x[k]; dy = y - s.y[k];
                if (dx * dx + dy * dy < 0.04f) ok = false;
            }
            if (ok) break;
        }
        s.x[r] = x; s.y[r] = y; s.vx[r] = 0.0f; s.vy[r] = 0.0f; s.om[r] = 0.0f;
        s.th[r] = wrap_pi(g.uniform(0.0f, 360.0f) * (1.0f / RS_DEG_F));
    }
}
// contested_possession.py:210-227 (two draws)
template <int R, class PP>
__device__ __forceinline__ void ssl_cp_place_stream(const PP &P, PlaceStream g, Scene<R> &s) {
    s.x[0] = 0.0f; s.y[0] = 0.0f; s.th[0] = 0.0f; s.vx[0] = 0.0f; s.vy[0] = 0.0f; s.om[0] = 0.0f;
    const float ex = g.uniform(P.pen_len, P.half_len - P.pen_len);
    const float ey = g.uniform(-P.half_pen_wid, P.half_pen_wid);
    s.bx = ex - 0.1f; s.by = ey; s.bvx = 0.0f; s.bvy = 0.0f;
#pragma unroll
    for (int r = 1; r < R; ++r) {
        s.x[r] = ex; s.y[r] = ey + 0.5f * (float)(r - 1); s.th[r] = RS_PI_F;
        s.vx[r] = 0.0f; s.vy[r] = 0.0f; s.om[r] = 0.0f;
    }
}
template <int TASK, int R, class PP>
__device__ __forceinline__ void place_from_stream(const PP &P, const PlaceStream &g, Scene<R> &s) {
    if constexpr (TASK == RS_TASK_VSS) vss_place_stream<R>(P, g, s);
    else if constexpr (TASK == RS_TASK_SSL_STATIC_DEFENDERS) ssl_sd_place_stream<R>(P, g, s);
    else ssl_cp_place_stream<R>(P, g, s);
}

// Warp-cooperative draw of the first 4 x navail words of the auto-reset placement stream of the
// match whose global id is `env_src`: lane j computes Philox block j and writes it to words
// [4 j, 4 j + 4) of `buf` (shared memory, 128 words per warp).  Every lane in `live` calls.
__device__ __forceinline__ void warp_placement_words(uint32_t *buf, const unsigned live, const uint64_t seed,
                                                     const uint32_t env_src, const uint32_t t) {
    const uint2 key = make_uint2((uint32_t)seed, (uint32_t)(seed >> 32));
    const uint4 blk = philox4x32_10(make_uint4(env_src, t, RS_STREAM_AUTORESET, (uint32_t)(threadIdx.x & 31)), key);
    __syncwarp(live);
    reinterpret_cast<uint4 *>(buf)[threadIdx.x & 31] = blk;
    __syncwarp(live);
}

// static_defenders.py:214-254
template <int RT>
__device__ __noinline__ void ssl_sd_place(const DevParams &P, Rng g, Scene<RT> &s) {
    const int R = RT > 0 ? RT : P.n_robots;
    const float hl = P.half_len, hw = P.half_wid;
    s.x[0] = 0.0f; s.y[0] = 0.0f; s.th[0] = 0.0f; s.vx[0] = 0.0f; s.vy[0] = 0.0f; s.om[0] = 0.0f;
    for (int tries = 0; tries < 64; ++tries) {
        s.bx = g.uniform(0.2f, hl - 0.1f); s.by = g.uniform(-hw + 0.1f, hw - 0.1f);
        if (!(s.bx > hl - P.pen_len && fabsf(s.by) < P.half_pen_wid)) break;
    }
    s.bvx = 0.0f; s.bvy = 0.0f;
    for (int r = 1; r < R; ++r) {
        float x = 0.0f, y = 0.0f;
        for (int tries = 0; tries < 64; ++tries) {
            x = g.uniform(0.2f, hl - 0.1f); y = g.uniform(-hw + 0.1f, hw - 0.1f);
            float dx = x - s.bx, dy = y - s.by;
            bool ok = !(dx * dx + dy * dy < 0.04f);
            for (int k = 0; k < r; ++k) {
                dx = x - s.x[k]; dy = y - s.y[k];
                if (dx * dx + dy * dy < 0.04f) ok = false;
            }
            if (ok) break;
        }
        s.x[r] = x; s.y[r] = y; s.vx[r] = 0.0f; s.vy[r] = 0.0f; s.om[r] = 0.0f;
        s.th[r] = wrap_pi(g.uniform(0.0f, 360.0f) * (1.0f / RS_DEG_F));
    }
}
// contested_possession.py:210-227
template <int RT>
__device__ __noinline__ void ssl_cp_place(const DevParams &P, Rng g, Scene<RT> &s) {
    const int R = RT > 0 ? RT : P.n_robots;
    s.x[0] = 0.0f; s.y[0] = 0.0f; s.th[0] = 0.0f; s.vx[0] = 0.0f; s.vy[0] = 0.0f; s.om[0] = 0.0f;
    const float ex = g.uniform(P.pen_len, P.half_len - P.pen_len);
    const float ey = g.uniform(-P.half_pen_wid, P.half_pen_wid);
    s.bx = ex - 0.1f; s.by = ey; s.bvx = 0.0f; s.bvy = 0.0f;
    for (int r = 1; r < R; ++r) {
        s.x[r] = ex; s.y[r] = ey + 0.5f * (float)(r - 1); s.th[r] = RS_PI_F;
        s.vx[r] = 0.0f; s.vy[r] = 0.0f; s.om[r] = 0.0f;
    }
}
// dribbling.py:187-202: fixed course, the robot at the origin facing -x with the ball in its mouth
template <int RT>
__device__ __noinline__ void ssl_drib_place(const DevParams &P, Rng g, Scene<RT> &s) {
    const int R = RT > 0 ? RT : P.n_robots;
    s.bx = -0.1f; s.by = 0.0f; s.bvx = 0.0f; s.bvy = 0.0f;
    for (int r = 0; r < R; ++r) {
        s.x[r] = r == 0 ? 0.0f : -0.5f * (float)r; s.y[r] = 0.0f; s.th[r] = RS_PI_F;
        s.vx[r] = 0.0f; s.vy[r] = 0.0f; s.om[r] = 0.0f;
    }
}
// pass_endurance.py:152-181: ball anywhere in [-1.5, 1.5]^2, the shooter 0.115 m behind it facing it
// along y, the receiver mirrored in y and at least 1 m away in x, facing the shooter
template <int RT>
__device__ __noinline__ void ssl_pass_place(const DevParams &P, Rng g, Scene<RT> &s) {
    s.bx = g.uniform(-1.5f, 1.5f); s.by = g.uniform(1.5f, -1.5f); s.bvx = 0.0f; s.bvy = 0.0f;
    const float factor = s.by < 0.0f ? -1.0f : 1.0f;
    s.x[0] = s.bx; s.y[0] = s.by + 0.115f * factor; s.th[0] = factor > 0.0f ? -0.5f * RS_PI_F : 0.5f * RS_PI_F;
    float rx = 0.0f;
    for (int tries = 0; tries < 64; ++tries) {
        rx = g.uniform(-1.5f, 1.5f);
        if (!(fabsf(rx - s.bx) < 1.0f)) break;
    }
    s.x[1] = rx; s.y[1] = -s.by;
    s.th[1] = wrap_pi(atan2f(s.y[1] - s.y[0], s.x[1] - s.x[0]) + RS_PI_F);
    for (int r = 0; r < 2; ++r) { s.vx[r] = 0.0f; s.vy[r] = 0.0f; s.om[r] = 0.0f; }
}
template <int TASK, int RT>
__device__ __forceinline__ void task_place(const DevParams &P, const Rng &g, Scene<RT> &s) {
    if (TASK == RS_TASK_VSS) vss_place<RT>(P, g, s);
    else if (TASK == RS_TASK_SSL_STATIC_DEFENDERS) ssl_sd_place<RT>(P, g, s);
    else if (TASK == RS_TASK_SSL_DRIBBLING) ssl_drib_place<RT>(P, g, s);
    else if (TASK == RS_TASK_SSL_PASS_ENDURANCE) ssl_pass_place<RT>(P, g, s);
    else ssl_cp_place<RT>(P, g, s);
}

// dribbling.py:75-105 (21 floats: checkpoint progress first, infrared as +-1) and
// pass_endurance.py:78-90 (16 floats: no robot velocities, infrared as 1 / 0); the scene type is
// generic so that k_task_reset (Scene<0>) shares it
template <int TASK, class SC>
__device__ __forceinline__ void ssl_hw_obs(const DevParams &P, const SC &s, const float counter, float *o) {
    const float inv_v = 1.0f / 2.5f, inv_w = RS_DEG_F / 10.0f;     // dribbling.py:66-67, pass_endurance.py:73-74
    const int NB = TASK == RS_TASK_SSL_DRIBBLING ? 1 : 2, R = TASK == RS_TASK_SSL_DRIBBLING ? 5 : 2;
    int k = 0;
    if (TASK == RS_TASK_SSL_DRIBBLING) o[k++] = ((counter / 6.0f) * 2.0f) - 1.0f;
    o[k++] = nrm(s.bx, P.inv_max_pos); o[k++] = nrm(s.by, P.inv_max_pos);
    o[k++] = nrm(s.bvx, inv_v); o[k++] = nrm(s.bvy, inv_v);
#pragma unroll
    for (int r = 0; r < NB; ++r) {
        float sn, cs;
        __sincosf(s.th[r], &sn, &cs);
        o[k++] = nrm(s.x[r], P.inv_max_pos); o[k++] = nrm(s.y[r], P.inv_max_pos);
        o[k++] = sn; o[k++] = cs;
        if (TASK == RS_TASK_SSL_DRIBBLING) { o[k++] = nrm(s.vx[r], inv_v); o[k++] = nrm(s.vy[r], inv_v); }
        o[k++] = nrm(s.om[r], inv_w);
        const bool ir = touching(P, s.x[r], s.y[r], cs, sn, s.bx, s.by);
        o[k++] = ir ? 1.0f : (TASK == RS_TASK_SSL_DRIBBLING ? -1.0f : 0.0f);
    }
#pragma unroll
    for (int r = NB; r < R; ++r) { o[k++] = nrm(s.x[r], P.inv_max_pos); o[k++] = nrm(s.y[r], P.inv_max_pos); }
}
