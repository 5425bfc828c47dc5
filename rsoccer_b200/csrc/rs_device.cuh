// rs_device.cuh -- device-side physics of the batched 2-D robot-soccer engine (sm_100a).
//
// One LANE per match ("env"): a warp advances 32 independent matches in lockstep,
// the whole <= 23-body scene of a match lives in that lane's registers for all five
// sub-steps, and HBM is touched once per control step (coalesced 128-bit loads /
// stores of the SoA state).  DESIGN.md section 4 explains why this mapping beats a
// warp-per-match mapping for the 7-body VSS scene (22 % lane use in the body phases).
//
// Replaces the arithmetic inside `robosim.VSS.step` / `robosim.SSL.step`
// (reference call sites rsoccer_gym/Simulators/rsim.py:102, :155).  The model is
// DESIGN.md section 3; oracle/rs_oracle.c restates it independently in fp64.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/rs_spec.h"

// Kernel-tuning experiment knobs (tools/variants.py builds one library per combination; the
// product build defines none of them).
#ifndef RS_X_SUBSTEPS
#define RS_X_SUBSTEPS RS_SUBSTEPS
#endif

#define RS_PI_F 3.14159265358979323846f
#define RS_DEG_F 57.29577951308232f

struct DevParams {
    static constexpr bool packed = false;   // see v_sub / v_mul / v_fma below
    int kind, n_blue, n_yellow, n_robots;
    float dt, h;
    // field
    float half_len, half_wid, goal_depth, pen_len, half_pen_wid, half_goal_wid;
    float ball_r, rbt_r, rw, inv_rw, wmax;
    // walls
    float x_out, y_out, x_near;
    int n_box;
    float box[RS_MAX_BOXES][4];
    // bodies
    float wb, wr, inv_wsum, fb, fr;   // inverse masses, 1/(wb+wr), wb/(wb+wr), wr/(wb+wr)
    float e_ball_wall, e_rbt_wall, e_ball_rbt, e_rbt_rbt, mu_ball_rbt;
    float ball_decel_h;
    float rs_br, rs_br2, rs_rr, rs_rr2;
    // drive
    float inv_2b, acc_fwd_h, acc_lat_h, acc_ang_h;
    float J[4][3], Jp[3][4];
    // kicker
    float dk, kick_centre, kick_reach, kick_hw, mouth_hc, kick_max;
    // task normalisers (vss_gym_base.py:52-58, 213-220)
    float inv_max_pos, max_v, inv_max_v, inv_max_w_rad;  // obs w = clamp(omega[rad/s] * inv_max_w_rad)
    float sqrt_dt;
};

// Compile-time parameter block of the benchmarked world (VSS, field_type 0, 25 ms): the same
// formulas as rs_params_fill / fill_dev_params evaluated by the compiler, so that the hot loop
// sees immediates instead of constant-bank loads.  rs_capi.cu checks it field by field against
// the run-time block before it selects a kernel built on it.
struct VssF0 {
    static constexpr bool packed = false;
    static constexpr double h_ = 0.025 / RS_SUBSTEPS, wb_ = 1.0 / 0.046, wr_ = 1.0 / 0.18;
    static constexpr double br_ = 0.0375 + 0.0215, rr_ = 2.0 * 0.0375;
    static constexpr int n_robots = 6;
    static constexpr float h = (float)h_;
    static constexpr float ball_r = (float)0.0215, rbt_r = (float)0.0375;
    static constexpr float x_out = (float)(1.5 / 2 + 0.1), y_out = (float)(1.3 / 2);
    static constexpr float wall_lx = (float)(1.5 / 2), wall_ly = (float)(0.4 / 2);
    static constexpr float wb = (float)wb_, wr = (float)wr_, inv_wsum = (float)(1.0 / (wb_ + wr_));
    static constexpr float fb = (float)(wb_ / (wb_ + wr_)), fr = (float)(wr_ / (wb_ + wr_));
    static constexpr float e_ball_wall = (float)0.6, e_rbt_wall = (float)0.0, e_ball_rbt = (float)0.4, e_rbt_rbt = (float)0.0;
    static constexpr float mu_ball_rbt = (float)0.3;
    static constexpr float ball_decel_h = (float)(0.05 * 9.81 * h_);
    static constexpr float rs_br = (float)br_, rs_br2 = (float)(br_ * br_), rs_rr = (float)rr_, rs_rr2 = (float)(rr_ * rr_);
    static constexpr float acc_fwd_h = (float)(6.0 * h_), acc_lat_h = (float)(9.0 * h_), acc_ang_h = (float)(200.0 * h_);
};
// VssF0 with the packed fp32x2 instruction forms (below): the kernels of worlds large enough to be
// bound by issue slots; RS_X_NOPACK = tuning knob, the scalar forms everywhere
struct VssF0P : VssF0 {
#ifndef RS_X_NOPACK
    static constexpr bool packed = true;
#endif
};
__device__ __forceinline__ float wall_lx(const DevParams &P) { return P.box[0][0]; }
__device__ __forceinline__ float wall_ly(const DevParams &P) { return P.box[0][1]; }
__device__ __forceinline__ constexpr float wall_lx(const VssF0 &) { return VssF0::wall_lx; }
__device__ __forceinline__ constexpr float wall_ly(const VssF0 &) { return VssF0::wall_ly; }

template <int RT> struct Cap { static constexpr int v = RT > 0 ? RT : RS_MAX_ROBOTS; };

// the scene of one match, register resident when RT > 0
template <int RT>
struct Scene {
    float bx, by, bvx, bvy;
    float x[Cap<RT>::v], y[Cap<RT>::v], vx[Cap<RT>::v], vy[Cap<RT>::v];
    float th[Cap<RT>::v], om[Cap<RT>::v];
};

// per-robot drive targets in the robot frame + kicker / dribbler commands
template <int RT>
struct Drive {
    float tf[Cap<RT>::v], tl[Cap<RT>::v], tw[Cap<RT>::v];
    float kick[Cap<RT>::v];
    uint32_t drib;   // bit r: dribbler on
};

__device__ __forceinline__ float clampf(float x, float lo, float hi) { return fminf(fmaxf(x, lo), hi); }
// Packed fp32x2 arithmetic of sm_100 (SASS FFMA2 / FADD2 / FMUL2; a scalar register or an
// immediate broadcasts into both halves, negation is an operand modifier): two IEEE results per
// issued instruction, each half rounded exactly like the scalar instruction.  Every (x, y) /
// (vx, vy) operation of a body -- the halves of the register quad its state was loaded into --
// can be issued as one instruction.  Measured (tools/microbench/pk_latency.cu): scalar latency (4.4
// cycles), two cycles of the FMA pipe per packed instruction -- the same fp32 throughput in half
// the issue slots.  It pays where a kernel is bound by issue slots (VSS-v0, one lane per match,
// >= ~20 000 matches: 16.65 -> 15.7 us at 65 536) and not where it is bound by the dependent
// latency of a few warps (8 192 matches: 7.90 -> 8.02 us), so the source is written once on
// float2 and the parameter class picks the instruction form (PP::packed).
__device__ __forceinline__ float2 pk_add(const float2 a, const float2 b) { return __fadd2_rn(a, b); }
__device__ __forceinline__ float2 pk_mul(const float2 a, const float2 b) { return __fmul2_rn(a, b); }
__device__ __forceinline__ float2 pk_fma(const float2 a, const float2 b, const float2 c) { return __ffma2_rn(a, b, c); }
// a - b as ONE instruction: written as add(a, -b) the compiler keeps a negated copy of every body that enters
// several pairs (two FADD each) instead of using the operand modifier
__device__ __forceinline__ float2 pk_sub(const float2 a, const float2 b) {
    float2 r;
    asm("{\n\t.reg .b64 pa, pb, pr;\n\tmov.b64 pa, {%2, %3};\n\tmov.b64 pb, {%4, %5};\n\t"
        "sub.rn.ftz.f32x2 pr, pa, pb;\n\tmov.b64 {%0, %1}, pr;\n\t}"
        : "=f"(r.x), "=f"(r.y) : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y));
    return r;
}
template <bool PK> __device__ __forceinline__ float2 v_add(const float2 a, const float2 b) {
    if constexpr (PK) return pk_add(a, b); else return make_float2(a.x + b.x, a.y + b.y);
}
template <bool PK> __device__ __forceinline__ float2 v_sub(const float2 a, const float2 b) {
    if constexpr (PK) return pk_sub(a, b); else return make_float2(a.x - b.x, a.y - b.y);
}
template <bool PK> __device__ __forceinline__ float2 v_mul(const float2 a, const float2 b) {
    if constexpr (PK) return pk_mul(a, b); else return make_float2(a.x * b.x, a.y * b.y);
}
template <bool PK> __device__ __forceinline__ float2 v_fma(const float2 a, const float2 b, const float2 c) {
    if constexpr (PK) return pk_fma(a, b, c); else return make_float2(fmaf(a.x, b.x, c.x), fmaf(a.y, b.y, c.y));
}
__device__ __forceinline__ float2 bc2(const float a) { return make_float2(a, a); }
__device__ __forceinline__ float wrap_pi(float a) {
    if (a > RS_PI_F) a -= 2.0f * RS_PI_F; else if (a <= -RS_PI_F) a += 2.0f * RS_PI_F;
    return a;
}

// ---------------------------------------------------------------- Philox4x32-10
__device__ __forceinline__ uint4 philox4x32_10(uint4 c, uint2 k) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint32_t hi0 = __umulhi(0xD2511F53u, c.x), lo0 = 0xD2511F53u * c.x;
        const uint32_t hi1 = __umulhi(0xCD9E8D57u, c.z), lo1 = 0xCD9E8D57u * c.z;
        c = make_uint4(hi1 ^ c.y ^ k.x, lo1, hi0 ^ c.w ^ k.y, lo0);
        k.x += 0x9E3779B9u; k.y += 0xBB67AE85u;
    }
    return c;
}
// ((x >> 8) + 0.5) / 2^24 as one FMA: scaling by a power of two commutes with rounding, so this is bit for
// bit the fp32 add-then-multiply form (= the oracle's fp64 value rounded to fp32)
__device__ __forceinline__ float u01(uint32_t x) { return fmaf((float)(x >> 8), 1.0f / 16777216.0f, 0.5f / 16777216.0f); }

struct Rng {   // sequential u32 stream for (env, t, stream): counter (env, t, stream, j)
    uint4 ctr; uint2 key; uint4 buf; int idx;
    __device__ __forceinline__ Rng(uint64_t seed, uint32_t env, uint32_t t, uint32_t stream) {
        key = make_uint2((uint32_t)seed, (uint32_t)(seed >> 32));
        ctr = make_uint4(env, t, stream, 0u); idx = 4;
    }
    __device__ __forceinline__ uint32_t next() {
        if (idx == 4) { buf = philox4x32_10(ctr, key); ctr.w++; idx = 0; }
        const uint32_t v = idx == 0 ? buf.x : idx == 1 ? buf.y : idx == 2 ? buf.z : buf.w;
        ++idx; return v;
    }
    __device__ __forceinline__ float uniform(float a, float b) { return a + (b - a) * u01(next()); }
};

// ---------------------------------------------------------------- geometry
// kicker "touching" box in the robot frame (lx, ly = ball centre, robot frame)
__device__ __forceinline__ bool touching_local(const DevParams &P, float lx, float ly) {
    return (fabsf(lx - P.kick_centre) < P.kick_reach) && (fabsf(ly) < P.kick_hw);
}
__device__ __forceinline__ bool touching(const DevParams &P, float rx, float ry, float c, float s,
                                         float bx, float by) {
    const float dx = bx - rx, dy = by - ry;
    return touching_local(P, c * dx + s * dy, -s * dx + c * dy);
}

// walls for one body, mirrored quadrant.
//
// In steady state ~8 % of the robots touch a wall at any time, so in a warp of 32 matches
// "some lane touches a wall" is true for almost every body and sub-step: a branchy wall
// routine runs at 2-4 active lanes and was > 50 % of all issued instructions (profiles/
// r1_steady_before_walls.txt).  The VSS region (field + goal recesses = rectangle minus one
// unbounded corner box per quadrant) therefore uses a branch-free form: the faces of the
// corner box are axis limits selected by where the centre is, only the rounded goal-post
// corner (both coordinates within r of the box corner) takes a branch.  Same arithmetic
// as the generic closest-point test of the oracle for every reachable state (and for the
// unreachable "centre inside the solid" state, which it resolves with the same
// least-penetration face).
template <int KIND, class PP>
__device__ __forceinline__ void walls(const PP &P, float r, float e, float &x, float &y,
                                      float &vx, float &vy) {
    // mirror into the positive quadrant: multiply by sign(x) = flip the sign bit (exact)
    const uint32_t sgx = __float_as_uint(x) & 0x80000000u, sgy = __float_as_uint(y) & 0x80000000u;
    float ax = fabsf(x), ay = fabsf(y);
    float avx = __uint_as_float(__float_as_uint(vx) ^ sgx), avy = __uint_as_float(__float_as_uint(vy) ^ sgy);
    if constexpr (KIND == RS_KIND_VSS) {
        const float Lh = wall_lx(P), Gh = wall_ly(P);
        const bool inx = ax < Lh, iny = ay < Gh;
        const float dx = ax - Lh, dy = ay - Gh;
        if (__builtin_expect(inx && iny && dx > -r && dy > -r, 0)) {   // goal-post corner (rare)
            const float d2 = dx * dx + dy * dy;
            if (d2 < r * r) {
                float nx = -1.0f, ny = 0.0f, pen = r;     // d2 <= 1e-12: oracle's interior rule, m = fxl = 0
                if (d2 > 1e-12f) { const float inv = rsqrtf(d2); nx = dx * inv; ny = dy * inv; pen = r - d2 * inv; }
                ax += pen * nx; ay += pen * ny;
                const float vn = avx * nx + avy * ny;
                if (vn < 0.0f) { avx -= (1.0f + e) * vn * nx; avy -= (1.0f + e) * vn * ny; }
            }
        }
        const bool interior = !inx && !iny, pick_y = dy < dx;
        const bool box_x = !iny && !(interior && pick_y);
        const bool box_y = !inx && !(interior && !pick_y);
        const float xlim = (box_x ? Lh : P.x_out) - r, ylim = (box_y ? Gh : P.y_out) - r;
        // v > 0 ? -e v : v  ==  min(v, -e v)  for e >= 0
        if (ax > xlim) { ax = xlim; avx = fminf(avx, -e * avx); }
        if (ay > ylim) { ay = ylim; avy = fminf(avy, -e * avy); }
    } else {
        if (ax + r > P.x_near) {                          // near a goal: finite solid boxes
#pragma unroll
            for (int k = 0; k < RS_MAX_BOXES; ++k) {
                if (k < P.n_box) {
                    const float lox = P.box[k][0], loy = P.box[k][1], hix = P.box[k][2], hiy = P.box[k][3];
                    const float qx = clampf(ax, lox, hix), qy = clampf(ay, loy, hiy);
                    const float dx = ax - qx, dy = ay - qy;
                    const float d2 = dx * dx + dy * dy;
                    if (d2 < r * r) {
                        float nx, ny, pen;
                        if (d2 > 1e-12f) {
                            const float inv = rsqrtf(d2);
                            nx = dx * inv; ny = dy * inv; pen = r - d2 * inv;
                        } else {
                            const float fxl = ax - lox, fxh = hix - ax, fyl = ay - loy, fyh = hiy - ay;
                            float m = fxl; nx = -1.0f; ny = 0.0f;
                            if (fxh < m) { m = fxh; nx = 1.0f; ny = 0.0f; }
                            if (fyl < m) { m = fyl; nx = 0.0f; ny = -1.0f; }
                            if (fyh < m) { m = fyh; nx = 0.0f; ny = 1.0f; }
                            pen = r + m;
                        }
                        ax += pen * nx; ay += pen * ny;
                        const float vn = avx * nx + avy * ny;
                        if (vn < 0.0f) { avx -= (1.0f + e) * vn * nx; avy -= (1.0f + e) * vn * ny; }
                    }
                }
            }
        }
        if (ax > P.x_out - r) { ax = P.x_out - r; avx = fminf(avx, -e * avx); }
        if (ay > P.y_out - r) { ay = P.y_out - r; avy = fminf(avy, -e * avy); }
    }
    x = __uint_as_float(__float_as_uint(ax) | sgx); y = __uint_as_float(__float_as_uint(ay) | sgy);
    vx = __uint_as_float(__float_as_uint(avx) ^ sgx); vy = __uint_as_float(__float_as_uint(avy) ^ sgy);
}

// SSL walls, lean form for the lane-per-match kernels.  The two goal-wall boxes start at |x| = x_near = L / 2
// and the outer x bound lies beyond them, so a body with |x| + r <= x_near can only meet the outer y bound:
// clamp, and reflect the velocity if it points outward (what walls<SSL>() does there, without mirroring).
// Anything near a goal line takes the generic routine -- rare: the envs end an episode when the ball or the
// robot gets there, and un-commanded defenders are placed inside the field.
template <class PP>
__device__ __forceinline__ void ssl_walls(const PP &P, const float r, const float e, float &x, float &y, float &vx, float &vy) {
    if (__builtin_expect(fabsf(x) + r > P.x_near, 0)) { walls<RS_KIND_SSL>(P, r, e, x, y, vx, vy); return; }
    const float YO = P.y_out - r;
    const bool hy = fabsf(y) > YO, out = vy * y > 0.0f;
    y = fmaxf(fminf(y, YO), -YO);
    if (hy && out) vy = -e * vy;
}

// VSS walls, lean form for the lane-per-match kernels (7 bodies x 5 sub-steps per step: the
// largest single block of the sub-step loop).  No mirroring of the velocity, no predicate
// chains: away from the goal post each axis has ONE limit chosen by the sign of the other
// axis' distance to the corner box (|y| < Gh: the goal mouth band lets x run to the back of
// the goal; |x| >= Lh: inside the goal y stays within the mouth), the position is clamped with
// min/max and the velocity component is reflected only if it points outward.  A centre within
// r of BOTH faces of the corner box -- the rounded post, or (unreachable) inside the solid --
// takes the one branch.  Same decisions and arithmetic as walls<VSS>() above.
template <class PP>
__device__ __forceinline__ float wall_bounce(const PP &, const float e, const float v) { return e == 0.0f ? 0.0f : -e * v; }
template <class PP>
__device__ __forceinline__ void vss_walls(const PP &P, const float r, const float e, float &x, float &y, float &vx, float &vy) {
    const float Lh = wall_lx(P), Gh = wall_ly(P);
    const float XO = P.x_out - r, XL = Lh - r, YO = P.y_out - r, YL = Gh - r;      // loop invariants / immediates
    const float dx = fabsf(x) - Lh, dy = fabsf(y) - Gh;      // |.| is an operand modifier of FADD, not of FADD2
    float xlim = dy < 0.0f ? XO : XL, ylim = dx < 0.0f ? YO : YL;
    // both within r of a face AND on the same side of both faces (a robot leaning on the end
    // wall is within r of both, but outside one and inside the other: no branch for it)
    if (__builtin_expect(dx > -r && dy > -r && dx * dy >= 0.0f, 0)) {
        if (dx < 0.0f && dy < 0.0f) {
            const float d2 = dx * dx + dy * dy;
            if (d2 < r * r) {                             // rounded goal-post corner, in the mirrored quadrant
                const float sx = copysignf(1.0f, x), sy = copysignf(1.0f, y);
                float nx = -1.0f, ny = 0.0f, pen = r;     // d2 <= 1e-12: oracle's interior rule, m = fxl = 0
                if (d2 > 1e-12f) { const float inv = rsqrtf(d2); nx = dx * inv; ny = dy * inv; pen = r - d2 * inv; }
                float avx = vx * sx, avy = vy * sy;
                const float vn = avx * nx + avy * ny;
                if (vn < 0.0f) { avx -= (1.0f + e) * vn * nx; avy -= (1.0f + e) * vn * ny; }
                x = (fabsf(x) + pen * nx) * sx; y = (fabsf(y) + pen * ny) * sy;
                vx = avx * sx; vy = avy * sy;
            }
        } else if (dx >= 0.0f && dy >= 0.0f) {            // inside the solid: least-penetration face only
            if (dy < dx) xlim = XO; else ylim = YO;
        }
    }
    const float2 oo = v_mul<PP::packed>(make_float2(vx, vy), make_float2(x, y));
    const float ox = oo.x, oy = oo.y;                     // > 0: moving outward
    const bool hx = fabsf(x) > xlim, hy = fabsf(y) > ylim;
    x = fmaxf(fminf(x, xlim), -xlim); y = fmaxf(fminf(y, ylim), -ylim);
    if (hx && ox > 0.0f) vx = wall_bounce(P, e, vx);
    if (hy && oy > 0.0f) vy = wall_bounce(P, e, vy);
}

// SSL robot <-> ball geometry: the robot is the disc of radius rbt_r cut by the chord x_local = dk (the flat
// kicker mouth; contested_possession.py:224-225 starts the ball 0.1 m ahead of the centre, inside the
// bounding circle).  (dx, dy) = ball - robot on the phase-start positions, d2 its squared length (already known
// to be below the bounding-circle contact distance).  Closest point of the shape to the ball centre ->
// unit normal robot->ball, penetration, contact point offset (world frame).  false: no contact.
template <class PP>
__device__ __forceinline__ bool ssl_mouth_contact(const PP &P, const float dx, const float dy, const float d2, const float rth,
                                                  float &nx, float &ny, float &pen, float &rcx, float &rcy) {
    float s, c;
    __sincosf(rth, &s, &c);
    const float lx = c * dx + s * dy, ly = -s * dx + c * dy;
    const float bn = sqrtf(d2);
    float q1x = lx, q1y = ly;
    if (bn > P.rbt_r) { const float k = P.rbt_r / bn; q1x = lx * k; q1y = ly * k; }
    float qx, qy;
    if (q1x <= P.dk) { qx = q1x; qy = q1y; }
    else { qx = P.dk; qy = clampf(ly, -P.mouth_hc, P.mouth_hc); }
    const float ex = lx - qx, ey = ly - qy;
    const float e2 = ex * ex + ey * ey;
    if (e2 >= P.ball_r * P.ball_r) return false;
    float lnx, lny;
    if (e2 > 1e-12f) {
        const float inv = rsqrtf(e2);
        lnx = ex * inv; lny = ey * inv; pen = P.ball_r - e2 * inv;
    } else {
        const float pr = P.rbt_r - bn, pf = P.dk - lx;
        if (pf < pr) { lnx = 1.0f; lny = 0.0f; pen = P.ball_r + pf; }
        else {
            if (bn > 1e-9f) { lnx = lx / bn; lny = ly / bn; } else { lnx = 1.0f; lny = 0.0f; }
            pen = P.ball_r + pr;
        }
    }
    nx = c * lnx - s * lny; ny = s * lnx + c * lny;
    rcx = c * qx - s * qy; rcy = s * qx + c * qy;
    return true;
}

// robot <-> ball: detect on (rx, ry, rth) vs (bx, by); impulse on velocities, position
// corrections accumulated into (cbx, cby) / (crx, cry)
template <int KIND, class PP>
__device__ __forceinline__ void ball_robot(const PP &P, float bx, float by, float &bvx,
                                           float &bvy, float rx, float ry, float rth, float &rvx,
                                           float &rvy, float rom, float &cbx, float &cby,
                                           float &crx, float &cry, bool &any) {
    const float dx = bx - rx, dy = by - ry;
    const float d2 = dx * dx + dy * dy;
    if (__builtin_expect(d2 >= P.rs_br2, 1)) return;
    float nx, ny, pen, rcx, rcy;
    if constexpr (KIND == RS_KIND_VSS) {
        float d = 0.0f; nx = 1.0f; ny = 0.0f;
        if (d2 > 1e-12f) { const float inv = rsqrtf(d2); nx = dx * inv; ny = dy * inv; d = d2 * inv; }
        pen = P.rs_br - d;
        rcx = nx * P.rbt_r; rcy = ny * P.rbt_r;
    } else {
        if (!ssl_mouth_contact(P, dx, dy, d2, rth, nx, ny, pen, rcx, rcy)) return;
    }
    const float sx = rvx - rom * rcy, sy = rvy + rom * rcx;   // robot surface velocity at contact
    const float relx = bvx - sx, rely = bvy - sy;
    const float vn = relx * nx + rely * ny;
    if (vn < 0.0f) {
        const float Jn = -(1.0f + P.e_ball_rbt) * vn * P.inv_wsum;
        bvx += Jn * P.wb * nx; bvy += Jn * P.wb * ny;
        rvx -= Jn * P.wr * nx; rvy -= Jn * P.wr * ny;
        const float tx = -ny, ty = nx;
        const float vt = relx * tx + rely * ty;
        const float Jt = clampf(-vt * P.inv_wsum, -P.mu_ball_rbt * Jn, P.mu_ball_rbt * Jn);
        bvx += Jt * P.wb * tx; bvy += Jt * P.wb * ty;
        rvx -= Jt * P.wr * tx; rvy -= Jt * P.wr * ty;
    }
    cbx += pen * P.fb * nx; cby += pen * P.fb * ny;
    crx -= pen * P.fr * nx; cry -= pen * P.fr * ny;
    any = true;
}

template <class PP>
__device__ __forceinline__ void robot_robot(const PP &P, float xi, float yi, float &vxi,
                                            float &vyi, float xj, float yj, float &vxj, float &vyj,
                                            float &cxi, float &cyi, float &cxj, float &cyj, bool &any) {
    const float dx = xj - xi, dy = yj - yi;
    const float d2 = dx * dx + dy * dy;
    if (__builtin_expect(d2 >= P.rs_rr2, 1)) return;
    float d = 0.0f, nx = 1.0f, ny = 0.0f;
    if (d2 > 1e-12f) { const float inv = rsqrtf(d2); nx = dx * inv; ny = dy * inv; d = d2 * inv; }
    const float pen = P.rs_rr - d;
    const float vn = (vxj - vxi) * nx + (vyj - vyi) * ny;
    if (vn < 0.0f) {
        const float Jw = -(1.0f + P.e_rbt_rbt) * vn * 0.5f;   // J * wr with equal masses
        vxi -= Jw * nx; vyi -= Jw * ny; vxj += Jw * nx; vyj += Jw * ny;
    }
    const float hp = 0.5f * pen;
    cxi -= hp * nx; cyi -= hp * ny; cxj += hp * nx; cyj += hp * ny;
    any = true;
}

// pair k of the lexicographic order [ball x robots 0..R-1, then robot pairs i < j] -> (i, j)
template <int R> __host__ __device__ constexpr int rr_pair_i(int k) {
    int i = 0;
    while (k >= R - 1 - i) { k -= R - 1 - i; ++i; }
    return i;
}
template <int R> __host__ __device__ constexpr int rr_pair_j(int k) {
    int i = 0;
    while (k >= R - 1 - i) { k -= R - 1 - i; ++i; }
    return i + 1 + k;
}
// warp-uniform jump to the register-static body of pair p in [LO, HI): a binary decision tree
template <int KIND, int RT, int LO, int HI, class PP>
struct PairDispatch {
    static __device__ __forceinline__ void run(const int p, const PP &P, Scene<RT> &s, float &cbx, float &cby,
                                               float (&cx)[Cap<RT>::v], float (&cy)[Cap<RT>::v], bool &any) {
        if constexpr (HI - LO == 1) {
            if constexpr (LO < RT) {
                ball_robot<KIND>(P, s.bx, s.by, s.bvx, s.bvy, s.x[LO], s.y[LO], s.th[LO], s.vx[LO], s.vy[LO], s.om[LO],
                                 cbx, cby, cx[LO], cy[LO], any);
            } else {
                constexpr int i = rr_pair_i<RT>(LO - RT), j = rr_pair_j<RT>(LO - RT);
                robot_robot(P, s.x[i], s.y[i], s.vx[i], s.vy[i], s.x[j], s.y[j], s.vx[j], s.vy[j], cx[i], cy[i], cx[j], cy[j], any);
            }
        } else {
            constexpr int MID = (LO + HI) / 2;
            if (p < MID) PairDispatch<KIND, RT, LO, MID, PP>::run(p, P, s, cbx, cby, cx, cy, any);
            else PairDispatch<KIND, RT, MID, HI, PP>::run(p, P, s, cbx, cby, cx, cy, any);
        }
    }
};

// ---- per-lane contact resolve through shared memory (VSS, lane-per-match kernels) ----
// A contact is rare per lane but not per warp (see physics_step): with a register-static body
// per pair, a warp of 32 matches executed ~2 of 21 different ~60-instruction bodies per
// sub-step at 1-2 active lanes each -- 25 % of the step time, mostly dependent-latency and
// instruction-cache misses (profiles/r1c_decomposition.txt).  Registers cannot be indexed by a
// run-time pair number, shared memory can: a lane that has a contact publishes its scene to its
// own column of a CTA scratch (no other lane ever touches that column: no barrier, no bank
// conflict as the row pitch is a multiple of 128 B), resolves ITS OWN lowest pair with one
// branch-free body for every pair type, repeats while it has pairs left (ascending =
// lexicographic order, the oracle's Gauss-Seidel order), and reads the scene back.  Different
// lanes resolve different pairs in the same pass.
//   q[b]  = (x, y, vx, vy) of body b: receives impulses and position corrections
//   p0    = (x, y) and omega of every body at phase start: every contact of the phase is detected on these
template <int R> __host__ __device__ constexpr uint64_t rr_table(bool second) {
    uint64_t t = 0;
    for (int k = 0; k < R * (R - 1) / 2; ++k) t |= (uint64_t)(second ? rr_pair_j<R>(k) : rr_pair_i<R>(k)) << (3 * k);
    return t;
}
// per pair type (0 = ball pair, 1 = robot pair): contact distance, lever arm of the surface
// velocity, (1 + e) / mass sum, impulse -> velocity of S and F, friction, correction split of S and F
template <class PP> struct ContactTab {
    float rs, rc, kn, wS, wF, mu, gS, gF;
};
template <class PP>
__device__ __forceinline__ ContactTab<PP> contact_tab(const PP &P, const bool ball) {
    ContactTab<PP> t;
    t.rs = ball ? P.rs_br : P.rs_rr;
    t.rc = ball ? P.rbt_r : 0.0f;
    t.kn = ball ? (1.0f + P.e_ball_rbt) * P.inv_wsum : (1.0f + P.e_rbt_rbt) * 0.5f;
    t.wS = ball ? P.wb : 1.0f; t.wF = ball ? P.wr : 1.0f;
    t.mu = ball ? P.mu_ball_rbt : 0.0f;
    t.gS = ball ? P.fb : 0.5f; t.gF = ball ? P.fr : 0.5f;
    return t;
}
// The same per pair, for the compile-time world (VssF0), as a table indexed by the pair number: body
// indices and the eight constants of the pair's type arrive with three 16-byte loads instead of
// a bit-table decode and a chain of selects (~30 instructions per resolved contact; 2-3 lanes of
// a warp are active here, each reading the entry of its own pair).
struct alignas(16) PairEnt { int F, S; float rs, rc, kn, wS, wF, mu, gS, gF, inv_wsum, pad; };
template <int RT> struct PairTab { PairEnt e[RT + RT * (RT - 1) / 2]; };
template <int RT> constexpr PairTab<RT> make_pair_tab_f0() {
    PairTab<RT> t{};
    for (int p = 0; p < RT + RT * (RT - 1) / 2; ++p) {
        const bool ball = p < RT;
        PairEnt &e = t.e[p];
        e.F = ball ? p + 1 : 1 + rr_pair_i<RT>(p - RT);
        e.S = ball ? 0 : 1 + rr_pair_j<RT>(p - RT);
        e.rs = ball ? VssF0::rs_br : VssF0::rs_rr;
        e.rc = ball ? VssF0::rbt_r : 0.0f;
        e.kn = ball ? (1.0f + VssF0::e_ball_rbt) * VssF0::inv_wsum : (1.0f + VssF0::e_rbt_rbt) * 0.5f;
        e.wS = ball ? VssF0::wb : 1.0f; e.wF = ball ? VssF0::wr : 1.0f;
        e.mu = ball ? VssF0::mu_ball_rbt : 0.0f;
        e.gS = ball ? VssF0::fb : 0.5f; e.gF = ball ? VssF0::fr : 0.5f;
        e.inv_wsum = VssF0::inv_wsum; e.pad = 0.0f;
    }
    return t;
}
// In global memory, read through the read-only path (LDG.CONSTANT, L1-resident after the first touch), NOT
// in __constant__ memory: a user constant bank in the module added 0.1-0.3 us to EVERY kernel
// launch of the library (7.27 -> 7.57 us for the 4 096-match VSS-v0 step, which never reads the
// table; profiles/r1_logs/run39.log), far more than the table saves.
template <int RT> __device__ const PairTab<RT> g_pair_tab_f0 = make_pair_tab_f0<RT>();

template <int RT, class PP>
__device__ __forceinline__ void contacts_via_smem(const PP &P, Scene<RT> &s, uint32_t m, float4 *q, float4 *p0, const int pitch) {
    static_assert(RT >= 1 && RT <= 7, "3-bit body indices, 21 robot pairs in 63 bits");
    // q[b] straight from the register quads of the scene; the phase-start copy as (x, y) pairs and
    // the angular velocities as scalars (no register shuffling to build (x, y, omega, -) quads)
    constexpr bool PK = PP::packed;
    const int lane = threadIdx.x & 31;               // q / p0 point at this lane's float4 column of its warp's region
    float2 *const pxy = reinterpret_cast<float2 *>(p0 - lane) + lane;
    float *const pom = reinterpret_cast<float *>(p0 - lane + (RT + 1) * pitch / 2) + lane;
    q[0] = make_float4(s.bx, s.by, s.bvx, s.bvy); pxy[0] = make_float2(s.bx, s.by);
#pragma unroll
    for (int r = 0; r < RT; ++r) {
        q[(r + 1) * pitch] = make_float4(s.x[r], s.y[r], s.vx[r], s.vy[r]);
        pxy[(r + 1) * pitch] = make_float2(s.x[r], s.y[r]);
        pom[(r + 1) * pitch] = s.om[r];
    }
    constexpr uint64_t TI = rr_table<RT>(false), TJ = rr_table<RT>(true);
    do {
        const int p = __ffs((int)m) - 1;
        m &= m - 1;
        // normal points F -> S.  ball pairs: F = robot p, S = ball (body 0)
        int F, S;
        ContactTab<PP> T;
#ifndef RS_X_NOPAIRTAB
        if constexpr (PP::packed) {                      // VssF0P
            const float4 *const e = reinterpret_cast<const float4 *>(&g_pair_tab_f0<RT>.e[p]);
            const float4 e0 = __ldg(e), e1 = __ldg(e + 1), e2 = __ldg(e + 2);
            F = __float_as_int(e0.x); S = __float_as_int(e0.y);
            T.rs = e0.z; T.rc = e0.w; T.kn = e1.x; T.wS = e1.y; T.wF = e1.z; T.mu = e1.w; T.gS = e2.x; T.gF = e2.y;
        } else
#endif
        {
            const bool ball = p < RT;
            const int k3 = 3 * (p - RT);
            F = ball ? p + 1 : 1 + (int)((TI >> k3) & 7u); S = ball ? 0 : 1 + (int)((TJ >> k3) & 7u);
            T = contact_tab(P, ball);
        }
        const float2 pf = pxy[F * pitch], ps = pxy[S * pitch];
        const float omf = pom[F * pitch];
        float4 qf = q[F * pitch], qs = q[S * pitch];
        const float2 dd = v_sub<PK>(ps, pf);
        const float d2 = dd.x * dd.x + dd.y * dd.y;
        const float inv = rsqrtf(d2);
        const bool ok = d2 > 1e-12f;
        float2 n = v_mul<PK>(dd, bc2(inv));
        n.x = ok ? n.x : 1.0f; n.y = ok ? n.y : 0.0f;
        const float d = ok ? d2 * inv : 0.0f;
        const float pen = T.rs - d;
        const float2 rc = v_mul<PK>(n, bc2(T.rc));
        const float2 rel = v_sub<PK>(make_float2(qs.z, qs.w), make_float2(qf.z - omf * rc.y, qf.w + omf * rc.x));   // S - F's surface velocity
        const float vn = fminf(rel.x * n.x + rel.y * n.y, 0.0f);       // separating: every impulse below is +-0
        // ball pair: Jn = -(1 + e) vn / (wb + wr), dv = Jn w n.  robot pair: equal masses, dv = -(1 + e) vn / 2 n
        const float Jn = -T.kn * vn;
        const float vt = rel.y * n.x - rel.x * n.y;                    // tangent (-ny, nx)
        const float Jt = clampf(-vt * P.inv_wsum, -T.mu * Jn, T.mu * Jn);
        float2 vs = v_fma<PK>(n, bc2(Jn * T.wS), make_float2(qs.z, qs.w)), vf = v_fma<PK>(n, bc2(-Jn * T.wF), make_float2(qf.z, qf.w));
        const float jS = Jt * T.wS, jF = Jt * T.wF;
        vs.x -= jS * n.y; vs.y += jS * n.x; vf.x += jF * n.y; vf.y -= jF * n.x;
        const float2 xs = v_fma<PK>(n, bc2(pen * T.gS), make_float2(qs.x, qs.y)), xf = v_fma<PK>(n, bc2(-pen * T.gF), make_float2(qf.x, qf.y));
        qs = make_float4(xs.x, xs.y, vs.x, vs.y); qf = make_float4(xf.x, xf.y, vf.x, vf.y);
        q[F * pitch] = qf; q[S * pitch] = qs;
    } while (m);
    {
        const float4 b = q[0];
        s.bx = b.x; s.by = b.y; s.bvx = b.z; s.bvy = b.w;
    }
#pragma unroll
    for (int r = 0; r < RT; ++r) {
        const float4 b = q[(r + 1) * pitch];
        s.x[r] = b.x; s.y[r] = b.y; s.vx[r] = b.z; s.vy[r] = b.w;
    }
}

// The same for SSL worlds (lane-per-match kernels, R <= 7): the bit of a ball pair only says "inside the
// bounding circle" -- whether the ball touches the mouth-cut shape is decided here (ssl_mouth_contact, needs
// the robot's heading, published next to its angular velocity); robot pairs are discs as in VSS.  Run-time
// constants, scalar forms, same arithmetic and order as ball_robot<SSL> / robot_robot.  Replaces the
// register-static dispatch (contacts_static) that kept the 1 v 6 task kernel at 231 registers.
//   q[b] = (x, y, vx, vy) live;  pxy[b] = (x, y) at phase start;  pang[b] = (omega, theta)
template <int RT, class PP>
__device__ __forceinline__ void contacts_via_smem_ssl(const PP &P, Scene<RT> &s, uint32_t m, float4 *q, float4 *p0, const int pitch) {
    static_assert(RT >= 1 && RT <= 7, "3-bit body indices, 21 robot pairs in 63 bits");
    const int lane = threadIdx.x & 31;
    float2 *const pxy = reinterpret_cast<float2 *>(p0 - lane) + lane;
    float2 *const pang = reinterpret_cast<float2 *>(p0 - lane + (RT + 1) * pitch / 2) + lane;
    q[0] = make_float4(s.bx, s.by, s.bvx, s.bvy); pxy[0] = make_float2(s.bx, s.by);
#pragma unroll
    for (int r = 0; r < RT; ++r) {
        q[(r + 1) * pitch] = make_float4(s.x[r], s.y[r], s.vx[r], s.vy[r]);
        pxy[(r + 1) * pitch] = make_float2(s.x[r], s.y[r]);
        pang[(r + 1) * pitch] = make_float2(s.om[r], s.th[r]);
    }
    constexpr uint64_t TI = rr_table<RT>(false), TJ = rr_table<RT>(true);
    do {
        const int p = __ffs((int)m) - 1;
        m &= m - 1;
        const bool ball = p < RT;
        const int k3 = 3 * (p - RT);
        // normal points F -> S.  ball pairs: F = robot p, S = ball (body 0)
        const int F = ball ? p + 1 : 1 + (int)((TI >> k3) & 7u), S = ball ? 0 : 1 + (int)((TJ >> k3) & 7u);
        const float2 pf = pxy[F * pitch], ps = pxy[S * pitch];
        const float2 af = pang[F * pitch];
        float4 qf = q[F * pitch], qs = q[S * pitch];
        const float dx = ps.x - pf.x, dy = ps.y - pf.y;
        const float d2 = dx * dx + dy * dy;
        float nx, ny, pen, rcx = 0.0f, rcy = 0.0f;
        if (ball) {
            if (!ssl_mouth_contact(P, dx, dy, d2, af.y, nx, ny, pen, rcx, rcy)) continue;
        } else {
            float d = 0.0f; nx = 1.0f; ny = 0.0f;
            if (d2 > 1e-12f) { const float inv = rsqrtf(d2); nx = dx * inv; ny = dy * inv; d = d2 * inv; }
            pen = P.rs_rr - d;
        }
        const float sx = qf.z - af.x * rcy, sy = qf.w + af.x * rcx;      // F's surface velocity at the contact
        const float relx = qs.z - sx, rely = qs.w - sy;
        const float vn = relx * nx + rely * ny;
        if (vn < 0.0f) {
            if (ball) {
                const float Jn = -(1.0f + P.e_ball_rbt) * vn * P.inv_wsum;
                qs.z += Jn * P.wb * nx; qs.w += Jn * P.wb * ny;
                qf.z -= Jn * P.wr * nx; qf.w -= Jn * P.wr * ny;
                const float tx = -ny, ty = nx;
                const float vt = relx * tx + rely * ty;
                const float Jt = clampf(-vt * P.inv_wsum, -P.mu_ball_rbt * Jn, P.mu_ball_rbt * Jn);
                qs.z += Jt * P.wb * tx; qs.w += Jt * P.wb * ty;
                qf.z -= Jt * P.wr * tx; qf.w -= Jt * P.wr * ty;
            } else {
                const float Jw = -(1.0f + P.e_rbt_rbt) * vn * 0.5f;       // J * wr with equal masses
                qf.z -= Jw * nx; qf.w -= Jw * ny; qs.z += Jw * nx; qs.w += Jw * ny;
            }
        }
        const float gS = ball ? P.fb : 0.5f, gF = ball ? P.fr : 0.5f;
        qs.x += pen * gS * nx; qs.y += pen * gS * ny;
        qf.x -= pen * gF * nx; qf.y -= pen * gF * ny;
        q[F * pitch] = qf; q[S * pitch] = qs;
    } while (m);
    {
        const float4 b = q[0];
        s.bx = b.x; s.by = b.y; s.bvx = b.z; s.bvy = b.w;
    }
#pragma unroll
    for (int r = 0; r < RT; ++r) {
        const float4 b = q[(r + 1) * pitch];
        s.x[r] = b.x; s.y[r] = b.y; s.vx[r] = b.z; s.vy[r] = b.w;
    }
}

// register-static resolve: the masks are OR-reduced over the warp (REDUX) and the warp walks
// the set bits in ascending (= lexicographic) order, jumping to the register-static body of
// that pair; lanes without that contact fail the body's own distance test.
template <int KIND, int RT, class PP>
__device__ __forceinline__ void contacts_static(const PP &P, Scene<RT> &s, const uint32_t mask, const unsigned live) {
    uint32_t wm = __reduce_or_sync(live, mask);
#ifdef RS_X_NORESOLVE
    if (wm == 0xdeadbeefu) {
#else
    if (wm) {
#endif
        float cbx = 0.0f, cby = 0.0f;
        float cx[Cap<RT>::v], cy[Cap<RT>::v];
#pragma unroll
        for (int r = 0; r < RT; ++r) { cx[r] = 0.0f; cy[r] = 0.0f; }
        bool any = false;
        do {
            const int p = __ffs((int)wm) - 1;
            wm &= wm - 1;
            PairDispatch<KIND, RT, 0, RT + RT * (RT - 1) / 2, PP>::run(p, P, s, cbx, cby, cx, cy, any);
        } while (wm);
        s.bx += cbx; s.by += cby;
#pragma unroll
        for (int r = 0; r < RT; ++r) { s.x[r] += cx[r]; s.y[r] += cy[r]; }
    }
}

// commands -> drive targets.  VSS: cmd = (wl, wr) rad/s (rsim.py:100-101)
template <bool PK = false>
__device__ __forceinline__ void vss_target(const DevParams &P, float wl, float wr, float &tf, float &tw) {
    wl = clampf(wl, -P.wmax, P.wmax); wr = clampf(wr, -P.wmax, P.wmax);
    const float2 t = v_mul<PK>(make_float2(wl + wr, wr - wl), make_float2(P.rw * 0.5f, P.rw * P.inv_2b));
    tf = t.x; tw = t.y;
}
// SSL: cmd = 8 floats (rsim.py:137-153)
__device__ __forceinline__ void ssl_target(const DevParams &P, const float (&cmd)[8], float &tf,
                                           float &tl, float &tw, float &kick, bool &drib) {
    float sp[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        float w;
        if (cmd[0] != 0.0f) w = cmd[1 + i];
        else w = (P.J[i][0] * cmd[1] + P.J[i][1] * cmd[2] + P.J[i][2] * cmd[3]) * P.inv_rw;
        sp[i] = clampf(w, -P.wmax, P.wmax) * P.rw;
    }
    tf = P.Jp[0][0] * sp[0] + P.Jp[0][1] * sp[1] + P.Jp[0][2] * sp[2] + P.Jp[0][3] * sp[3];
    tl = P.Jp[1][0] * sp[0] + P.Jp[1][1] * sp[1] + P.Jp[1][2] * sp[2] + P.Jp[1][3] * sp[3];
    tw = P.Jp[2][0] * sp[0] + P.Jp[2][1] * sp[1] + P.Jp[2][2] * sp[2] + P.Jp[2][3] * sp[3];
    kick = fminf(cmd[5], P.kick_max);
    drib = cmd[7] != 0.0f;
}

// ---------------------------------------------------------------- one control step
template <int KIND, int RT, class PP>
__device__ __forceinline__ void physics_step(const PP &P, Scene<RT> &s, const Drive<RT> &d,
                                             const unsigned live /* lanes of this warp that call */,
                                             float4 *cq = nullptr /* this lane's column of the contact scratch */,
                                             float4 *cp0 = nullptr, const int cpitch = 0) {
    constexpr bool PK = PP::packed;
    const int R = RT > 0 ? RT : P.n_robots;
    const float h = P.h;
    uint32_t kicked = 0;
    if constexpr (KIND == RS_KIND_SSL) {
        // kick: once per control step, robots in row order
#pragma unroll
        for (int r = 0; r < R; ++r) {
            if (d.kick[r] > 0.0f) {
                float sn, cs;
                __sincosf(s.th[r], &sn, &cs);
                if (touching(P, s.x[r], s.y[r], cs, sn, s.bx, s.by)) {
                    s.bvx = cs * d.kick[r]; s.bvy = sn * d.kick[r];
                    kicked |= 1u << r;
                }
            }
        }
    }
#pragma unroll 1
    for (int k = 0; k < RS_X_SUBSTEPS; ++k) {
        int holder = -1; float hx = 0.0f, hy = 0.0f;
        // (a) drive, (b) dribbler latch
#ifndef RS_X_NODRIVE
#pragma unroll
        for (int r = 0; r < R; ++r) {
            float sn, cs;
            __sincosf(s.th[r], &sn, &cs);
            // world -> robot frame on the (vx, vy) pair
            const float2 t = v_mul<PK>(make_float2(s.vx[r], s.vy[r]), bc2(cs));
            float vf = fmaf(sn, s.vy[r], t.x), vl = fmaf(-sn, s.vx[r], t.y);
            if constexpr (KIND == RS_KIND_VSS) {
                // v + clamp(t - v, -a, a) == clamp(t, v - a, v + a): independent adds, then min / max
                const float2 f = make_float2(vf, vl), a = make_float2(P.acc_fwd_h, P.acc_lat_h);
                const float2 hi = v_add<PK>(f, a), lo = v_sub<PK>(f, a);
                vf = fmaxf(fminf(d.tf[r], hi.x), lo.x); vl = fmaxf(fminf(d.tl[r], hi.y), lo.y);
            } else {
                const float df = d.tf[r] - vf, dl = d.tl[r] - vl;
                const float n2 = df * df + dl * dl;
                float sc = 1.0f;
                if (n2 > P.acc_fwd_h * P.acc_fwd_h) sc = P.acc_fwd_h * rsqrtf(n2);
                vf += df * sc; vl += dl * sc;
            }
            s.om[r] = fmaxf(fminf(d.tw[r], s.om[r] + P.acc_ang_h), s.om[r] - P.acc_ang_h);
            const float2 u = v_mul<PK>(make_float2(cs, sn), bc2(vf));        // robot -> world frame
            s.vx[r] = fmaf(-sn, vl, u.x); s.vy[r] = fmaf(cs, vl, u.y);
            if constexpr (KIND == RS_KIND_SSL) {
                if (holder < 0 && ((d.drib >> r) & 1u) && !((kicked >> r) & 1u)) {
                    const float dx = s.bx - s.x[r], dy = s.by - s.y[r];
                    const float lx = cs * dx + sn * dy, ly = -sn * dx + cs * dy;
                    if (touching_local(P, lx, ly)) { holder = r; hx = lx; hy = ly; }
                }
            }
        }
#endif
        // (c) ball rolling friction
        if (holder < 0) {
            const float sp2 = s.bvx * s.bvx + s.bvy * s.bvy;
            const float sc = fmaxf(1.0f - P.ball_decel_h * rsqrtf(sp2 + 1e-12f), 0.0f);
            const float2 bv = v_mul<PK>(make_float2(s.bvx, s.bvy), bc2(sc));
            s.bvx = bv.x; s.bvy = bv.y;
        }
        // (d) integrate
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const float2 p = v_fma<PK>(make_float2(s.vx[r], s.vy[r]), bc2(h), make_float2(s.x[r], s.y[r]));
            s.x[r] = p.x; s.y[r] = p.y;
            s.th[r] += s.om[r] * h;      // |omega| dt < 2 pi: wrapped once, after the last sub-step
        }
        if (holder < 0) {
            const float2 p = v_fma<PK>(make_float2(s.bvx, s.bvy), bc2(h), make_float2(s.bx, s.by));
            s.bx = p.x; s.by = p.y;
        }
        else {
#pragma unroll
            for (int r = 0; r < R; ++r) {
                if (r == holder) {
                    float sn, cs;
                    __sincosf(s.th[r], &sn, &cs);
                    const float ox = cs * hx - sn * hy, oy = sn * hx + cs * hy;
                    s.bx = s.x[r] + ox; s.by = s.y[r] + oy;
                    s.bvx = s.vx[r] - s.om[r] * oy; s.bvy = s.vy[r] + s.om[r] * ox;
                }
            }
        }
        // (e) pairs, lexicographic; contacts detected on the positions at phase start,
        // velocity impulses applied sequentially, position corrections applied on top of them.
        // A contact is rare per lane (~7 % of the matches have one at any time,
        // tests/contact_stats.py) but in a warp of 32 matches "some lane has one" holds for
        // ~95 % of the sub-steps, spread over the pairs.  Up to 28 pairs (R <= 7):
        //   1. detection is straight-line: one FFMA pair and one funnel shift per pair (the sign
        //      bit of d^2 - rs^2 is the mask bit), no compare, no branch, full ILP;
        //   2. resolve: VSS task kernels -- each lane with a contact resolves its own pairs
        //      through shared memory (contacts_via_smem); everything else -- the warp walks
        //      the OR of the masks and jumps to register-static bodies (contacts_static).
#ifdef RS_X_NOPAIRS
        if constexpr (false) {
#else
        if constexpr (RT > 0 && RT <= 7) {
#endif
            // the sign bit of (d^2 - rs^2) enters the mask from the right, pairs visited in
            // descending order so that pair 0 ends in bit 0; two chains for ILP
            uint32_t mask = 0, mrr = 0;
#pragma unroll
            for (int i = R - 2; i >= 0; --i) {
#pragma unroll
                for (int j = R - 1; j > i; --j) {
                    const float2 dd = v_sub<PK>(make_float2(s.x[j], s.y[j]), make_float2(s.x[i], s.y[i]));
                    const float dx = dd.x, dy = dd.y;
                    mrr = __funnelshift_l(__float_as_uint(fmaf(dx, dx, fmaf(dy, dy, -P.rs_rr2))), mrr, 1);
                }
            }
#pragma unroll
            for (int r = R - 1; r >= 0; --r) {
                const float2 dd = v_sub<PK>(make_float2(s.bx, s.by), make_float2(s.x[r], s.y[r]));
                const float dx = dd.x, dy = dd.y;
                mask = __funnelshift_l(__float_as_uint(fmaf(dx, dx, fmaf(dy, dy, -P.rs_br2))), mask, 1);
            }
            mask |= mrr << R;
            if (cpitch > 0) {                            // constants once inlined
                if constexpr (KIND == RS_KIND_VSS) { if (mask) contacts_via_smem<RT>(P, s, mask, cq, cp0, cpitch); }
                else { if (mask) contacts_via_smem_ssl<RT>(P, s, mask, cq, cp0, cpitch); }
            } else {
                contacts_static<KIND, RT>(P, s, mask, live);
            }
#ifdef RS_X_NOPAIRS
        } else if (false) {
#else
        } else {
#endif
            float cbx = 0.0f, cby = 0.0f;
            float cx[Cap<RT>::v], cy[Cap<RT>::v];
#pragma unroll
            for (int r = 0; r < R; ++r) { cx[r] = 0.0f; cy[r] = 0.0f; }
            bool any = false;
#pragma unroll
            for (int r = 0; r < R; ++r)
                ball_robot<KIND>(P, s.bx, s.by, s.bvx, s.bvy, s.x[r], s.y[r], s.th[r], s.vx[r],
                                 s.vy[r], s.om[r], cbx, cby, cx[r], cy[r], any);
#pragma unroll
            for (int i = 0; i < R; ++i) {
#pragma unroll
                for (int j = i + 1; j < R; ++j)
                    robot_robot(P, s.x[i], s.y[i], s.vx[i], s.vy[i], s.x[j], s.y[j], s.vx[j],
                                s.vy[j], cx[i], cy[i], cx[j], cy[j], any);
            }
            if (any) {
                s.bx += cbx; s.by += cby;
#pragma unroll
                for (int r = 0; r < R; ++r) { s.x[r] += cx[r]; s.y[r] += cy[r]; }
            }
        }
        // (f) walls
#ifndef RS_X_NOWALLS
        if constexpr (KIND == RS_KIND_VSS) {
            vss_walls(P, P.ball_r, P.e_ball_wall, s.bx, s.by, s.bvx, s.bvy);
#pragma unroll
            for (int r = 0; r < R; ++r) vss_walls(P, P.rbt_r, P.e_rbt_wall, s.x[r], s.y[r], s.vx[r], s.vy[r]);
        } else {
            ssl_walls(P, P.ball_r, P.e_ball_wall, s.bx, s.by, s.bvx, s.bvy);
#pragma unroll
            for (int r = 0; r < R; ++r) ssl_walls(P, P.rbt_r, P.e_rbt_wall, s.x[r], s.y[r], s.vx[r], s.vy[r]);
        }
#endif
    }
#pragma unroll
    for (int r = 0; r < R; ++r) {
        float a = s.th[r];
        // more than one turn per control step needs time_step_ms > 100 at the motor limits
        if (fabsf(a) > 3.0f * RS_PI_F) a = fmaf(-rintf(a * (0.5f / RS_PI_F)), 2.0f * RS_PI_F, a);
        s.th[r] = wrap_pi(a);
    }
}

// ---------------------------------------------------------------- state I/O (SoA in HBM)
struct StatePtrs {
    float4 *body;     // [R+1][Np]
    float2 *ang;      // [R][Np]
    float2 *ou;       // [R-1][Np]
    float *prev;      // [Np]
    int *steps;       // [Np]
    float *info;      // [RS_SSL_INFO][Np]
    uint32_t *aux;    // [2 + RS_SSL_INFO][Np] = prev, steps, info seen as one array of task words (rs_lanes.cuh)
    int n;            // envs
    int np;           // padded env count (array pitch)
};

// State loads bypass L1 (ld.global.cg): every state word is read exactly once per step, and a
// step kernel that overlaps its predecessor (tile_acquire below) must never hit a line an
// earlier launch left in this SM's L1.
template <int RT>
__device__ __forceinline__ void load_scene(const DevParams &P, const StatePtrs &S, int e, Scene<RT> &s) {
    const int R = RT > 0 ? RT : P.n_robots;
    const float4 b = __ldcg(S.body + e);
    s.bx = b.x; s.by = b.y; s.bvx = b.z; s.bvy = b.w;
#pragma unroll
    for (int r = 0; r < R; ++r) {
        const float4 q = __ldcg(S.body + (size_t)(r + 1) * S.np + e);
        const float2 a = __ldcg(S.ang + (size_t)r * S.np + e);
        s.x[r] = q.x; s.y[r] = q.y; s.vx[r] = q.z; s.vy[r] = q.w; s.th[r] = a.x; s.om[r] = a.y;
    }
}
template <int RT>
__device__ __forceinline__ void store_scene(const DevParams &P, const StatePtrs &S, int e, const Scene<RT> &s) {
    const int R = RT > 0 ? RT : P.n_robots;
    S.body[e] = make_float4(s.bx, s.by, s.bvx, s.bvy);
#pragma unroll
    for (int r = 0; r < R; ++r) {
        S.body[(size_t)(r + 1) * S.np + e] = make_float4(s.x[r], s.y[r], s.vx[r], s.vy[r]);
        S.ang[(size_t)r * S.np + e] = make_float2(s.th[r], s.om[r]);
    }
}

// ---------------------------------------------------------------- smem tile -> global rows
// Each lane has written its `row_floats` outputs to smem row `tid`; the 32 rows of a warp
// are one contiguous span of global memory, so each warp ships its own span as ONE TMA
// bulk copy (cp.async.bulk.global.shared::cta, SASS UBLKCP) -- no CTA-wide barrier --
// when the span is 16-byte granular, else as a coalesced warp copy.
// gdst / stile point at row 0 of the CALLING WARP; rows = valid rows of this warp.
// bulk = false (launches that take part in the step-overlap protocol): always the warp copy -- the tile
// hand-over at the end of the kernel orders generic-proxy stores (fence + flag), not the async proxy's.
__device__ __forceinline__ void warp_tile_store(float *gdst, const float *stile, int rows, int row_floats, const bool bulk = true) {
    const uint32_t bytes = (uint32_t)rows * (uint32_t)row_floats * 4u;
    const int lane = threadIdx.x & 31;
    if (rows <= 0) return;
    if (bulk && (bytes & 15u) == 0u && ((reinterpret_cast<uintptr_t>(gdst) & 15u) == 0u)) {
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncwarp();
        if (lane == 0) {
            const uint32_t saddr = (uint32_t)__cvta_generic_to_shared(stile);
            asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
                         :: "l"(gdst), "r"(saddr), "r"(bytes) : "memory");
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
        }
    } else {
        __syncwarp();
        const int total = rows * row_floats;
        for (int i = lane; i < total; i += 32) gdst[i] = stile[i];
    }
}

// World step counter t (Philox counter word 1) in device memory, so that a captured CUDA
// graph replays with fresh counters.  One copy per group of RS_CTR_GROUP consecutive matches,
// all copies equal (every step kernel advances every match): the first lane of a group
// reads its copy, the group gets it by shuffle, and the same lane stores t + 1 at the end
// of the kernel.  No atomics, no CTA barrier, no hot address (a single shared counter with a
// "last CTA bumps it" ticket cost 8 % of the stall samples, profiles/r1_steady_final.txt),
// and no race: reader and writer of a copy are the same thread.
//   GL = lanes per counter group = RS_CTR_GROUP x lanes per match (a divisor of 32).
#define RS_CTR_GROUP 4
#define RS_T_BUSY 0x80000000u      // bit 31 of a counter copy: a step holds the tile that starts at this copy (step overlap)
#define RS_T_MASK 0x7fffffffu      // the counter proper: 31 bits
// The counter words are the only input of the Philox / Box-Muller block, which is meant to
// run under the latency of the state loads: they are loaded and stored with an L2
// evict_last policy so that this 64 KB array survives in L2 between two steps of a world
// while the state of other worlds streams through (an L2 hit returns ~1 us before HBM does).
__device__ __forceinline__ uint64_t l2_keep_policy() {
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
template <int GL>
__device__ __forceinline__ uint32_t step_counter_read(const uint32_t *ctr, const int e, const unsigned live = 0xffffffffu) {
    const int lane = threadIdx.x & 31;
    uint32_t t = 0u;
    if ((lane & (GL - 1)) == 0) {
        asm volatile("ld.global.cg.L2::cache_hint.u32 %0, [%1], %2;" : "=r"(t) : "l"(ctr + e / RS_CTR_GROUP), "l"(l2_keep_policy()) : "memory");
    }
    return __shfl_sync(live, t, lane & ~(GL - 1)) & RS_T_MASK;      // a group's first lane is live whenever any of its lanes is
}
// skip_lane0: the warp's first copy is the tile lock of the step-overlap protocol and is written by step_end
template <int GL>
__device__ __forceinline__ void step_counter_bump(uint32_t *ctr, const int e, const uint32_t t, const bool skip_lane0 = false) {
    if (((threadIdx.x & 31) & (GL - 1)) == 0 && !(skip_lane0 && (threadIdx.x & 31) == 0)) {
        asm volatile("st.global.L2::cache_hint.u32 [%0], %1, %2;" :: "l"(ctr + e / RS_CTR_GROUP), "r"((t + 1u) & RS_T_MASK), "l"(l2_keep_policy()) : "memory");
    }
}

// Programmatic dependent launch (sm_90+): a step kernel launched with the
// programmaticStreamSerialization attribute may start while its predecessor in the stream is
// still draining.  Everything before pdl_wait() must touch only kernel parameters; after it
// the predecessor's writes are visible.  pdl_release() lets the NEXT launch begin its own
// pre-wait part (launch latency, parameter fetch, CTA placement hide under this kernel).
// Both are no-ops for launches without the attribute.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_release() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

// ---------------------------------------------------------------- step-to-step overlap
// A step kernel normally begins with griddepcontrol.wait: nothing of step k+1 runs before the
// whole grid of step k has finished AND flushed, so launch ramp, state loads and the store
// drain of consecutive steps are serialised (~7 of 15.5 us at 65 536 matches).  But match i at
// step k+1 depends only on match i at step k.  With RS_OPT_STEP_OVERLAP the dependency is
// tracked per TILE = what one warp of a step kernel works on (32 matches in the lane-per-match
// kernels, 32 / L in the lane-per-body ones).  The tile's lock is bit 31 of the FIRST step-counter
// copy of its matches (every warp covers whole counter groups), so taking the tile and reading
// the step counter are one L2 round trip:
//   start of a warp:  lane 0: old = atomicOr(copy, BUSY).  BUSY clear in old: the tile is ours and
//                     old is the step counter t.  Else a previous step still holds it: poll with
//                     plain L2 loads until BUSY clears, retry.  When all warps of the CTA hold their
//                     tiles the CTA executes griddepcontrol.launch_dependents -- so step k+2 cannot
//                     launch before every tile of step k has been handed to step k+1;
//   end of a warp:    the other counter copies get t + 1 as always; __syncwarp; lane 0 writes
//                     t + 1 (BUSY clear) to the lock copy with a release store (cumulative: covers
//                     the stores of the whole warp).
// Visibility.  The producer's release makes its stores visible at L2 before the lock word flips.
// The consumer reads the lock with an L2 atomic and everything a predecessor may have written --
// state, task words -- with loads that go around L1 (ld.global.cg, issued only after the atomic
// has returned: the loop exit depends on its value), so no stale line of this SM's non-coherent L1
// can be observed and no L1 invalidation is needed.  (ld.acquire.gpu / __threadfence() each
// compile to a CCTL.IVALL -- the whole L1 -- and a separate flag word costs a second L2 round
// trip before the first state load can issue: 7 % + 4 % of the warp time at 1 M matches,
// profiles/r2_overlap.txt.)
// No deadlock: a programmatic launch starts only after EVERY CTA of the predecessor has
// executed launch_dependents, i.e. is resident, so a polling warp always waits for a running
// one.  The poll is bounded anyway (RS_SPIN_LIMIT polls, then the error word is bumped and the
// warp goes on): a protocol error must not hang the GPU.
#ifndef RS_SPIN_LIMIT
#define RS_SPIN_LIMIT (1 << 19)
#endif
struct StepTile {
    uint32_t *lock;     // the lock copy of this warp's tile; null: the launch does not take part in the protocol
    uint32_t t;         // world step counter of this step (valid when lock != null)
};
// Start of every task step kernel.  ctr: the step counter copies; err: error counter of the protocol, null =
// the launch does not take part; w0: first match of this warp.
//   chain == 0: grid-wide wait first (griddepcontrol.wait), then take the tile (always free)
//   chain != 0: no grid-wide wait here, the tile is the dependency
template <int BS>
__device__ __forceinline__ StepTile step_begin(uint32_t *ctr, uint32_t *err, const int chain, const int w0) {
    StepTile T;
    T.lock = nullptr; T.t = 0u;
    if (chain == 0) pdl_wait();
    if (err) {
        T.lock = ctr + w0 / RS_CTR_GROUP;
        uint32_t old = 0u;
        if ((threadIdx.x & 31) == 0) {
            int spins = 0;
            for (;;) {
                asm volatile("atom.relaxed.gpu.global.or.b32 %0, [%1], %2;" : "=r"(old) : "l"(T.lock), "r"(RS_T_BUSY) : "memory");
                if (!(old & RS_T_BUSY)) break;
                // held by the previous step: poll with plain L2 reads (a thousand warps spinning on atomics
                // would queue up in front of the very store that frees the tile), then try again
                uint32_t v;
                do {
                    __nanosleep(64);
                    asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(T.lock) : "memory");
                } while ((v & RS_T_BUSY) && ++spins <= RS_SPIN_LIMIT);
                if (spins > RS_SPIN_LIMIT) { atomicAdd(err, 1u); old = v; break; }   // timed out: a dead predecessor, or one world on two streams
            }
        }
        T.t = __shfl_sync(0xffffffffu, old, 0) & RS_T_MASK;
        // the CTA's trigger must not fire before ALL its warps hold their tiles (a CTA counts as triggered
        // once any of its threads has executed launch_dependents): otherwise step k+2 could start and
        // wait for a tile that step k+1 has not taken yet
        if (BS > 32) __syncthreads();
    }
    pdl_release();
    return T;
}
// End of the kernel, after the warp's last global store.  A release store (MEMBAR.ALL.GPU + ST.STRONG.GPU),
// cumulative over the warp through the __syncwarp.  Not __threadfence() + store: that compiles to MEMBAR.SC.GPU
// ... CCTL.IVALL, a sequentially consistent fence plus an invalidation of the whole L1 that nothing here needs.
__device__ __forceinline__ void step_end(const StepTile &T) {
    if (!T.lock) return;
    __syncwarp();
    if ((threadIdx.x & 31) == 0)
        asm volatile("st.release.gpu.global.u32 [%0], %1;" :: "l"(T.lock), "r"((T.t + 1u) & RS_T_MASK) : "memory");
}
