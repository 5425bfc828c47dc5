/*
 * rs_oracle.c -- CPU ORACLE (TEST INFRASTRUCTURE, NOT PRODUCT CODE).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * `--impl reference` legs may load this file's library.  The product path
 * (rsoccer_b200/) never imports, links or executes anything under oracle/.
 *
 * What it restates, in plain scalar fp64 C, array-of-structs, one env at a time:
 *
 *  (1) The physics of one `robosim.VSS.step` / `robosim.SSL.step` call as defined
 *      by the 2-D model of DESIGN.md section 3 (call sites: reference
 *      rsoccer_gym/Simulators/rsim.py:102 and :155; reset :38; get_state :105,
 *      :158; field params :50).
 *      PARITY UNPINNED for this part: the reference's arithmetic lives in the
 *      third-party package rc-robosim >= 1.2.0 (setup.py:15) whose source is not
 *      under /root/reference and which cannot be installed here; the reference
 *      tree holds no golden vector for it (its only test is Utils/kdtree_test.py).
 *      The model is anchored on the reference's call sites, wire layout
 *      (Entities/Frame.py:17-93) and the behavioural contract K1-K12 of SURVEY.md
 *      appendix C (tests/test_contract.py).
 *
 *  (2) The task logic of the in-tree reference envs -- command conversion,
 *      observation, reward, done -- which IS pinned: tests/golden/*.npz are
 *      produced by running the unmodified reference classes
 *      (rsoccer_gym/vss/env_vss/vss_gym.py, ssl/ssl_hw_challenge/
 *      static_defenders.py, contested_possession.py) with this oracle behind a
 *      `robosim` shim (tests/golden/make_golden.py), and tests/test_golden.py
 *      checks the functions below against them.
 *
 * Each function cites the reference lines it follows.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include "../include/rs_spec.h"

#ifdef _OPENMP
#include <omp.h>
#endif

#define PI 3.14159265358979323846
#define DEG (180.0 / PI)

typedef struct { double x, y, vx, vy; } o_ball;
typedef struct { double x, y, th, vx, vy, om; } o_robot;
typedef struct { double tf, tl, tw, kick, drib; int kicked; } o_target;

typedef struct orc_world {
    rs_params p;
    int n;
    int n_threads;
    o_ball *ball;          /* [n] */
    o_robot *rob;          /* [n][R] */
    /* task state */
    double *ou;            /* [n][2*(R-1)]   OU process state, vss_gym.py:75-79 */
    double *prev_pot;      /* [n]            previous_ball_potential, vss_gym.py:70 */
    int *has_prev;         /* [n] */
    int *steps;            /* [n]            episode step counter (TimeLimit) */
    double *info;          /* [n][RS_SSL_INFO] reward_shaping_total */
    double *margin;        /* [n] smallest distance to a discrete decision boundary in the last call */
    uint64_t seed;
    uint64_t t;            /* world step counter (Philox counter word 1) */
    int64_t env_offset;    /* global id of env 0 (multi-GPU shards) */
} orc_world;

/* ------------------------------------------------------------------------- */
/* Philox4x32-10 (Salmon et al., SC'11; Random123 v1.09 constants).  Checked
 * against the Random123 known-answer vectors in tests/test_oracle.py.         */
static void philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) {
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
void orc_philox(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) {
    philox4x32_10(ctr, key, out);
}

/* stream of u32 for (env, t, stream): counter = (env, t, stream, j), j = 0,1,2.. */
typedef struct { uint32_t ctr[4], key[2], buf[4]; int idx; } o_rng;
static void rng_init(o_rng *g, uint64_t seed, uint64_t env, uint64_t t, uint32_t stream) {
    g->key[0] = (uint32_t)seed; g->key[1] = (uint32_t)(seed >> 32);
    g->ctr[0] = (uint32_t)env; g->ctr[1] = (uint32_t)t;
    g->ctr[2] = stream;             /* env ids and step counters are < 2^32 by contract */
    g->ctr[3] = 0; g->idx = 4;
}
static uint32_t rng_u32(o_rng *g) {
    if (g->idx == 4) { philox4x32_10(g->ctr, g->key, g->buf); g->ctr[3]++; g->idx = 0; }
    return g->buf[g->idx++];
}
/* 24-bit uniform in (0,1), exactly representable in fp32 */
static double u01(uint32_t x) { return ((double)(x >> 8) + 0.5) * (1.0 / 16777216.0); }
static double rng_uniform(o_rng *g, double a, double b) { return a + (b - a) * u01(rng_u32(g)); }

/* ------------------------------------------------------------------------- */
static double wrap_pi(double a) {
    if (a > PI) a -= 2.0 * PI; else if (a <= -PI) a += 2.0 * PI;
    return a;
}
static double clampd(double x, double lo, double hi) { return x < lo ? lo : (x > hi ? hi : x); }
/* a contact normal n = delta/|delta| amplifies an input perturbation eps to eps/|delta|;
 * scaled so that |delta| < 5 mm (only reachable by injecting overlapping bodies, the
 * dynamics keep centres >= one radius apart) counts as ill conditioned at margin 2e-5 */
#define NORMAL_COND 4e-3
static void note(double *margin, double m) { m = fabs(m); if (m < *margin) *margin = m; }
/* A contact / wall decision is discontinuous only in the velocity impulse it triggers: the
 * position correction it adds is the penetration, which vanishes AT the boundary.  So a decision
 * decided by a hair matters only when the impulse on the other side of it is not negligible:
 * `jump` = the normal velocity change the contact would apply [m/s].  A robot resting on a wall
 * (velocity already zeroed, centre exactly on the limit) or two bodies sliding apart are decided
 * by 0 m and yet give the same result either way -- they are not ill conditioned. */
#define JUMP_TOL 2e-5
static void note_jump(double *margin, double m, double jump) { if (fabs(jump) > JUMP_TOL) note(margin, m); }

/* kicker "touching" box, in the robot frame (grSim isTouchingBall analogue).
 * K7: true with the ball 0.1 m ahead (dribbling.py:193-195, contested_possession.py:224-225). */
static int touching(const rs_params *p, const o_robot *r, const o_ball *b, double *margin) {
    double c = cos(r->th), s = sin(r->th);
    double dx = b->x - r->x, dy = b->y - r->y;
    double lx = c * dx + s * dy, ly = -s * dx + c * dy;
    double gx = p->kick_reach - fabs(lx - p->kick_centre);
    double gy = p->kick_half_width - fabs(ly);
    if (margin) {
        /* only the boundary of the box matters: distance to it */
        if (gx > 0 && gy > 0) note(margin, gx < gy ? gx : gy);
        else if (gx <= 0 && gy > 0) note(margin, gx);
        else if (gy <= 0 && gx > 0) note(margin, gy);
        else note(margin, sqrt(gx * gx + gy * gy));
    }
    return gx > 0 && gy > 0;
}

/* robot <-> ball contact.  shape: disc radius R, cut by the chord x_local = dk when
 * kind == SSL (flat kicker mouth; contested_possession.py:224-225 starts the ball
 * 0.1 m ahead, inside the bounding circle).  Writes the unit normal robot->ball,
 * penetration, and the contact point offset from the robot centre (world frame). */
static int ball_robot_contact(const rs_params *p, const o_robot *r, const o_ball *b,
                              double *nx, double *ny, double *pen, double *rcx, double *rcy,
                              double *margin) {
    double dx = b->x - r->x, dy = b->y - r->y;
    double R = p->rbt_radius, rb = p->ball_radius;
    double d2 = dx * dx + dy * dy;
    if (p->kind == RS_KIND_VSS) {
        double rs = R + rb;
        {   /* discs: the lever arm is parallel to the normal, omega does not enter the normal velocity */
            double d = sqrt(d2), vn = d > 0 ? ((b->vx - r->vx) * dx + (b->vy - r->vy) * dy) / d : -1.0;
            note_jump(margin, d - rs, vn < 0 ? (1.0 + p->e_ball_rbt) * vn : 0.0);
        }
        if (d2 >= rs * rs) return 0;
        double d = sqrt(d2);
        note(margin, d * NORMAL_COND);
        if (d2 > 1e-12) { *nx = dx / d; *ny = dy / d; } else { *nx = 1.0; *ny = 0.0; d = 0.0; }
        *pen = rs - d;
        *rcx = *nx * R; *rcy = *ny * R;
        return 1;
    }
    if (d2 >= (R + rb) * (R + rb)) {
        double d = sqrt(d2), vn = ((b->vx - r->vx) * dx + (b->vy - r->vy) * dy) / d;
        note_jump(margin, d - (R + rb), vn < 0 ? (1.0 + p->e_ball_rbt) * vn : 0.0);
        return 0;
    }
    double c = cos(r->th), s = sin(r->th);
    double lx = c * dx + s * dy, ly = -s * dx + c * dy;
    double dk = p->rbt_distance_center_kicker, ch = p->mouth_half_chord;
    double bn = sqrt(d2);
    double q1x = lx, q1y = ly;
    if (bn > R) { q1x = lx * (R / bn); q1y = ly * (R / bn); }
    double qx, qy;
    if (q1x <= dk) { qx = q1x; qy = q1y; }
    else { qx = dk; qy = clampd(ly, -ch, ch); }
    double ex = lx - qx, ey = ly - qy;
    double e2 = ex * ex + ey * ey;
    if (e2 > 1e-12) {   /* the impulse this contact applies (or would apply): surface velocity at the closest point */
        double e = sqrt(e2), wnx = (c * ex - s * ey) / e, wny = (s * ex + c * ey) / e;
        double wrx = c * qx - s * qy, wry = s * qx + c * qy;
        double vn = (b->vx - (r->vx - r->om * wry)) * wnx + (b->vy - (r->vy + r->om * wrx)) * wny;
        note_jump(margin, e - rb, vn < 0 ? (1.0 + p->e_ball_rbt) * vn : 0.0);
    } else note(margin, sqrt(e2) - rb);
    if (e2 >= rb * rb) return 0;
    double lnx, lny;
    if (e2 > 1e-12) {
        double e = sqrt(e2);
        note(margin, e * NORMAL_COND);
        lnx = ex / e; lny = ey / e; *pen = rb - e;
    } else {                     /* ball centre inside the robot shape: least-penetration exit */
        double pr = R - bn, pf = dk - lx;
        note(margin, 0.0);       /* unreachable by the dynamics; never well conditioned */
        if (pf < pr) { lnx = 1.0; lny = 0.0; *pen = rb + pf; }
        else {
            if (bn > 1e-9) { lnx = lx / bn; lny = ly / bn; } else { lnx = 1.0; lny = 0.0; }
            *pen = rb + pr;
        }
    }
    *nx = c * lnx - s * lny; *ny = s * lnx + c * lny;
    *rcx = c * qx - s * qy; *rcy = s * qx + c * qy;
    return 1;
}

/* walls for one body, mirrored quadrant */
static void walls(const rs_params *p, double r, double e, double *x, double *y,
                  double *vx, double *vy, double *margin) {
    double sx = *x < 0 ? -1.0 : 1.0, sy = *y < 0 ? -1.0 : 1.0;
    double ax = fabs(*x), ay = fabs(*y);
    double avx = sx * *vx, avy = sy * *vy;
    if (ax + r > p->x_near) {
        for (int k = 0; k < p->n_box; ++k) {
            const double *bx = p->box[k];
            double qx = clampd(ax, bx[0], bx[2]), qy = clampd(ay, bx[1], bx[3]);
            double dx = ax - qx, dy = ay - qy;
            double d2 = dx * dx + dy * dy;
            if (d2 > 1e-12) {
                double d = sqrt(d2), vn = (avx * dx + avy * dy) / d;
                note_jump(margin, d - r, vn < 0 ? (1.0 + e) * vn : 0.0);
            } else note(margin, sqrt(d2) - r);
            if (d2 >= r * r) continue;
            double nx, ny, pen;
            if (d2 > 1e-12) {
                double d = sqrt(d2);
                note(margin, d * NORMAL_COND);
                nx = dx / d; ny = dy / d; pen = r - d;
            } else {
                note(margin, 0.0);
                double fxl = ax - bx[0], fxh = bx[2] - ax, fyl = ay - bx[1], fyh = bx[3] - ay;
                double m = fxl; nx = -1.0; ny = 0.0;
                if (fxh < m) { m = fxh; nx = 1.0; ny = 0.0; }
                if (fyl < m) { m = fyl; nx = 0.0; ny = -1.0; }
                if (fyh < m) { m = fyh; nx = 0.0; ny = 1.0; }
                pen = r + m;
            }
            ax += pen * nx; ay += pen * ny;
            double vn = avx * nx + avy * ny;
            if (vn < 0) { avx -= (1.0 + e) * vn * nx; avy -= (1.0 + e) * vn * ny; }
        }
    }
    note_jump(margin, ax - (p->x_out - r), avx > 0 ? (1.0 + e) * avx : 0.0);
    if (ax > p->x_out - r) { ax = p->x_out - r; if (avx > 0) avx = -e * avx; }
    note_jump(margin, ay - (p->y_out - r), avy > 0 ? (1.0 + e) * avy : 0.0);
    if (ay > p->y_out - r) { ay = p->y_out - r; if (avy > 0) avy = -e * avy; }
    *x = sx * ax; *y = sy * ay; *vx = sx * avx; *vy = sy * avy;
}

/* commands -> per-robot drive targets (robot frame) */
static void drive_targets(const rs_params *p, const double *cmd, o_target *t) {
    double wmax = p->wheel_max_rad_s, rw = p->rbt_wheel_radius;
    if (p->kind == RS_KIND_VSS) {
        /* rsim.py:100-101: col 0 = v_wheel0 (left), col 1 = v_wheel1 (right), rad/s */
        double wl = clampd(cmd[0], -wmax, wmax), wr = clampd(cmd[1], -wmax, wmax);
        t->tf = rw * (wl + wr) * 0.5; t->tl = 0.0;
        t->tw = rw * (wr - wl) / (2.0 * p->half_track);  /* K4: vss_gym_base.py:57-58 */
        t->kick = 0.0; t->drib = 0.0;
    } else {
        /* rsim.py:137-153: [flag, w0..w3 | vx vy vtheta 0, kick_v_x, kick_v_z, dribbler] */
        double sp[4];
        for (int i = 0; i < 4; ++i) {
            double w;
            if (cmd[0] != 0.0) w = cmd[1 + i];
            else w = (p->omni_J[i][0] * cmd[1] + p->omni_J[i][1] * cmd[2] + p->omni_J[i][2] * cmd[3]) / rw;
            sp[i] = clampd(w, -wmax, wmax) * rw;
        }
        double tw[3];
        for (int a = 0; a < 3; ++a) {
            tw[a] = 0.0;
            for (int i = 0; i < 4; ++i) tw[a] += p->omni_Jpinv[a][i] * sp[i];
        }
        t->tf = tw[0]; t->tl = tw[1]; t->tw = tw[2];
        t->kick = cmd[5] < p->kick_speed_max ? cmd[5] : p->kick_speed_max;
        t->drib = cmd[7];
    }
    t->kicked = 0;
}

/* One control step (RS_SUBSTEPS sub-steps) of one env.  cmd: [R][C] doubles. */
static void step_env(const rs_params *p, o_ball *b, o_robot *rb, const double *cmd, double *margin) {
    const int R = p->n_robots;
    const int C = p->kind == RS_KIND_VSS ? RS_CMD_VSS : RS_CMD_SSL;
    const double h = p->h;
    o_target tg[RS_MAX_ROBOTS];
    for (int r = 0; r < R; ++r) drive_targets(p, cmd + r * C, &tg[r]);

    /* kick: once per control step, before the sub-steps, robots in row order (K6) */
    if (p->kind == RS_KIND_SSL) {
        for (int r = 0; r < R; ++r) {
            if (tg[r].kick > 0.0 && touching(p, &rb[r], b, margin)) {
                b->vx = cos(rb[r].th) * tg[r].kick; b->vy = sin(rb[r].th) * tg[r].kick;
                tg[r].kicked = 1;
            }
        }
    }

    for (int k = 0; k < RS_SUBSTEPS; ++k) {
        /* (a) drive: track the target twist in the robot frame, traction limited */
        for (int r = 0; r < R; ++r) {
            o_robot *q = &rb[r];
            double c = cos(q->th), s = sin(q->th);
            double vf = c * q->vx + s * q->vy, vl = -s * q->vx + c * q->vy;
            if (p->kind == RS_KIND_VSS) {
                vf += clampd(tg[r].tf - vf, -p->acc_fwd * h, p->acc_fwd * h);
                vl += clampd(tg[r].tl - vl, -p->acc_lat * h, p->acc_lat * h);
            } else {
                double df = tg[r].tf - vf, dl = tg[r].tl - vl;
                double n2 = df * df + dl * dl, lim = p->acc_fwd * h;
                double sc = 1.0;
                if (n2 > lim * lim) sc = lim / sqrt(n2);
                vf += df * sc; vl += dl * sc;
            }
            q->om += clampd(tg[r].tw - q->om, -p->acc_ang * h, p->acc_ang * h);
            q->vx = c * vf - s * vl; q->vy = s * vf + c * vl;
        }
        /* (b) dribbler: the first robot in row order with dribbler on, not kicking this
         * step and touching the ball holds it rigidly (rSim attaches a joint [MEM]) */
        int holder = -1; double hx = 0, hy = 0;
        if (p->kind == RS_KIND_SSL) {
            for (int r = 0; r < R && holder < 0; ++r) {
                if (tg[r].drib != 0.0 && !tg[r].kicked && touching(p, &rb[r], b, margin)) {
                    double c = cos(rb[r].th), s = sin(rb[r].th);
                    double dx = b->x - rb[r].x, dy = b->y - rb[r].y;
                    hx = c * dx + s * dy; hy = -s * dx + c * dy;
                    holder = r;
                }
            }
        }
        /* (c) ball rolling friction: Coulomb deceleration, exact stop */
        if (holder < 0) {
            double sp2 = b->vx * b->vx + b->vy * b->vy;
            double sc = 1.0 - p->ball_decel * h / sqrt(sp2 + 1e-12);
            if (sc < 0.0) sc = 0.0;
            b->vx *= sc; b->vy *= sc;
        }
        /* (d) integrate (semi-implicit Euler: new velocity, then position) */
        for (int r = 0; r < R; ++r) {
            o_robot *q = &rb[r];
            q->x += q->vx * h; q->y += q->vy * h;
            q->th = wrap_pi(q->th + q->om * h);
        }
        if (holder < 0) { b->x += b->vx * h; b->y += b->vy * h; }
        else {
            o_robot *q = &rb[holder];
            double c = cos(q->th), s = sin(q->th);
            double ox = c * hx - s * hy, oy = s * hx + c * hy;
            b->x = q->x + ox; b->y = q->y + oy;
            b->vx = q->vx - q->om * oy; b->vy = q->vy + q->om * ox;
        }
        /* (e) pairs, lexicographic over [ball, robot 0, .. robot R-1]; contacts are
         * detected on the positions at phase start, velocity impulses are applied
         * sequentially, position corrections are summed and applied afterwards */
        double cbx = 0, cby = 0, cx[RS_MAX_ROBOTS], cy[RS_MAX_ROBOTS];
        for (int r = 0; r < R; ++r) { cx[r] = 0; cy[r] = 0; }
        double wb = 1.0 / p->ball_mass, wr = 1.0 / p->rbt_mass;
        for (int r = 0; r < R; ++r) {
            double nx, ny, pen, rcx, rcy;
            if (!ball_robot_contact(p, &rb[r], b, &nx, &ny, &pen, &rcx, &rcy, margin)) continue;
            o_robot *q = &rb[r];
            /* robot surface velocity at the contact point */
            double sx = q->vx - q->om * rcy, sy = q->vy + q->om * rcx;
            double rvx = b->vx - sx, rvy = b->vy - sy;
            double vn = rvx * nx + rvy * ny;
            if (vn < 0.0) {
                double J = -(1.0 + p->e_ball_rbt) * vn / (wb + wr);
                b->vx += J * wb * nx; b->vy += J * wb * ny;
                q->vx -= J * wr * nx; q->vy -= J * wr * ny;
                double tx = -ny, ty = nx;
                double vt = rvx * tx + rvy * ty;
                double Jt = clampd(-vt / (wb + wr), -p->mu_ball_rbt * J, p->mu_ball_rbt * J);
                b->vx += Jt * wb * tx; b->vy += Jt * wb * ty;
                q->vx -= Jt * wr * tx; q->vy -= Jt * wr * ty;
            }
            cbx += pen * (wb / (wb + wr)) * nx; cby += pen * (wb / (wb + wr)) * ny;
            cx[r] -= pen * (wr / (wb + wr)) * nx; cy[r] -= pen * (wr / (wb + wr)) * ny;
        }
        for (int i = 0; i < R; ++i) for (int j = i + 1; j < R; ++j) {
            double dx = rb[j].x - rb[i].x, dy = rb[j].y - rb[i].y;
            double d2 = dx * dx + dy * dy, rs = 2.0 * p->rbt_radius;
            {
                double d = sqrt(d2), vn = d > 0 ? ((rb[j].vx - rb[i].vx) * dx + (rb[j].vy - rb[i].vy) * dy) / d : -1.0;
                note_jump(margin, d - rs, vn < 0 ? (1.0 + p->e_rbt_rbt) * vn : 0.0);
            }
            if (d2 >= rs * rs) continue;
            double d = sqrt(d2), nx, ny;
            note(margin, d * NORMAL_COND);
            if (d2 > 1e-12) { nx = dx / d; ny = dy / d; } else { nx = 1.0; ny = 0.0; d = 0.0; }
            double pen = rs - d;
            double vn = (rb[j].vx - rb[i].vx) * nx + (rb[j].vy - rb[i].vy) * ny;
            if (vn < 0.0) {
                double J = -(1.0 + p->e_rbt_rbt) * vn / (wr + wr);
                rb[i].vx -= J * wr * nx; rb[i].vy -= J * wr * ny;
                rb[j].vx += J * wr * nx; rb[j].vy += J * wr * ny;
            }
            cx[i] -= 0.5 * pen * nx; cy[i] -= 0.5 * pen * ny;
            cx[j] += 0.5 * pen * nx; cy[j] += 0.5 * pen * ny;
        }
        b->x += cbx; b->y += cby;
        for (int r = 0; r < R; ++r) { rb[r].x += cx[r]; rb[r].y += cy[r]; }
        /* (f) walls */
        walls(p, p->ball_radius, p->e_ball_wall, &b->x, &b->y, &b->vx, &b->vy, margin);
        for (int r = 0; r < R; ++r)
            walls(p, p->rbt_radius, p->e_rbt_wall, &rb[r].x, &rb[r].y, &rb[r].vx, &rb[r].vy, margin);
    }
}

/* get_state() row of one env: Entities/Frame.py:20-47 / :55-93 */
static void state_row(const rs_params *p, const o_ball *b, const o_robot *rb, double *out) {
    const int R = p->n_robots;
    const int K = p->kind == RS_KIND_VSS ? RS_STATE_VSS_ROBOT : RS_STATE_SSL_ROBOT;
    out[0] = b->x; out[1] = b->y; out[2] = p->ball_radius; out[3] = b->vx; out[4] = b->vy;
    for (int r = 0; r < R; ++r) {
        double *o = out + RS_STATE_BALL + K * r;
        o[0] = rb[r].x; o[1] = rb[r].y; o[2] = rb[r].th * DEG;
        o[3] = rb[r].vx; o[4] = rb[r].vy; o[5] = rb[r].om * DEG;
        if (p->kind == RS_KIND_SSL) {
            o[6] = touching(p, &rb[r], b, NULL) ? 1.0 : 0.0;
            double c = cos(rb[r].th), s = sin(rb[r].th);
            double vf = c * rb[r].vx + s * rb[r].vy, vl = -s * rb[r].vx + c * rb[r].vy;
            for (int i = 0; i < 4; ++i)
                o[7 + i] = (p->omni_J[i][0] * vf + p->omni_J[i][1] * vl + p->omni_J[i][2] * rb[r].om)
                           / p->rbt_wheel_radius;
        }
    }
}

/* ------------------------------------------------------------------------- */
/* C API of the oracle (loaded by oracle/oracle.py through ctypes)             */

orc_world *orc_create(int kind, int field_type, int n_blue, int n_yellow, int time_step_ms,
                      int n_envs, uint64_t seed, int64_t env_offset) {
    orc_world *w = (orc_world *)calloc(1, sizeof(orc_world));
    if (!w) return NULL;
    if (rs_params_fill(&w->p, kind, field_type, n_blue, n_yellow, time_step_ms) != 0 || n_envs < 0) {
        free(w); return NULL;
    }
    const int R = w->p.n_robots;
    w->n = n_envs; w->seed = seed; w->t = 0; w->env_offset = env_offset;
    w->n_threads = 1;
    size_t n = (size_t)(n_envs > 0 ? n_envs : 1);
    w->ball = (o_ball *)calloc(n, sizeof(o_ball));
    w->rob = (o_robot *)calloc(n * R, sizeof(o_robot));
    w->ou = (double *)calloc(n * 2 * RS_MAX_ROBOTS, sizeof(double));
    w->prev_pot = (double *)calloc(n, sizeof(double));
    w->has_prev = (int *)calloc(n, sizeof(int));
    w->steps = (int *)calloc(n, sizeof(int));
    w->info = (double *)calloc(n * RS_SSL_INFO, sizeof(double));
    w->margin = (double *)calloc(n, sizeof(double));
    /* rsim.py:19-24 dummy initial poses */
    for (int e = 0; e < n_envs; ++e) {
        for (int r = 0; r < R; ++r) {
            o_robot *q = &w->rob[(size_t)e * R + r];
            q->x = r < n_blue ? -0.2 * (r + 1) : 0.2 * (r - n_blue + 1);
        }
        w->margin[e] = 1e30;
    }
    return w;
}
void orc_destroy(orc_world *w) {
    if (!w) return;
    free(w->ball); free(w->rob); free(w->ou); free(w->prev_pot); free(w->has_prev);
    free(w->steps); free(w->info); free(w->margin); free(w);
}
void orc_set_threads(orc_world *w, int n) { w->n_threads = n > 0 ? n : 1; }
void orc_field_params(const orc_world *w, double out[RS_FIELD_KEYS]) { rs_params_field(&w->p, out); }
const rs_params *orc_params(const orc_world *w) { return &w->p; }
/* the DERIVED block of include/rs_spec.h (rs_params_fill) in a fixed order, for tests/test_spec_independent.py,
 * which re-derives it with numpy and hand-computed literals (the header is shared by the library and this oracle,
 * so a CUDA <-> oracle comparison cannot see a mistake in it).  Returns the number of doubles written (64). */
int orc_params_dump(const orc_world *w, double *out) {
    const rs_params *p = &w->p;
    int k = 0;
    out[k++] = p->dt; out[k++] = p->h; out[k++] = p->x_out; out[k++] = p->y_out; out[k++] = p->x_near;
    out[k++] = (double)p->n_box;
    for (int b = 0; b < RS_MAX_BOXES; ++b) for (int i = 0; i < 4; ++i) out[k++] = p->box[b][i];
    out[k++] = p->ball_mass; out[k++] = p->rbt_mass;
    out[k++] = p->e_ball_wall; out[k++] = p->e_rbt_wall; out[k++] = p->e_ball_rbt; out[k++] = p->e_rbt_rbt;
    out[k++] = p->mu_ball_rbt; out[k++] = p->ball_decel;
    out[k++] = p->wheel_max_rad_s; out[k++] = p->half_track;
    out[k++] = p->acc_fwd; out[k++] = p->acc_lat; out[k++] = p->acc_ang;
    for (int i = 0; i < 4; ++i) for (int a = 0; a < 3; ++a) out[k++] = p->omni_J[i][a];
    for (int a = 0; a < 3; ++a) for (int i = 0; i < 4; ++i) out[k++] = p->omni_Jpinv[a][i];
    out[k++] = p->kick_centre; out[k++] = p->kick_reach; out[k++] = p->kick_half_width;
    out[k++] = p->mouth_half_chord; out[k++] = p->kick_speed_max;
    return k;
}
uint64_t orc_get_t(const orc_world *w) { return w->t; }
void orc_set_t(orc_world *w, uint64_t t) { w->t = t; }

static void clear_task(orc_world *w, int e) {
    memset(w->ou + (size_t)e * 2 * RS_MAX_ROBOTS, 0, sizeof(double) * 2 * RS_MAX_ROBOTS);
    w->has_prev[e] = 0; w->prev_pot[e] = 0.0; w->steps[e] = 0;
}

/* robosim.reset(ball[4], blue[nb][3], yellow[ny][3]) -- rsim.py:36-38, 52-75.
 * ball = [x, y, vx, vy]; robots = [x, y, theta_deg]; robot velocities zeroed (K1).
 * mask (nullable): only envs with mask[e] != 0 are reset. */
void orc_reset(orc_world *w, const double *ball, const double *blue, const double *yellow,
               const uint8_t *mask) {
    const int R = w->p.n_robots, nb = w->p.n_blue, ny = w->p.n_yellow;
    for (int e = 0; e < w->n; ++e) {
        if (mask && !mask[e]) continue;
        o_ball *b = &w->ball[e];
        b->x = ball[4 * e]; b->y = ball[4 * e + 1]; b->vx = ball[4 * e + 2]; b->vy = ball[4 * e + 3];
        for (int r = 0; r < R; ++r) {
            const double *src = r < nb ? blue + ((size_t)e * nb + r) * 3
                                       : yellow + ((size_t)e * ny + (r - nb)) * 3;
            o_robot *q = &w->rob[(size_t)e * R + r];
            q->x = src[0]; q->y = src[1];
            q->th = remainder(src[2], 360.0) / DEG;
            if (q->th <= -PI) q->th += 2.0 * PI;
            q->vx = q->vy = q->om = 0.0;
        }
        clear_task(w, e);
    }
}

/* full internal state in / out, for re-synced parity tests: [n][4 + 6R]
 * (ball x y vx vy; robot x y theta_rad vx vy omega) */
void orc_set_raw(orc_world *w, const double *s) {
    const int R = w->p.n_robots; const int S = 4 + 6 * R;
    for (int e = 0; e < w->n; ++e) {
        const double *r = s + (size_t)e * S;
        w->ball[e].x = r[0]; w->ball[e].y = r[1]; w->ball[e].vx = r[2]; w->ball[e].vy = r[3];
        for (int k = 0; k < R; ++k) memcpy(&w->rob[(size_t)e * R + k], r + 4 + 6 * k, sizeof(o_robot));
    }
}
void orc_get_raw(const orc_world *w, double *s) {
    const int R = w->p.n_robots; const int S = 4 + 6 * R;
    for (int e = 0; e < w->n; ++e) {
        double *r = s + (size_t)e * S;
        r[0] = w->ball[e].x; r[1] = w->ball[e].y; r[2] = w->ball[e].vx; r[3] = w->ball[e].vy;
        for (int k = 0; k < R; ++k) memcpy(r + 4 + 6 * k, &w->rob[(size_t)e * R + k], sizeof(o_robot));
    }
}

/* robosim.step(cmds[R][C]) for every env -- rsim.py:102 / :155 */
void orc_step(orc_world *w, const double *cmds) {
    const int R = w->p.n_robots;
    const int C = w->p.kind == RS_KIND_VSS ? RS_CMD_VSS : RS_CMD_SSL;
#pragma omp parallel for num_threads(w->n_threads) schedule(static)
    for (int e = 0; e < w->n; ++e) {
        w->margin[e] = 1e30;
        step_env(&w->p, &w->ball[e], &w->rob[(size_t)e * R], cmds + (size_t)e * R * C, &w->margin[e]);
    }
    w->t++;
}

/* robosim.get_state() for every env -- rsim.py:105 / :158; out [n][5 + K R] */
void orc_get_state(const orc_world *w, double *out) {
    const int R = w->p.n_robots;
    const int K = w->p.kind == RS_KIND_VSS ? RS_STATE_VSS_ROBOT : RS_STATE_SSL_ROBOT;
    for (int e = 0; e < w->n; ++e)
        state_row(&w->p, &w->ball[e], &w->rob[(size_t)e * R], out + (size_t)e * (RS_STATE_BALL + K * R));
}
void orc_get_margin(const orc_world *w, double *out) { memcpy(out, w->margin, sizeof(double) * w->n); }
void orc_get_task_state(const orc_world *w, double *ou, double *prev_pot, int *has_prev, int *steps,
                        double *info) {
    const int R = w->p.n_robots;
    for (int e = 0; e < w->n; ++e) {
        if (ou) memcpy(ou + (size_t)e * 2 * (R - 1), w->ou + (size_t)e * 2 * RS_MAX_ROBOTS,
                       sizeof(double) * 2 * (R - 1));
        if (prev_pot) prev_pot[e] = w->prev_pot[e];
        if (has_prev) has_prev[e] = w->has_prev[e];
        if (steps) steps[e] = w->steps[e];
        if (info) memcpy(info + (size_t)e * RS_SSL_INFO, w->info + (size_t)e * RS_SSL_INFO,
                         sizeof(double) * RS_SSL_INFO);
    }
}
void orc_set_task_state(orc_world *w, const double *ou, const double *prev_pot, const int *has_prev,
                        const int *steps, const double *info) {
    const int R = w->p.n_robots;
    for (int e = 0; e < w->n; ++e) {
        if (ou) memcpy(w->ou + (size_t)e * 2 * RS_MAX_ROBOTS, ou + (size_t)e * 2 * (R - 1),
                       sizeof(double) * 2 * (R - 1));
        if (prev_pot) w->prev_pot[e] = prev_pot[e];
        if (has_prev) w->has_prev[e] = has_prev[e];
        if (steps) w->steps[e] = steps[e];
        if (info) memcpy(w->info + (size_t)e * RS_SSL_INFO, info + (size_t)e * RS_SSL_INFO,
                         sizeof(double) * RS_SSL_INFO);
    }
}

/* ------------------------------------------------------------------------- */
/* base-env normalisers: vss_gym_base.py:52-58, 213-220; ssl_gym_base.py:53-59 */
static double max_pos_of(const rs_params *p) {
    double a = p->width / 2, b = p->length / 2 + p->penalty_length;
    return a > b ? a : b;
}
static double max_v_of(const rs_params *p) {
    return (p->rbt_motor_max_rpm / 60.0) * 2.0 * PI * p->rbt_wheel_radius;
}
static double nrm(double v, double m) { return clampd(v / m, -RS_NORM_BOUNDS, RS_NORM_BOUNDS); }

/* vss_gym.py:235-254 _actions_to_v_wheels */
static void vss_action_to_wheels(const rs_params *p, double a0, double a1, double *wl, double *wr,
                                 double *margin) {
    double mv = max_v_of(p);
    double l = clampd(a0 * mv, -mv, mv), r = clampd(a1 * mv, -mv, mv);
    note(margin, fabs(l) - RS_VSS_DEADZONE); note(margin, fabs(r) - RS_VSS_DEADZONE);
    if (-RS_VSS_DEADZONE < l && l < RS_VSS_DEADZONE) l = 0.0;
    if (-RS_VSS_DEADZONE < r && r < RS_VSS_DEADZONE) r = 0.0;
    *wl = l / p->rbt_wheel_radius; *wr = r / p->rbt_wheel_radius;
}

/* vss_gym.py:93-117 _frame_to_observations (n_obs = 4 + 7 nb + 5 ny) */
static void vss_obs(const rs_params *p, const o_ball *b, const o_robot *rb, double *o) {
    double mp = max_pos_of(p), mv = max_v_of(p), mw = (mv / 0.04) * DEG;
    int k = 0;
    o[k++] = nrm(b->x, mp); o[k++] = nrm(b->y, mp); o[k++] = nrm(b->vx, mv); o[k++] = nrm(b->vy, mv);
    for (int r = 0; r < p->n_blue; ++r) {
        o[k++] = nrm(rb[r].x, mp); o[k++] = nrm(rb[r].y, mp);
        o[k++] = sin(rb[r].th); o[k++] = cos(rb[r].th);
        o[k++] = nrm(rb[r].vx, mv); o[k++] = nrm(rb[r].vy, mv); o[k++] = nrm(rb[r].om * DEG, mw);
    }
    for (int r = p->n_blue; r < p->n_robots; ++r) {
        o[k++] = nrm(rb[r].x, mp); o[k++] = nrm(rb[r].y, mp);
        o[k++] = nrm(rb[r].vx, mv); o[k++] = nrm(rb[r].vy, mv); o[k++] = nrm(rb[r].om * DEG, mw);
    }
}

/* random non-overlapping placement: vss_gym.py:194-233 (min_dist 0.1; exact all-pairs
 * distance instead of the reference's KD-tree) */
static void vss_place(const rs_params *p, o_rng *g, o_ball *b, o_robot *rb) {
    double hl = p->length / 2, hw = p->width / 2;
    double px[1 + RS_MAX_ROBOTS], py[1 + RS_MAX_ROBOTS];
    b->x = rng_uniform(g, -hl + 0.1, hl - 0.1); b->y = rng_uniform(g, -hw + 0.1, hw - 0.1);
    b->vx = b->vy = 0.0;
    px[0] = b->x; py[0] = b->y;
    for (int r = 0; r < p->n_robots; ++r) {
        double x = 0, y = 0;
        for (int tries = 0; tries < 64; ++tries) {
            x = rng_uniform(g, -hl + 0.1, hl - 0.1); y = rng_uniform(g, -hw + 0.1, hw - 0.1);
            int ok = 1;
            for (int k = 0; k <= r; ++k) {
                double dx = x - px[k], dy = y - py[k];
                if (dx * dx + dy * dy < 0.1 * 0.1) ok = 0;
            }
            if (ok) break;
        }
        px[r + 1] = x; py[r + 1] = y;
        rb[r].x = x; rb[r].y = y; rb[r].vx = rb[r].vy = rb[r].om = 0.0;
        rb[r].th = wrap_pi(rng_uniform(g, 0.0, 360.0) / DEG);
    }
}

/* static_defenders.py:214-254 */
static void ssl_sd_place(const rs_params *p, o_rng *g, o_ball *b, o_robot *rb) {
    double hl = p->length / 2, hw = p->width / 2, pl = p->penalty_length, hpw = p->penalty_width / 2;
    double px[2 + RS_MAX_ROBOTS], py[2 + RS_MAX_ROBOTS];
    rb[0].x = rb[0].y = rb[0].th = rb[0].vx = rb[0].vy = rb[0].om = 0.0;
    for (int tries = 0; tries < 64; ++tries) {
        b->x = rng_uniform(g, 0.2, hl - 0.1); b->y = rng_uniform(g, -hw + 0.1, hw - 0.1);
        if (!(b->x > hl - pl && fabs(b->y) < hpw)) break;
    }
    b->vx = b->vy = 0.0;
    px[0] = b->x; py[0] = b->y; px[1] = 0.0; py[1] = 0.0;
    for (int r = 1; r < p->n_robots; ++r) {
        double x = 0, y = 0;
        for (int tries = 0; tries < 64; ++tries) {
            x = rng_uniform(g, 0.2, hl - 0.1); y = rng_uniform(g, -hw + 0.1, hw - 0.1);
            int ok = 1;
            for (int k = 0; k <= r; ++k) {
                double dx = x - px[k], dy = y - py[k];
                if (dx * dx + dy * dy < 0.2 * 0.2) ok = 0;
            }
            if (ok) break;
        }
        px[r + 1] = x; py[r + 1] = y;
        rb[r].x = x; rb[r].y = y; rb[r].vx = rb[r].vy = rb[r].om = 0.0;
        rb[r].th = wrap_pi(rng_uniform(g, 0.0, 360.0) / DEG);
    }
}

/* contested_possession.py:210-227 */
static void ssl_cp_place(const rs_params *p, o_rng *g, o_ball *b, o_robot *rb) {
    double hl = p->length / 2, pl = p->penalty_length, hpw = p->penalty_width / 2;
    rb[0].x = rb[0].y = rb[0].th = rb[0].vx = rb[0].vy = rb[0].om = 0.0;
    double ex = rng_uniform(g, pl, hl - pl), ey = rng_uniform(g, -hpw, hpw);
    b->x = ex - 0.1; b->y = ey; b->vx = b->vy = 0.0;
    for (int r = 1; r < p->n_robots; ++r) {
        rb[r].x = ex; rb[r].y = ey + 0.5 * (r - 1); rb[r].th = PI;
        rb[r].vx = rb[r].vy = rb[r].om = 0.0;
    }
}

/* dribbling.py:187-202: fixed course, robot at the origin facing -x with the ball in its mouth */
static void ssl_drib_place(const rs_params *p, o_rng *g, o_ball *b, o_robot *rb) {
    (void)g;
    b->x = -0.1; b->y = 0.0; b->vx = b->vy = 0.0;
    for (int r = 0; r < p->n_robots; ++r) {
        rb[r].x = r == 0 ? 0.0 : -0.5 * r; rb[r].y = 0.0; rb[r].th = PI;
        rb[r].vx = rb[r].vy = rb[r].om = 0.0;
    }
}
/* pass_endurance.py:152-181: ball anywhere in [-1.5, 1.5]^2, the shooter 0.115 m behind it facing
 * the ball along y, the receiver mirrored in y at least 1 m away in x, facing the shooter */
static void ssl_pass_place(const rs_params *p, o_rng *g, o_ball *b, o_robot *rb) {
    (void)p;
    b->x = rng_uniform(g, -1.5, 1.5); b->y = rng_uniform(g, 1.5, -1.5); b->vx = b->vy = 0.0;
    const double factor = b->y < 0.0 ? -1.0 : 1.0;
    rb[0].x = b->x; rb[0].y = b->y + 0.115 * factor; rb[0].th = factor > 0.0 ? -0.5 * PI : 0.5 * PI;
    double rx = 0.0;
    for (int tries = 0; tries < 64; ++tries) {
        rx = rng_uniform(g, -1.5, 1.5);
        if (!(fabs(rx - b->x) < 1.0)) break;
    }
    rb[1].x = rx; rb[1].y = -b->y;
    rb[1].th = wrap_pi(atan2(rb[1].y - rb[0].y, rb[1].x - rb[0].x) + PI);
    for (int r = 0; r < 2; ++r) rb[r].vx = rb[r].vy = rb[r].om = 0.0;
}

static void place(const orc_world *w, int task, int e, uint32_t stream) {
    o_rng g;
    rng_init(&g, w->seed, (uint64_t)(w->env_offset + e), w->t, stream);
    const int R = w->p.n_robots;
    if (task == RS_TASK_VSS) vss_place(&w->p, &g, &w->ball[e], &w->rob[(size_t)e * R]);
    else if (task == RS_TASK_SSL_STATIC_DEFENDERS) ssl_sd_place(&w->p, &g, &w->ball[e], &w->rob[(size_t)e * R]);
    else if (task == RS_TASK_SSL_DRIBBLING) ssl_drib_place(&w->p, &g, &w->ball[e], &w->rob[(size_t)e * R]);
    else if (task == RS_TASK_SSL_PASS_ENDURANCE) ssl_pass_place(&w->p, &g, &w->ball[e], &w->rob[(size_t)e * R]);
    else ssl_cp_place(&w->p, &g, &w->ball[e], &w->rob[(size_t)e * R]);
}

/* env.reset() for the envs selected by mask (nullable = all): draws the initial
 * frame on stream RS_STREAM_RESET at the current world step counter */
void orc_task_reset(orc_world *w, int task, const uint8_t *mask) {
    for (int e = 0; e < w->n; ++e) {
        if (mask && !mask[e]) continue;
        place(w, task, e, RS_STREAM_RESET);
        clear_task(w, e);
    }
}

/* vss_gym.py: one VSSEnv.step() per env (step :89-91 via vss_gym_base.py:72-90)
 *  actions [n][2] (agent = blue 0), normals [n][2(R-1)] or NULL (=> Philox stream 0)
 *  outputs obs [n][n_obs], reward [n], done [n], trunc [n], cmds_out [n][R][2] (nullable)   */
void orc_vss_env_step(orc_world *w, const float *actions, const double *normals, int auto_reset,
                      int max_steps, double *obs, double *reward, uint8_t *done, uint8_t *trunc,
                      double *cmds_out) {
    const rs_params *p = &w->p;
    const int R = p->n_robots, NZ = 2 * (R - 1);
    const int n_obs = 4 + 7 * p->n_blue + 5 * p->n_yellow;
#pragma omp parallel for num_threads(w->n_threads) schedule(static)
    for (int e = 0; e < w->n; ++e) {
        o_ball *b = &w->ball[e]; o_robot *rb = &w->rob[(size_t)e * R];
        double *ou = w->ou + (size_t)e * 2 * RS_MAX_ROBOTS;
        double *info = w->info + (size_t)e * RS_SSL_INFO;
        double cmd[RS_MAX_ROBOTS * 2];
        double *mg = &w->margin[e]; *mg = 1e30;
        if (w->steps[e] == 0) memset(info, 0, sizeof(double) * RS_SSL_INFO);
        w->steps[e] += 1;                                            /* vss_gym_base.py:73 */
        /* --- _get_commands, vss_gym.py:119-142 --- */
        vss_action_to_wheels(p, actions[2 * e], actions[2 * e + 1], &cmd[0], &cmd[1], mg);
        double z[2 * RS_MAX_ROBOTS];
        if (normals) memcpy(z, normals + (size_t)e * NZ, sizeof(double) * NZ);
        else {
            o_rng g; rng_init(&g, w->seed, (uint64_t)(w->env_offset + e), w->t, RS_STREAM_OU);
            for (int k = 0; k < NZ; k += 2) {
                double u1 = u01(rng_u32(&g)), u2 = u01(rng_u32(&g));
                double rr = sqrt(-2.0 * log(u1));
                z[k] = rr * cos(2.0 * PI * u2);
                if (k + 1 < NZ) z[k + 1] = rr * sin(2.0 * PI * u2);
            }
        }
        for (int r = 1; r < R; ++r) {
            /* Utils/Utils.py:14-21: x = x_prev + theta (mu - x_prev) dt + sigma sqrt(dt) N(0,1) */
            for (int k = 0; k < 2; ++k) {
                double *x = &ou[2 * (r - 1) + k];
                *x = *x + RS_OU_THETA * (0.0 - *x) * p->dt + RS_OU_SIGMA * sqrt(p->dt) * z[2 * (r - 1) + k];
            }
            vss_action_to_wheels(p, ou[2 * (r - 1)], ou[2 * (r - 1) + 1], &cmd[2 * r], &cmd[2 * r + 1], mg);
        }
        if (cmds_out) memcpy(cmds_out + (size_t)e * R * 2, cmd, sizeof(double) * R * 2);
        /* --- rsim.send_commands + get_frame, vss_gym_base.py:77-82 --- */
        step_env(p, b, rb, cmd, mg);
        /* --- _calculate_reward_and_done, vss_gym.py:144-192 --- */
        double rew = 0.0; int goal = 0;
        note(mg, fabs(b->x) - p->length / 2);
        if (b->x > p->length / 2) { info[0] += 1; info[4] += 1; rew = 10.0; goal = 1; }
        else if (b->x < -p->length / 2) { info[0] -= 1; info[5] += 1; rew = -10.0; goal = 1; }
        else {
            /* __ball_grad, vss_gym.py:256-283 */
            double length_cm = p->length * 100, hl = p->length / 2.0 + p->goal_depth;
            double dx_d = (hl + b->x) * 100, dx_a = (hl - b->x) * 100, dy = b->y * 100;
            double pot = ((-sqrt(dx_a * dx_a + 2 * dy * dy) + sqrt(dx_d * dx_d + 2 * dy * dy)) / length_cm - 1) / 2;
            double grad = 0.0;
            if (w->has_prev[e]) grad = clampd((pot - w->prev_pot[e]) * 3 / p->dt, -5.0, 5.0);
            w->prev_pot[e] = pot; w->has_prev[e] = 1;
            /* __move_reward, vss_gym.py:285-303 */
            double rx = b->x - rb[0].x, ry = b->y - rb[0].y, rn = sqrt(rx * rx + ry * ry);
            double move = clampd((rx / rn * rb[0].vx + ry / rn * rb[0].vy) / 0.4, -5.0, 5.0);
            /* __energy_penalty, vss_gym.py:305-311 */
            double energy = -(fabs(cmd[0]) + fabs(cmd[1]));
            rew = 0.2 * move + 0.8 * grad + 2e-4 * energy;
            info[1] += 0.2 * move; info[2] += 0.8 * grad; info[3] += 2e-4 * energy;
        }
        reward[e] = rew; done[e] = (uint8_t)goal;
        int tr = w->steps[e] >= max_steps;                          /* TimeLimit, __init__.py:4 */
        trunc[e] = (uint8_t)tr;
        if (auto_reset && (goal || tr)) { place(w, RS_TASK_VSS, e, RS_STREAM_AUTORESET); clear_task(w, e); }
        vss_obs(p, b, rb, obs + (size_t)e * n_obs);
    }
    w->t++;
}

/* static_defenders.py / contested_possession.py: one env.step() per env.
 * actions [n][5]; obs [n][4 + 8 nb + 2 ny]; cmds_out [n][R][8] nullable */
void orc_ssl_env_step(orc_world *w, int task, const float *actions, int auto_reset, int max_steps,
                      double *obs, double *reward, uint8_t *done, uint8_t *trunc, double *cmds_out) {
    const rs_params *p = &w->p;
    const int R = p->n_robots;
    const int n_obs = 4 + 8 * p->n_blue + 2 * p->n_yellow;
    const double max_v = 2.5, max_w = 10.0, kick_speed = 5.0;       /* static_defenders.py:76-78 */
    const double mp = max_pos_of(p);
    const double hl = p->length / 2, hw = p->width / 2, pl = p->penalty_length;
    const double hpw = p->penalty_width / 2, hgw = p->goal_width / 2;
    const double ball_dist_scale = sqrt(p->width * p->width + hl * hl);          /* :65 */
    const double ball_grad_scale = sqrt(hw * hw + hl * hl) / 4;                   /* :66-68 */
    const double energy_scale = 160.0 * 4 * (task == RS_TASK_SSL_STATIC_DEFENDERS ? 1000 : 1200); /* :71-73 */
#pragma omp parallel for num_threads(w->n_threads) schedule(static)
    for (int e = 0; e < w->n; ++e) {
        o_ball *b = &w->ball[e]; o_robot *rb = &w->rob[(size_t)e * R];
        double *info = w->info + (size_t)e * RS_SSL_INFO;
        double cmd[RS_MAX_ROBOTS * RS_CMD_SSL];
        double *mg = &w->margin[e]; *mg = 1e30;
        memset(cmd, 0, sizeof(cmd));
        if (w->steps[e] == 0) memset(info, 0, sizeof(double) * RS_SSL_INFO);
        w->steps[e] += 1;
        const float *a = actions + (size_t)e * RS_SSL_ACT;
        /* _get_commands + convert_actions, static_defenders.py:114-148 */
        double ang = rb[0].th;
        double vx = a[0] * max_v, vy = a[1] * max_v, vth = a[2] * max_w;
        double lx = vx * cos(ang) + vy * sin(ang), ly = -vx * sin(ang) + vy * cos(ang);
        double vn = sqrt(lx * lx + ly * ly);
        double c = vn < max_v ? 1.0 : max_v / vn;
        cmd[0] = 0.0; cmd[1] = lx * c; cmd[2] = ly * c; cmd[3] = vth;
        cmd[5] = a[3] > 0 ? kick_speed : 0.0; cmd[6] = 0.0; cmd[7] = a[4] > 0 ? 1.0 : 0.0;
        if (cmds_out) memcpy(cmds_out + (size_t)e * R * RS_CMD_SSL, cmd, sizeof(double) * R * RS_CMD_SSL);
        double lbx = b->x, lby = b->y, lrx = rb[0].x, lry = rb[0].y;  /* last_frame */
        step_env(p, b, rb, cmd, mg);
        /* _calculate_reward_and_done, static_defenders.py:150-212 / contested_possession.py:136-208 */
        double rew = 0.0; int dn = 0;
        if (task == RS_TASK_SSL_CONTESTED_POSSESSION) {
            for (int r = p->n_blue; r < R; ++r) {
                note(mg, fabs(rb[r].vx) - 0.1); note(mg, fabs(rb[r].vy) - 0.1);
                if (fabs(rb[r].vx) > 0.1 || fabs(rb[r].vy) > 0.1) { info[8] += 1; dn = 1; }
            }
        }
        note(mg, rb[0].x + 0.2); note(mg, fabs(rb[0].y) - hw); note(mg, rb[0].x - (hl - pl));
        note(mg, fabs(rb[0].y) - hpw); note(mg, b->x); note(mg, fabs(b->y) - hw);
        note(mg, b->x - hl); note(mg, fabs(b->y) - hgw);
        if (rb[0].x < -0.2 || fabs(rb[0].y) > hw) { dn = 1; info[4] += 1; }
        else if (rb[0].x > hl - pl && fabs(rb[0].y) < hpw) { dn = 1; info[1] += 1; }
        else if (b->x < 0 || fabs(b->y) > hw) { dn = 1; info[2] += 1; }
        else if (b->x > hl) {
            dn = 1;
            if (fabs(b->y) < hgw) { rew = 5.0; info[0] += 1; } else { rew = 0.0; info[3] += 1; }
        } else {
            double ld = sqrt((lrx - lbx) * (lrx - lbx) + (lry - lby) * (lry - lby));
            double nd = sqrt((rb[0].x - b->x) * (rb[0].x - b->x) + (rb[0].y - b->y) * (rb[0].y - b->y));
            double bd = clampd(ld - nd, -1, 1) / ball_dist_scale;
            double lg = sqrt((hl - lbx) * (hl - lbx) + lby * lby);
            double ng = sqrt((hl - b->x) * (hl - b->x) + b->y * b->y);
            double bg = clampd(lg - ng, -1, 1) / ball_grad_scale;
            double st[RS_STATE_BALL + RS_STATE_SSL_ROBOT * RS_MAX_ROBOTS];
            state_row(p, b, rb, st);
            double en = fabs(st[5 + 7]) + fabs(st[5 + 8]) + fabs(st[5 + 9]) + fabs(st[5 + 10]);
            double er = -en / energy_scale;
            info[5] += bd; info[6] += bg; info[7] += er;
            rew = rew + bd + bg + er;
        }
        reward[e] = rew; done[e] = (uint8_t)dn;
        int tr = w->steps[e] >= max_steps;
        trunc[e] = (uint8_t)tr;
        if (auto_reset && (dn || tr)) { place(w, task, e, RS_STREAM_AUTORESET); clear_task(w, e); }
        /* _frame_to_observations, static_defenders.py:90-112 */
        double *o = obs + (size_t)e * n_obs; int k = 0;
        o[k++] = nrm(b->x, mp); o[k++] = nrm(b->y, mp); o[k++] = nrm(b->vx, max_v); o[k++] = nrm(b->vy, max_v);
        for (int r = 0; r < p->n_blue; ++r) {
            o[k++] = nrm(rb[r].x, mp); o[k++] = nrm(rb[r].y, mp);
            o[k++] = sin(rb[r].th); o[k++] = cos(rb[r].th);
            o[k++] = nrm(rb[r].vx, max_v); o[k++] = nrm(rb[r].vy, max_v);
            o[k++] = nrm(rb[r].om * DEG, max_w);
            o[k++] = touching(p, &rb[r], b, mg) ? 1.0 : 0.0;
        }
        for (int r = p->n_blue; r < R; ++r) { o[k++] = nrm(rb[r].x, mp); o[k++] = nrm(rb[r].y, mp); }
    }
    w->t++;
}

/* ---- SSLDribbling-v0 / SSLPassEndurance-v0 (ssl_hw_challenge/dribbling.py, pass_endurance.py) ----
 * The per-episode counter of each task (dribbling: checkpoints_count; pass endurance:
 * stopped_steps) lives in prev_pot[e].  info[0..1] of pass endurance = reward_shaping_total
 * {reversed_dist, ball_grad} (pass_endurance.py:113-114); its holding_steps is never incremented
 * by the reference (initialised :56, reset :92, only compared :121), so `> 15` never fires.   */
static void ssl_hw_obs(const rs_params *p, int task, const o_ball *b, const o_robot *rb, double counter,
                       double *o, double *mg) {
    const double mp = max_pos_of(p), max_v = 2.5, max_w = 10.0;       /* dribbling.py:66-67, pass_endurance.py:73-74 */
    int k = 0;
    if (task == RS_TASK_SSL_DRIBBLING) {                               /* dribbling.py:75-105 */
        o[k++] = ((counter / 6) * 2) - 1;
        o[k++] = nrm(b->x, mp); o[k++] = nrm(b->y, mp); o[k++] = nrm(b->vx, max_v); o[k++] = nrm(b->vy, max_v);
        for (int r = 0; r < p->n_blue; ++r) {
            o[k++] = nrm(rb[r].x, mp); o[k++] = nrm(rb[r].y, mp);
            o[k++] = sin(rb[r].th); o[k++] = cos(rb[r].th);
            o[k++] = nrm(rb[r].vx, max_v); o[k++] = nrm(rb[r].vy, max_v);
            o[k++] = nrm(rb[r].om * DEG, max_w);
            o[k++] = touching(p, &rb[r], b, mg) ? 1.0 : -1.0;
        }
        for (int r = p->n_blue; r < p->n_robots; ++r) { o[k++] = nrm(rb[r].x, mp); o[k++] = nrm(rb[r].y, mp); }
    } else {                                                           /* pass_endurance.py:78-90 */
        o[k++] = nrm(b->x, mp); o[k++] = nrm(b->y, mp); o[k++] = nrm(b->vx, max_v); o[k++] = nrm(b->vy, max_v);
        for (int r = 0; r < p->n_blue; ++r) {
            o[k++] = nrm(rb[r].x, mp); o[k++] = nrm(rb[r].y, mp);
            o[k++] = sin(rb[r].th); o[k++] = cos(rb[r].th);
            o[k++] = nrm(rb[r].om * DEG, max_w);
            o[k++] = touching(p, &rb[r], b, mg) ? 1.0 : 0.0;
        }
    }
}

/* actions [n][4] (dribbling) or [n][3] (pass endurance); obs [n][21] or [n][16];
 * cmds_out [n][R][8] nullable */
void orc_ssl_hw_env_step(orc_world *w, int task, const float *actions, int auto_reset, int max_steps,
                         double *obs, double *reward, uint8_t *done, uint8_t *trunc, double *cmds_out) {
    const rs_params *p = &w->p;
    const int R = p->n_robots;
    const int n_act = task == RS_TASK_SSL_DRIBBLING ? RS_DRIB_ACT : RS_PASS_ACT;
    const int n_obs = task == RS_TASK_SSL_DRIBBLING ? RS_DRIB_OBS : RS_PASS_OBS;
    const double max_v = 2.5, max_w = 10.0, max_kick_x = 5.0;
    const double hl = p->length / 2, hw = p->width / 2;
    const double ball_grad_scale = sqrt(hw * hw + hl * hl) / 4;       /* pass_endurance.py:68-70 */
#pragma omp parallel for num_threads(w->n_threads) schedule(static)
    for (int e = 0; e < w->n; ++e) {
        o_ball *b = &w->ball[e]; o_robot *rb = &w->rob[(size_t)e * R];
        double *info = w->info + (size_t)e * RS_SSL_INFO;
        double cmd[RS_MAX_ROBOTS * RS_CMD_SSL];
        double *mg = &w->margin[e]; *mg = 1e30;
        memset(cmd, 0, sizeof(cmd));
        if (w->steps[e] == 0) { memset(info, 0, sizeof(double) * RS_SSL_INFO); w->prev_pot[e] = 0.0; }
        w->steps[e] += 1;
        const float *a = actions + (size_t)e * n_act;
        if (task == RS_TASK_SSL_DRIBBLING) {                           /* dribbling.py:107-133 */
            double ang = rb[0].th;
            double vx = a[0] * max_v, vy = a[1] * max_v, vth = a[2] * max_w;
            double lx = vx * cos(ang) + vy * sin(ang), ly = -vx * sin(ang) + vy * cos(ang);
            double vn = sqrt(lx * lx + ly * ly);
            double c = vn < max_v ? 1.0 : max_v / vn;
            cmd[1] = lx * c; cmd[2] = ly * c; cmd[3] = vth; cmd[7] = a[3] > 0 ? 1.0 : 0.0;
        } else {                                                       /* pass_endurance.py:100-124 */
            double a1 = fabsf(a[1]) > 0.5f ? a[1] : 0.0;
            cmd[3] = a[0] * max_w; cmd[5] = a1 * max_kick_x; cmd[7] = a[2] > 0 ? 1.0 : 0.0;
            cmd[RS_CMD_SSL + 7] = 1.0;                                 /* receiver: dribbler on, everything else 0 */
        }
        if (cmds_out) memcpy(cmds_out + (size_t)e * R * RS_CMD_SSL, cmd, sizeof(double) * R * RS_CMD_SSL);
        const double lbx = b->x, lby = b->y;                           /* last_frame.ball */
        step_env(p, b, rb, cmd, mg);
        /* ssl_gym_base.py:83-85: the observation is taken BEFORE the reward updates the counter */
        double *o = obs + (size_t)e * n_obs;
        ssl_hw_obs(p, task, b, rb, w->prev_pot[e], o, mg);
        double rew = 0.0; int dn = 0;
        if (task == RS_TASK_SSL_DRIBBLING) {                           /* dribbling.py:135-185 */
            const double n0 = -0.5, n1 = -1.0, n2 = -1.5, n3 = -2.0, fm = 1.0;
            int cc = (int)w->prev_pot[e];
            for (int r = p->n_blue; r < R; ++r) {
                note(mg, fabs(rb[r].vx) - 0.05); note(mg, fabs(rb[r].vy) - 0.05);
                if (fabs(rb[r].vx) > 0.05 || fabs(rb[r].vy) > 0.05) dn = 1;
            }
            note(mg, rb[0].x - (n3 - fm)); note(mg, rb[0].x - fm); note(mg, fabs(rb[0].y) - fm);
            {   /* margins of the checkpoint test: the x window of this counter value matters only if
                 * the ball crossed y = 0, the sign of the new y only inside the window (the old y is
                 * an input, identical in every precision) */
                const double wlo = cc == 0 ? n1 : cc == 1 ? n2 : (cc % 2 == 0 ? n3 : n3 - fm);
                const double whi = cc == 0 ? n0 : cc == 1 ? n1 : (cc % 2 == 0 ? n2 : n3);
                if ((lby >= 0) != (b->y >= 0)) { note(mg, b->x - wlo); note(mg, b->x - whi); }
                if (b->x < whi && b->x > wlo) note(mg, b->y);
            }
            if (rb[0].x < n3 - fm || rb[0].x > fm || fabs(rb[0].y) > fm) dn = 1;
            else if (cc == 0) {
                if (b->x < n0 && b->x > n1 && lby >= 0 && b->y < 0) { rew = 1; cc += 1; }
            } else if (cc == 1) {
                if (b->x < n1 && b->x > n2 && lby < 0 && b->y >= 0) { rew = 1; cc += 1; }
            } else if (cc % 2 == 0) {
                if (b->x < n2 && b->x > n3) {
                    if (lby >= 0 && b->y < 0) { rew = 1; cc += 1; if (cc == 7) dn = 1; }
                    else if (lby < 0 && b->y >= 0) dn = 1;
                }
            } else {
                if (b->x > n3 - fm && b->x < n3 && lby < 0 && b->y >= 0) { rew = 1; cc += 1; }
            }
            w->prev_pot[e] = (double)cc;
        } else {                                                       /* pass_endurance.py:126-150, 183-233 */
            const o_robot *sh = &rb[0], *rc = &rb[1];
            const double ld = sqrt((lbx - rc->x) * (lbx - rc->x) + (lby - rc->y) * (lby - rc->y));
            const double nd = sqrt((b->x - rc->x) * (b->x - rc->x) + (b->y - rc->y) * (b->y - rc->y));
            if (touching(p, rc, b, mg)) { rew += 1; dn = 1; }
            else {
                const double g = clampd(ld - nd, -1, 1) / ball_grad_scale;
                rew = g; info[1] += g;
            }
            /* __wrong_ball: centimetre-truncated box between shooter and receiver, 20-step stall counter */
            const double cb[2] = {b->x * 100, b->y * 100}, cs[2] = {sh->x * 100, sh->y * 100}, cr[2] = {rc->x * 100, rc->y * 100};
            int inside = 1;
            for (int k = 0; k < 2; ++k) {
                const int ib = (int)cb[k], is = (int)cs[k], ir = (int)cr[k];
                note(mg, (cb[k] - floor(cb[k])) / 100); note(mg, (ceil(cb[k]) - cb[k]) / 100);
                note(mg, (cs[k] - floor(cs[k])) / 100); note(mg, (ceil(cs[k]) - cs[k]) / 100);
                note(mg, (cr[k] - floor(cr[k])) / 100); note(mg, (ceil(cr[k]) - cr[k]) / 100);
                const int lo = is < ir ? is : ir, hi = is < ir ? ir : is;
                if (!(lo <= ib && ib <= hi)) inside = 0;
            }
            note(mg, fabs(ld - nd) - 0.01);
            int stopped_steps = (int)w->prev_pot[e];
            stopped_steps = fabs(ld - nd) < 0.01 ? stopped_steps + 1 : 0;
            w->prev_pot[e] = (double)stopped_steps;
            if (stopped_steps > 20 || !inside) { rew -= 1; dn = 1; }
            if (dn) {
                const double dr = sqrt((rc->x - sh->x) * (rc->x - sh->x) + (rc->y - sh->y) * (rc->y - sh->y));
                info[0] = (dr - nd) / dr;
            }
        }
        reward[e] = rew; done[e] = (uint8_t)dn;
        int tr = w->steps[e] >= max_steps;
        trunc[e] = (uint8_t)tr;
        if (auto_reset && (dn || tr)) {
            place(w, task, e, RS_STREAM_AUTORESET); clear_task(w, e);
            ssl_hw_obs(p, task, b, rb, 0.0, o, mg);
        }
    }
    w->t++;
}

/* obs of the current frame without stepping (env.reset() return value) */
void orc_task_obs(orc_world *w, int task, double *obs) {
    const rs_params *p = &w->p; const int R = p->n_robots;
    if (task == RS_TASK_VSS) {
        const int n_obs = 4 + 7 * p->n_blue + 5 * p->n_yellow;
        for (int e = 0; e < w->n; ++e) vss_obs(p, &w->ball[e], &w->rob[(size_t)e * R], obs + (size_t)e * n_obs);
        return;
    }
    if (task == RS_TASK_SSL_DRIBBLING || task == RS_TASK_SSL_PASS_ENDURANCE) {
        const int n_hw = task == RS_TASK_SSL_DRIBBLING ? RS_DRIB_OBS : RS_PASS_OBS;
        for (int e = 0; e < w->n; ++e)
            ssl_hw_obs(p, task, &w->ball[e], &w->rob[(size_t)e * R], w->steps[e] == 0 ? 0.0 : w->prev_pot[e],
                       obs + (size_t)e * n_hw, NULL);
        return;
    }
    const int n_obs = 4 + 8 * p->n_blue + 2 * p->n_yellow;
    const double mp = max_pos_of(p), max_v = 2.5, max_w = 10.0;
    for (int e = 0; e < w->n; ++e) {
        const o_ball *b = &w->ball[e]; const o_robot *rb = &w->rob[(size_t)e * R];
        double *o = obs + (size_t)e * n_obs; int k = 0;
        o[k++] = nrm(b->x, mp); o[k++] = nrm(b->y, mp); o[k++] = nrm(b->vx, max_v); o[k++] = nrm(b->vy, max_v);
        for (int r = 0; r < p->n_blue; ++r) {
            o[k++] = nrm(rb[r].x, mp); o[k++] = nrm(rb[r].y, mp);
            o[k++] = sin(rb[r].th); o[k++] = cos(rb[r].th);
            o[k++] = nrm(rb[r].vx, max_v); o[k++] = nrm(rb[r].vy, max_v);
            o[k++] = nrm(rb[r].om * DEG, max_w);
            o[k++] = touching(p, &rb[r], b, NULL) ? 1.0 : 0.0;
        }
        for (int r = p->n_blue; r < R; ++r) { o[k++] = nrm(rb[r].x, mp); o[k++] = nrm(rb[r].y, mp); }
    }
}

int orc_max_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

int orc_has_openmp(void) {
#ifdef _OPENMP
    return 1;
#else
    return 0;
#endif
}
