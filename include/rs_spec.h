/*
 * rs_spec.h -- the physics / world SPEC of the batched 2-D robot-soccer engine.
 *
 * This header is the contract both the CUDA library (rsoccer_b200/csrc) and the
 * CPU oracle (oracle/) implement.  It holds ONLY constants and the parameter
 * struct; no algorithm lives here (the oracle and the kernels restate the
 * algorithm independently from DESIGN.md section 3).
 *
 * What it replaces: the world construction done inside the third-party
 * `robosim.VSS(...)` / `robosim.SSL(...)` constructors that the reference calls at
 * rsoccer_gym/Simulators/rsim.py:116-124 and :169-177, and the 17-key dictionary
 * returned by `get_field_params()` (rsim.py:49-50, Entities/Field.py:4-21).
 *
 * PARITY UNPINNED: robosim (rc-robosim 1.2.0, an ODE 3-D simulator) is not in
 * /root/reference and cannot be installed, so every number below marked [MEM] is
 * the builder's recollection of rSim's config and every number marked [SPEC] is
 * a choice of this 2-D model.  Numbers marked [REF] are confirmed by the
 * reference tree itself (file:line given).
 */
#ifndef RS_SPEC_H
#define RS_SPEC_H

#ifdef __cplusplus
extern "C" {
#endif

#define RS_KIND_VSS 0
#define RS_KIND_SSL 1

#define RS_SUBSTEPS 5          /* [MEM] rSim advances dt/5 five times per step() */
#define RS_MAX_ROBOTS 22       /* 11 v 11 */
#define RS_MAX_BOXES 2

/* wire layout of get_state(): Entities/Frame.py:20-47 (VSS), :55-93 (SSL) [REF] */
#define RS_STATE_BALL 5        /* x y z vx vy */
#define RS_STATE_VSS_ROBOT 6   /* x y theta_deg vx vy vtheta_deg */
#define RS_STATE_SSL_ROBOT 11  /* + infrared w0 w1 w2 w3 */
/* command row width: rsim.py:92-93 (VSS, [wl, wr]) and :129-153 (SSL, 8) [REF] */
#define RS_CMD_VSS 2
#define RS_CMD_SSL 8

/* the 17 Field keys, in Entities/Field.py:5-21 order [REF] */
#define RS_FIELD_KEYS 17

typedef struct rs_params {
    int kind, field_type, n_blue, n_yellow, n_robots, time_step_ms;
    double dt;                 /* control step [s] = time_step_ms/1000 */
    double h;                  /* sub-step [s] = dt / RS_SUBSTEPS */

    /* ---- Field (Entities/Field.py) ---- */
    double length, width, penalty_length, penalty_width, goal_width, goal_depth;
    double ball_radius;
    double rbt_distance_center_kicker, rbt_kicker_thickness, rbt_kicker_width;
    double rbt_wheel_angle[4]; /* degrees */
    double rbt_radius, rbt_wheel_radius, rbt_motor_max_rpm;

    /* ---- walls, expressed in the mirrored quadrant (|x|, |y|) ---- */
    double x_out, y_out;       /* outer half extents (hard bounds for body centres + r) */
    double x_near;             /* no solid box starts before this |x| (quick reject) */
    int n_box;
    double box[RS_MAX_BOXES][4]; /* solid AABBs: lo.x lo.y hi.x hi.y */

    /* ---- bodies ---- */
    double ball_mass, rbt_mass;
    double e_ball_wall, e_rbt_wall, e_ball_rbt, e_rbt_rbt; /* restitution */
    double mu_ball_rbt;        /* Coulomb coefficient of the tangential ball<->robot impulse */
    double ball_decel;         /* rolling deceleration mu_roll * g  [m/s^2] */

    /* ---- drive ---- */
    double wheel_max_rad_s;    /* rpm * 2pi / 60 */
    double half_track;         /* VSS: wheel lateral offset [m], vss_gym_base.py:57-58 [REF] */
    double acc_fwd, acc_lat;   /* traction limited linear accel [m/s^2] (SSL: acc_lat == acc_fwd, isotropic) */
    double acc_ang;            /* traction limited angular accel [rad/s^2] */
    double omni_J[4][3];       /* SSL: wheel surface speed = J . (vx, vy, w) (robot frame) */
    double omni_Jpinv[3][4];   /* SSL: least-squares twist from 4 wheel surface speeds */

    /* ---- kicker / dribbler (SSL) ---- */
    double kick_centre;        /* |b.x - kick_centre| < kick_reach  and |b.y| < kick_half_width  => touching */
    double kick_reach;
    double kick_half_width;
    double mouth_half_chord;   /* sqrt(R^2 - dk^2) */
    double kick_speed_max;
} rs_params;

/* Fills *p for (kind, field_type, n_blue, n_yellow, time_step_ms).
 * Returns 0, or -1 on an unknown world / bad sizes.  Implemented in rs_spec.c
 * style below as static so that the oracle and the library need no extra TU. */
static inline int rs_params_fill(rs_params *p, int kind, int field_type, int n_blue,
                                 int n_yellow, int time_step_ms);

/* ------------------------------------------------------------------------- */
static inline double rs__sqrt(double x) {          /* Newton, avoids <math.h> in a header */
    if (x <= 0.0) return 0.0;
    double r = x > 1.0 ? x : 1.0;
    for (int i = 0; i < 64; ++i) r = 0.5 * (r + x / r);
    return r;
}
/* sin/cos of an angle in degrees by Taylor series after reduction; used only to build J */
static inline void rs__sincos_deg(double deg, double *s, double *c) {
    const double PI = 3.14159265358979323846;
    while (deg > 180.0) deg -= 360.0;
    while (deg <= -180.0) deg += 360.0;
    double x = deg * PI / 180.0, x2 = x * x;
    double ts = x, tc = 1.0, ss = 0.0, cs = 0.0;
    for (int k = 0; k < 20; ++k) {
        ss += ts; cs += tc;
        tc = -tc * x2 / ((2.0 * k + 1.0) * (2.0 * k + 2.0));
        ts = -ts * x2 / ((2.0 * k + 2.0) * (2.0 * k + 3.0));
    }
    *s = ss; *c = cs;
}

static inline int rs_params_fill(rs_params *p, int kind, int field_type, int n_blue,
                                 int n_yellow, int time_step_ms) {
    const double PI = 3.14159265358979323846;
    const double G = 9.81;                       /* [MEM] rSim gravity */
    const double BIG = 1.0e9;
    if (!p) return -1;
    if (n_blue < 0 || n_yellow < 0 || n_blue + n_yellow < 1 ||
        n_blue + n_yellow > RS_MAX_ROBOTS || time_step_ms <= 0)
        return -1;
    char *z = (char *)p;
    for (unsigned i = 0; i < sizeof(*p); ++i) z[i] = 0;
    p->kind = kind; p->field_type = field_type;
    p->n_blue = n_blue; p->n_yellow = n_yellow; p->n_robots = n_blue + n_yellow;
    p->time_step_ms = time_step_ms;
    p->dt = time_step_ms / 1000.0;
    p->h = p->dt / RS_SUBSTEPS;

    if (kind == RS_KIND_VSS) {
        if (field_type == 0) {          /* 3v3: Render/field.py:189-199 [REF] */
            p->length = 1.5; p->width = 1.3; p->penalty_length = 0.15;
            p->penalty_width = 0.7; p->goal_width = 0.4; p->goal_depth = 0.1;
        } else if (field_type == 1) {   /* 5v5 (vss/README.md:4) [MEM] */
            p->length = 2.2; p->width = 1.8; p->penalty_length = 0.15;
            p->penalty_width = 0.8; p->goal_width = 0.4; p->goal_depth = 0.15;
        } else return -1;
        p->ball_radius = 0.0215;         /* Render/ball.py:6 [REF] */
        p->rbt_radius = 0.0375;          /* vss_gym_base.py:57 comment [REF] */
        p->rbt_wheel_radius = 0.026;     /* [MEM] */
        p->rbt_motor_max_rpm = 440.0;    /* [MEM] => max_v = 1.198 m/s */
        p->rbt_distance_center_kicker = 0.0; p->rbt_kicker_thickness = 0.0;
        p->rbt_kicker_width = 0.0;
        p->rbt_wheel_angle[0] = 90.0; p->rbt_wheel_angle[1] = 270.0;
        p->rbt_wheel_angle[2] = 0.0; p->rbt_wheel_angle[3] = 0.0;
        p->half_track = 0.04;            /* vss_gym_base.py:57-58 [REF] */
        p->ball_mass = 0.046; p->rbt_mass = 0.18;              /* [MEM] */
        p->e_ball_wall = 0.6; p->e_rbt_wall = 0.0;             /* [SPEC] */
        p->e_ball_rbt = 0.4; p->e_rbt_rbt = 0.0;               /* [SPEC] */
        p->mu_ball_rbt = 0.3;                                  /* [SPEC] */
        p->ball_decel = 0.05 * G;                              /* [MEM] mu 0.05 */
        p->acc_fwd = 6.0; p->acc_lat = 9.0; p->acc_ang = 200.0; /* [SPEC] */
        /* walls: the playable region is the field rectangle plus the two goal
         * recesses; in the mirrored quadrant that is the rectangle
         * [0, L/2+gd] x [0, W/2] minus the solid corner [L/2, inf) x [gw/2, inf) */
        p->x_out = p->length / 2 + p->goal_depth; p->y_out = p->width / 2;
        p->x_near = p->length / 2;
        p->n_box = 1;
        p->box[0][0] = p->length / 2; p->box[0][1] = p->goal_width / 2;
        p->box[0][2] = BIG; p->box[0][3] = BIG;
        p->box[1][0] = BIG; p->box[1][1] = BIG; p->box[1][2] = 2 * BIG; p->box[1][3] = 2 * BIG;
    } else if (kind == RS_KIND_SSL) {
        if (field_type == 0) {          /* div B, Render/field.py:252-263 [REF]; ssl/README.md:4 */
            p->length = 9.0; p->width = 6.0; p->penalty_length = 1.0;
            p->penalty_width = 2.0; p->goal_width = 1.0; p->goal_depth = 0.18;
        } else if (field_type == 1) {   /* div A [MEM] */
            p->length = 12.0; p->width = 9.0; p->penalty_length = 1.8;
            p->penalty_width = 3.6; p->goal_width = 1.8; p->goal_depth = 0.18;
        } else if (field_type == 2) {   /* 2021 hardware challenge field [MEM] */
            p->length = 6.0; p->width = 4.0; p->penalty_length = 0.8;
            p->penalty_width = 1.8; p->goal_width = 0.8; p->goal_depth = 0.18;
        } else return -1;
        p->ball_radius = 0.0215;                               /* [MEM] */
        p->rbt_radius = 0.09;            /* ssl_gym_base.py:58 comment [REF] */
        p->rbt_wheel_radius = 0.02475;                         /* [MEM] */
        p->rbt_motor_max_rpm = 160.0 * 60.0 / (2.0 * PI);      /* 160 rad/s: static_defenders.py:71 [REF] */
        p->rbt_distance_center_kicker = 0.073;                 /* [MEM] */
        p->rbt_kicker_thickness = 0.005; p->rbt_kicker_width = 0.08; /* [MEM] */
        p->rbt_wheel_angle[0] = 60.0; p->rbt_wheel_angle[1] = 135.0;
        p->rbt_wheel_angle[2] = 225.0; p->rbt_wheel_angle[3] = 300.0; /* [MEM] */
        p->half_track = 0.0;
        p->ball_mass = 0.043; p->rbt_mass = 2.2;               /* [MEM] */
        p->e_ball_wall = 0.5; p->e_rbt_wall = 0.0;             /* [SPEC] */
        p->e_ball_rbt = 0.3; p->e_rbt_rbt = 0.0;               /* [SPEC] */
        p->mu_ball_rbt = 0.3;                                  /* [SPEC] */
        p->ball_decel = 0.05 * G;                              /* [MEM] */
        p->acc_fwd = 5.0; p->acc_lat = 5.0; p->acc_ang = 50.0; /* [SPEC] */
        {
            const double margin = 0.7;   /* field margin 0.3 + referee margin 0.4 [MEM] */
            const double t = 0.02;       /* goal wall thickness [MEM] */
            p->x_out = p->length / 2 + margin; p->y_out = p->width / 2 + margin;
            p->x_near = p->length / 2;
            p->n_box = 2;
            /* goal side wall */
            p->box[0][0] = p->length / 2;               p->box[0][1] = p->goal_width / 2;
            p->box[0][2] = p->length / 2 + p->goal_depth + t; p->box[0][3] = p->goal_width / 2 + t;
            /* goal back wall (spans y=0 in the mirrored quadrant) */
            p->box[1][0] = p->length / 2 + p->goal_depth; p->box[1][1] = -(p->goal_width / 2 + t);
            p->box[1][2] = p->length / 2 + p->goal_depth + t; p->box[1][3] = p->goal_width / 2 + t;
        }
        for (int i = 0; i < 4; ++i) {
            double s, c;
            rs__sincos_deg(p->rbt_wheel_angle[i], &s, &c);
            p->omni_J[i][0] = -s; p->omni_J[i][1] = c; p->omni_J[i][2] = p->rbt_radius;
        }
        /* Jpinv = (J^T J)^-1 J^T, 3x3 inverse by cofactors */
        {
            double A[3][3];
            for (int a = 0; a < 3; ++a) for (int b = 0; b < 3; ++b) {
                A[a][b] = 0.0;
                for (int i = 0; i < 4; ++i) A[a][b] += p->omni_J[i][a] * p->omni_J[i][b];
            }
            double det = A[0][0] * (A[1][1] * A[2][2] - A[1][2] * A[2][1])
                       - A[0][1] * (A[1][0] * A[2][2] - A[1][2] * A[2][0])
                       + A[0][2] * (A[1][0] * A[2][1] - A[1][1] * A[2][0]);
            double inv[3][3];
            inv[0][0] =  (A[1][1] * A[2][2] - A[1][2] * A[2][1]) / det;
            inv[0][1] = -(A[0][1] * A[2][2] - A[0][2] * A[2][1]) / det;
            inv[0][2] =  (A[0][1] * A[1][2] - A[0][2] * A[1][1]) / det;
            inv[1][0] = -(A[1][0] * A[2][2] - A[1][2] * A[2][0]) / det;
            inv[1][1] =  (A[0][0] * A[2][2] - A[0][2] * A[2][0]) / det;
            inv[1][2] = -(A[0][0] * A[1][2] - A[0][2] * A[1][0]) / det;
            inv[2][0] =  (A[1][0] * A[2][1] - A[1][1] * A[2][0]) / det;
            inv[2][1] = -(A[0][0] * A[2][1] - A[0][1] * A[2][0]) / det;
            inv[2][2] =  (A[0][0] * A[1][1] - A[0][1] * A[1][0]) / det;
            for (int a = 0; a < 3; ++a) for (int i = 0; i < 4; ++i) {
                p->omni_Jpinv[a][i] = 0.0;
                for (int b = 0; b < 3; ++b) p->omni_Jpinv[a][i] += inv[a][b] * p->omni_J[i][b];
            }
        }
        /* grSim-style `isTouchingBall` box in front of the flat mouth [MEM] */
        p->kick_centre = p->rbt_distance_center_kicker + 0.5 * p->rbt_kicker_thickness;
        p->kick_reach = 2.0 * p->rbt_kicker_thickness + p->ball_radius;
        p->kick_half_width = 0.5 * p->rbt_kicker_width;
        p->mouth_half_chord = rs__sqrt(p->rbt_radius * p->rbt_radius -
            p->rbt_distance_center_kicker * p->rbt_distance_center_kicker);
        p->kick_speed_max = 6.5;         /* SSL rule limit [SPEC] */
    } else return -1;

    p->wheel_max_rad_s = p->rbt_motor_max_rpm * 2.0 * PI / 60.0;
    return 0;
}

/* Field dict values in Entities/Field.py:5-21 key order */
static inline void rs_params_field(const rs_params *p, double out[RS_FIELD_KEYS]) {
    out[0] = p->length; out[1] = p->width; out[2] = p->penalty_length;
    out[3] = p->penalty_width; out[4] = p->goal_width; out[5] = p->goal_depth;
    out[6] = p->ball_radius; out[7] = p->rbt_distance_center_kicker;
    out[8] = p->rbt_kicker_thickness; out[9] = p->rbt_kicker_width;
    out[10] = p->rbt_wheel_angle[0]; out[11] = p->rbt_wheel_angle[1];
    out[12] = p->rbt_wheel_angle[2]; out[13] = p->rbt_wheel_angle[3];
    out[14] = p->rbt_radius; out[15] = p->rbt_wheel_radius; out[16] = p->rbt_motor_max_rpm;
}

/* ---- task-level constants (the "env.step()" half of the path) ---- */
#define RS_VSS_OBS 40                 /* vss_gym.py:65-67 [REF] */
#define RS_VSS_ACT 2                  /* vss_gym.py:64 [REF] */
#define RS_VSS_NOISE 10               /* 5 OU robots x 2 wheels, vss_gym.py:128-140 [REF] */
#define RS_VSS_INFO 6                 /* goal_score move ball_grad energy goals_blue goals_yellow, vss_gym.py:150-158 */
#define RS_VSS_MAX_STEPS 1200         /* rsoccer_gym/__init__.py:4 [REF] */
#define RS_OU_THETA 0.17              /* Utils/Utils.py:6 [REF] */
#define RS_OU_SIGMA 0.5               /* (high - mu)/2 with Box(-1,1): Utils.py:8-9 [REF] */
#define RS_VSS_DEADZONE 0.05          /* vss_gym.py:73 [REF] */
#define RS_NORM_BOUNDS 1.2            /* vss_gym_base.py:26 [REF] */
#define RS_SSL_ACT 5                  /* static_defenders.py:54 [REF] */
#define RS_SSL_INFO 9                 /* goal rbt_in_gk_area done_ball_out done_ball_out_right done_rbt_out ball_dist ball_grad energy collision */
#define RS_TASK_VSS 0
#define RS_TASK_SSL_STATIC_DEFENDERS 1
#define RS_TASK_SSL_CONTESTED_POSSESSION 2
#define RS_TASK_SSL_DRIBBLING 3        /* SSLDribbling-v0: 1 blue + 4 yellow, dribbling.py:47-49 [REF] */
#define RS_TASK_SSL_PASS_ENDURANCE 4   /* SSLPassEndurance-v0: 2 blue, pass_endurance.py:46-52 [REF] */
#define RS_DRIB_ACT 4                  /* dribbling.py:50-51 [REF] */
#define RS_DRIB_OBS 21                 /* 5 + 8 nb + 2 ny, dribbling.py:53 [REF] */
#define RS_DRIB_MAX_STEPS 4800         /* rsoccer_gym/__init__.py:17 [REF] */
#define RS_PASS_ACT 3                  /* pass_endurance.py:53 [REF] */
#define RS_PASS_OBS 16                 /* 4 + 6 nb, pass_endurance.py:55 [REF] */
#define RS_PASS_MAX_STEPS 1200         /* rsoccer_gym/__init__.py:29 [REF] */

/* Philox stream ids (counter word 2) */
#define RS_STREAM_OU 0u
#define RS_STREAM_AUTORESET 1u
#define RS_STREAM_RESET 2u

#ifdef __cplusplus
}
#endif
#endif /* RS_SPEC_H */
