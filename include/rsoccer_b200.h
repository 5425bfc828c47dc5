/*
 * rsoccer_b200.h -- C ABI of the B200 batched robot-soccer engine
 * (librsoccer_b200.so, built from rsoccer_b200/csrc/ for sm_100a).
 *
 * Drop-in boundary.  The reference crosses into native code at exactly five
 * pybind11 methods of the third-party `robosim` module, all called from
 * rsoccer_gym/Simulators/rsim.py.  Each entry point below names the reference
 * call it replaces.  The ABI is plain C: opaque handle, raw pointers, sizes,
 * `void *stream` (a cudaStream_t), int status.  No torch / pybind types.
 *
 *   robosim.VSS(...) / robosim.SSL(...)   rsim.py:116-124, 169-177  -> rs_create
 *   del simulator (RSim.stop)             rsim.py:40-41             -> rs_destroy
 *   simulator.get_field_params()          rsim.py:49-50             -> rs_field_params
 *   simulator.reset(ball, blue, yellow)   rsim.py:36-38, 52-75      -> rs_reset
 *   simulator.step(cmds)                  rsim.py:102, 155          -> rs_step
 *   simulator.get_state()                 rsim.py:105, 158          -> rs_get_state
 *
 * and, one level up (SURVEY section 8(f) "next" rows 1-2), the whole
 * `env.step()` of the three benchmarked task envs as ONE fused launch:
 *
 *   VSSEnv.step            vss/env_vss/vss_gym.py:89-91 (+ vss_gym_base.py:72-90)
 *                                                                    -> rs_vss_env_step
 *   SSLHWStaticDefendersEnv.step   ssl/ssl_hw_challenge/static_defenders.py:86-88
 *   SSLContestedPossessionEnv.step ssl/ssl_hw_challenge/contested_possession.py:74-76
 *                                                                    -> rs_ssl_env_step
 *   env.reset()            vss_gym_base.py:92-106, vss_gym.py:194-233 -> rs_task_reset
 *
 * Conventions
 *  - one rs_world = N independent matches ("envs") of one world kind on one GPU.
 *  - every `d_*` pointer is a DEVICE pointer owned by the caller (e.g. a torch
 *    tensor's data_ptr()); every `h_*` pointer is a HOST pointer.  The library
 *    never retains a caller pointer past the call, except the state buffer given
 *    to rs_bind_state, which the caller keeps alive until rs_destroy.
 *  - calls are asynchronous on `stream` unless stated; a handle is not thread
 *    safe (one caller at a time; handles are independent of each other); there is
 *    no global state besides the per-thread error string -- every tuning switch is
 *    per handle, and the process environment is read once, in rs_create.
 *  - every call runs on the world's own CUDA device (rs_create's `device`) whatever
 *    device is current in the calling thread, and restores the caller's device.
 *  - return 0 on success, <0 on error (RS_E_*); rs_last_error() has the text.
 *  - there is NO CPU fallback: rs_create fails if no CUDA device is usable.
 *  - units on the wire follow the reference (Entities/Frame.py:8): m, m/s,
 *    degrees, degrees/s; robot rows are blue ids first, then yellow
 *    (rsim.py:96-99).
 */
#ifndef RSOCCER_B200_H
#define RSOCCER_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RS_OK 0
#define RS_E_INVALID (-1)     /* bad argument */
#define RS_E_CUDA (-2)        /* CUDA runtime error */
#define RS_E_STATE (-3)       /* state buffer not bound */
#define RS_E_UNSUPPORTED (-4) /* world/task combination not built */

typedef struct rs_world rs_world;

/* names of the SoA arrays inside the state buffer, for rs_layout */
enum {
    RS_ARR_BODY = 0,   /* float4 [R+1][Np]  (x, y, vx, vy); body 0 = ball, then robots */
    RS_ARR_ANG = 1,    /* float2 [R][Np]    (theta [rad, (-pi, pi]], omega [rad/s]) */
    RS_ARR_OU = 2,     /* float2 [R-1][Np]  OU process state of the non-agent robots (VSS-v0) */
    RS_ARR_PREV = 3,   /* float  [Np]       previous ball potential (VSS-v0) */
    RS_ARR_STEPS = 4,  /* int32  [Np]       episode step counter | has_prev << 24 */
    RS_ARR_INFO = 5,   /* float  [9][Np]    reward_shaping_total accumulators */
    RS_ARR_COUNT = 6
};

int rs_version(void);
const char *rs_last_error(void);

/* robosim.VSS / robosim.SSL constructor (rsim.py:116-124, 169-177).
 * kind: 0 VSS, 1 SSL.  seed/env_offset key the on-device Philox streams by
 * GLOBAL env id = env_offset + local index, so results do not depend on how
 * envs are sharded over GPUs.  `device` < 0 keeps the current CUDA device. */
int rs_create(int kind, int field_type, int n_blue, int n_yellow, int time_step_ms,
              int n_envs, int device, uint64_t seed, int64_t env_offset, rs_world **out);
/* RSim.stop (rsim.py:40-41) */
int rs_destroy(rs_world *w);

/* state memory is caller-owned: allocate rs_state_bytes() bytes of device memory
 * (256-byte aligned), then bind it.  rs_bind_state zero-fills it and places the
 * robots at the reference's dummy initial poses (rsim.py:19-24). */
size_t rs_state_bytes(const rs_world *w);
int rs_bind_state(rs_world *w, void *d_state, void *stream);
/* byte offset of each RS_ARR_* array inside the state buffer and the padded env
 * count Np (out_offsets[RS_ARR_COUNT], *out_np) so the caller can build views */
int rs_layout(const rs_world *w, int64_t *out_offsets, int64_t *out_np);

/* get_field_params (rsim.py:49-50): the 17 Field values in Entities/Field.py:5-21 order */
int rs_field_params(const rs_world *w, double out[17]);

/* simulator.reset (rsim.py:38): d_ball [N][4] = x y vx vy; d_blue [N][nb][3],
 * d_yellow [N][ny][3] = x y theta_deg.  Robot velocities are zeroed, task state
 * (OU, potential, step counter) cleared.  d_mask (nullable) [N] uint8: only
 * envs with mask != 0 are touched. */
int rs_reset(rs_world *w, const float *d_ball, const float *d_blue, const float *d_yellow,
             const uint8_t *d_mask, void *stream);

/* simulator.step (rsim.py:102 VSS: d_cmds [N][R][2] = wheel rad/s left,right;
 * rsim.py:155 SSL: d_cmds [N][R][8] = flag, w0..w3 | vx vy vtheta 0, kick_x, kick_z, dribbler).
 * Advances every env by one control step (5 sub-steps). */
int rs_step(rs_world *w, const float *d_cmds, void *stream);

/* simulator.get_state (rsim.py:105, 158): d_out [N][5 + K R] floats in the
 * Entities/Frame.py layout (K = 6 VSS, 11 SSL) */
int rs_get_state(const rs_world *w, float *d_out, void *stream);

/* raw internal state in/out, [N][4 + 6 R] = ball x y vx vy, robot x y theta_rad
 * vx vy omega (superset of reset: also sets robot velocities; checkpoint/restore
 * and re-synced parity tests) */
int rs_set_raw(rs_world *w, const float *d_in, void *stream);
int rs_get_raw(const rs_world *w, float *d_out, void *stream);

/* World step counter t = Philox counter word 1: 31 bits, wraps modulo 2^31 (rs_set_t rejects
 * larger values; bit 31 of its device copies is the tile lock of RS_OPT_STEP_OVERLAP).  It is advanced ON THE DEVICE by every task-level step (rs_vss_env_step,
 * rs_ssl_env_step) -- so a captured CUDA graph replays with fresh noise -- and by nothing else:
 * rs_step draws no random numbers and leaves it alone.  rs_get_t returns the host mirror,
 * exact unless launches were replayed from a graph; rs_sync_t (blocking) reads the device
 * value back, or, if an rs_set_t is still pending, writes it.  A pending rs_set_t is applied
 * by the next step launch; that launch fails with RS_E_STATE inside a stream capture (the
 * write would be replayed with the graph and rewind the counter). */
uint64_t rs_get_t(const rs_world *w);
int rs_set_t(rs_world *w, uint64_t t);
int rs_sync_t(rs_world *w, void *stream);

/* ---- per-handle options ---- */
/* RS_OPT_STEP_OVERLAP (default 0; rs_create reads RS_STEP_OVERLAP from the environment):
 * consecutive rs_vss_env_step launches of one world on one stream synchronise per 32-match
 * tile instead of grid-wide, so that the tail of step k (slow tiles, store drain) overlaps the
 * head of step k+1 (launch, state loads, noise).  Results are bit-identical to mode 0.
 *   0  off: every step begins with a grid-wide wait on the previous kernel of the stream.
 *   1  the world state is synchronised per tile; caller buffers (d_actions, d_normals) are
 *      read only after a grid-wide wait.  Always safe.
 *   2  per tile only, no grid-wide wait at all.  The caller vouches that d_actions /
 *      d_normals were completely written before the PREVIOUS step launch of this world was
 *      enqueued (fixed or pre-generated action buffers, action repeat).  Any other work
 *      enqueued between two steps (a policy kernel, a copy) serialises them as usual.
 *   3  as 2, with the step kernel built for 11 instead of 7 resident CTAs per SM: for several
 *      worlds stepped round-robin on one stream (consecutive launches are then independent
 *      and fill the GPU together); slower than 2 when one world is stepped again and again.
 * Whoever writes the state buffer behind the library's back (through the zero-copy views)
 * sets the option again afterwards: the next step then starts with a grid-wide wait.
 * RS_OPT_PDL (default 1): launch step kernels with programmatic stream serialization.
 * RS_OPT_OVERLAP_ERRORS (read only, blocking): tiles whose wait timed out -- always 0 unless
 * one world was stepped from two streams at once.
 * RS_OPT_HOST_COPY_ACTIONS (default -1; RS_HOST_COPY_ACTIONS in the environment): how the *_host steps
 * fetch PINNED host actions.  0: the step kernel reads them in place over PCIe (no H2D copy).  1: staged
 * with a copy.  -1: in place only when a match's action row is 8 bytes (VSS-v0: coalesced), else the copy
 * (the 20-byte rows of the SSL tasks would cross the link five times).  Pageable memory is always copied. */
#define RS_OPT_STEP_OVERLAP 1
#define RS_OPT_PDL 2
#define RS_OPT_OVERLAP_ERRORS 3
#define RS_OPT_HOST_COPY_ACTIONS 4
int rs_set_option(rs_world *w, int option, int64_t value);
int rs_get_option(const rs_world *w, int option, int64_t *value, void *stream);

/* ---- task level: the env.step() of the benchmarked reference envs ---- */
#define RS_TASK_VSS_V0 0
#define RS_TASK_SSL_STATIC_DEFENDERS_V0 1
#define RS_TASK_SSL_CONTESTED_POSSESSION_V0 2
#define RS_TASK_SSL_DRIBBLING_V0 3          /* world: SSL, field_type 2, 1 blue + 4 yellow (dribbling.py:47-49) */
#define RS_TASK_SSL_PASS_ENDURANCE_V0 4     /* world: SSL, field_type 2, 2 blue (pass_endurance.py:46-52) */

/* observation width of a task for this world (40 / 24 / 14 / 21 / 16 at the reference sizes) */
int rs_task_obs_dim(const rs_world *w, int task);
/* action width of a task: 2 / 5 / 5 / 4 / 3 */
int rs_task_act_dim(int task);

/* env.reset() for the envs selected by d_mask (nullable = all): random initial
 * frame per the task's _get_initial_positions_frame, drawn on device; writes the
 * first observation of the reset envs into d_obs (nullable) [N][obs_dim]. */
int rs_task_reset(rs_world *w, int task, const uint8_t *d_mask, float *d_obs, void *stream);

/* VSSEnv.step for all N envs in one launch: OU commands for the 5 other robots,
 * action -> wheel speeds, physics, observation, reward, done, TimeLimit
 * truncation, reward_shaping_total, masked auto-reset.
 *  d_actions [N][2] in [-1,1]; d_normals (nullable) [N][2(R-1)] standard normals
 *  that replace the on-device Philox draw (parity harness); d_obs [N][40], 16-byte aligned
 *  (d_actions 8-byte aligned); d_reward [N]; d_done [N]; d_trunc [N]; d_cmds_out (nullable)
 *  [N][R][2]. */
int rs_vss_env_step(rs_world *w, const float *d_actions, const float *d_normals, int auto_reset,
                    int max_steps, float *d_obs, float *d_reward, uint8_t *d_done,
                    uint8_t *d_trunc, float *d_cmds_out, void *stream);

/* SSLHWStaticDefendersEnv.step / SSLContestedPossessionEnv.step (d_actions [N][5]),
 * SSLHWDribblingEnv.step (d_actions [N][4], obs [N][21]) and SSLPassEnduranceEnv.step
 * (d_actions [N][3], obs [N][16]); rows are rs_task_act_dim / rs_task_obs_dim wide */
int rs_ssl_env_step(rs_world *w, int task, const float *d_actions, int auto_reset, int max_steps,
                    float *d_obs, float *d_reward, uint8_t *d_done, uint8_t *d_trunc,
                    float *d_cmds_out, void *stream);

/* End-to-end variants with HOST buffers (what a binding holding numpy arrays
 * calls): H2D copy of the actions, the fused step, D2H copy of obs / reward /
 * done / trunc, then a stream synchronize.  Host buffers should be pinned. */
int rs_vss_env_step_host(rs_world *w, const float *h_actions, int auto_reset, int max_steps,
                         float *h_obs, float *h_reward, uint8_t *h_done, uint8_t *h_trunc,
                         void *stream);
int rs_ssl_env_step_host(rs_world *w, int task, const float *h_actions, int auto_reset,
                         int max_steps, float *h_obs, float *h_reward, uint8_t *h_done,
                         uint8_t *h_trunc, void *stream);

/* Split-phase host steps (the step_async / step_wait pair of gymnasium's VectorEnv): *_begin enqueues the actions'
 * fetch, the fused step and the D2H copies on `stream` and returns at once; rs_host_step_wait blocks until the
 * outputs of that step have landed in the host buffers (no-op when nothing is pending).  One step may be pending
 * per world (a second *_begin / *_host returns RS_E_STATE); h_actions and the output buffers belong to the library
 * until the wait returns.  What it is for: double-buffered env groups -- two worlds on two streams, one computes
 * while the other's observations cross PCIe, the consumer works on one group while the other steps (bench.py
 * `e2e.pipelined`) -- or overlapping the transfer with the caller's own host work. */
int rs_vss_env_step_host_begin(rs_world *w, const float *h_actions, int auto_reset, int max_steps,
                               float *h_obs, float *h_reward, uint8_t *h_done, uint8_t *h_trunc,
                               void *stream);
int rs_ssl_env_step_host_begin(rs_world *w, int task, const float *h_actions, int auto_reset,
                               int max_steps, float *h_obs, float *h_reward, uint8_t *h_done,
                               uint8_t *h_trunc, void *stream);
int rs_host_step_wait(rs_world *w);

/* number of kernels this handle has launched so far (bench.py gpu_launches); atomic, so
 * the const getters above may run concurrently with each other */
uint64_t rs_launch_count(const rs_world *w);

/* diagnostics: launches a kernel with the grid, CTA size, parameter block and launch attributes of a
 * lane-per-match step of this world and NO work (chain == 0: it waits for its predecessor like a serialised
 * step; != 0: it only triggers its dependents, like an overlapped one) -- the launch floor a step is measured
 * against (tools/launch_floor.py) */
int rs_debug_empty_step(rs_world *w, int chain, void *stream);

/* which kernels step this world (diagnostics; the results do not depend on it):
 *   bit 0  task kernels run one lane per BODY (else one lane per MATCH)
 *   bit 1  rs_step runs one lane per BODY
 *   bit 2  physics constants are compile-time immediates (VSS, field_type 0, 25 ms)
 *   bit 3  the VSS-v0 task kernel uses the packed fp32x2 instruction forms (worlds of >= 20 480 matches, and every
 *          size under RS_OPT_STEP_OVERLAP = 3; same results) */
int rs_kernel_flags(const rs_world *w);

#ifdef __cplusplus
}
#endif
#endif /* RSOCCER_B200_H */
