/*
 * glg_oracle.h -- CPU parity oracle for the GreenLight env-step path.  TEST INFRASTRUCTURE ONLY.
 *
 * Plain-C fp64 restatement of the reference algorithm; only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs may link or call it.  The product (libglgym.so)
 * never does.
 *
 * Parity pinning status:
 *   - RHS (R1,R2: aux_states.hpp:96-1271, ode.hpp:6-124): PINNED against golden vectors produced from the
 *     reference's own source text (tests/golden/make_golden.py -> tests/golden/rhs_golden.npz).
 *   - Step semantics (S1-S8: tomato_env.py, observations.py, rewards.py, noise.py, utils.py:init_state):
 *     PINNED against traces produced by executing the reference's own Python env on top of this
 *     oracle's evalF (reference-shell, tests/golden/make_golden.py -> tests/golden/shell_trace.npz).
 *   - Integrator (R3: greenlight_model.cpp:43-63): the reference integrates with CasADi 3.6.7 / SUNDIALS
 *     CVODES (BDF, abstol=reltol=1e-6), a third-party dependency absent from /root/reference and from
 *     this image => PARITY UNPINNED at the CVODES boundary.  The contract here is BASELINE.json's: fixed
 *     step classical RK4, n_sub substeps, u/d/p held constant over [0,dt] (zero-order hold, same as
 *     greenlight_model.cpp:59-63).  Method error vs. a tight implicit solve is measured in tests.
 */
#ifndef GLG_ORACLE_H
#define GLG_ORACLE_H

#ifdef __cplusplus
extern "C" {
#endif

#define GLGO_NX 28
#define GLGO_NU 6
#define GLGO_ND 10
#define GLGO_NP 208
#define GLGO_NA 239
#define GLGO_NOBS_FIXED 23 /* obs = 23 + 5*Np */
#define GLGO_NINFO 11

/* R1: update() -- 239 auxiliary values. a may be NULL. dxdt may be NULL. */
void glgo_aux_rhs(const double *x, const double *u, const double *d, const double *p, double *a, double *dxdt);
/* R2: ODE() */
void glgo_rhs(const double *x, const double *u, const double *d, const double *p, double *dxdt);
/* glgo_evalf with options: stiff_guard != 0 adds the transient-stiffness micro-step rule; n_micro (may be NULL) receives the
 * number of RK4 micro-steps executed. */
int glgo_evalf_ex(const double *x, const double *u, const double *d, const double *p, double dt, int n_sub, int stiff_guard,
                  double *x_next, long *n_micro);
/* R3/R4: x_next = RK4^n_sub(x; u,d,p const).  returns 0, or 1 if the result is not finite */
int glgo_evalf(const double *x, const double *u, const double *d, const double *p, double dt, int n_sub,
               double *x_next);
/* CVODES-class CPU baseline (glg_oracle_bdf.c): adaptive variable-order BDF/NDF 1..5, simplified Newton, dense LU, forward-
 * difference Jacobian.  J_io [28*28] + jac_valid_io carry the Jacobian between calls (both may be NULL); stats[4] (may be
 * NULL) += {rhs evaluations, Jacobian evaluations, LU factorisations, accepted steps}.  Not a parity checker. */
int glgo_evalf_bdf(const double *x, const double *u, const double *d, const double *p, double dt, double rtol, double atol,
                   double *x_next, double *J_io, int *jac_valid_io, long *stats);
/* batched, AoS rows: x[B][28] u[B][6] d[B][10] p[B][208] (p_stride = 0 => shared p) ; OpenMP over envs */
int glgo_evalf_batch(const double *x, const double *u, const double *d, const double *p, int p_stride, double dt,
                     int n_sub, double *x_next, int B, int n_threads);

/* ---- step semantics (S1..S8) ---- */
typedef struct glgo_env_cfg {
    double dt;              /* 900 */
    int n_sub;              /* RK4 substeps per step */
    int N;                  /* season_length*86400/dt, last step index */
    int Np;                 /* forecast horizon in steps */
    double delta_u_max_f32; /* float32(0.1) widened */
    double u_min[6], u_max[6];
    double con_low[3], con_high[3]; /* co2 ppm, temp, rh */
    double elec_price, heating_price, co2_price, fruit_price, dmfm;
    double uncertainty_scale;
    double fixed_costs;     /* rewards.py:69-70,154: yearly/365/(86400//dt); reported in info only */
    int stiff_guard;        /* integrator of glgo_env_step: 0 = the fixed-step RK4 contract; bits 0/1 = graded RK4 rules
                               (glgo_evalf_ex); 16 = adaptive implicit BDF at rtol = atol = 1e-6 (glgo_evalf_bdf, the
                               CVODES-class CPU baseline; n_micro then counts right-hand-side evaluations) */
    int obs_modules[8];     /* ordered observation module ids (tomato_env.py:77-96), 0-terminated; empty = default stack
                               {2,3,4,5,6,7}: 1 State(27) 2 IndoorClimate(4) 3 BasicCrop(3) 4 Control(6) 5 Weather(5) 6 Time(5)
                               7 WeatherForecast(5 Np) */
} glgo_env_cfg;

typedef struct glgo_env {
    double x[28], x_prev[28], u[6];
    double day_of_year, hour_of_day;
    int timestep;
    int terminated;
    const double *weather; /* [rows][10] */
    int weather_rows;
    long n_micro;          /* RK4 micro-steps executed by the last step (implicit solver: right-hand-side evaluations) */
    double *jac;           /* optional [28*28] Jacobian carried between steps by the implicit solver (stiff_guard bit 5) */
    int jac_valid;
    const double *state_obs; /* optional [27]: the StateObservations entries of the next observation (random in the reference) */
} glgo_env;

void glgo_init_state(const double *d0, double *x);                                   /* utils.py:13-46 */
void glgo_env_reset(glgo_env *e, const double *weather, int rows, double start_day); /* tomato_env.py:231-270 */
/* S2: noise.py:3-23 given the 34 uniform draws n_i in (-s/2, s/2) (already scaled). p_out float32-rounded. */
void glgo_param_noise(const double *p_nom, const double *noise34, double *p_out);
/* length of the observation row for c->obs_modules (23 + 5 Np for the default stack) */
int glgo_obs_dim(const glgo_env_cfg *c);
/* obs for the current (x,u,timestep,time) -- observations.py:59-182 ; obs has glgo_obs_dim(c) entries */
void glgo_env_obs(const glgo_env_cfg *c, const glgo_env *e, double *obs);
/* tomato_env.py:115-146.  action: 6 float32 values (raw_control=0) or 6 doubles u (raw_control=1,
 * step_raw_control :148-173).  noise34 may be NULL (scale 0).  Outputs obs[23+5Np], reward, info[11].
 * returns terminated flag. */
int glgo_env_step(const glgo_env_cfg *c, glgo_env *e, const double *p_nom, const void *action, int raw_control,
                  const double *noise34, double *obs, double *reward, double *info);

/* multi-threaded rollout used for the CPU baseline: B envs, each stepped n_steps with the given actions
 * (float32 [n_steps][B][6]); envs auto-reset on termination. returns total env-steps executed. */
long glgo_rollout(const glgo_env_cfg *c, const double *p_nom, const double *weather, int rows, int B, int n_steps,
                  const float *actions, int n_threads, double *reward_sum_out);

/* persistent batch of B reference-semantics envs stepped by n_threads host threads (auto-reset on termination):
 * the CPU counterpart of the SubprocVecEnv-of-TomatoEnv stack, used as the measured CPU baseline. */
/* Rule-based controller: environments/baseline.py:68-227, settings configs/agents/rule_based.yml.
 * s[29] in the order of GLGO_CTRL_NAMES below; x = state before the step, d = weather row of the current timestep
 * (all 10 columns), hod / doy = env clock before the step (experiments/evaluate_baseline.py:21-23). */
#define GLGO_NCTRL 29
/* lamps_on, lamps_off, lamps_day_start, lamps_day_stop, lamps_off_sun, lamp_rad_sum_limit, temp_setpoint_day,
 * temp_setpoint_night, heat_correction, heat_deadzone, co2_day, vent_heat_Pband, rh_max, mech_dehumid_Pband,
 * vent_rh_Pband, t_vent_off, vent_cold_Pband, thScrSpDay, thScrSpNight, thScrPband, thScrDeadZone, thScrRh,
 * thScrRhPband, lampExtraHeat, blScrExtraRh, rhMax, tHeatBand, co2Band, useBlScr */
void glgo_rule_control(const double *s, const double *x, const double *d, double hod, double doy, double *u);
/* one env step with the controller in the loop: u = rule_control(x, weather[k], clock) ; step_raw_control(u) */
int glgo_env_step_rule(const glgo_env_cfg *c, glgo_env *e, const double *p_nom, const double *ctrl29, const double *noise34,
                       double *obs, double *reward, double *info);

typedef struct glgo_batch glgo_batch;
glgo_batch *glgo_batch_create(const glgo_env_cfg *c, const double *p_nom, const double *weather, int rows, int B);
/* actions float32 [B][6]; reward [B] doubles; done [B] bytes; obs_f32 may be NULL or float [B][23+5Np] */
void glgo_batch_step(glgo_batch *b, const float *actions, float *obs_f32, double *reward, unsigned char *done, int n_threads);
/* cumulative work of the batch: RK4 micro-steps (x4 = right-hand-side evaluations), or right-hand-side evaluations of the
 * implicit solver including those spent on finite-difference Jacobians */
long glgo_batch_work(const glgo_batch *b);
void glgo_batch_destroy(glgo_batch *b);

#ifdef __cplusplus
}
#endif
#endif
