/*
 * ORACLE (test infrastructure, not product code).
 *
 * Plain-C restatement of the CPU algorithms on the reference's data-parallel PPO hot path, used as
 * the checker for the CUDA kernels and as the fast leg of the CPU baseline.  It is a second,
 * independent statement of what oracle/envs.py (numpy) and the reference's Python do:
 *
 *   env physics      gymnasium==1.1.1 classic_control (third-party, absent -> PARITY UNPINNED vs
 *                    gymnasium; pinned bit-for-bit against oracle/envs.py by tests/test_oracle.py)
 *   orc_rollout      /root/reference/AsyncTools/AsyncPPO.py:117-146 (AsyncPPO.worker) with a taped
 *                    action source, + utils.buffer_append (utils.py:17-36) and
 *                    utils.buffer_to_target_buffer_transfer (utils.py:45-50): env-major flat buffer
 *   orc_gae          /root/reference/PPO/PPO.py:107-120 (PPO.compute_gae), float32 end to end
 *   orc_adv_norm     /root/reference/PPO/PPO.py:198-199
 *
 * Elementary functions are glibc libm's sin/cos/pow/powf/fmod - exactly what numpy calls for
 * float64 scalars on this image.  Build with -O2 -ffp-contract=off (no FMA contraction).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs may link or call this.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

enum { ORC_CARTPOLE = 0, ORC_PENDULUM = 1, ORC_ACROBOT = 2, ORC_MOUNTAINCAR = 3, ORC_MOUNTAINCARCONT = 4 };

int orc_env_dims(int env, int *S, int *O, int *A, int *cont, int *max_steps) {
    switch (env) {
    case ORC_CARTPOLE: *S = 4; *O = 4; *A = 2; *cont = 0; *max_steps = 500; return 0;
    case ORC_PENDULUM: *S = 2; *O = 3; *A = 1; *cont = 1; *max_steps = 200; return 0;
    case ORC_ACROBOT:  *S = 4; *O = 6; *A = 3; *cont = 0; *max_steps = 500; return 0;
    case ORC_MOUNTAINCAR: *S = 2; *O = 2; *A = 3; *cont = 0; *max_steps = 200; return 0;
    case ORC_MOUNTAINCARCONT: *S = 3; *O = 2; *A = 1; *cont = 1; *max_steps = 999; return 0;
    }
    return -1;
}

/* numpy float64 scalar `%` (npy_divmod): python-style modulo built on fmod */
static double np_mod(double a, double b) {
    double mod = fmod(a, b);
    if (mod != 0.0) {
        if ((b < 0) != (mod < 0)) mod += b;
    } else {
        mod = copysign(0.0, b);
    }
    return mod;
}

/* ---------------------------------------------------------------- CartPole-v1 */
static void cartpole_obs(const double *s, float *o) {
    for (int i = 0; i < 4; ++i) o[i] = (float)s[i];
}

static int cartpole_step(double *s, int action, float *obs, double *reward) {
    const double gravity = 9.8, masscart = 1.0, masspole = 0.1, length = 0.5, force_mag = 10.0, tau = 0.02;
    const double total_mass = masspole + masscart;
    const double polemass_length = masspole * length;
    const double theta_thr = 12 * 2 * 3.141592653589793 / 360;
    const double x_thr = 2.4;
    double x = s[0], x_dot = s[1], theta = s[2], theta_dot = s[3];
    double force = (action == 1) ? force_mag : -force_mag;
    double costheta = cos(theta), sintheta = sin(theta);
    double temp = (force + polemass_length * (theta_dot * theta_dot) * sintheta) / total_mass;
    double thetaacc = (gravity * sintheta - costheta * temp) /
                      (length * (4.0 / 3.0 - masspole * (costheta * costheta) / total_mass));
    double xacc = temp - polemass_length * thetaacc * costheta / total_mass;
    x = x + tau * x_dot;
    x_dot = x_dot + tau * xacc;
    theta = theta + tau * theta_dot;
    theta_dot = theta_dot + tau * thetaacc;
    s[0] = x; s[1] = x_dot; s[2] = theta; s[3] = theta_dot;
    cartpole_obs(s, obs);
    *reward = 1.0;
    return (x < -x_thr) || (x > x_thr) || (theta < -theta_thr) || (theta > theta_thr);
}

/* ---------------------------------------------------------------- Pendulum-v1 */
static void pendulum_obs(const double *s, float *o) {
    o[0] = (float)cos(s[0]); o[1] = (float)sin(s[0]); o[2] = (float)s[1];
}

static int pendulum_step(double *s, const float *action, float *obs, double *reward) {
    const double g = 10.0, m = 1.0, l = 1.0, dt = 0.05;
    const double pi = 3.141592653589793;
    double th = s[0], thdot = s[1];
    float u = action[0];
    u = fminf(fmaxf(u, -2.0f), 2.0f);                 /* np.clip on a float32 array stays float32 */
    double an = np_mod(th + pi, 2 * pi) - pi;
    float ucost = 0.001f * powf(u, 2.0f);             /* python float * np.float32 -> float32 */
    double costs = pow(an, 2.0) + 0.1 * pow(thdot, 2.0) + (double)ucost;
    float tq = (float)(3.0 / (m * (l * l))) * u;      /* python float * np.float32 -> float32 */
    double newthdot = thdot + (3 * g / (2 * l) * sin(th) + (double)tq) * dt;
    if (newthdot < -8.0) newthdot = -8.0;             /* np.clip(newthdot, -8, 8) */
    if (newthdot > 8.0) newthdot = 8.0;
    double newth = th + newthdot * dt;
    s[0] = newth; s[1] = newthdot;
    pendulum_obs(s, obs);
    *reward = -costs;
    return 0;
}

/* ---------------------------------------------------------------- Acrobot-v1 */
static void acrobot_obs(const double *s, float *o) {
    o[0] = (float)cos(s[0]); o[1] = (float)sin(s[0]);
    o[2] = (float)cos(s[1]); o[3] = (float)sin(s[1]);
    o[4] = (float)s[2]; o[5] = (float)s[3];
}

static void acrobot_dsdt(const double *sa, double *out) {
    const double pi = 3.141592653589793;
    const double m1 = 1.0, m2 = 1.0, l1 = 1.0, lc1 = 0.5, lc2 = 0.5, I1 = 1.0, I2 = 1.0, g = 9.8;
    double a = sa[4];
    double theta1 = sa[0], theta2 = sa[1], dtheta1 = sa[2], dtheta2 = sa[3];
    double c2 = cos(theta2), s2 = sin(theta2);
    double d1 = m1 * (lc1 * lc1) + m2 * ((l1 * l1) + (lc2 * lc2) + 2 * l1 * lc2 * c2) + I1 + I2;
    double d2 = m2 * ((lc2 * lc2) + l1 * lc2 * c2) + I2;
    double phi2 = m2 * lc2 * g * cos(theta1 + theta2 - pi / 2.0);
    double phi1 = -m2 * l1 * lc2 * pow(dtheta2, 2.0) * s2
                  - 2 * m2 * l1 * lc2 * dtheta2 * dtheta1 * s2
                  + (m1 * lc1 + m2 * l1) * g * cos(theta1 - pi / 2)
                  + phi2;
    double ddtheta2 = (a + d2 / d1 * phi1 - m2 * l1 * lc2 * pow(dtheta1, 2.0) * s2 - phi2) /
                      (m2 * (lc2 * lc2) + I2 - pow(d2, 2.0) / d1);
    double ddtheta1 = -(d2 * ddtheta2 + phi1) / d1;
    out[0] = dtheta1; out[1] = dtheta2; out[2] = ddtheta1; out[3] = ddtheta2; out[4] = 0.0;
}

static int acrobot_step(double *s, int action, float *obs, double *reward) {
    const double pi = 3.141592653589793;
    static const double torque[3] = {-1.0, 0.0, 1.0};
    const double dt = 0.2 - 0, dt2 = dt / 2.0;
    double y0[5] = {s[0], s[1], s[2], s[3], torque[action]};
    double k1[5], k2[5], k3[5], k4[5], y[5];
    acrobot_dsdt(y0, k1);
    for (int i = 0; i < 5; ++i) y[i] = y0[i] + dt2 * k1[i];
    acrobot_dsdt(y, k2);
    for (int i = 0; i < 5; ++i) y[i] = y0[i] + dt2 * k2[i];
    acrobot_dsdt(y, k3);
    for (int i = 0; i < 5; ++i) y[i] = y0[i] + dt * k3[i];
    acrobot_dsdt(y, k4);
    double ns[4];
    for (int i = 0; i < 4; ++i) ns[i] = y0[i] + dt / 6.0 * (k1[i] + 2 * k2[i] + 2 * k3[i] + k4[i]);
    const double diff = pi - (-pi);
    for (int i = 0; i < 2; ++i) {
        while (ns[i] > pi) ns[i] = ns[i] - diff;
        while (ns[i] < -pi) ns[i] = ns[i] + diff;
    }
    const double mv1 = 4 * pi, mv2 = 9 * pi;
    /* python: min(max(x, m), M) */
    { double v = (-mv1 > ns[2]) ? -mv1 : ns[2]; ns[2] = (mv1 < v) ? mv1 : v; }
    { double v = (-mv2 > ns[3]) ? -mv2 : ns[3]; ns[3] = (mv2 < v) ? mv2 : v; }
    memcpy(s, ns, sizeof ns);
    int terminated = (-cos(ns[0]) - cos(ns[1] + ns[0])) > 1.0;
    acrobot_obs(s, obs);
    *reward = terminated ? 0.0 : -1.0;
    return terminated;
}

/* ---------------------------------------------------------------- MountainCar-v0 */
static void mountaincar_obs(const double *s, float *o) {
    o[0] = (float)s[0];
    o[1] = (float)s[1];
}

static int mountaincar_step(double *s, int action, float *obs, double *reward) {
    const double min_position = -1.2, max_position = 0.6, max_speed = 0.07, goal_position = 0.5, goal_velocity = 0.0;
    const double force = 0.001, gravity = 0.0025;
    double position = s[0], velocity = s[1];
    velocity += (double)(action - 1) * force + cos(3 * position) * (-gravity);
    velocity = velocity < -max_speed ? -max_speed : (velocity > max_speed ? max_speed : velocity);
    position += velocity;
    position = position < min_position ? min_position : (position > max_position ? max_position : position);
    if (position == min_position && velocity < 0) velocity = 0;
    s[0] = position;
    s[1] = velocity;
    mountaincar_obs(s, obs);
    *reward = -1.0;
    return position >= goal_position && velocity >= goal_velocity;
}

/* ---------------------------------------------------------------- MountainCarContinuous-v0 */
/* gymnasium's continuous_mountain_car.py with numpy's typing spelled out (oracle/envs.py has the story): state = {position,
 * velocity, stepped}.  stepped == 0: the state is the float64 array reset() made, the step computes in double; afterwards the state
 * is a float32 array and the step computes in float, with python-float intermediates rounded to float32 where numpy would
 * (NEP 50).  A clamped force / velocity / position is a python float (a double) again. */
static void mccont_obs(const double *s, float *o) {
    o[0] = (float)s[0];
    o[1] = (float)s[1];
}

static int mccont_step(double *s, const float *action, float *obs, double *reward) {
    const int first = s[2] == 0.0;
    const float a = action[0];
    /* force = min(max(action[0], -1.0), 1.0): the float32 scalar itself unless it is outside, then the python float bound */
    int force_py = 0;
    double force_d = 0.0;
    float force_f = a;
    if (-1.0f > a) { force_py = 1; force_d = -1.0; }
    else if (1.0f < a) { force_py = 1; force_d = 1.0; }
    /* 3 * position: float64 on the first step, float32 afterwards; math.cos works on the double value of either */
    const double p3 = first ? 3.0 * s[0] : (double)(3.0f * (float)s[0]);
    const double t = 0.0025 * cos(p3);                       /* python float */
    int vel_py = 0, pos_py = 0;                               /* the value is a python float (just clamped) */
    double vel, pos;
    if (first) {                                              /* np.float64 arithmetic */
        const double X = force_py ? force_d * 0.0015 - t : (double)((float)(force_f * 0.0015f) - (float)t);
        vel = s[1] + X;
        if (vel > 0.07) { vel = 0.07; vel_py = 1; }
        if (vel < -0.07) { vel = -0.07; vel_py = 1; }
        pos = s[0] + vel;
        if (pos > 0.6) { pos = 0.6; pos_py = 1; }
        if (pos < -1.2) { pos = -1.2; pos_py = 1; }
    } else {                                                  /* np.float32 arithmetic */
        const float X = force_py ? (float)(force_d * 0.0015 - t) : (float)(force_f * 0.0015f) - (float)t;
        float v = (float)s[1] + X;
        vel = (double)v;
        if (v > (float)0.07) { vel = 0.07; vel_py = 1; }
        if ((vel_py ? (vel < -0.07) : (v < (float)-0.07))) { vel = -0.07; vel_py = 1; }
        float q = (float)s[0] + (vel_py ? (float)vel : v);
        pos = (double)q;
        if (q > (float)0.6) { pos = 0.6; pos_py = 1; }
        if ((pos_py ? (pos < -1.2) : (q < (float)-1.2))) { pos = -1.2; pos_py = 1; }
    }
    /* position == min_position: a python float -1.2 equals itself; a float32 / float64 scalar is compared in its own precision */
    const int at_min = pos_py ? (pos == -1.2) : (first ? (pos == -1.2) : ((float)pos == (float)-1.2));
    if (at_min && vel < 0) vel = 0.0;
    const int over = pos_py ? (pos >= 0.45) : (first ? (pos >= 0.45) : ((float)pos >= (float)0.45));
    const int terminated = over && vel >= 0.0;
    double r = terminated ? 100.0 : 0.0;
    r -= pow((double)a, 2.0) * 0.1;
    s[0] = (double)(float)pos;                                /* self.state = np.array([position, velocity], dtype=np.float32) */
    s[1] = (double)(float)vel;
    s[2] = 1.0;
    mccont_obs(s, obs);
    *reward = r;
    return terminated;
}

/* ---------------------------------------------------------------- dispatch */
void orc_env_obs(int env, const double *state, float *obs) {
    if (env == ORC_CARTPOLE) cartpole_obs(state, obs);
    else if (env == ORC_PENDULUM) pendulum_obs(state, obs);
    else if (env == ORC_MOUNTAINCAR) mountaincar_obs(state, obs);
    else if (env == ORC_MOUNTAINCARCONT) mccont_obs(state, obs);
    else acrobot_obs(state, obs);
}

/* one env, one step.  `action` points at an int32 (discrete) or float[A] (continuous). */
int orc_env_step(int env, double *state, const void *action, float *obs, double *reward) {
    if (env == ORC_CARTPOLE) return cartpole_step(state, *(const int32_t *)action, obs, reward);
    if (env == ORC_PENDULUM) return pendulum_step(state, (const float *)action, obs, reward);
    if (env == ORC_MOUNTAINCAR) return mountaincar_step(state, *(const int32_t *)action, obs, reward);
    if (env == ORC_MOUNTAINCARCONT) return mccont_step(state, (const float *)action, obs, reward);
    return acrobot_step(state, *(const int32_t *)action, obs, reward);
}

/*
 * Teacher-forced AsyncPPO.worker(): every env runs ONE episode from init_state, driven by the
 * taped action of (t, env) - tape[t*E + e] (int32) or tape[(t*E + e)*A .. ] (float32).  An episode
 * ends on terminated | truncated (TimeLimit: elapsed >= max_steps).  Output is the env-major,
 * time-minor flat buffer that utils.buffer_to_target_buffer_transfer hands to PPO.memory: the
 * stored state is the PRE-step observation, actions are stored as float32, dones = done|trunc.
 * Returns N = sum of episode lengths (the caller sizes outputs for E*max_steps).
 */
int64_t orc_rollout(int env, int E, int max_steps, const double *init_state, const void *tape,
                    float *flat_states, float *flat_actions, float *flat_rewards, float *flat_dones,
                    int32_t *lengths, double *final_state, double *reward_sum) {
    int S, O, A, cont, dflt;
    if (orc_env_dims(env, &S, &O, &A, &cont, &dflt)) return -1;
    const int AS = cont ? A : 1;
    int64_t n = 0;
    double rsum = 0.0;
    for (int e = 0; e < E; ++e) {
        double st[4];
        float obs[8], nobs[8];
        memcpy(st, init_state + (size_t)e * S, S * sizeof(double));
        orc_env_obs(env, st, obs);
        int t = 0;
        for (;;) {
            double r;
            int term;
            if (cont) {
                const float *a = (const float *)tape + ((size_t)t * E + e) * A;
                term = orc_env_step(env, st, a, nobs, &r);
                for (int i = 0; i < A; ++i) flat_actions[n * AS + i] = a[i];
            } else {
                int32_t a = ((const int32_t *)tape)[(size_t)t * E + e];
                term = orc_env_step(env, st, &a, nobs, &r);
                flat_actions[n] = (float)a;
            }
            ++t;
            int trunc = t >= max_steps;
            memcpy(flat_states + n * O, obs, O * sizeof(float));
            flat_rewards[n] = (float)r;
            flat_dones[n] = (term | trunc) ? 1.0f : 0.0f;
            rsum += r;
            ++n;
            memcpy(obs, nobs, O * sizeof(float));
            if (term | trunc) break;
        }
        lengths[e] = t;
        if (final_state) memcpy(final_state + (size_t)e * S, st, S * sizeof(double));
    }
    if (reward_sum) *reward_sum = rsum;
    return n;
}

/* PPO.compute_gae (PPO.py:107-120): float32 throughout, python-float hyper-parameters rounded to
 * float32 once (NEP 50), left-to-right evaluation order of lines 113-114. */
void orc_gae(const float *rewards, const float *dones, const float *values, float next_value,
             double gamma, double gae_lambda, int64_t N, float *returns) {
    const float g = (float)gamma;
    const float gl = (float)(gamma * gae_lambda);
    float gae = 0.0f;
    for (int64_t t = N - 1; t >= 0; --t) {
        float nd = 1.0f - dones[t];
        float delta = rewards[t] + g * next_value * nd - values[t];
        gae = delta + gl * nd * gae;
        returns[t] = gae + values[t];
        next_value = values[t];
    }
}

/* PPO.py:198-199: adv = ret - v; (adv - mean) / (std_unbiased + 1e-8).  Statistics in double so
 * that the oracle is summation-order independent; the tolerance for this step is 1e-5 relative. */
void orc_adv_norm(const float *returns, const float *values, int64_t N, float *adv, double *mean_out,
                  double *std_out) {
    double s = 0.0;
    for (int64_t i = 0; i < N; ++i) s += (double)(float)(returns[i] - values[i]);
    double mean = s / (double)N, q = 0.0;
    for (int64_t i = 0; i < N; ++i) {
        double d = (double)(float)(returns[i] - values[i]) - mean;
        q += d * d;
    }
    double sd = sqrt(q / (double)(N - 1));
    for (int64_t i = 0; i < N; ++i)
        adv[i] = (float)(((double)(float)(returns[i] - values[i]) - mean) / (sd + 1e-8));
    if (mean_out) *mean_out = mean;
    if (std_out) *std_out = sd;
}

/* libm pass-throughs: what numpy's float64 / float32 SCALAR `x ** 2` evaluates (npy_pow / npy_powf).  The exponent is
 * passed at run time and the file is built with -fno-builtin so the call cannot be folded into x*x. */
void orc_pow_scalar(const double *x, double y, double *out, const float *xf, float yf, float *outf, int64_t n) {
    for (int64_t i = 0; i < n; ++i) {
        if (x) out[i] = pow(x[i], y);
        if (xf) outf[i] = powf(xf[i], yf);
    }
}
