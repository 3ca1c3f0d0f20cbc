// Classic-control dynamics, one env per thread, fp64 with every rounding explicit (no FMA contraction), so
// that states / observations / rewards carry the same bits as the numpy arithmetic gymnasium executes.
//
// Reference call sites: /root/reference/AsyncTools/AsyncPPO.py:53 (env.reset) and :76 (env.step) inside
// EnvVectorizer; the physics itself is gymnasium==1.1.1 classic_control (third-party):
//   CartPoleEnv.step  Euler, tau 0.02;  PendulumEnv.step  float32 torque entering fp64;  AcrobotEnv.step  RK4 "book";
//   MountainCarEnv.step  clipped velocity / position updates;  Continuous_MountainCarEnv.step  the same in float32 after the first step.
#pragma once
#include "common.cuh"
#include "pow_glibc.cuh"
#include "trig_glibc.cuh"

namespace prl {

__device__ __forceinline__ double dmul(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ double dadd(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ double dsub(double a, double b) { return __dsub_rn(a, b); }
__device__ __forceinline__ double ddiv(double a, double b) { return __ddiv_rn(a, b); }
using prl_trig::cos_glibc;
using prl_trig::sin_glibc;

struct CartPole {
    static constexpr int ID = PRL_ENV_CARTPOLE, S = 4, O = 4, A = 2, AS = 1, MAX_STEPS = 500;
    static constexpr bool CONT = false;
    using Action = int;
    __host__ __device__ static constexpr double reset_hi(int) { return 0.05; }   // reset: U(-0.05, 0.05)^4
    __host__ __device__ static constexpr double reset_lo(int) { return -0.05; }
    static constexpr bool RESET_F32 = false;
    static constexpr int RESET_DRAWS = S;   // state components drawn at reset (the rest start at 0)

    __device__ static __forceinline__ void obs(const double (&s)[S], float (&o)[O]) {
#pragma unroll
        for (int i = 0; i < 4; ++i) o[i] = (float)s[i];
    }
    // returns terminated; reward is 1.0 on every step including the terminating one
    __device__ static __forceinline__ bool step(double (&s)[S], int action, double &reward) {
        constexpr double gravity = 9.8, masspole = 0.1, total_mass = 0.1 + 1.0, length = 0.5;
        constexpr double polemass_length = 0.1 * 0.5, force_mag = 10.0, tau = 0.02;
        constexpr double theta_thr = 12 * 2 * 3.141592653589793 / 360, x_thr = 2.4;
        const double x = s[0], x_dot = s[1], theta = s[2], theta_dot = s[3];
        const double force = action == 1 ? force_mag : -force_mag;
        const double costheta = cos_glibc(theta), sintheta = sin_glibc(theta);
        const double temp = ddiv(dadd(force, dmul(dmul(polemass_length, dmul(theta_dot, theta_dot)), sintheta)), total_mass);
        const double thetaacc =
            ddiv(dsub(dmul(gravity, sintheta), dmul(costheta, temp)),
                 dmul(length, dsub(4.0 / 3.0, ddiv(dmul(masspole, dmul(costheta, costheta)), total_mass))));
        const double xacc = dsub(temp, ddiv(dmul(dmul(polemass_length, thetaacc), costheta), total_mass));
        s[0] = dadd(x, dmul(tau, x_dot));
        s[1] = dadd(x_dot, dmul(tau, xacc));
        s[2] = dadd(theta, dmul(tau, theta_dot));
        s[3] = dadd(theta_dot, dmul(tau, thetaacc));
        reward = 1.0;
        return (s[0] < -x_thr) || (s[0] > x_thr) || (s[2] < -theta_thr) || (s[2] > theta_thr);
    }
};

struct Pendulum {
    static constexpr int ID = PRL_ENV_PENDULUM, S = 2, O = 3, A = 1, AS = 1, MAX_STEPS = 200;
    static constexpr bool CONT = true;
    using Action = float;
    __host__ __device__ static constexpr double reset_hi(int i) { return i == 0 ? 3.141592653589793 : 1.0; }  // U(-[pi,1], [pi,1])
    __host__ __device__ static constexpr double reset_lo(int i) { return i == 0 ? -3.141592653589793 : -1.0; }
    static constexpr bool RESET_F32 = false;
    static constexpr int RESET_DRAWS = S;

    __device__ static __forceinline__ void obs(const double (&s)[S], float (&o)[O]) {
        o[0] = (float)cos_glibc(s[0]);
        o[1] = (float)sin_glibc(s[0]);
        o[2] = (float)s[1];
    }
    __device__ static __forceinline__ bool step(double (&s)[S], float action, double &reward) {
        constexpr double pi = 3.141592653589793, dt = 0.05;
        const double th = s[0], thdot = s[1];
        const float u = fminf(fmaxf(action, -2.0f), 2.0f);  // np.clip on float32 stays float32
        // angle_normalize: ((x + pi) % (2 pi)) - pi with numpy's python-style float modulo
        double mod = fmod(dadd(th, pi), 2 * pi);
        if (mod != 0.0) {
            if (mod < 0) mod = dadd(mod, 2 * pi);
        } else {
            mod = 0.0;
        }
        const double an = dsub(mod, pi);
        const float ucost = __fmul_rn(0.001f, powf2_glibc(u));  // python float * np.float32 -> float32
        const double costs = dadd(dadd(pow2_glibc(an), dmul(0.1, pow2_glibc(thdot))), (double)ucost);
        const float tq = __fmul_rn(3.0f, u);
        double newthdot = dadd(thdot, dmul(dadd(dmul(15.0, sin_glibc(th)), (double)tq), dt));
        newthdot = newthdot < -8.0 ? -8.0 : (newthdot > 8.0 ? 8.0 : newthdot);
        s[0] = dadd(th, dmul(newthdot, dt));
        s[1] = newthdot;
        reward = -costs;
        return false;
    }
};

struct Acrobot {
    static constexpr int ID = PRL_ENV_ACROBOT, S = 4, O = 6, A = 3, AS = 1, MAX_STEPS = 500;
    static constexpr bool CONT = false;
    using Action = int;
    __host__ __device__ static constexpr double reset_hi(int) { return 0.1; }   // reset: U(-0.1, 0.1)^4
    __host__ __device__ static constexpr double reset_lo(int) { return -0.1; }
    static constexpr bool RESET_F32 = true;  // gymnasium casts the drawn state to float32
    static constexpr int RESET_DRAWS = S;

    __device__ static __forceinline__ void obs(const double (&s)[S], float (&o)[O]) {
        o[0] = (float)cos_glibc(s[0]);
        o[1] = (float)sin_glibc(s[0]);
        o[2] = (float)cos_glibc(s[1]);
        o[3] = (float)sin_glibc(s[1]);
        o[4] = (float)s[2];
        o[5] = (float)s[3];
    }
    // derivative of [theta1, theta2, dtheta1, dtheta2] under torque a ("book" variant)
    __device__ static __forceinline__ void dsdt(const double (&y)[4], double a, double (&k)[4]) {
        constexpr double pi = 3.141592653589793;
        const double theta1 = y[0], theta2 = y[1], dtheta1 = y[2], dtheta2 = y[3];
        const double c2 = cos_glibc(theta2), s2 = sin_glibc(theta2);
        // m1*lc1^2 + m2*(l1^2 + lc2^2 + 2*l1*lc2*cos(theta2)) + I1 + I2, unit masses/lengths, lc = 0.5
        const double d1 = dadd(dadd(dadd(0.25, dadd(1.25, c2)), 1.0), 1.0);
        const double d2 = dadd(dadd(0.25, dmul(0.5, c2)), 1.0);
        const double phi2 = dmul(0.5 * 9.8, cos_glibc(dsub(dadd(theta1, theta2), pi / 2.0)));
        const double t1 = dmul(dmul(-0.5, pow2_glibc(dtheta2)), s2);
        const double t2 = dmul(dmul(dmul(1.0, dtheta2), dtheta1), s2);
        const double t3 = dmul(1.5 * 9.8, cos_glibc(dsub(theta1, pi / 2)));
        const double phi1 = dadd(dadd(dsub(t1, t2), t3), phi2);
        const double num = dsub(dsub(dadd(a, dmul(ddiv(d2, d1), phi1)), dmul(dmul(0.5, pow2_glibc(dtheta1)), s2)), phi2);
        const double den = dsub(1.25, ddiv(pow2_glibc(d2), d1));
        const double ddtheta2 = ddiv(num, den);
        const double ddtheta1 = ddiv(-dadd(dmul(d2, ddtheta2), phi1), d1);
        k[0] = dtheta1; k[1] = dtheta2; k[2] = ddtheta1; k[3] = ddtheta2;
    }
    __device__ static __forceinline__ bool step(double (&s)[S], int action, double &reward) {
        constexpr double pi = 3.141592653589793, dt = 0.2, dt2 = 0.2 / 2.0, dt6 = 0.2 / 6.0;
        const double a = (double)(action - 1);  // AVAIL_TORQUE = [-1, 0, +1]
        double k1[4], k2[4], k3[4], k4[4], y[4];
        dsdt(s, a, k1);
#pragma unroll
        for (int i = 0; i < 4; ++i) y[i] = dadd(s[i], dmul(dt2, k1[i]));
        dsdt(y, a, k2);
#pragma unroll
        for (int i = 0; i < 4; ++i) y[i] = dadd(s[i], dmul(dt2, k2[i]));
        dsdt(y, a, k3);
#pragma unroll
        for (int i = 0; i < 4; ++i) y[i] = dadd(s[i], dmul(dt, k3[i]));
        dsdt(y, a, k4);
        double ns[4];
#pragma unroll
        for (int i = 0; i < 4; ++i)
            ns[i] = dadd(s[i], dmul(dt6, dadd(dadd(dadd(k1[i], dmul(2.0, k2[i])), dmul(2.0, k3[i])), k4[i])));
        constexpr double diff = pi - (-pi);
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            while (ns[i] > pi) ns[i] = dsub(ns[i], diff);
            while (ns[i] < -pi) ns[i] = dadd(ns[i], diff);
        }
        constexpr double mv1 = 4 * pi, mv2 = 9 * pi;
        { const double v = (-mv1 > ns[2]) ? -mv1 : ns[2]; ns[2] = (mv1 < v) ? mv1 : v; }
        { const double v = (-mv2 > ns[3]) ? -mv2 : ns[3]; ns[3] = (mv2 < v) ? mv2 : v; }
#pragma unroll
        for (int i = 0; i < 4; ++i) s[i] = ns[i];
        const bool terminated = dsub(-cos_glibc(ns[0]), cos_glibc(dadd(ns[1], ns[0]))) > 1.0;
        reward = terminated ? 0.0 : -1.0;
        return terminated;
    }
};

// MountainCar-v0 (gymnasium/envs/classic_control/mountain_car.py): state (position, velocity) as python floats.
struct MountainCar {
    static constexpr int ID = PRL_ENV_MOUNTAINCAR, S = 2, O = 2, A = 3, AS = 1, MAX_STEPS = 200;
    static constexpr bool CONT = false;
    using Action = int;
    __host__ __device__ static constexpr double reset_hi(int) { return -0.4; }   // reset: position U(-0.6, -0.4), velocity 0
    __host__ __device__ static constexpr double reset_lo(int) { return -0.6; }
    static constexpr bool RESET_F32 = false;
    static constexpr int RESET_DRAWS = 1;

    __device__ static __forceinline__ void obs(const double (&s)[S], float (&o)[O]) {
        o[0] = (float)s[0];
        o[1] = (float)s[1];
    }
    __device__ static __forceinline__ bool step(double (&s)[S], int action, double &reward) {
        constexpr double min_position = -1.2, max_position = 0.6, max_speed = 0.07, goal_position = 0.5, goal_velocity = 0.0;
        constexpr double force = 0.001, gravity = 0.0025;
        double position = s[0], velocity = s[1];
        // velocity += (action - 1) * force + cos(3 * position) * (-gravity)
        velocity = dadd(velocity, dadd(dmul((double)(action - 1), force), dmul(cos_glibc(dmul(3.0, position)), -gravity)));
        velocity = velocity < -max_speed ? -max_speed : (velocity > max_speed ? max_speed : velocity);
        position = dadd(position, velocity);
        position = position < min_position ? min_position : (position > max_position ? max_position : position);
        if (position == min_position && velocity < 0) velocity = 0.0;
        s[0] = position;
        s[1] = velocity;
        reward = -1.0;
        return position >= goal_position && velocity >= goal_velocity;
    }
};

// MountainCarContinuous-v0 (gymnasium/envs/classic_control/continuous_mountain_car.py).  State {position, velocity, stepped}:
// reset() leaves a FLOAT64 state array, step() a FLOAT32 one, the action is a float32 scalar and the constants are python floats,
// so (NEP 50) the first step of an episode computes in double and every later one in float, with python-float intermediates
// rounded to float32 where numpy does it; a clamped force / velocity / position is a python float (a double) again.  Operation
// by operation the same as oracle/c/prl_oracle.c::mccont_step and oracle/envs.py::MountainCarContinuous.
struct MountainCarContinuous {
    static constexpr int ID = PRL_ENV_MOUNTAINCARCONT, S = 3, O = 2, A = 1, AS = 1, MAX_STEPS = 999;
    static constexpr bool CONT = true;
    using Action = float;
    __host__ __device__ static constexpr double reset_hi(int) { return -0.4; }   // reset: position U(-0.6, -0.4), velocity 0, stepped 0
    __host__ __device__ static constexpr double reset_lo(int) { return -0.6; }
    static constexpr bool RESET_F32 = false;
    static constexpr int RESET_DRAWS = 1;

    __device__ static __forceinline__ void obs(const double (&s)[S], float (&o)[O]) {
        o[0] = (float)s[0];
        o[1] = (float)s[1];
    }
    __device__ static __forceinline__ bool step(double (&s)[S], float a, double &reward) {
        const bool first = s[2] == 0.0;
        // force = min(max(action[0], -1.0), 1.0): the float32 scalar unless it lies outside, then the python-float bound
        const bool force_py = (-1.0f > a) || (1.0f < a);
        const double force_d = (-1.0f > a) ? -1.0 : 1.0;
        const double p3 = first ? dmul(3.0, s[0]) : (double)__fmul_rn(3.0f, (float)s[0]);
        const double t = dmul(0.0025, cos_glibc(p3));                                   // python float
        bool vel_py = false, pos_py = false;                                            // the value is a python float (just clamped)
        double vel, pos;
        if (first) {                                                                    // np.float64 arithmetic
            const double X = force_py ? dsub(dmul(force_d, 0.0015), t) : (double)__fsub_rn(__fmul_rn(a, 0.0015f), (float)t);
            vel = dadd(s[1], X);
            if (vel > 0.07) { vel = 0.07; vel_py = true; }
            if (vel < -0.07) { vel = -0.07; vel_py = true; }
            pos = dadd(s[0], vel);
            if (pos > 0.6) { pos = 0.6; pos_py = true; }
            if (pos < -1.2) { pos = -1.2; pos_py = true; }
        } else {                                                                        // np.float32 arithmetic
            const float X = force_py ? (float)dsub(dmul(force_d, 0.0015), t) : __fsub_rn(__fmul_rn(a, 0.0015f), (float)t);
            const float v = __fadd_rn((float)s[1], X);
            vel = (double)v;
            if (v > (float)0.07) { vel = 0.07; vel_py = true; }
            if (vel_py ? (vel < -0.07) : (v < (float)-0.07)) { vel = -0.07; vel_py = true; }
            const float q = __fadd_rn((float)s[0], vel_py ? (float)vel : v);
            pos = (double)q;
            if (q > (float)0.6) { pos = 0.6; pos_py = true; }
            if (pos_py ? (pos < -1.2) : (q < (float)-1.2)) { pos = -1.2; pos_py = true; }
        }
        const bool at_min = (pos_py || first) ? (pos == -1.2) : ((float)pos == (float)-1.2);
        if (at_min && vel < 0) vel = 0.0;
        const bool over = (pos_py || first) ? (pos >= 0.45) : ((float)pos >= (float)0.45);
        const bool terminated = over && vel >= 0.0;
        reward = dsub(terminated ? 100.0 : 0.0, dmul(pow2_glibc((double)a), 0.1));      // reward -= math.pow(action[0], 2) * 0.1
        s[0] = (double)(float)pos;                                                      // np.array([position, velocity], dtype=np.float32)
        s[1] = (double)(float)vel;
        s[2] = 1.0;
        return terminated;
    }
};

// dispatch a functor templated on the env type
template <typename F>
inline int dispatch_env(int env_id, F &&f) {
    switch (env_id) {
        case PRL_ENV_CARTPOLE: return f(CartPole{});
        case PRL_ENV_PENDULUM: return f(Pendulum{});
        case PRL_ENV_ACROBOT: return f(Acrobot{});
        case PRL_ENV_MOUNTAINCAR: return f(MountainCar{});
        case PRL_ENV_MOUNTAINCARCONT: return f(MountainCarContinuous{});
    }
    set_error("unknown env id %d", env_id);
    return PRL_ERR_INVALID;
}

}  // namespace prl
