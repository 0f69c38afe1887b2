// State-space model functions (device).  Each restates one dyn_fcn / meas_fcn of the reference's
// ssmod.py for additive noise; the noise vector q / r is passed explicitly so the same code
// serves the filter (zero noise, TransitionModel.dyn_eval ssmod.py:129-166,
// MeasurementModel.meas_eval ssmod.py:960-1009) and the simulators (ssmod.py:168-244, 1011-1039).
#pragma once
#include <type_traits>

#include "ssm_common.cuh"

namespace ssm {

// Dyn::TIME_TERM (optional member): the dynamics have an additive / multiplicative term that depends on time only
template <class D, class = void>
struct HasTimeTerm {
    static constexpr bool value = false;
};
template <class D>
struct HasTimeTerm<D, std::void_t<decltype(D::TIME_TERM)>> {
    static constexpr bool value = D::TIME_TERM;
};

// ---- UNGMTransition.dyn_fcn, ssmod.py:268-269 -------------------------------------------------
struct DynUngm {
    static constexpr int DX = 1, DQ = 1, ID = SSM_DYN_UNGM;
    static constexpr bool ADDITIVE = true;
    static constexpr bool HAS_CONT = false;
    // The forcing term depends on the time step only.  It is the same for every sigma point and -- unless trajectories
    // carry their own time offsets -- for every trajectory, so the forward pass reads it from a per-step table filled by
    // one tiny launch (time_tab_kernel) instead of evaluating an fp64 cosine per trajectory-step (a fifth of the
    // instructions of a UNGM UKF step).  The product is rounded on its own: table and in-line evaluation agree bit for bit.
    static constexpr bool TIME_TERM = true;
    SSM_DEV static double time_term(double t) { return __dmul_rn(8.0, cos(1.2 * t)); }
    template <bool NOISE>
    SSM_DEV static void f_tt(const double *, const double (&x)[1], const double (&q)[1], double tt, double (&o)[1]) {
        o[0] = 0.5 * x[0] + 25.0 * (x[0] / (1.0 + x[0] * x[0])) + tt;
        if (NOISE) o[0] += q[0];
    }
    template <bool NOISE>
    SSM_DEV static void f(const double *par, const double (&x)[1], const double (&q)[1], double t, double (&o)[1]) {
        f_tt<NOISE>(par, x, q, time_term(t), o);
    }
    SSM_DEV static void fc(const double *, const double (&)[1], const double (&)[1], double, double (&o)[1]) { o[0] = 0.0; }
};

// ---- UNGMNATransition.dyn_fcn, ssmod.py:299-300: NON-additive noise 8 q cos(1.2 k) ------------------
// The filter feeds it through the augmented state [x; q] (ssinf.py:271-272); f<NOISE = false> is the model with q = 0.
struct DynUngmNA {
    static constexpr int DX = 1, DQ = 1, ID = SSM_DYN_UNGMNA;
    static constexpr bool ADDITIVE = false;
    static constexpr bool HAS_CONT = false;
    static constexpr bool TIME_TERM = true;   // see DynUngm
    SSM_DEV static double time_term(double t) { return cos(1.2 * t); }
    template <bool NOISE>
    SSM_DEV static void f_tt(const double *, const double (&x)[1], const double (&q)[1], double tt, double (&o)[1]) {
        o[0] = 0.5 * x[0] + 25.0 * (x[0] / (1.0 + x[0] * x[0])) + 8.0 * (NOISE ? q[0] : 0.0) * tt;
    }
    template <bool NOISE>
    SSM_DEV static void f(const double *par, const double (&x)[1], const double (&q)[1], double t, double (&o)[1]) {
        f_tt<NOISE>(par, x, q, time_term(t), o);
    }
    SSM_DEV static void fc(const double *, const double (&)[1], const double (&)[1], double, double (&o)[1]) { o[0] = 0.0; }
};

// ---- Pendulum2DTransition.dyn_fcn, ssmod.py:357-358; par[0] = dt, g = 9.81 (ssmod.py:351) ------
struct DynPendulum {
    static constexpr int DX = 2, DQ = 2, ID = SSM_DYN_PENDULUM;
    static constexpr bool ADDITIVE = true;
    static constexpr bool HAS_CONT = false;
    template <bool NOISE>
    SSM_DEV static void f(const double *par, const double (&x)[2], const double (&q)[2], double, double (&o)[2]) {
        const double dt = par[0];
        o[0] = x[0] + x[1] * dt;
        if (NOISE) o[0] += q[0];
        o[1] = x[1] - 9.81 * dt * sin(x[0]);
        if (NOISE) o[1] += q[1];
    }
    SSM_DEV static void fc(const double *, const double (&)[2], const double (&)[2], double, double (&o)[2]) { o[0] = o[1] = 0.0; }
};

// ---- ReentryVehicle2DTransition, ssmod.py:521-584; par[0] = dt ---------------------------------
// constants ssmod.py:523-526; the 3-dimensional noise enters components 2..4 (G = [0; I3], :527)
struct DynReentry {
    static constexpr int DX = 5, DQ = 3, ID = SSM_DYN_REENTRY;
    static constexpr bool ADDITIVE = true;
    static constexpr bool HAS_CONT = true;
    SSM_DEV static void forces(const double (&x)[5], double &D, double &G) {
        const double R0 = 6374.0, H0 = 13.406, Gm0 = 3.9860e5, b0 = -0.59783;
#ifdef SSM_DUP_EXP
        const double b = b0 * SSM_DUP(m_exp, m_exp(x[4]), x[4] * 1.0000001);
#else
        const double b = b0 * m_exp(x[4]);
#endif
        // R = sqrt(r2) and 1/R^3 from ONE out-of-line call: ir = 1/sqrt(r2), R = r2 ir, R^-3 = ir^3 (a few ulp
        // from the reference's sqrt + pow + divide, far below the parity tolerance; r2 = 0 gives NaN/inf in both)
        const double r2 = x[0] * x[0] + x[1] * x[1];
#ifdef SSM_DUP_RSQRT
        const double ir = SSM_DUP(m_rsqrt, m_rsqrt(r2), r2 * 1.0000001);
#else
        const double ir = m_rsqrt(r2);
#endif
        const double R = r2 * ir;
        const double V = m_sqrt(x[2] * x[2] + x[3] * x[3]);
        D = b * m_exp((R0 - R) * (1.0 / H0)) * V;  // (R0 - R) / H0 up to 1 ulp: no division call
        G = -Gm0 * (ir * ir * ir);
    }
    template <bool NOISE>
    SSM_DEV static void f(const double *par, const double (&x)[5], const double (&q)[3], double, double (&o)[5]) {
        const double dt = par[0];
        double D, G;
        forces(x, D, G);
        o[0] = x[0] + dt * x[2];
        o[1] = x[1] + dt * x[3];
        o[2] = x[2] + dt * (D * x[2] + G * x[0]);
        if (NOISE) o[2] += q[0];
        o[3] = x[3] + dt * (D * x[3] + G * x[1]);
        if (NOISE) o[3] += q[1];
        o[4] = x[4];
        if (NOISE) o[4] += q[2];
    }
    // dyn_fcn_cont, ssmod.py:569-584
    SSM_DEV static void fc(const double *, const double (&x)[5], const double (&q)[3], double, double (&o)[5]) {
        double D, G;
        forces(x, D, G);
        o[0] = x[2];
        o[1] = x[3];
        o[2] = D * x[2] + G * x[0] + q[0];
        o[3] = D * x[3] + G * x[1] + q[1];
        o[4] = q[2];
    }
};

// ---- CoordinatedTurnTransition.dyn_fcn, ssmod.py:675-690; par[0] = dt --------------------------
// No omega == 0 guard: sin(0)/0 = NaN, and because the reference multiplies the
// full 5x5 matrix, the NaN entries c, d contaminate rows 0 and 2 only (0 * NaN terms do not occur
// in rows 1, 3, 4: their c/d coefficients are structural zeros of mdyn but numpy still multiplies
// 0 * x, which is finite).  Rows 0 and 2 become NaN, exactly as below.
struct DynCoordTurn {
    static constexpr int DX = 5, DQ = 5, ID = SSM_DYN_COORDTURN;
    static constexpr bool ADDITIVE = true;
    static constexpr bool HAS_CONT = false;
    template <bool NOISE>
    SSM_DEV static void f(const double *par, const double (&x)[5], const double (&q)[5], double, double (&o)[5]) {
        const double dt = par[0];
        const double om = x[4];
        const Pair2 sc = m_sincos(om * dt);
        const double a = sc.u, b = sc.v;
        const Pair2 cd = m_div2(a, 1.0 - b, om);
        const double c = cd.u, d = cd.v;
        o[0] = x[0] + c * x[1] - d * x[3];
        if (NOISE) o[0] += q[0];
        o[1] = b * x[1] - a * x[3];
        if (NOISE) o[1] += q[1];
        o[2] = d * x[1] + x[2] + c * x[3];
        if (NOISE) o[2] += q[2];
        o[3] = a * x[1] + b * x[3];
        if (NOISE) o[3] += q[3];
        o[4] = x[4];
        if (NOISE) o[4] += q[4];
    }
    SSM_DEV static void fc(const double *, const double (&)[5], const double (&)[5], double, double (&o)[5]) {
#pragma unroll
        for (int i = 0; i < 5; ++i) o[i] = 0.0;
    }
};

// ---- ReentryVehicle1DTransition, ssmod.py:418-426; par[0] = dt, Gamma = 1 / 6.096 (ssmod.py:416) ----
// state [altitude, velocity, ballistic coefficient], additive 3-D noise
struct DynReentry1D {
    static constexpr int DX = 3, DQ = 3, ID = SSM_DYN_REENTRY1D;
    static constexpr bool ADDITIVE = true;
    static constexpr bool HAS_CONT = true;
    SSM_DEV static double drag(const double (&x)[3]) { return m_exp(-(1.0 / 6.096) * x[0]) * (x[1] * x[1]) * x[2]; }
    template <bool NOISE>
    SSM_DEV static void f(const double *par, const double (&x)[3], const double (&q)[3], double, double (&o)[3]) {
        const double dt = par[0];
        o[0] = x[0] - dt * x[1];
        if (NOISE) o[0] += q[0];
        o[1] = x[1] - dt * m_exp(-(1.0 / 6.096) * x[0]) * (x[1] * x[1]) * x[2];
        if (NOISE) o[1] += q[1];
        o[2] = x[2];
        if (NOISE) o[2] += q[2];
    }
    // dyn_fcn_cont, ssmod.py:423-426
    SSM_DEV static void fc(const double *, const double (&x)[3], const double (&q)[3], double, double (&o)[3]) {
        o[0] = -x[1] + q[0];
        o[1] = -drag(x) + q[1];
        o[2] = q[2];
    }
};

// ---- ConstantVelocity.dyn_fcn, ssmod.py:831-846; par[0] = dt -------------------------------------
// state [x, vx, y, vy]; the 2-D noise enters through the gain [[dt^2/2, 0], [dt, 0], [0, dt^2/2], [0, dt]] (:833-836)
struct DynConstVel {
    static constexpr int DX = 4, DQ = 2, ID = SSM_DYN_CONSTVEL;
    static constexpr bool ADDITIVE = true;
    static constexpr bool HAS_CONT = false;
    template <bool NOISE>
    SSM_DEV static void f(const double *par, const double (&x)[4], const double (&q)[2], double, double (&o)[4]) {
        const double dt = par[0], h = 0.5 * dt * dt;
        o[0] = x[0] + dt * x[1];
        o[1] = x[1];
        o[2] = x[2] + dt * x[3];
        o[3] = x[3];
        if (NOISE) {
            o[0] += h * q[0];
            o[1] += dt * q[0];
            o[2] += h * q[1];
            o[3] += dt * q[1];
        }
    }
    SSM_DEV static void fc(const double *, const double (&)[4], const double (&)[2], double, double (&o)[4]) {
#pragma unroll
        for (int i = 0; i < 4; ++i) o[i] = 0.0;
    }
};

// ---- ConstantTurnRateSpeed.dyn_fcn, ssmod.py:755-774; par[0] = dt: NON-additive noise ----------------
// state [x, y, speed, heading, yaw rate], noise [acceleration, yaw acceleration].  Restated as written, including
// the heading increment dt * x[3] (not dt * x[4]) of both branches and the noise-free position of the x[4] == 0 branch.
struct DynCtrs {
    static constexpr int DX = 5, DQ = 2, ID = SSM_DYN_CTRS;
    static constexpr bool ADDITIVE = false;
    static constexpr bool HAS_CONT = false;
    template <bool NOISE>
    SSM_DEV static void f(const double *par, const double (&x)[5], const double (&q)[2], double, double (&o)[5]) {
        const double dt = par[0], h = 0.5 * dt * dt;
        const double q0 = NOISE ? q[0] : 0.0, q1 = NOISE ? q[1] : 0.0;
        const Pair2 sc3 = m_sincos(x[3]);
        const double s3 = sc3.u, c3 = sc3.v;
        double f0, f1;
        if (x[4] == 0.0) {
            f0 = dt * x[2] * c3;
            f1 = dt * x[2] * s3;
        } else {
            const double c = m_div(x[2], x[4]);
            const Pair2 sc34 = m_sincos(x[3] + x[4] * dt);
            const double s34 = sc34.u, c34 = sc34.v;
            f0 = c * (s34 - s3) + h * c3 * q0;
            f1 = c * (-c34 + c3) + h * s3 * q0;
        }
        o[0] = x[0] + f0;
        o[1] = x[1] + f1;
        o[2] = x[2] + dt * q0;
        o[3] = x[3] + (dt * x[3] + h * q1);
        o[4] = x[4] + dt * q1;
    }
    SSM_DEV static void fc(const double *, const double (&)[5], const double (&)[2], double, double (&o)[5]) {
#pragma unroll
        for (int i = 0; i < 5; ++i) o[i] = 0.0;
    }
};

// ---- UNGMMeasurement.meas_fcn, ssmod.py:1060-1061 ----------------------------------------------
template <int DXS, int I0>
struct ObsUngm {
    static constexpr int DX = DXS, DY = 1, ID = SSM_OBS_UNGM;
    static constexpr bool ADDITIVE = true;
    template <bool NOISE>
    SSM_DEV static void h(const double *, const double (&x)[DXS], const double (&r)[1], double, double (&o)[1]) {
        o[0] = 0.05 * x[I0] * x[I0];
        if (NOISE) o[0] += r[0];
    }
};

// ---- UNGMNAMeasurement.meas_fcn, ssmod.py:1085-1086: NON-additive noise, z = 0.05 r x^2 --------------
template <int DXS, int I0>
struct ObsUngmNA {
    static constexpr int DX = DXS, DY = 1, ID = SSM_OBS_UNGMNA;
    static constexpr bool ADDITIVE = false;
    template <bool NOISE>
    SSM_DEV static void h(const double *, const double (&x)[DXS], const double (&r)[1], double, double (&o)[1]) {
        o[0] = 0.05 * (NOISE ? r[0] : 0.0) * (x[I0] * x[I0]);
    }
};

// ---- Pendulum2DMeasurement.meas_fcn, ssmod.py:1114-1115 ----------------------------------------
template <int DXS, int I0>
struct ObsPendulum {
    static constexpr int DX = DXS, DY = 1, ID = SSM_OBS_PENDULUM;
    static constexpr bool ADDITIVE = true;
    template <bool NOISE>
    SSM_DEV static void h(const double *, const double (&x)[DXS], const double (&r)[1], double, double (&o)[1]) {
        o[0] = sin(x[I0]);
        if (NOISE) o[0] += r[0];
    }
};

// ---- RangeMeasurement.meas_fcn, ssmod.py:1146-1148; par[0..1] = sensor position (sx, sy) ---------
template <int DXS, int I0>
struct ObsRange {
    static constexpr int DX = DXS, DY = 1, ID = SSM_OBS_RANGE;
    static constexpr bool ADDITIVE = true;
    template <bool NOISE>
    SSM_DEV static void h(const double *par, const double (&x)[DXS], const double (&r)[1], double, double (&o)[1]) {
        const double ey = x[I0] - par[1];
        o[0] = m_sqrt(par[0] * par[0] + ey * ey);
        if (NOISE) o[0] += r[0];
    }
};

// ---- Radar2DMeasurement.meas_fcn, ssmod.py:1227-1252; par[0..1] = radar_loc --------------------
// I0, I1 = state_index (compile-time so that sigma points sharing both coordinates share the
// sqrt / atan2 through common-subexpression elimination)
template <int DXS, int I0, int I1>
struct ObsRadar {
    static constexpr int DX = DXS, DY = 2, ID = SSM_OBS_RADAR;
    static constexpr bool ADDITIVE = true;
    template <bool NOISE>
    SSM_DEV static void h(const double *par, const double (&x)[DXS], const double (&r)[2], double, double (&o)[2]) {
        const double ex = x[I0] - par[0], ey = x[I1] - par[1];
#ifdef SSM_DUP_SQRT
        o[0] = SSM_DUP(m_sqrt, m_sqrt(ex * ex + ey * ey), ex * ex + ey * ey * 1.0000001);
#else
        o[0] = m_sqrt(ex * ex + ey * ey);
#endif
        if (NOISE) o[0] += r[0];
#ifdef SSM_DUP_ATAN2
        o[1] = SSM_DUP(m_atan2, m_atan2(ey, ex), ey * 1.0000001, ex);
#else
        o[1] = m_atan2(ey, ex);
#endif
        if (NOISE) o[1] += r[1];
    }
};

// ---- BearingMeasurement.meas_fcn, ssmod.py:1189-1195; 4 sensors, par[2 i], par[2 i + 1] = position of sensor i ------
template <int DXS, int I0, int I1>
struct ObsBearing4 {
    static constexpr int DX = DXS, DY = 4, ID = SSM_OBS_BEARING;
    static constexpr bool ADDITIVE = true;
    template <bool NOISE>
    SSM_DEV static void h(const double *par, const double (&x)[DXS], const double (&r)[4], double, double (&o)[4]) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            o[i] = m_atan2(x[I1] - par[2 * i + 1], x[I0] - par[2 * i]);
            if (NOISE) o[i] += r[i];
        }
    }
};

}  // namespace ssm
