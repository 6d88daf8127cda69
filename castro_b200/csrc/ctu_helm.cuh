// ctu_helm.cuh -- the Helmholtz ("stellar") equation of state behind the same eos_* interface as the gamma law.
//
// Restates Microphysics EOS/helmholtz/actual_eos.H (Timmes & Swesty 2000): radiation + ideal ions + the tabulated
// electron-positron Helmholtz free energy (biquintic Hermite interpolation on the 541 x 201 (rho Ye, T) table) + Coulomb
// corrections (Yakovlev & Shalybkov 1989), and the Newton iteration on T (or rho) for the (rho, e), (rho, p), (T, p) input modes
// (:1233-1296), in the reference's operation order.  One source for host and device: the host copy serves ctu_params_finalize
// (small_pres / small_ener = EOS(small_dens, small_temp)), the problem-hook shim and the CPU check against values the reference
// executable wrote; the device copy is what the kernels call.  Transcendentals (log10, pow, cbrt, log, tanh) come from the CUDA math
// library on the device, so device results agree with the reference's to rounding (1e-15), not bit for bit.
//
// Only what an eos_rep_t / eos_re_t state of the reference carries is computed (pressure, energy and their rho / T derivatives,
// cv, cp, gam1, cs, dpde, dpdr_e): the hot path never asks for entropy, enthalpy or the chemical potential.
#pragma once
#include <cmath>

#if defined(__CUDACC__)
#define CTU_HD __host__ __device__ __forceinline__
#else
#define CTU_HD inline
#endif
#if defined(__CUDA_ARCH__)
#define CTU_HELM_UNROLL _Pragma("unroll")
#else
#define CTU_HELM_UNROLL
#endif

namespace ctu {
namespace helm {

constexpr int IMAX = 541, JMAX = 201;                       // actual_eos_data.H:18-44
constexpr double TLO = 3.0, THI = 13.0, TSTP = (THI - TLO) / (double)(JMAX - 1), TSTPI = 1.0 / TSTP;
constexpr double DLO = -12.0, DHI = 15.0, DSTP = (DHI - DLO) / (double)(IMAX - 1), DSTPI = 1.0 / DSTP;
constexpr double H_PLANCK = 6.6260689633e-27, AVO = 6.0221417930e23, KERG = 1.380650424e-16, AMU = 1.66053878283e-24;

// the table and the arrays actual_eos_init derives from it (pointers into host OR device memory)
struct Table {
    const double* f;        // [JMAX][IMAX][9]: f, ft, ftt, fd, fdd, fdt, fddt, fdtt, fddtt
    const double* dpdf;     // [JMAX][IMAX][4]
    const double *d, *t;    // [IMAX], [JMAX]
    const double *dt_sav, *dt2_sav, *dti_sav, *dt2i_sav;    // [JMAX]
    const double *dd_sav, *dd2_sav, *ddi_sav, *dd2i_sav;    // [IMAX]
    int do_coulomb, input_is_constant;
    double ttol, dtol;
    double mintemp, maxtemp, mindens, maxdens;               // EOSData after eos_init (interfaces/eos.H:17-50)
};

struct State {
    double rho, T, e, p;
    double dpdT, dpdr, dedT, dedr, dpde, dpdr_e;
    double cv, cp, gam1, cs;
    double abar, zbar, y_e, mu_e, mu;
};

enum Input { IN_RT = 0, IN_RP = 1, IN_RE = 2, IN_TP = 3 };

// quintic / cubic Hermite basis functions (:50-135)
CTU_HD double psi0(double z) { return z * z * z * (z * (-6.0 * z + 15.0) - 10.0) + 1.0; }
CTU_HD double dpsi0(double z) { return z * z * (z * (-30.0 * z + 60.0) - 30.0); }
CTU_HD double ddpsi0(double z) { return z * (z * (-120.0 * z + 180.0) - 60.0); }
CTU_HD double psi1(double z) { return z * (z * z * (z * (-3.0 * z + 8.0) - 6.0) + 1.0); }
CTU_HD double dpsi1(double z) { return z * z * (z * (-15.0 * z + 32.0) - 18.0) + 1.0; }
CTU_HD double ddpsi1(double z) { return z * (z * (-60.0 * z + 96.0) - 36.0); }
CTU_HD double psi2(double z) { return 0.5 * z * z * (z * (z * (-z + 3.0) - 3.0) + 1.0); }
CTU_HD double dpsi2(double z) { return 0.5 * z * (z * (z * (-5.0 * z + 12.0) - 9.0) + 2.0); }
CTU_HD double ddpsi2(double z) { return 0.5 * (z * (z * (-20.0 * z + 36.0) - 18.0) + 2.0); }
CTU_HD double xpsi0(double z) { return z * z * (2.0 * z - 3.0) + 1.0; }
CTU_HD double xpsi1(double z) { return z * (z * (z - 2.0) + 1.0); }

CTU_HD void fwt(const double* fi, const double* wt, double* r) {
    r[0] = fi[0] * wt[0] + fi[1] * wt[1] + fi[2] * wt[2] + fi[18] * wt[3] + fi[19] * wt[4] + fi[20] * wt[5];
    r[1] = fi[3] * wt[0] + fi[5] * wt[1] + fi[7] * wt[2] + fi[21] * wt[3] + fi[23] * wt[4] + fi[25] * wt[5];
    r[2] = fi[4] * wt[0] + fi[6] * wt[1] + fi[8] * wt[2] + fi[22] * wt[3] + fi[24] * wt[4] + fi[26] * wt[5];
    r[3] = fi[9] * wt[0] + fi[10] * wt[1] + fi[11] * wt[2] + fi[27] * wt[3] + fi[28] * wt[4] + fi[29] * wt[5];
    r[4] = fi[12] * wt[0] + fi[14] * wt[1] + fi[16] * wt[2] + fi[30] * wt[3] + fi[32] * wt[4] + fi[34] * wt[5];
    r[5] = fi[13] * wt[0] + fi[15] * wt[1] + fi[17] * wt[2] + fi[31] * wt[3] + fi[33] * wt[4] + fi[35] * wt[5];
}

CTU_HD double hmax2(double a, double b) { return (a < b) ? b : a; }      // amrex::max
CTU_HD double hmin2(double a, double b) { return (b < a) ? b : a; }      // amrex::min
CTU_HD int iclamp(int v, int lo, int hi) { return v < lo ? lo : (hi < v ? hi : v); }
CTU_HD double dclamp(double v, double lo, double hi) { return v < lo ? lo : (hi < v ? hi : v); }

// apply_radiation (:531-612); sets instead of adding (it comes first).  The optional pressure limiter (eos.prad_limiter_*) is off by default.
CTU_HD void apply_radiation(State& s) {
    constexpr double clight = 2.99792458e10, ssol = 5.67051e-5, asol = 4.0 * ssol / clight, asoli3 = asol / 3.0;
    const double deni = 1.0 / s.rho, tempi = 1.0 / s.T;
    const double prad = asoli3 * s.T * s.T * s.T * s.T;
    const double dpraddd = 0.0;
    const double dpraddt = 4.0 * prad * tempi;
    const double erad = 3.0 * prad * deni;
    const double deraddd = -erad * deni;
    const double deraddt = 3.0 * dpraddt * deni;
    s.p = prad; s.dpdr = dpraddd; s.dpdT = dpraddt;
    s.e = erad; s.dedr = deraddd; s.dedT = deraddt;
}

// apply_ions (:450-527)
CTU_HD void apply_ions(State& s) {
    const double deni = 1.0 / s.rho;
    const double ytot1 = 1.0 / s.abar;
    const double xni = AVO * ytot1 * s.rho;
    const double dxnidd = AVO * ytot1;
    const double kt = KERG * s.T;
    const double pion = xni * kt;
    const double dpiondd = dxnidd * kt;
    const double dpiondt = xni * KERG;
    const double eion = 1.5 * pion * deni;
    const double deiondd = (1.5 * dpiondd - eion) * deni;
    const double deiondt = 1.5 * dpiondt * deni;
    s.p = s.p + pion; s.dpdT = s.dpdT + dpiondt; s.dpdr = s.dpdr + dpiondd;
    s.e = s.e + eion; s.dedT = s.dedT + deiondt; s.dedr = s.dedr + deiondd;
}

// apply_electrons (:153-446)
CTU_HD void apply_electrons(const Table& tb, State& s) {
    const double din = s.y_e * s.rho;
    int jat = (int)((log10(s.T) - TLO) * TSTPI) + 1;
    jat = iclamp(jat, 1, JMAX - 1) - 1;
    int iat = (int)((log10(din) - DLO) * DSTPI) + 1;
    iat = iclamp(iat, 1, IMAX - 1) - 1;
    double fi[36];
    {
        const double* f00 = tb.f + ((long)jat * IMAX + iat) * 9;
        const double* f10 = tb.f + ((long)(jat + 1) * IMAX + iat) * 9;
CTU_HELM_UNROLL
        for (int i = 0; i < 9; ++i) {
            fi[i] = f00[i];
            fi[i + 9] = f00[9 + i];
            fi[i + 18] = f10[i];
            fi[i + 27] = f10[9 + i];
        }
    }
    const double xt = hmax2((s.T - tb.t[jat]) * tb.dti_sav[jat], 0.0);
    const double xd = hmax2((din - tb.d[iat]) * tb.ddi_sav[iat], 0.0);
    const double mxt = 1.0 - xt, mxd = 1.0 - xd;
    const double dd_s = tb.dd_sav[iat], dd2_s = tb.dd2_sav[iat], ddi_s = tb.ddi_sav[iat];
    const double dt_s = tb.dt_sav[jat], dt2_s = tb.dt2_sav[jat], dti_s = tb.dti_sav[jat], dt2i_s = tb.dt2i_sav[jat];
    double sid[6], dsid[6], fwtr[6], sit[6];
    sid[0] = psi0(xd); sid[1] = psi1(xd) * dd_s; sid[2] = psi2(xd) * dd2_s;
    sid[3] = psi0(mxd); sid[4] = -psi1(mxd) * dd_s; sid[5] = psi2(mxd) * dd2_s;
    dsid[0] = dpsi0(xd) * ddi_s; dsid[1] = dpsi1(xd); dsid[2] = dpsi2(xd) * dd_s;
    dsid[3] = -dpsi0(mxd) * ddi_s; dsid[4] = dpsi1(mxd); dsid[5] = -dpsi2(mxd) * dd_s;
    sit[0] = psi0(xt); sit[1] = psi1(xt) * dt_s; sit[2] = psi2(xt) * dt2_s;
    sit[3] = psi0(mxt); sit[4] = -psi1(mxt) * dt_s; sit[5] = psi2(mxt) * dt2_s;
    fwt(fi, sit, fwtr);
    double free = 0.0, df_d = 0.0;
CTU_HELM_UNROLL
    for (int i = 0; i <= 5; ++i) { free = free + fwtr[i] * sid[i]; df_d = df_d + fwtr[i] * dsid[i]; }
    sit[0] = dpsi0(xt) * dti_s; sit[1] = dpsi1(xt); sit[2] = dpsi2(xt) * dt_s;
    sit[3] = -dpsi0(mxt) * dti_s; sit[4] = dpsi1(mxt); sit[5] = -dpsi2(mxt) * dt_s;
    fwt(fi, sit, fwtr);
    double df_t = 0.0, df_dt = 0.0;
CTU_HELM_UNROLL
    for (int i = 0; i <= 5; ++i) { df_t += fwtr[i] * sid[i]; df_dt += fwtr[i] * dsid[i]; }
    sit[0] = ddpsi0(xt) * dt2i_s; sit[1] = ddpsi1(xt) * dti_s; sit[2] = ddpsi2(xt);
    sit[3] = ddpsi0(mxt) * dt2i_s; sit[4] = -ddpsi1(mxt) * dti_s; sit[5] = ddpsi2(mxt);
    fwt(fi, sit, fwtr);
    double df_tt = 0.0;
CTU_HELM_UNROLL
    for (int i = 0; i <= 5; ++i) df_tt = df_tt + fwtr[i] * sid[i];
    // pressure derivative with density: bicubic interpolation of the dpdf table (:299-336)
    sit[0] = xpsi0(xt); sit[1] = xpsi1(xt) * dt_s; sit[2] = xpsi0(mxt); sit[3] = -xpsi1(mxt) * dt_s;
    sid[0] = xpsi0(xd); sid[1] = xpsi1(xd) * dd_s; sid[2] = xpsi0(mxd); sid[3] = -xpsi1(mxd) * dd_s;
    double wdt[16];
CTU_HELM_UNROLL
    for (int i = 0; i <= 3; ++i) { wdt[i] = sid[0] * sit[i]; wdt[i + 4] = sid[1] * sit[i]; wdt[i + 8] = sid[2] * sit[i]; wdt[i + 12] = sid[3] * sit[i]; }
    double dpepdd = 0.0;
    {
        // electron_table_indexing::map: k -> (dj, di, n)
        const int mdj[16] = {0, 0, 1, 1, 0, 0, 1, 1, 0, 0, 1, 1, 0, 0, 1, 1};
        const int mdi[16] = {0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 1, 1, 1, 1};
        const int mn[16] = {0, 1, 0, 1, 2, 3, 2, 3, 0, 1, 0, 1, 2, 3, 2, 3};
CTU_HELM_UNROLL
        for (int k = 0; k < 16; ++k) dpepdd += tb.dpdf[((long)(jat + mdj[k]) * IMAX + (iat + mdi[k])) * 4 + mn[k]] * wdt[k];
    }
    dpepdd = hmax2(s.y_e * dpepdd, 0.0);
    double x = din * din;
    const double pele = x * df_d;
    const double dpepdt = x * df_dt;
    x = s.y_e * s.y_e;
    const double sele = -df_t * s.y_e;
    const double dsepdt = -df_tt * s.y_e;
    const double dsepdd = -df_dt * x;
    const double eele = s.y_e * free + s.T * sele;
    const double deepdt = s.T * dsepdt;
    const double deepdd = x * df_d + s.T * dsepdd;
    s.p = s.p + pele; s.dpdT = s.dpdT + dpepdt; s.dpdr = s.dpdr + dpepdd;
    s.e = s.e + eele; s.dedT = s.dedT + deepdt; s.dedr = s.dedr + deepdd;
}

// apply_coulomb_corrections (:616-820).  HASP: the state carries a pressure (eos_rep_t) -- only then does a negative total pressure
// switch the corrections off (:771-779).
template <bool HASP>
CTU_HD void apply_coulomb(State& st) {
    constexpr double a1 = -0.898004, b1 = 0.96786, c1 = 0.220703, d1 = -0.86097;
    constexpr double a2 = 0.29561, b2 = 1.9885, c2 = 0.288675;
    constexpr double qe = 4.8032042712e-10, esqu = qe * qe, onethird = 1.0 / 3.0, forth = 4.0 / 3.0;
    constexpr double pi = 3.141592653589793238462643383279502884;
    double pcoul = 0.0, dpcouldd = 0.0, dpcouldt = 0.0, ecoul = 0.0, decouldd = 0.0, decouldt = 0.0;
    double s, z;
    const double ytot1 = 1.0 / st.abar;
    const double xni = AVO * ytot1 * st.rho;
    const double dxnidd = AVO * ytot1;
    const double kt = KERG * st.T;
    const double ktinv = 1.0 / kt;
    z = forth * pi;
    s = z * xni;
    const double dsdd = z * dxnidd;
    const double inv_lami = cbrt(s);
    const double lami = 1.0 / inv_lami;
    z = -onethird * lami;
    const double lamidd = z * dsdd / s;
    const double plasg = st.zbar * st.zbar * esqu * ktinv * inv_lami;
    z = -plasg * inv_lami;
    const double plasgdd = z * lamidd;
    const double plasgdt = -plasg * ktinv * KERG;
    if (plasg >= 1.0) {
        const double x = pow(plasg, 0.25);
        double y = AVO * ytot1 * KERG;
        ecoul = y * st.T * (a1 * plasg + b1 * x + c1 / x + d1);
        pcoul = onethird * st.rho * ecoul;
        y = AVO * ytot1 * kt * (a1 + 0.25 / plasg * (b1 * x - c1 / x));
        decouldd = y * plasgdd;
        decouldt = y * plasgdt + ecoul / st.T;
        y = onethird * st.rho;
        dpcouldd = onethird * ecoul + y * decouldd;
        dpcouldt = y * decouldt;
    } else if (plasg < 1.0) {
        const double pion = xni * kt;
        const double dpiondd = dxnidd * kt;
        const double dpiondt = xni * KERG;
        const double x = plasg * sqrt(plasg);
        const double y = pow(plasg, b2);
        z = c2 * x - onethird * a2 * y;
        pcoul = -pion * z;
        ecoul = 3.0 * pcoul / st.rho;
        s = 1.5 * c2 * x / plasg - onethird * a2 * b2 * y / plasg;
        dpcouldd = -dpiondd * z - pion * s * plasgdd;
        dpcouldt = -dpiondt * z - pion * s * plasgdt;
        s = 3.0 / st.rho;
        decouldd = s * dpcouldd - ecoul / st.rho;
        decouldt = s * dpcouldt;
    }
    const double p_temp = HASP ? st.p + pcoul : 1.7976931348623157e308;
    const double e_temp = st.e + ecoul;
    if (p_temp <= 0.0 || e_temp <= 0.0) {
        pcoul = 0.0; dpcouldd = 0.0; dpcouldt = 0.0; ecoul = 0.0; decouldd = 0.0; decouldt = 0.0;
    }
    st.p = st.p + pcoul; st.dpdr = st.dpdr + dpcouldd; st.dpdT = st.dpdT + dpcouldt;
    st.e = st.e + ecoul; st.dedr = st.dedr + decouldd; st.dedT = st.dedT + decouldt;
}

// actual_eos (:1233-1296): Newton iteration (at most 100) on T (rho for the (T, p) mode), one more evaluation after convergence,
// then finalize_state (:1135-1229).  composition (abar, zbar, y_e, mu_e) must be set in the state.
template <bool HASP>
CTU_HD void actual_eos(const Table& tb, const int input, State& s) {
    double v_want = 0.0;
    if (input == IN_RP || input == IN_TP) v_want = s.p;
    else if (input == IN_RE) v_want = s.e;
    bool converged = (input == IN_RT);
    for (int iter = 1; iter <= 100; ++iter) {
        apply_radiation(s);
        apply_ions(s);
        apply_electrons(tb, s);
        if (tb.do_coulomb) apply_coulomb<HASP>(s);
        if (converged) break;
        // single_iter_update (:919-1020)
        double x, v, dvdx, xtol, smallx;
        if (input == IN_TP) { x = s.rho; smallx = tb.mindens; xtol = tb.dtol; v = s.p; dvdx = s.dpdr; }
        else {
            x = s.T; smallx = tb.mintemp; xtol = tb.ttol;
            if (input == IN_RP) { v = s.p; dvdx = s.dpdT; } else { v = s.e; dvdx = s.dedT; }
        }
        double xnew = x - (v - v_want) / dvdx;
        xnew = dclamp(xnew, 0.5 * x, 2.0 * x);
        xnew = hmax2(smallx, xnew);
        if (input == IN_TP) s.rho = xnew; else s.T = xnew;
        const double error = fabs(xnew - x);
        if (error < xtol * x) converged = true;
    }
    s.mu = 1.0 / (1.0 / s.abar + 1.0 / s.mu_e);
    s.dpde = s.dpdT / s.dedT;
    s.dpdr_e = s.dpdr - s.dpdT * s.dedr / s.dedT;
    s.cv = s.dedT;
    {
        const double chit = s.T / s.p * s.dpdT;
        const double chid = s.dpdr * s.rho / s.p;
        s.gam1 = (chit * (s.p / s.rho)) * (chit / (s.T * s.cv)) + chid;
        s.cp = s.cv * s.gam1 / chid;
    }
    s.cs = sqrt(s.gam1 * s.p / s.rho);
    if (tb.input_is_constant) {
        if (input == IN_RP || input == IN_TP) s.p = v_want;
        else if (input == IN_RE) s.e = v_want;
    }
}

// eos() (interfaces/eos.H:395-431) after composition: reset_inputs (:141-206; an energy / pressure outside [1e-200, 1e200] falls back to
// the clamped (rho, T) pair), then actual_eos
template <bool HASP>
CTU_HD void eos(const Table& tb, int input, State& s) {
    if (input != IN_TP) s.rho = hmin2(tb.maxdens, hmax2(tb.mindens, s.rho));
    if (input == IN_RT || input == IN_TP) s.T = hmin2(tb.maxtemp, hmax2(tb.mintemp, s.T));
    const bool bad_e = (input == IN_RE) && (s.e < 1.e-200 || s.e > 1.e200);
    const bool bad_p = (input == IN_RP || input == IN_TP) && (s.p < 1.e-200 || s.p > 1.e200);
    if (bad_e || bad_p) {
        s.T = hmin2(tb.maxtemp, hmax2(tb.mintemp, s.T));
        s.rho = hmin2(tb.maxdens, hmax2(tb.mindens, s.rho));
        input = IN_RT;
    }
    actual_eos<HASP>(tb, input, s);
}

}  // namespace helm
}  // namespace ctu
