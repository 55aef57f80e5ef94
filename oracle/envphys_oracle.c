/* envphys_oracle.c -- CPU restatement (TEST INFRASTRUCTURE ONLY) of the dew-point and wet-bulb routines that consume the
 * distributed grids (wet bulb: smrf/envphys/core/iwbt.c:14-134, at the end of this file); dew point: smrf/envphys/core/dewpt.c:14-229 with the saturation curves of smrf/envphys/core/topotherm.c:109-176.
 * Restated, not copied: one state struct, iterative form; the arithmetic order of every expression follows the reference
 * so that the Brent iteration takes the same path.  Pinned against the reference C compiled in place
 * (oracle/_ref/libenvphys_ref.so, tests/test_oracle.py). */
#include <math.h>

#define K_FREEZE 2.7316e2      /* envphys.h:31 */
#define K_BOIL 3.7315e2        /* envphys.h:32 */
#define P_SEA 1.013246e5       /* envphys.h:74 */
#define EPS_D 2.2204460492503131e-16

static double sat_water(double tk) {   /* topotherm.c:109-137 */
    double l10 = log(1.e1);
    double x = -7.90298 * (K_BOIL / tk - 1.) + 5.02808 * log(K_BOIL / tk) / l10 -
               1.3816e-7 * (pow(1.e1, 1.1344e1 * (1. - tk / K_BOIL)) - 1.) +
               8.1328e-3 * (pow(1.e1, -3.49149 * (K_BOIL / tk - 1.)) - 1.) + log(P_SEA) / l10;
    return pow(1.e1, x);
}

static double sat_ice(double tk) {     /* topotherm.c:143-176 */
    double l10, x;
    if (tk > K_FREEZE) return sat_water(tk);
    l10 = log(1.e1);
    x = pow(1.e1, -9.09718 * ((K_FREEZE / tk) - 1.) - 3.56654 * log(K_FREEZE / tk) / l10 + 8.76793e-1 * (1. - (tk / K_FREEZE)) +
                      log(6.1071) / l10);
    return x * 1.e2;
}

typedef struct {
    double a, b, c, d, e, fa, fb, fc;
} brent_t;

/* dewpt.c:95-229: root of f(t) = ea - sat_ice(t) between the spanning guesses lo, hi */
static double brent_root(double lo, double hi, double t, double ea) {
    brent_t s = {lo, hi, 0., 0., 0., 0., 0., 0.};
    double tol, m, p, q, r, w;
    int budget;
    if (s.a == s.b) return s.a;
    w = (fabs(s.b) >= fabs(s.a)) ? fabs(s.b) : fabs(s.a);
    tol = 5.e-1 * t + 2 * EPS_D * w;
    w = log(fabs(s.b - s.a) / tol) / log(2.);
    budget = (int)(w * w + 1);
    s.fa = ea - sat_ice(s.a);
    s.fc = s.fb = ea - sat_ice(s.b);
    if (fabs(s.fb) <= tol) return s.b;
    if (fabs(s.fa) <= tol) return s.a;
    if (s.fb * s.fa > 0) return 0.;
    while (budget--) {
        if ((s.fb > 0 && s.fc > 0) || (s.fb <= 0 && s.fc <= 0)) {
            s.c = s.a;
            s.fc = s.fa;
            s.d = s.e = s.b - s.a;
        }
        if (fabs(s.fc) < fabs(s.fb)) {
            s.a = s.b; s.b = s.c; s.c = s.a;
            s.fa = s.fb; s.fb = s.fc; s.fc = s.fa;
        }
        tol = EPS_D * fabs(s.b) + t;
        m = (s.c - s.b) / 2;
        if (fabs(m) < tol || s.fb == 0) return s.b;
        if (fabs(s.e) < tol || fabs(s.fa) <= fabs(s.fb)) {
            s.d = s.e = m;                                    /* bisection forced */
        } else {
            w = s.fb / s.fa;
            if (s.a == s.c) {                                 /* secant step */
                p = 2 * m * w;
                q = 1 - w;
            } else {                                          /* inverse quadratic step */
                q = s.fa / s.fc;
                r = s.fb / s.fc;
                p = w * (2 * m * q * (q - r) - (s.b - s.a) * (r - 1));
                q -= 1;
                q *= (r - 1) * (w - 1);
            }
            if (p > 0) q = -q; else p = -p;
            w = s.e;
            s.e = s.d;
            if (2 * p < 3 * m * q - fabs(tol * q) && p < fabs(w * q / 2)) s.d = p / q;
            else s.d = s.e = m;
        }
        s.a = s.b;
        s.fa = s.fb;
        if (fabs(s.d) > tol) s.b += s.d;
        else if (m > 0) s.b += tol;
        else s.b -= tol;
        s.fb = ea - sat_ice(s.b);
    }
    return 0.;
}

/* dewpt.c:14-47, 63-92: per cell, spanning guesses, root, truncation to float, K -> C */
void oracle_dew_point(long n, const double* ea, double tol, double* dpt) {
    long i;
    for (i = 0; i < n; i++) {
        double lo = K_FREEZE, hi = K_FREEZE + 15;
        float v;
        if (ea[i] != ea[i]) { dpt[i] = ea[i]; continue; }
        while (ea[i] < sat_ice(lo)) lo *= .75;
        while (ea[i] > sat_ice(hi)) hi *= 1.25;
        v = (float)brent_root(lo, hi, tol, ea[i]);
        v -= K_FREEZE;
        dpt[i] = v;
    }
}

/* iwbt.c:14-81: Newton iteration on the psychrometer equation, K in, K out; *bad where the reference calls exit(-1) */
static double wet_bulb_k(double ta, double dpt, double press, double tol, int* bad) {
    const double rh2o = 461.5, epsw = 18.0153 / 28.9644, cp_air = 1.005e3;
    double xlh, xlhv, xlhf, ea, esat, psyc, step = 1.0, ti = ta, prev, dedt, pf, dpdt, mid = (ta + dpt) / 2.0;
    int it = 0;
    if (ta <= K_FREEZE) {
        xlhv = 2.5e6 - 2.95573e3 * (mid - K_FREEZE);
        xlhf = 3.336e5 + 1.6667e2 * (K_FREEZE - mid);
        xlh = xlhv + xlhf;
    } else if (dpt <= K_FREEZE) {
        xlhv = 2.5e6 - 2.95573e3 * (mid - K_FREEZE);
        xlhf = 3.336e5 + 1.6667e2 * (K_FREEZE - ((K_FREEZE + dpt) / 2.0));
        xlh = xlhv + (((K_FREEZE - dpt) / (ta - dpt)) * xlhf);
    } else {
        xlh = 2.5e6 - 2.95573e3 * (((ta + dpt) / 2) - K_FREEZE);
    }
    ea = sat_ice(dpt);
    esat = sat_ice(ta);
    psyc = epsw * (xlh / (cp_air * press));
    while (step > tol) {
        prev = ti;
        if (ti != ta) esat = sat_ice(ti);
        dedt = xlh * (esat / (rh2o * (ti * ti)));
        pf = (ti - ta) + (psyc * (esat - ea));
        dpdt = 1.0 + (psyc * dedt);
        ti = ti - (pf / dpdt);
        step = prev - ti;
        if (++it > 10) {
            *bad |= 1;
            break;
        }
    }
    return ti;
}

/* iwbt.c:84-134; returns the flags (1: no convergence, 2: below 0 K) where the reference exits */
int oracle_wet_bulb(long n, const double* ta, const double* td, const double* z, double tol, double* tw) {
    long i;
    int bad = 0;
    for (i = 0; i < n; i++) {
        double pa = P_SEA, tak = ta[i] + K_FREEZE, tdk = td[i] + K_FREEZE;
        if (z[i] != 0.0) pa = P_SEA * pow(2.88e2 / (2.88e2 + -6.5 * (z[i] / 1000.0)), 9.80665 * 28.9644 / (8.31432e3 * -6.5 * 1.e-3));
        if (tak < 0 || tdk < 0) {
            bad |= 2;
            tw[i] = NAN;
            continue;
        }
        tw[i] = wet_bulb_k(tak, tdk, pa, tol, &bad) - K_FREEZE;
    }
    return bad;
}
