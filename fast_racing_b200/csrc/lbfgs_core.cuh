// Scalar core of the L-BFGS line searches, usable from host code and from CUDA kernels.
//
// The reference drives its optimiser with lbfgs::lbfgs_optimize (lbfgs.hpp:1103), whose
// More-Thuente search (lbfgs.hpp:730-938, update_trial_interval :520-728) and backtracking
// fallback (:940-1033) call the evaluator from inside nested loops.  A batched device solver
// cannot call back, so the searches are re-cast here as a resumable state machine: everything
// between two evaluations is one transition, the O(n) vector work (dot products, x = xp + t*d,
// the two-loop recursion) is done by the caller (cooperatively by a CTA on the device, by plain
// loops on the host), and this file holds only the O(1) scalar logic.
#pragma once
#include <math.h>

#if defined(__CUDACC__)
#define FRX_HD __host__ __device__ __forceinline__
#else
#define FRX_HD inline
#endif

namespace frx {

// lbfgs.hpp:148-205 numeric values (the ones the solver can return)
enum : int {
    LBFGS_CONVERGENCE = 0,
    LBFGS_STOP = 1,
    LBFGS_ALREADY_MINIMIZED = 2,
    LBFGSERR_UNKNOWNERROR = -1024,
    LBFGSERR_LOGICERROR = -1023,
    LBFGSERR_INVALID_N = -1021,
    LBFGSERR_OUTOFINTERVAL = -1010,
    LBFGSERR_INCORRECT_TMINMAX = -1009,
    LBFGSERR_ROUNDING_ERROR = -1008,
    LBFGSERR_MINIMUMSTEP = -1007,
    LBFGSERR_MAXIMUMSTEP = -1006,
    LBFGSERR_MAXIMUMLINESEARCH = -1005,
    LBFGSERR_MAXIMUMITERATION = -1004,
    LBFGSERR_WIDTHTOOSMALL = -1003,
    LBFGSERR_INVALIDPARAMETERS = -1002,
    LBFGSERR_INCREASEGRADIENT = -1001,
};

struct LbfgsConfig {  // lbfgs.hpp:19-123; defaults :128-140
    int mem_size = 8;
    double g_epsilon = 1e-5;
    int past = 0;
    double delta = 1e-5;
    int max_iterations = 0;
    int max_linesearch = 40;
    double min_step = 1e-20;
    double max_step = 1e20;
    double f_dec_coeff = 1e-4;
    double s_curv_coeff = 0.9;
    double xtol = 1.0e-16;
};

struct Bracket {
    double st, f, dg;
};

// Interpolation formulas of lbfgs.hpp:322-405.
FRX_HD double interp_cubic(double u, double fu, double du, double v, double fv, double dv) {
    const double d = v - u;
    const double theta = (fu - fv) * 3 / d + du + dv;
    double p = fabs(theta), q = fabs(du), r = fabs(dv);
    double s = p >= q ? p : q;
    s = s >= r ? s : r;
    const double a = theta / s;
    double gamm = s * sqrt(a * a - (du / s) * (dv / s));
    if (v < u) gamm = -gamm;
    p = gamm - du + theta;
    q = gamm - du + gamm + dv;
    return u + (p / q) * d;
}
FRX_HD double interp_cubic_bounded(double u, double fu, double du, double v, double fv, double dv, double lo,
                                   double hi) {
    const double d = v - u;
    const double theta = (fu - fv) * 3 / d + du + dv;
    double p = fabs(theta), q = fabs(du), r = fabs(dv);
    double s = p >= q ? p : q;
    s = s >= r ? s : r;
    const double a = theta / s;
    double gamm = a * a - (du / s) * (dv / s);
    gamm = gamm > 0 ? s * sqrt(gamm) : 0;
    if (u < v) gamm = -gamm;
    p = gamm - dv + theta;
    q = gamm - dv + gamm + du;
    r = p / q;
    if (r < 0. && gamm != 0.) return v - r * d;
    return a < 0 ? hi : lo;
}
FRX_HD double interp_quad(double u, double fu, double du, double v, double fv) {
    const double a = v - u;
    return u + du / ((fu - fv) / a + du) / 2 * a;
}
FRX_HD double interp_secant(double u, double du, double v, double dv) {
    const double a = u - v;
    return v + dv / (dv - du) * a;
}

// lbfgs.hpp:520-728.  Returns 0 or an LBFGSERR code; t.st receives the next trial.
FRX_HD int refine_bracket(Bracket& x, Bracket& y, Bracket& t, double tmin, double tmax, int& brackt) {
    const int dsign = t.dg * (x.dg / fabs(x.dg)) < 0.;
    if (brackt) {
        const double lo = x.st <= y.st ? x.st : y.st;
        const double hi = x.st >= y.st ? x.st : y.st;
        if (t.st <= lo || hi <= t.st) return LBFGSERR_OUTOFINTERVAL;
        if (0. <= x.dg * (t.st - x.st)) return LBFGSERR_INCREASEGRADIENT;
        if (tmax < tmin) return LBFGSERR_INCORRECT_TMINMAX;
    }
    int bound;
    double nt;
    if (x.f < t.f) {
        brackt = 1, bound = 1;
        const double mc = interp_cubic(x.st, x.f, x.dg, t.st, t.f, t.dg);
        const double mq = interp_quad(x.st, x.f, x.dg, t.st, t.f);
        nt = fabs(mc - x.st) < fabs(mq - x.st) ? mc : mc + 0.5 * (mq - mc);
    } else if (dsign) {
        brackt = 1, bound = 0;
        const double mc = interp_cubic(x.st, x.f, x.dg, t.st, t.f, t.dg);
        const double mq = interp_secant(x.st, x.dg, t.st, t.dg);
        nt = fabs(mc - t.st) > fabs(mq - t.st) ? mc : mq;
    } else if (fabs(t.dg) < fabs(x.dg)) {
        bound = 1;
        const double mc = interp_cubic_bounded(x.st, x.f, x.dg, t.st, t.f, t.dg, tmin, tmax);
        const double mq = interp_secant(x.st, x.dg, t.st, t.dg);
        if (brackt)
            nt = fabs(t.st - mc) < fabs(t.st - mq) ? mc : mq;
        else
            nt = fabs(t.st - mc) > fabs(t.st - mq) ? mc : mq;
    } else {
        bound = 0;
        if (brackt)
            nt = interp_cubic(t.st, t.f, t.dg, y.st, y.f, y.dg);
        else
            nt = x.st < t.st ? tmax : tmin;
    }
    if (x.f < t.f) {
        y = t;
    } else {
        if (dsign) y = x;
        x = t;
    }
    if (tmax < nt) nt = tmax;
    if (nt < tmin) nt = tmin;
    if (brackt && bound) {
        const double mq = x.st + 0.66 * (y.st - x.st);
        if (x.st < y.st) {
            if (mq < nt) nt = mq;
        } else {
            if (nt < mq) nt = mq;
        }
    }
    t.st = nt;
    return 0;
}

// Outcome of feeding one evaluation to a line search.
enum : int { LS_CONTINUE = 0, LS_DONE = 1, LS_FAILED = 2 };

// State of one line search (either kind).  `step` is the trial step whose point the caller must
// evaluate next whenever a transition returns LS_CONTINUE.
struct LineSearch {
    double step, step_min, step_max;
    double finit, dginit, dgtest;
    double width, prev_width, stmin, stmax;
    Bracket bx, by;
    int count, brackt, stage1, uinfo;
    int code;  // >0: evaluations used on success; <0: LBFGSERR on failure

    // lbfgs.hpp:743-787.  Returns LS_CONTINUE (evaluate xp + step*d) or LS_FAILED.
    FRX_HD int mt_begin(double step0, double f0, double dg0, double smin, double smax, const LbfgsConfig& P) {
        step = step0, step_min = smin, step_max = smax;
        count = 0, uinfo = 0;
        if (step <= 0.) return fail(LBFGSERR_INVALIDPARAMETERS);
        dginit = dg0;
        if (0 < dginit) return fail(LBFGSERR_INCREASEGRADIENT);
        brackt = 0, stage1 = 1;
        finit = f0;
        dgtest = P.f_dec_coeff * dginit;
        width = step_max - step_min;
        prev_width = 2.0 * width;
        bx = {0., finit, dginit};
        by = bx;
        mt_choose(P);
        return LS_CONTINUE;
    }
    // lbfgs.hpp:788-818
    FRX_HD void mt_choose(const LbfgsConfig& P) {
        if (brackt) {
            stmin = bx.st <= by.st ? bx.st : by.st;
            stmax = bx.st >= by.st ? bx.st : by.st;
        } else {
            stmin = bx.st;
            stmax = step + 4.0 * (step - bx.st);
        }
        if (step < step_min) step = step_min;
        if (step_max < step) step = step_max;
        if ((brackt && ((step <= stmin || stmax <= step) || P.max_linesearch <= count + 1 || uinfo != 0)) ||
            (brackt && (stmax - stmin <= P.xtol * stmax)))
            step = bx.st;
    }
    // lbfgs.hpp:829-935: f, dg = value and directional derivative at xp + step*d.
    FRX_HD int mt_feed(double f, double dg, const LbfgsConfig& P) {
        const double ftest1 = finit + step * dgtest;
        ++count;
        if ((isinf(f) || isnan(f)) || (brackt && ((step <= stmin || stmax <= step) || uinfo != 0)))
            return fail(LBFGSERR_ROUNDING_ERROR);
        if (step == step_max && f <= ftest1 && dg <= dgtest) return fail(LBFGSERR_MAXIMUMSTEP);
        if (step == step_min && (ftest1 < f || dgtest <= dg)) return fail(LBFGSERR_MINIMUMSTEP);
        if (brackt && (stmax - stmin) <= P.xtol * stmax) return fail(LBFGSERR_WIDTHTOOSMALL);
        if (P.max_linesearch <= count) return fail(LBFGSERR_MAXIMUMLINESEARCH);
        if (f <= ftest1 && fabs(dg) <= P.s_curv_coeff * (-dginit)) {
            code = count;
            return LS_DONE;
        }
        const double mc = P.f_dec_coeff <= P.s_curv_coeff ? P.f_dec_coeff : P.s_curv_coeff;
        if (stage1 && f <= ftest1 && mc * dginit <= dg) stage1 = 0;
        Bracket t{step, f, dg};
        if (stage1 && ftest1 < f && f <= bx.f) {
            Bracket xm{bx.st, bx.f - bx.st * dgtest, bx.dg - dgtest};
            Bracket ym{by.st, by.f - by.st * dgtest, by.dg - dgtest};
            Bracket tm{step, f - step * dgtest, dg - dgtest};
            uinfo = refine_bracket(xm, ym, tm, stmin, stmax, brackt);
            bx = {xm.st, xm.f + xm.st * dgtest, xm.dg + dgtest};
            by = {ym.st, ym.f + ym.st * dgtest, ym.dg + dgtest};
            step = tm.st;
        } else {
            uinfo = refine_bracket(bx, by, t, stmin, stmax, brackt);
            step = t.st;
        }
        if (brackt) {
            if (0.66 * prev_width <= fabs(by.st - bx.st)) step = bx.st + 0.5 * (by.st - bx.st);
            prev_width = width;
            width = fabs(by.st - bx.st);
        }
        mt_choose(P);
        return LS_CONTINUE;
    }

    // lbfgs.hpp:954-978
    FRX_HD int bt_begin(double step0, double f0, double dg0, double smin, double smax, const LbfgsConfig& P) {
        step = step0, step_min = smin, step_max = smax;
        count = 0;
        if (step <= 0.) return fail(LBFGSERR_INVALIDPARAMETERS);
        dginit = dg0;
        if (0 < dginit) return fail(LBFGSERR_INCREASEGRADIENT);
        finit = f0;
        dgtest = P.f_dec_coeff * dginit;
        return LS_CONTINUE;
    }
    // lbfgs.hpp:984-1031.  dg is only consulted when the sufficient-decrease test passes (the
    // reference computes it lazily; passing it always is equivalent).
    FRX_HD int bt_feed(double f, double dg, const LbfgsConfig& P) {
        ++count;
        double w;
        if (f > finit + step * dgtest) {
            w = 0.5;
        } else if (dg < P.s_curv_coeff * dginit) {
            w = 2.1;
        } else if (dg > -P.s_curv_coeff * dginit) {
            w = 0.5;
        } else {
            code = count;
            return LS_DONE;
        }
        if (step < step_min) return fail(LBFGSERR_MINIMUMSTEP);
        if (step > step_max) return fail(LBFGSERR_MAXIMUMSTEP);
        if (P.max_linesearch <= count) return fail(LBFGSERR_MAXIMUMLINESEARCH);
        step *= w;
        return LS_CONTINUE;
    }
    FRX_HD int fail(int err) {
        code = err;
        return LS_FAILED;
    }
};

}  // namespace frx
