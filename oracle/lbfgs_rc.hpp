// ORACLE — TEST INFRASTRUCTURE ONLY. Nothing under oracle/ is part of the product path; only
// tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may use it.
//
// Restatement of the reference's L-BFGS with More-Thuente line search and backtracking fallback
// (reference: src/plan_manage/include/se3gcopter/lbfgs.hpp:1103-1444 driver, :730-938 More-Thuente,
// :520-728 trial-interval update, :940-1033 backtracking, :422-487 vector helpers), written as a
// REVERSE-COMMUNICATION state machine: the caller owns the evaluator and feeds (f, g) back in.
// The arithmetic (operation order of every dot/axpy, every branch of the line search) follows the
// reference so that, driven by the same evaluator, the iterates are bit-identical to the verbatim
// lbfgs.hpp (checked by tests/test_oracle_lbfgs.py through oracle/_ref when /root/reference exists).
//
// Parity status: pinned against the verbatim reference lbfgs.hpp (which compiles stand-alone).
#pragma once
#include <cmath>
#include <cstring>
#include <vector>

namespace frx_oracle {

// Same numeric values as the enum at lbfgs.hpp:148-205.
enum LbfgsStatus : int {
    LB_CONVERGENCE = 0,
    LB_STOP = 1,
    LB_ALREADY_MINIMIZED = 2,
    LBERR_UNKNOWNERROR = -1024,
    LBERR_LOGICERROR,
    LBERR_CANCELED,
    LBERR_INVALID_N,
    LBERR_INVALID_MEMSIZE,
    LBERR_INVALID_GEPSILON,
    LBERR_INVALID_TESTPERIOD,
    LBERR_INVALID_DELTA,
    LBERR_INVALID_MINSTEP,
    LBERR_INVALID_MAXSTEP,
    LBERR_INVALID_FDECCOEFF,
    LBERR_INVALID_SCURVCOEFF,
    LBERR_INVALID_XTOL,
    LBERR_INVALID_MAXLINESEARCH,
    LBERR_OUTOFINTERVAL,
    LBERR_INCORRECT_TMINMAX,
    LBERR_ROUNDING_ERROR,
    LBERR_MINIMUMSTEP,
    LBERR_MAXIMUMSTEP,
    LBERR_MAXIMUMLINESEARCH,
    LBERR_MAXIMUMITERATION,
    LBERR_WIDTHTOOSMALL,
    LBERR_INVALIDPARAMETERS,
    LBERR_INCREASEGRADIENT,
};

// lbfgs.hpp:19-123 (fields) and :128-140 (defaults).
struct LbfgsParams {
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

// Sequential dot product, lbfgs.hpp:467-475.
inline double seq_dot(const double* a, const double* b, int n) {
    double s = 0.;
    for (int i = 0; i < n; ++i) s += a[i] * b[i];
    return s;
}

// Minimiser of the cubic through (u,fu,du),(v,fv,dv): lbfgs.hpp:322-338.
inline double cubic_min(double u, double fu, double du, double v, double fv, double dv) {
    const double d = v - u;
    const double theta = (fu - fv) * 3 / d + du + dv;
    double p = std::fabs(theta), q = std::fabs(du), r = std::fabs(dv);
    double s = p >= q ? p : q;
    s = s >= r ? s : r;
    const double a = theta / s;
    double gamm = s * std::sqrt(a * a - (du / s) * (dv / s));
    if (v < u) gamm = -gamm;
    p = gamm - du + theta;
    q = gamm - du + gamm + dv;
    r = p / q;
    return u + r * d;
}

// Safeguarded cubic minimiser: lbfgs.hpp:352-380.
inline double cubic_min2(double u, double fu, double du, double v, double fv, double dv,
                         double xmin, double xmax) {
    const double d = v - u;
    const double theta = (fu - fv) * 3 / d + du + dv;
    double p = std::fabs(theta), q = std::fabs(du), r = std::fabs(dv);
    double s = p >= q ? p : q;
    s = s >= r ? s : r;
    const double a = theta / s;
    double gamm = a * a - (du / s) * (dv / s);
    gamm = gamm > 0 ? s * std::sqrt(gamm) : 0;
    if (u < v) gamm = -gamm;
    p = gamm - dv + theta;
    q = gamm - dv + gamm + du;
    r = p / q;
    if (r < 0. && gamm != 0.) return v - r * d;
    if (a < 0) return xmax;
    return xmin;
}

// lbfgs.hpp:391-393 and :403-405.
inline double quad_min(double u, double fu, double du, double v, double fv) {
    const double a = v - u;
    return u + du / ((fu - fv) / a + du) / 2 * a;
}
inline double quad_min2(double u, double du, double v, double dv) {
    const double a = u - v;
    return v + dv / (dv - du) * a;
}

// One bracket endpoint of the More-Thuente search.
struct MtPoint { double st, f, dg; };

// Trial-interval update, lbfgs.hpp:520-728.  x = best point, y = other endpoint, t = trial.
inline int mt_update_interval(MtPoint& x, MtPoint& y, MtPoint& t, double tmin, double tmax, int& brackt) {
    int bound;
    const int dsign = t.dg * (x.dg / std::fabs(x.dg)) < 0.;
    double mc, mq, newt;
    if (brackt) {
        if (t.st <= (x.st <= y.st ? x.st : y.st) || (x.st >= y.st ? x.st : y.st) <= t.st)
            return LBERR_OUTOFINTERVAL;
        if (0. <= x.dg * (t.st - x.st)) return LBERR_INCREASEGRADIENT;
        if (tmax < tmin) return LBERR_INCORRECT_TMINMAX;
    }
    if (x.f < t.f) {  // case 1: higher value, minimum bracketed
        brackt = 1;
        bound = 1;
        mc = cubic_min(x.st, x.f, x.dg, t.st, t.f, t.dg);
        mq = quad_min(x.st, x.f, x.dg, t.st, t.f);
        newt = std::fabs(mc - x.st) < std::fabs(mq - x.st) ? mc : mc + 0.5 * (mq - mc);
    } else if (dsign) {  // case 2: lower value, derivative sign change
        brackt = 1;
        bound = 0;
        mc = cubic_min(x.st, x.f, x.dg, t.st, t.f, t.dg);
        mq = quad_min2(x.st, x.dg, t.st, t.dg);
        newt = std::fabs(mc - t.st) > std::fabs(mq - t.st) ? mc : mq;
    } else if (std::fabs(t.dg) < std::fabs(x.dg)) {  // case 3: derivative magnitude decreases
        bound = 1;
        mc = cubic_min2(x.st, x.f, x.dg, t.st, t.f, t.dg, tmin, tmax);
        mq = quad_min2(x.st, x.dg, t.st, t.dg);
        if (brackt)
            newt = std::fabs(t.st - mc) < std::fabs(t.st - mq) ? mc : mq;
        else
            newt = std::fabs(t.st - mc) > std::fabs(t.st - mq) ? mc : mq;
    } else {  // case 4
        bound = 0;
        if (brackt)
            newt = cubic_min(t.st, t.f, t.dg, y.st, y.f, y.dg);
        else if (x.st < t.st)
            newt = tmax;
        else
            newt = tmin;
    }
    if (x.f < t.f) {
        y = t;
    } else {
        if (dsign) y = x;
        x = t;
    }
    if (tmax < newt) newt = tmax;
    if (newt < tmin) newt = tmin;
    if (brackt && bound) {
        mq = x.st + 0.66 * (y.st - x.st);
        if (x.st < y.st) {
            if (mq < newt) newt = mq;
        } else {
            if (newt < mq) newt = mq;
        }
    }
    t.st = newt;
    return 0;
}

// Reverse-communication solver.  Usage:
//   LbfgsRC s; s.begin(n, x0, params);           // s.x() is the point to evaluate
//   while (s.wants_eval()) { f = eval(s.x(), s.g()); s.feed(f); }
//   s.status(), s.fx(), s.x() (final point)
class LbfgsRC {
public:
    int begin(int n_, const double* x0, const LbfgsParams& p) {
        n = n_;
        par = p;
        ret = LBERR_UNKNOWNERROR;
        phase = PH_DONE;
        if (n <= 0) return ret = LBERR_INVALID_N;
        if (par.mem_size <= 0) return ret = LBERR_INVALID_MEMSIZE;
        if (par.g_epsilon < 0.) return ret = LBERR_INVALID_GEPSILON;
        if (par.past < 0) return ret = LBERR_INVALID_TESTPERIOD;
        if (par.delta < 0.) return ret = LBERR_INVALID_DELTA;
        if (par.min_step < 0.) return ret = LBERR_INVALID_MINSTEP;
        if (par.max_step < par.min_step) return ret = LBERR_INVALID_MAXSTEP;
        if (par.f_dec_coeff < 0.) return ret = LBERR_INVALID_FDECCOEFF;
        if (par.s_curv_coeff <= par.f_dec_coeff || 1. <= par.s_curv_coeff) return ret = LBERR_INVALID_SCURVCOEFF;
        if (par.xtol < 0.) return ret = LBERR_INVALID_XTOL;
        if (par.max_linesearch <= 0) return ret = LBERR_INVALID_MAXLINESEARCH;
        m = par.mem_size;
        xv.assign(x0, x0 + n);
        gv.assign(n, 0.);
        xp.assign(n, 0.);
        gp.assign(n, 0.);
        d.assign(n, 0.);
        S.assign((size_t)m * n, 0.);
        Y.assign((size_t)m * n, 0.);
        alpha.assign(m, 0.);
        ysv.assign(m, 0.);
        pf.assign(par.past > 0 ? par.past : 0, 0.);
        fx_ = 0.;
        n_evals = 0;
        k = 0;
        phase = PH_INIT;
        return 0;
    }
    bool wants_eval() const { return phase != PH_DONE; }
    double* x() { return xv.data(); }
    double* g() { return gv.data(); }
    double fx() const { return fx_; }
    int status() const { return ret; }
    int iterations() const { return k; }
    int evaluations() const { return n_evals; }

    // Feed the objective value at x() (gradient already written into g()).
    void feed(double f) {
        ++n_evals;
        switch (phase) {
        case PH_INIT: on_initial(f); break;
        case PH_MT: on_mt_eval(f); break;
        case PH_BT: on_bt_eval(f); break;
        default: break;
        }
    }

private:
    enum Phase { PH_INIT, PH_MT, PH_BT, PH_DONE };

    // lbfgs.hpp:1211-1247
    void on_initial(double f) {
        fx_ = f;
        if (!pf.empty()) pf[0] = fx_;
        for (int i = 0; i < n; ++i) d[i] = -gv[i];
        double xnorm = std::sqrt(seq_dot(xv.data(), xv.data(), n));
        double gnorm = std::sqrt(seq_dot(gv.data(), gv.data(), n));
        if (xnorm < 1.0) xnorm = 1.0;
        if (gnorm / xnorm <= par.g_epsilon) {
            ret = LB_ALREADY_MINIMIZED;
            phase = PH_DONE;
            return;
        }
        step = 1.0 / std::sqrt(seq_dot(d.data(), d.data(), n));
        k = 1;
        end = 0;
        start_iteration();
    }

    // lbfgs.hpp:1248-1268: save (xp, gp, stepp, fp) and enter the More-Thuente search.
    void start_iteration() {
        step_min = par.min_step;
        step_max = par.max_step;
        xp = xv;
        gp = gv;
        stepp = step;
        fp = fx_;
        int r = mt_begin();
        if (r < 0) mt_failed(r);
    }

    // Head of line_search_morethuente, lbfgs.hpp:743-787.
    int mt_begin() {
        count = 0;
        uinfo = 0;
        if (step <= 0.) return LBERR_INVALIDPARAMETERS;
        dginit = seq_dot(gp.data(), d.data(), n);
        if (0 < dginit) return LBERR_INCREASEGRADIENT;
        brackt = 0;
        stage1 = 1;
        finit = fx_;
        dgtest = par.f_dec_coeff * dginit;
        width = step_max - step_min;
        prev_width = 2.0 * width;
        bx = {0., finit, dginit};
        by = {0., finit, dginit};
        mt_propose();
        return 0;
    }

    // Top half of the loop body, lbfgs.hpp:788-826: choose the trial step and form x.
    void mt_propose() {
        if (brackt) {
            stmin = bx.st <= by.st ? bx.st : by.st;
            stmax = bx.st >= by.st ? bx.st : by.st;
        } else {
            stmin = bx.st;
            stmax = step + 4.0 * (step - bx.st);
        }
        if (step < step_min) step = step_min;
        if (step_max < step) step = step_max;
        if ((brackt && ((step <= stmin || stmax <= step) || par.max_linesearch <= count + 1 || uinfo != 0)) ||
            (brackt && (stmax - stmin <= par.xtol * stmax))) {
            step = bx.st;
        }
        for (int i = 0; i < n; ++i) xv[i] = xp[i];
        for (int i = 0; i < n; ++i) xv[i] += step * d[i];
        phase = PH_MT;
    }

    // Bottom half of the loop body, lbfgs.hpp:829-935.
    void on_mt_eval(double f) {
        fx_ = f;
        double dg = seq_dot(gv.data(), d.data(), n);
        const double ftest1 = finit + step * dgtest;
        ++count;
        int r = 0;
        bool finished = false;
        if ((std::isinf(fx_) || std::isnan(fx_)) || (brackt && ((step <= stmin || stmax <= step) || uinfo != 0))) {
            r = LBERR_ROUNDING_ERROR, finished = true;
        } else if (step == step_max && fx_ <= ftest1 && dg <= dgtest) {
            r = LBERR_MAXIMUMSTEP, finished = true;
        } else if (step == step_min && (ftest1 < fx_ || dgtest <= dg)) {
            r = LBERR_MINIMUMSTEP, finished = true;
        } else if (brackt && (stmax - stmin) <= par.xtol * stmax) {
            r = LBERR_WIDTHTOOSMALL, finished = true;
        } else if (par.max_linesearch <= count) {
            r = LBERR_MAXIMUMLINESEARCH, finished = true;
        } else if (fx_ <= ftest1 && std::fabs(dg) <= par.s_curv_coeff * (-dginit)) {
            r = count, finished = true;
        }
        if (finished) {
            if (r < 0)
                mt_failed(r);
            else
                line_search_done(r);
            return;
        }
        const double mincoef = par.f_dec_coeff <= par.s_curv_coeff ? par.f_dec_coeff : par.s_curv_coeff;
        if (stage1 && fx_ <= ftest1 && mincoef * dginit <= dg) stage1 = 0;
        MtPoint t{step, fx_, dg};
        if (stage1 && ftest1 < fx_ && fx_ <= bx.f) {
            // modified function values, lbfgs.hpp:893-913
            MtPoint xm{bx.st, bx.f - bx.st * dgtest, bx.dg - dgtest};
            MtPoint ym{by.st, by.f - by.st * dgtest, by.dg - dgtest};
            MtPoint tm{step, fx_ - step * dgtest, dg - dgtest};
            uinfo = mt_update_interval(xm, ym, tm, stmin, stmax, brackt);
            bx = {xm.st, xm.f + xm.st * dgtest, xm.dg + dgtest};
            by = {ym.st, ym.f + ym.st * dgtest, ym.dg + dgtest};
            step = tm.st;
        } else {
            uinfo = mt_update_interval(bx, by, t, stmin, stmax, brackt);
            step = t.st;
            // NB the reference passes f and dg by pointer; update_trial_interval never writes
            // *ft/*dt, so fx_ is unchanged here.
        }
        if (brackt) {
            if (0.66 * prev_width <= std::fabs(by.st - bx.st)) step = bx.st + 0.5 * (by.st - bx.st);
            prev_width = width;
            width = std::fabs(by.st - bx.st);
        }
        mt_propose();
    }

    // lbfgs.hpp:1271-1282: More-Thuente failed -> restore and run the backtracking search.
    void mt_failed(int ls_err) {
        ls_mt_err = ls_err;
        step = stepp;
        fx_ = fp;
        int r = bt_begin();
        if (r < 0) line_search_failed(r);
    }

    // line_search_backtracking head, lbfgs.hpp:954-978.
    int bt_begin() {
        count = 0;
        if (step <= 0.) return LBERR_INVALIDPARAMETERS;
        dginit = seq_dot(gp.data(), d.data(), n);
        if (0 < dginit) return LBERR_INCREASEGRADIENT;
        finit = fx_;
        dgtest = par.f_dec_coeff * dginit;
        bt_propose();
        return 0;
    }
    void bt_propose() {
        for (int i = 0; i < n; ++i) xv[i] = xp[i];
        for (int i = 0; i < n; ++i) xv[i] += step * d[i];
        phase = PH_BT;
    }
    // lbfgs.hpp:984-1031
    void on_bt_eval(double f) {
        fx_ = f;
        ++count;
        double wdt;
        if (fx_ > finit + step * dgtest) {
            wdt = 0.5;
        } else {
            const double dg = seq_dot(gv.data(), d.data(), n);
            if (dg < par.s_curv_coeff * dginit) {
                wdt = 2.1;
            } else if (dg > -par.s_curv_coeff * dginit) {
                wdt = 0.5;
            } else {
                line_search_done(count);
                return;
            }
        }
        if (step < step_min) return line_search_failed(LBERR_MINIMUMSTEP);
        if (step > step_max) return line_search_failed(LBERR_MAXIMUMSTEP);
        if (par.max_linesearch <= count) return line_search_failed(LBERR_MAXIMUMLINESEARCH);
        step *= wdt;
        bt_propose();
    }

    // lbfgs.hpp:1284-1292: revert to the previous point and stop.
    void line_search_failed(int err) {
        xv = xp;
        gv = gp;
        ret = err;
        phase = PH_DONE;
    }

    // lbfgs.hpp:1294-1416: convergence tests, history update, two-loop recursion.
    void line_search_done(int /*ls*/) {
        double xnorm = std::sqrt(seq_dot(xv.data(), xv.data(), n));
        const double gnorm = std::sqrt(seq_dot(gv.data(), gv.data(), n));
        if (xnorm < 1.0) xnorm = 1.0;
        if (gnorm / xnorm <= par.g_epsilon) {
            ret = LB_CONVERGENCE;
            phase = PH_DONE;
            return;
        }
        if (!pf.empty()) {
            if (par.past <= k) {
                const double rate = (pf[k % par.past] - fx_) / fx_;
                if (std::fabs(rate) < par.delta) {
                    ret = LB_STOP;
                    phase = PH_DONE;
                    return;
                }
            }
            pf[k % par.past] = fx_;
        }
        if (par.max_iterations != 0 && par.max_iterations < k + 1) {
            ret = LBERR_MAXIMUMITERATION;
            phase = PH_DONE;
            return;
        }
        double* s_end = &S[(size_t)end * n];
        double* y_end = &Y[(size_t)end * n];
        for (int i = 0; i < n; ++i) s_end[i] = xv[i] - xp[i];
        for (int i = 0; i < n; ++i) y_end[i] = gv[i] - gp[i];
        const double ys = seq_dot(y_end, s_end, n);
        const double yy = seq_dot(y_end, y_end, n);
        ysv[end] = ys;
        const int bound = (m <= k) ? m : k;
        ++k;
        end = (end + 1) % m;
        for (int i = 0; i < n; ++i) d[i] = -gv[i];
        int j = end;
        for (int i = 0; i < bound; ++i) {
            j = (j + m - 1) % m;
            const double* sj = &S[(size_t)j * n];
            const double* yj = &Y[(size_t)j * n];
            alpha[j] = seq_dot(sj, d.data(), n);
            alpha[j] /= ysv[j];
            const double c = -alpha[j];
            for (int q = 0; q < n; ++q) d[q] += c * yj[q];
        }
        {
            const double c = ys / yy;
            for (int q = 0; q < n; ++q) d[q] *= c;
        }
        for (int i = 0; i < bound; ++i) {
            const double* sj = &S[(size_t)j * n];
            const double* yj = &Y[(size_t)j * n];
            double beta = seq_dot(yj, d.data(), n);
            beta /= ysv[j];
            const double c = alpha[j] - beta;
            for (int q = 0; q < n; ++q) d[q] += c * sj[q];
            j = (j + 1) % m;
        }
        step = 1.0;
        start_iteration();
    }

    int n = 0, m = 0;
    LbfgsParams par;
    Phase phase = PH_DONE;
    int ret = LBERR_UNKNOWNERROR;
    std::vector<double> xv, gv, xp, gp, d, S, Y, alpha, ysv, pf;
    double fx_ = 0., fp = 0., step = 0., stepp = 0., step_min = 0., step_max = 0.;
    int k = 0, end = 0, n_evals = 0;
    // line-search state
    int count = 0, brackt = 0, stage1 = 0, uinfo = 0, ls_mt_err = 0;
    double dginit = 0., finit = 0., dgtest = 0., width = 0., prev_width = 0., stmin = 0., stmax = 0.;
    MtPoint bx{0, 0, 0}, by{0, 0, 0};
};

// Convenience driver with the callback signature of lbfgs.hpp:223-226.
typedef double (*eval_fn)(void* instance, const double* x, double* g, const int n);
inline int lbfgs_minimize(int n, double* x, double* ptr_fx, eval_fn f, void* inst, const LbfgsParams& p,
                          int* iters = nullptr, int* evals = nullptr) {
    LbfgsRC s;
    int r = s.begin(n, x, p);
    if (r < 0) return r;
    while (s.wants_eval()) {
        const double fx = f(inst, s.x(), s.g(), n);
        s.feed(fx);
    }
    std::memcpy(x, s.x(), sizeof(double) * n);
    if (ptr_fx) *ptr_fx = s.fx();
    if (iters) *iters = s.iterations();
    if (evals) *evals = s.evaluations();
    return s.status();
}

}  // namespace frx_oracle
