// Ask/tell COBYLA for the unconstrained problems of the path (SURVEY.md section 8 f-2 companion).
//
// The reference calls scipy.optimize.minimize(cost, x0, method="COBYLA", options={"maxiter": 1000}) with NO constraints
// (environments/environment_qulacs.py:429-445).  This is Powell's COBYLA (M. J. D. Powell, "A direct search optimization
// method that models the objective and constraint functions by linear interpolation", 1994) restated for m = 0
// constraints: a simplex of n + 1 points carries a linear model of the objective; every iteration either takes the
// trust-region step x_pole - rho * g / |g| or replaces a vertex to restore the simplex geometry (the alpha / beta / gamma /
// delta rules), and rho is halved when neither helps, down to rhoend.  With m = 0 the linear-programming subproblem has
// the closed-form solution above and the penalty parameter stays zero.
//
// Why native: under scipy >= 1.16 COBYLA is pure Python (2-12 ms of interpreter time per iteration, under the GIL), which
// caps what a lock-step multi-environment driver can gain from evaluating B energies in one launch.  Here an iteration is
// O(n^2) flops in C++, and the ask/tell form lets one host loop drive B optimisers: ask all -> one batched GPU launch ->
// tell all.  The optimiser's trajectory is NOT scipy's (different COBYLA lineage: scipy >= 1.16 ships PRIMA's), so it sits
// behind a switch and scipy stays the default (the reference's episodes are pinned to scipy's path).
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <new>
#include <vector>

#include "../../include/tqsim.h"

namespace {

struct Cobyla {
    int n = 0, np = 0, maxfun = 0, nfvals = 0;
    double rho = 0, rhoend = 0;
    std::vector<double> sim, simi, fv, a, vsig, veta, sigbar, dx, x, w, xbest_seen;
    double fbest_seen = 0;
    bool have_best = false;
    int ibrnch = 0, iflag = 0, jdrop = 0, ifull = 0;
    double prerem = 0, parsig = 0, pareta = 0;
    int status = 0;   // 0 running, 1 converged (rho reached rhoend), 2 maxfun reached, 3 rounding errors
    bool waiting = false;   // an x has been handed out and its value is pending
    double f_final = 0;

    double& S(int i, int j) { return sim[(size_t)i * np + j]; }     // n x (n + 1); column n = the pole
    double& SI(int i, int j) { return simi[(size_t)i * n + j]; }    // n x n
};

constexpr double kAlpha = 0.25, kBeta = 2.1, kGamma = 0.5, kDelta = 1.1;

// label 40 of Powell's code: hand out x unless the budget is spent
bool next_eval(Cobyla& c) {
    if (c.nfvals >= c.maxfun && c.nfvals > 0) {
        c.status = 2;
        return false;
    }
    ++c.nfvals;
    c.waiting = true;
    return true;
}

void finish_from_pole(Cobyla& c) {
    for (int i = 0; i < c.n; ++i) c.x[i] = c.S(i, c.n);
    c.f_final = c.fv[c.n];
}

// replace vertex jdrop by pole + dx and keep simi = inverse of the displacement matrix
void replace_vertex(Cobyla& c, int jdrop) {
    const int n = c.n;
    double temp = 0.0;
    for (int i = 0; i < n; ++i) {
        c.S(i, jdrop) = c.dx[i];
        temp += c.SI(jdrop, i) * c.dx[i];
    }
    for (int i = 0; i < n; ++i) c.SI(jdrop, i) /= temp;
    for (int j = 0; j < n; ++j) {
        if (j == jdrop) continue;
        double t = 0.0;
        for (int i = 0; i < n; ++i) t += c.SI(j, i) * c.dx[i];
        for (int i = 0; i < n; ++i) c.SI(j, i) -= t * c.SI(jdrop, i);
    }
}

// everything between two function evaluations; returns true when c.x holds the next point to evaluate
bool advance(Cobyla& c, double f) {
    const int n = c.n;
    if (!c.have_best || f < c.fbest_seen) {
        c.have_best = true;
        c.fbest_seen = f;
        c.xbest_seen = c.x;
    }
    bool from_eval = true;
    if (c.ibrnch == 1) goto L440;
    c.fv[c.jdrop] = f;
    if (c.nfvals > c.np) goto L130;
    // building the initial simplex: swap the new vertex with the pole if it is better
    if (c.jdrop < n) {
        const int jd = c.jdrop;
        if (c.fv[n] <= f) {
            c.x[jd] = c.S(jd, n);
        } else {
            c.S(jd, n) = c.x[jd];
            c.fv[jd] = c.fv[n];
            c.fv[n] = f;
            for (int k = 0; k <= jd; ++k) {
                c.S(jd, k) = -c.rho;
                double temp = 0.0;
                for (int i = k; i <= jd; ++i) temp -= c.SI(i, k);
                c.SI(jd, k) = temp;
            }
        }
    }
    if (c.nfvals <= n) {
        c.jdrop = c.nfvals - 1;
        c.x[c.jdrop] += c.rho;
        return next_eval(c);
    }
L130:
    c.ibrnch = 1;
L140: {
    // best vertex into pole position
    int nbest = n;
    double phimin = c.fv[n];
    for (int j = 0; j < n; ++j)
        if (c.fv[j] < phimin) { nbest = j; phimin = c.fv[j]; }
    if (nbest < n) {
        std::swap(c.fv[n], c.fv[nbest]);
        for (int i = 0; i < n; ++i) {
            const double temp = c.S(i, nbest);
            c.S(i, nbest) = 0.0;
            c.S(i, n) += temp;
            double tempa = 0.0;
            for (int k = 0; k < n; ++k) {
                c.S(i, k) -= temp;
                tempa -= c.SI(k, i);
            }
            c.SI(nbest, i) = tempa;
        }
    }
    // simi must still be the inverse of the displacement matrix
    double error = 0.0;
    for (int i = 0; i < n; ++i)
        for (int j = 0; j < n; ++j) {
            double temp = (i == j) ? -1.0 : 0.0;
            for (int k = 0; k < n; ++k) temp += c.SI(i, k) * c.S(k, j);
            error = std::max(error, std::fabs(temp));
        }
    if (error > 0.1) {
        c.status = 3;
        finish_from_pole(c);
        return false;
    }
    // a = minus the gradient of the linear model
    for (int j = 0; j < n; ++j) c.w[j] = c.fv[j] - c.fv[n];
    for (int i = 0; i < n; ++i) {
        double temp = 0.0;
        for (int j = 0; j < n; ++j) temp += c.w[j] * c.SI(j, i);
        c.a[i] = -temp;
    }
    // acceptability of the simplex
    c.iflag = 1;
    c.parsig = kAlpha * c.rho;
    c.pareta = kBeta * c.rho;
    for (int j = 0; j < n; ++j) {
        double wsig = 0.0, weta = 0.0;
        for (int i = 0; i < n; ++i) {
            wsig += c.SI(j, i) * c.SI(j, i);
            weta += c.S(i, j) * c.S(i, j);
        }
        c.vsig[j] = 1.0 / std::sqrt(wsig);
        c.veta[j] = std::sqrt(weta);
        if (c.vsig[j] < c.parsig || c.veta[j] > c.pareta) c.iflag = 0;
    }
    if (c.ibrnch == 1 || c.iflag == 1) goto L370;
    // geometry step: drop the vertex that is too far away, else the one closest to the opposite face
    int jd = -1;
    double temp = c.pareta;
    for (int j = 0; j < n; ++j)
        if (c.veta[j] > temp) { jd = j; temp = c.veta[j]; }
    if (jd < 0)
        for (int j = 0; j < n; ++j)
            if (c.vsig[j] < temp) { jd = j; temp = c.vsig[j]; }
    temp = kGamma * c.rho * c.vsig[jd];
    double sum = 0.0;
    for (int i = 0; i < n; ++i) {
        c.dx[i] = temp * c.SI(jd, i);
        sum += c.a[i] * c.dx[i];
    }
    if (0.0 > sum + sum)   // the model predicts an increase along dx: go the other way
        for (int i = 0; i < n; ++i) c.dx[i] = -c.dx[i];
    replace_vertex(c, jd);
    for (int j = 0; j < n; ++j) c.x[j] = c.S(j, n) + c.dx[j];
    c.jdrop = jd;
    return next_eval(c);
}
L370: {
    // trust-region step of the linear model (no constraints: steepest descent to the boundary)
    double g2 = 0.0;
    for (int i = 0; i < n; ++i) g2 += c.a[i] * c.a[i];
    if (g2 > 0.0) {
        const double s = c.rho / std::sqrt(g2);
        for (int i = 0; i < n; ++i) c.dx[i] = s * c.a[i];
        c.ifull = 1;
    } else {
        for (int i = 0; i < n; ++i) c.dx[i] = 0.0;
        c.ifull = 0;
    }
    if (c.ifull == 0) {
        c.ibrnch = 1;
        from_eval = false;
        goto L550;
    }
    double sum = 0.0;
    for (int i = 0; i < n; ++i) sum -= c.a[i] * c.dx[i];
    c.prerem = -sum;   // predicted reduction of the objective
    for (int i = 0; i < n; ++i) c.x[i] = c.S(i, n) + c.dx[i];
    c.ibrnch = 1;
    return next_eval(c);
}
L440: {
    double trured = c.fv[n] - f;
    if (f == c.fv[n]) { c.prerem = 0.0; trured = 0.0; }
    double ratio = (trured <= 0.0) ? 1.0 : 0.0;
    int jd = -1;
    for (int j = 0; j < n; ++j) {
        double temp = 0.0;
        for (int i = 0; i < n; ++i) temp += c.SI(j, i) * c.dx[i];
        temp = std::fabs(temp);
        if (temp > ratio) { jd = j; ratio = temp; }
        c.sigbar[j] = temp * c.vsig[j];
    }
    double edgmax = kDelta * c.rho;
    int l = -1;
    for (int j = 0; j < n; ++j) {
        if (c.sigbar[j] >= c.parsig || c.sigbar[j] >= c.vsig[j]) {
            double temp = c.veta[j];
            if (trured > 0.0) {
                temp = 0.0;
                for (int i = 0; i < n; ++i) temp += (c.dx[i] - c.S(i, j)) * (c.dx[i] - c.S(i, j));
                temp = std::sqrt(temp);
            }
            if (temp > edgmax) { l = j; edgmax = temp; }
        }
    }
    if (l >= 0) jd = l;
    if (jd >= 0) {
        replace_vertex(c, jd);
        c.fv[jd] = f;
        if (trured > 0.0 && trured >= 0.1 * c.prerem) goto L140;
    }
}
L550:
    (void)from_eval;
    if (c.iflag == 0) {
        c.ibrnch = 0;
        goto L140;
    }
    if (c.rho > c.rhoend) {
        c.rho *= 0.5;
        if (c.rho <= 1.5 * c.rhoend) c.rho = c.rhoend;
        goto L140;
    }
    c.status = 1;
    finish_from_pole(c);
    return false;
}

}  // namespace

struct tq_cobyla { Cobyla c; };

// (exception barrier: nothing throws across the C ABI -- include/tqsim.h)
extern "C" {

int tq_cobyla_create(int n, const double* x0, double rhobeg, double rhoend, int maxfun, tq_cobyla_handle* out) {
    try {
    if (!out) return TQ_EINVAL;
    *out = nullptr;
    if (n < 1 || !x0 || !(rhobeg > 0.0) || !(rhoend > 0.0) || rhoend > rhobeg || maxfun < 1) return TQ_EINVAL;
    tq_cobyla* h = new tq_cobyla();
    Cobyla& c = h->c;
    c.n = n;
    c.np = n + 1;
    c.maxfun = maxfun;
    c.rho = rhobeg;
    c.rhoend = rhoend;
    c.sim.assign((size_t)n * (n + 1), 0.0);
    c.simi.assign((size_t)n * n, 0.0);
    c.fv.assign(n + 1, 0.0);
    for (auto* v : {&c.a, &c.vsig, &c.veta, &c.sigbar, &c.dx, &c.w}) v->assign(n, 0.0);
    c.x.assign(x0, x0 + n);
    for (int i = 0; i < n; ++i) {
        c.S(i, n) = x0[i];
        c.S(i, i) = rhobeg;
        c.SI(i, i) = 1.0 / rhobeg;
    }
    c.jdrop = n;
    c.ibrnch = 0;
    next_eval(c);   // the first point is x0
    *out = h;
    return TQ_OK;
    } catch (const std::bad_alloc&) {
        return TQ_ENOMEM;
    } catch (...) {
        return TQ_EINVAL;
    }
}

int tq_cobyla_destroy(tq_cobyla_handle h) {
    try {
    delete h;
    return TQ_OK;
    } catch (const std::bad_alloc&) {
        return TQ_ENOMEM;
    } catch (...) {
        return TQ_EINVAL;
    }
}

int tq_cobyla_ask(tq_cobyla_handle h, double* x_out) {
    try {
    if (!h || !x_out) return TQ_EINVAL;
    if (!h->c.waiting) return 1;   // finished: nothing to evaluate
    std::copy(h->c.x.begin(), h->c.x.end(), x_out);
    return 0;
    } catch (const std::bad_alloc&) {
        return TQ_ENOMEM;
    } catch (...) {
        return TQ_EINVAL;
    }
}

int tq_cobyla_tell(tq_cobyla_handle h, double f) {
    try {
    if (!h) return TQ_EINVAL;
    Cobyla& c = h->c;
    if (!c.waiting) return TQ_ESTATE;
    c.waiting = false;
    if (!std::isfinite(f)) f = 1e300;   // keep the simplex arithmetic finite
    if (!advance(c, f)) {
        if (c.status == 2) finish_from_pole(c);   // budget spent: best vertex of the simplex
        // never hand back something worse than the best point evaluated
        if (c.have_best && c.fbest_seen < c.f_final) {
            c.x = c.xbest_seen;
            c.f_final = c.fbest_seen;
        }
        return 1;
    }
    return 0;
    } catch (const std::bad_alloc&) {
        return TQ_ENOMEM;
    } catch (...) {
        return TQ_EINVAL;
    }
}

int tq_cobyla_result(tq_cobyla_handle h, double* x_out, double* f_out, int* nfev_out, int* status_out) {
    try {
    if (!h) return TQ_EINVAL;
    const Cobyla& c = h->c;
    if (c.waiting) return TQ_ESTATE;
    if (x_out) std::copy(c.x.begin(), c.x.end(), x_out);
    if (f_out) *f_out = c.f_final;
    if (nfev_out) *nfev_out = c.nfvals;
    if (status_out) *status_out = c.status;
    return TQ_OK;
    } catch (const std::bad_alloc&) {
        return TQ_ENOMEM;
    } catch (...) {
        return TQ_EINVAL;
    }
}

}  // extern "C"
