/*
 * sdp_oracle.c - CPU restatement of the reference's hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under stodynprog_b200/ may import, call,
 * link or execute this file; only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs use it, as the checker.
 *
 * Reference: pierre-haessig/stodynprog.  Each function cites the file:line it
 * restates.  Compile with -O2 -ffp-contract=off (oracle/build.py): plain fp64,
 * true divisions, no FMA - the reference's .so contains no FMA instruction
 * either.  The double->int cast is written out with the x86
 * cvttsd2si semantics instead of relying on C undefined behaviour.
 *
 * Pinning: tests/test_oracle.py checks these functions against
 *   - the reference's own known-answer vectors (stodynprog/tests/test_dolointerp.py:17-40, :45-93),
 *   - the compiled reference routine itself (oracle/_ref, bit-exact on random
 *     and adversarial inputs) when it is present,
 *   - golden fixtures generated from the unmodified reference (tests/golden/).
 */
#include <limits.h>
#include <math.h>
#include <stdint.h>
#include <string.h>

#define ORACLE_MAX_D 4

/* (int)t as the x86-64 build of the reference does it: truncate toward zero,
 * INT_MIN for NaN / out-of-range (cvttsd2si "integer indefinite"). */
static int cast_int(double t) {
    if (!(t > -2147483649.0 && t < 2147483648.0)) return INT_MIN;
    return (int)t;
}

/* multilinear_cython.pyx:117-131 (and the 1-D/3-D/4-D analogues):
 *   sn = (s - smin)/(smax - smin)
 *   q  = max(min(<int>(sn*(order-1)), order-2), 0)
 *   lam = sn*(order-1) - q                                  (not clamped) */
static void cell_1d(double s, double smin, double smax, int order, int* q, double* lam) {
    double sn = (s - smin) / (smax - smin);
    double t = sn * (order - 1);
    int qi = cast_int(t);
    if (qi > order - 2) qi = order - 2;
    if (qi < 0) qi = 0;
    *q = qi;
    *lam = t - qi;
}

/* nested lerp, last axis innermost: pyx:88 (1-D), :140 (2-D), :208 (3-D), :300 (4-D) */
static double lerp_rec(const double* V, int base, const int* stride, const double* lam, int d, int k) {
    if (k == d) return V[base];
    double a = lerp_rec(V, base, stride, lam, d, k + 1);
    double b = lerp_rec(V, base + stride[k], stride, lam, d, k + 1);
    return (1 - lam[k]) * a + lam[k] * b;
}

static int strides_of(int d, const int64_t* orders, int* stride) {
    int64_t n = 1;
    for (int k = d - 1; k >= 0; --k) {
        stride[k] = (int)n; /* M_k = prod_{j>k} order_j, pyx:109,164-165,235-237 */
        n *= orders[k];
    }
    return (int)n;
}

/* multilinear_interpolation(smin, smax, orders, values, s)  pyx:17-49
 * values [n_v][prod(orders)], s [d][n_s], out [n_v][n_s].
 * returns -1 for d outside 1..4 (the reference raises there, pyx:46-47). */
int oracle_interp(int d, const double* smin, const double* smax, const int64_t* orders,
                  int64_t n_v, const double* values, int64_t n_s, const double* s, double* out) {
    if (d < 1 || d > ORACLE_MAX_D) return -1;
    int stride[ORACLE_MAX_D];
    int64_t n_grid = strides_of(d, orders, stride);
    for (int64_t v = 0; v < n_v; ++v) {
        const double* V = values + v * n_grid;
        for (int64_t i = 0; i < n_s; ++i) {
            int base = 0;
            double lam[ORACLE_MAX_D];
            for (int k = 0; k < d; ++k) {
                int q;
                cell_1d(s[k * n_s + i], smin[k], smax[k], (int)orders[k], &q, &lam[k]);
                base += stride[k] * q;
            }
            out[v * n_s + i] = lerp_rec(V, base, stride, lam, d, 0);
        }
    }
    return 0;
}

/* fp32 specialisation of the fused type (pyx:12-14): every operation in float */
static int cast_int_f(float t) {
    if (!(t >= -2147483648.0f && t < 2147483648.0f)) return INT_MIN;
    return (int)t;
}
/* The generated C of the float specialisation writes the weights as
 * `(1.0 - lam_k)` with a DOUBLE literal (Cython turns the pyx's `1` into 1.0),
 * so by C's usual arithmetic conversions the `(1.0-lam)*a` terms and the sums
 * are evaluated in double, while the innermost `lam*v` products of two floats
 * stay float products; the result is rounded to float on the final store.
 * Restated here with explicit types. */
static double lerp_rec_f(const float* V, int base, const int* stride, const float* lam, int d, int k) {
    if (k == d - 1) {
        float v0 = V[base], v1 = V[base + stride[k]];
        float t2 = lam[k] * v1;                       /* float * float */
        return (1.0 - (double)lam[k]) * (double)v0 + (double)t2;
    }
    double a = lerp_rec_f(V, base, stride, lam, d, k + 1);
    double b = lerp_rec_f(V, base + stride[k], stride, lam, d, k + 1);
    return (1.0 - (double)lam[k]) * a + (double)lam[k] * b;
}
int oracle_interp_f32(int d, const float* smin, const float* smax, const int64_t* orders,
                      int64_t n_v, const float* values, int64_t n_s, const float* s, float* out) {
    if (d < 1 || d > ORACLE_MAX_D) return -1;
    int stride[ORACLE_MAX_D];
    int64_t n_grid = strides_of(d, orders, stride);
    for (int64_t v = 0; v < n_v; ++v) {
        const float* V = values + v * n_grid;
        for (int64_t i = 0; i < n_s; ++i) {
            int base = 0;
            float lam[ORACLE_MAX_D];
            for (int k = 0; k < d; ++k) {
                float sn = (s[k * n_s + i] - smin[k]) / (smax[k] - smin[k]);
                float t = sn * (float)((int)orders[k] - 1);
                int q = cast_int_f(t);
                if (q > (int)orders[k] - 2) q = (int)orders[k] - 2;
                if (q < 0) q = 0;
                lam[k] = t - (float)q;
                base += stride[k] * q;
            }
            out[v * n_s + i] = (float)lerp_rec_f(V, base, stride, lam, d, 0);
        }
    }
    return 0;
}

/* cell search only: the integer base index and the d weights the GPU setup
 * kernel must reproduce bit-for-bit. s [d][n] -> cell [n], lam [d][n] */
int oracle_cell_search(int d, const double* smin, const double* smax, const int64_t* orders,
                       int64_t n, const double* s, int32_t* cell, double* lam) {
    if (d < 1 || d > ORACLE_MAX_D) return -1;
    int stride[ORACLE_MAX_D];
    strides_of(d, orders, stride);
    for (int64_t i = 0; i < n; ++i) {
        int base = 0;
        for (int k = 0; k < d; ++k) {
            int q;
            cell_1d(s[k * n + i], smin[k], smax[k], (int)orders[k], &q, &lam[k * n + i]);
            base += stride[k] * q;
        }
        cell[i] = base;
    }
    return 0;
}

/* numpy's DOUBLE_argmin (first minimum, a NaN stops the scan and wins):
 * what `J.argmin()` does at stodynprog.py:686 */
int64_t oracle_argmin(const double* J, int64_t n) {
    double mp = J[0];
    int64_t idx = 0;
    if (mp != mp) return 0;
    for (int64_t i = 1; i < n; ++i) {
        double v = J[i];
        if (!(v >= mp)) { /* v < mp or v is NaN */
            mp = v;
            idx = i;
            if (mp != mp) break;
        }
    }
    return idx;
}

/* One state's backup, stodynprog.py:674-690 after the callables returned:
 *   Jg[u][w] = g[u][w] + interp(x_next[.][u][w])          :677
 *   J[u] = sum_w Jg[u][w]*p[w]  (sequential, index order)   :682  (np.inner; see note)
 *        or J[u] = Jg[u][0] when p == NULL (deterministic)  :679-680
 *   ind = argmin(J)  (first minimum)                        :686
 * Coordinates and cost come un-broadcast: array k has element (u,w) at
 * x[k][u*us[k] + w*ws[k]].  J_all (may be NULL) receives J[u].
 * Note: np.inner's summation order / FMA use inside OpenBLAS is not part of
 * the reference's source; the plain ordered sum here is the documented
 * restatement and the tolerance on J is 1e-10 relative. */
int oracle_backup(int d, const double* smin, const double* smax, const int64_t* orders,
                  const double* J_next, int64_t U, int64_t W,
                  const double* const* x, const int64_t* us, const int64_t* ws,
                  const double* g, int64_t g_us, int64_t g_ws, const double* p,
                  double* J_opt, int64_t* ind_opt, double* J_all) {
    if (d < 1 || d > ORACLE_MAX_D || U < 1 || W < 1) return -1;
    int stride[ORACLE_MAX_D];
    strides_of(d, orders, stride);
    double best = 0;
    int64_t best_i = -1;
    int stop = 0;
    for (int64_t u = 0; u < U; ++u) {
        double acc = 0.0;
        for (int64_t w = 0; w < W; ++w) {
            int base = 0;
            double lam[ORACLE_MAX_D];
            for (int k = 0; k < d; ++k) {
                int q;
                cell_1d(x[k][u * us[k] + w * ws[k]], smin[k], smax[k], (int)orders[k], &q, &lam[k]);
                base += stride[k] * q;
            }
            double jg = g[u * g_us + w * g_ws] + lerp_rec(J_next, base, stride, lam, d, 0);
            if (p) acc = acc + jg * p[w];
            else acc = jg;
        }
        if (J_all) J_all[u] = acc;
        if (!stop) {
            if (best_i < 0) { best = acc; best_i = 0; if (acc != acc) stop = 1; }
            else if (!(acc >= best)) { best = acc; best_i = u; if (acc != acc) stop = 1; }
        }
    }
    *J_opt = best;
    *ind_opt = best_i;
    return 0;
}

/* Fixed-policy backup over all states, stodynprog.py:752-757:
 *   J_out[i] = sum_w (g[i][w] + interp(x_next[.][i][w])) * p[w]
 * coordinates dense: s [d][n*W] with point index i*W + w; g [n][W] or [n] (g_ws=0). */
int oracle_policy_backup(int d, const double* smin, const double* smax, const int64_t* orders,
                         const double* J_in, int64_t n, int64_t W, const double* s,
                         const double* g, int64_t g_ws, const double* p, double* J_out) {
    if (d < 1 || d > ORACLE_MAX_D) return -1;
    int stride[ORACLE_MAX_D];
    strides_of(d, orders, stride);
    int64_t npts = n * W;
    for (int64_t i = 0; i < n; ++i) {
        double acc = 0.0;
        for (int64_t w = 0; w < W; ++w) {
            int base = 0;
            double lam[ORACLE_MAX_D];
            for (int k = 0; k < d; ++k) {
                int q;
                cell_1d(s[k * npts + i * W + w], smin[k], smax[k], (int)orders[k], &q, &lam[k]);
                base += stride[k] * q;
            }
            double gv = g_ws ? g[i * W + w] : g[i];
            acc = acc + (gv + lerp_rec(J_in, base, stride, lam, d, 0)) * p[w];
        }
        J_out[i] = acc;
    }
    return 0;
}
