/*
 * oracle/pde_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT (see pde_oracle.h).
 *
 * Plain-C restatement of the seven PDE layers of MariMamgo/CNN-with-PDE, forward and
 * hand-derived adjoint.  Compiled twice by oracle/Makefile:
 *     -DREAL=float  -DSUF=f32     (mirrors the reference's fp32 op order, no FMA contraction)
 *     -DREAL=double -DSUF=f64     (the "truth" twin)
 * Parameter gradients are accumulated in double in both builds.
 *
 * Nothing here was copied from the reference: each routine re-derives the arithmetic the
 * cited reference lines perform.  The only simplification is that quantities that are
 * batch-independent in the reference (tridiagonal pivots, c*) are computed once per sweep
 * instead of once per line of every sample -- the values are bit-identical, because the
 * reference computes them from batch-expanded copies of the same coefficient map.
 */
#include "pde_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#ifndef REAL
#error "compile with -DREAL=float|double -DSUF=f32|f64"
#endif
#define CAT_(a, b) a##_##b
#define CAT(a, b) CAT_(a, b)
#define FN(name) CAT(name, SUF)

#define R(x) ((REAL)(x))

static REAL FN(r_exp)(REAL x) { return sizeof(REAL) == 4 ? (REAL)expf((float)x) : (REAL)exp((double)x); }
static REAL FN(r_sin)(REAL x) { return sizeof(REAL) == 4 ? (REAL)sinf((float)x) : (REAL)sin((double)x); }
static REAL FN(r_cos)(REAL x) { return sizeof(REAL) == 4 ? (REAL)cosf((float)x) : (REAL)cos((double)x); }

static int FN(pick_threads)(int want) {
#ifdef _OPENMP
    return want > 0 ? want : omp_get_max_threads();
#else
    (void)want;
    return 1;
#endif
}

/* ===================================================================================== */
/* Implicit ADI family                                                                   */
/* ===================================================================================== */

typedef struct {
    int axis;      /* 0: lines along W using alpha; 1: lines along H using beta */
    double t;      /* current_time when the coefficient maps are evaluated      */
    double dts;    /* time step handed to the sweep                             */
    double h;      /* spacing handed to the sweep                               */
} FN(sweep);

/* Step schedule and the Python-double accumulation of current_time:
 * Strang  mnist_test.py:49-63, fashion_mnist.py:53-67, SVHN.py:54-68, cifar10.py:84-110
 * Lie     cifar_2version.py:81-101 */
static int FN(build_schedule)(const oracle_adi_desc *d, FN(sweep) *sw) {
    double t = 0.0;
    int s = 0;
    for (int step = 0; step < d->steps; ++step) {
        if (!d->lie) {
            sw[s].axis = 0; sw[s].t = t; sw[s].dts = d->dt / 2; sw[s].h = d->hx; ++s;
            t += d->dt / 2;
            sw[s].axis = 1; sw[s].t = t; sw[s].dts = d->dt;     sw[s].h = d->hy; ++s;
            t += d->dt / 2;
            sw[s].axis = 0; sw[s].t = t; sw[s].dts = d->dt / 2; sw[s].h = d->hx; ++s;
        } else {
            sw[s].axis = 0; sw[s].t = t; sw[s].dts = d->dt / 2; sw[s].h = d->hx; ++s;
            t += d->dt / 2;
            sw[s].axis = 1; sw[s].t = t; sw[s].dts = d->dt / 2; sw[s].h = d->hy; ++s;
            t += d->dt / 2;
        }
    }
    return s;
}

static inline size_t FN(cell)(int axis, int N, int c, int l, int i) {
    /* line coordinates (l, i) -> offset into a [C][N][N] map */
    return axis == 0 ? ((size_t)c * N + l) * N + i : ((size_t)c * N + i) * N + l;
}

/* Coefficient map at time t, clamp, optional smoothing, r = coef*dt/h^2, tridiagonal
 * rows and the Thomas pivots with eps on every pivot.
 *   get_alpha_beta_at_time   mnist_test.py:33-42, cifar10.py:53-63, SVHN.py:38-47
 *   smooth_coefficients      mnist_test.py:135-149 (replicate pad, conv1d with ones/3)
 *   rows / Neumann ends      mnist_test.py:83-93, cifar10.py:130-145
 *   pivots                   mnist_test.py:169,177-181, cifar10.py:188-198
 * Tables are stored in LINE coordinates: [c][line][i]. */
static void FN(build_tables)(const oracle_adi_desc *d, const FN(sweep) *sw, const REAL *base,
                             const REAL *tc, REAL *r, REAL *den, REAL *cst, unsigned char *mask) {
    const int N = d->N, C = d->C;
    const REAL tt = R(sw->t), dts = R(sw->dts), h2 = R(sw->h * sw->h), eps = R(d->eps);
    const REAL lo = R(d->cmin), hi = R(d->cmax);
    const REAL third = R(1) / R(3);
    REAL *kap = (REAL *)malloc(sizeof(REAL) * (size_t)N * 2);
    REAL *ks = kap + N;
    for (int c = 0; c < C; ++c)
        for (int l = 0; l < N; ++l) {
            size_t o = ((size_t)c * N + l) * N;
            for (int i = 0; i < N; ++i) {
                size_t q = FN(cell)(sw->axis, N, c, l, i);
                REAL prod = tc[q] * tt;
                REAL raw = base[q] + prod;
                int m = raw >= lo;
                REAL k = raw < lo ? lo : raw;
                if (d->has_max) {
                    m = m && raw <= hi;
                    k = k > hi ? hi : k;
                }
                kap[i] = k;
                mask[o + i] = (unsigned char)m;
            }
            for (int i = 0; i < N; ++i) {
                if (d->smooth) {
                    REAL a0 = kap[i > 0 ? i - 1 : 0] * third;
                    REAL a1 = kap[i] * third;
                    REAL a2 = kap[i < N - 1 ? i + 1 : N - 1] * third;
                    REAL s01 = a0 + a1;
                    ks[i] = s01 + a2;
                } else
                    ks[i] = kap[i];
            }
            for (int i = 0; i < N; ++i) {
                REAL p = ks[i] * dts;
                r[o + i] = p / h2;
            }
            for (int i = 0; i < N; ++i) {
                REAL ri = r[o + i];
                REAL b;
                if (i == 0 || i == N - 1)
                    b = R(1) + ri;
                else {
                    REAL two = R(2) * ri;
                    b = R(1) + two;
                }
                REAL dn;
                if (i == 0)
                    dn = b + eps;
                else {
                    REAL ac = (-ri) * cst[o + i - 1];
                    REAL bm = b - ac;
                    dn = bm + eps;
                }
                den[o + i] = dn;
                cst[o + i] = i < N - 1 ? (-ri) / dn : R(0);
            }
        }
    free(kap);
}

/* One implicit sweep over one plane, in place.  thomas_solver_batch mnist_test.py:151-198,
 * thomas_solver_batch_optimized cifar10.py:179-211. */
static void FN(sweep_plane)(int axis, int N, REAL *x, const REAL *r, const REAL *den,
                            const REAL *cst, REAL *ds) {
    for (int l = 0; l < N; ++l) {
        const REAL *rl = r + (size_t)l * N, *dl = den + (size_t)l * N, *cl = cst + (size_t)l * N;
        const size_t st = axis == 0 ? 1 : (size_t)N;
        REAL *p = axis == 0 ? x + (size_t)l * N : x + l;
        ds[0] = p[0] / dl[0];
        for (int i = 1; i < N; ++i) {
            REAL ad = (-rl[i]) * ds[i - 1];
            REAL num = p[i * st] - ad;
            ds[i] = num / dl[i];
        }
        p[(N - 1) * st] = ds[N - 1];
        for (int i = N - 2; i >= 0; --i) {
            REAL cx = cl[i] * p[(i + 1) * st];
            p[i * st] = ds[i] - cx;
        }
    }
}

/* Adjoint of one sweep over one plane: g <- (A+eps I)^{-T} g, and
 * Gr[line][i] += lambda_i * (L_N x)_i with x the sweep OUTPUT. */
static void FN(sweep_plane_adj)(int axis, int N, REAL *g, const REAL *x, const REAL *r,
                                const REAL *den, const REAL *cst, REAL *w, double *Gr) {
    for (int l = 0; l < N; ++l) {
        const REAL *rl = r + (size_t)l * N, *dl = den + (size_t)l * N, *cl = cst + (size_t)l * N;
        const size_t st = axis == 0 ? 1 : (size_t)N;
        REAL *p = axis == 0 ? g + (size_t)l * N : g + l;
        const REAL *xs = axis == 0 ? x + (size_t)l * N : x + l;
        double *gr = Gr + (size_t)l * N;
        /* U^T w = g */
        w[0] = p[0];
        for (int i = 1; i < N; ++i) {
            REAL cw = cl[i - 1] * w[i - 1];
            w[i] = p[i * st] - cw;
        }
        /* L^T lambda = w */
        REAL lam = w[N - 1] / dl[N - 1];
        p[(N - 1) * st] = lam;
        for (int i = N - 2; i >= 0; --i) {
            REAL al = (-rl[i + 1]) * lam;
            REAL num = w[i] - al;
            lam = num / dl[i];
            p[i * st] = lam;
        }
        for (int i = 0; i < N; ++i) {
            REAL xc = xs[i * st];
            REAL lx;
            if (N == 1)
                lx = R(0);
            else if (i == 0)
                lx = xs[st] - xc;
            else if (i == N - 1)
                lx = xs[(N - 2) * st] - xc;
            else
                lx = (xs[(i - 1) * st] - xc) + (xs[(i + 1) * st] - xc);
            gr[i] += (double)p[i * st] * (double)lx;
        }
    }
}

/* u[c][p] <- sum_d Mat[c][d] u[d][p]   (cifar10.py:65-72 pre-step; SVHN.py:78-86 post-step) */
static void FN(chan_apply)(int C, size_t P, const REAL *Mat, const REAL *in, REAL *out) {
    for (size_t p = 0; p < P; ++p)
        for (int c = 0; c < C; ++c) {
            REAL acc = R(0);
            for (int dd = 0; dd < C; ++dd) {
                REAL pr = Mat[c * C + dd] * in[(size_t)dd * P + p];
                acc = acc + pr;
            }
            out[(size_t)c * P + p] = acc;
        }
}

static void FN(chan_adjoint)(int C, size_t P, const REAL *Mat, const REAL *uin, REAL *g,
                             REAL *tmp, double *gMat) {
    for (int c = 0; c < C; ++c)
        for (int dd = 0; dd < C; ++dd) {
            double acc = 0.0;
            for (size_t p = 0; p < P; ++p) acc += (double)g[(size_t)c * P + p] * (double)uin[(size_t)dd * P + p];
            gMat[c * C + dd] += acc;
        }
    for (size_t p = 0; p < P; ++p)
        for (int dd = 0; dd < C; ++dd) {
            REAL acc = R(0);
            for (int c = 0; c < C; ++c) {
                REAL pr = Mat[c * C + dd] * g[(size_t)c * P + p];
                acc = acc + pr;
            }
            tmp[(size_t)dd * P + p] = acc;
        }
    memcpy(g, tmp, sizeof(REAL) * (size_t)C * P);
}

typedef struct {
    int S;
    FN(sweep) *sw;
    REAL *r, *den, *cst;       /* [S][C][N][N] line coordinates */
    unsigned char *mask;
} FN(adi_tables);

static int FN(adi_tables_make)(const oracle_adi_desc *d, const REAL *ab, const REAL *bb,
                               const REAL *atc, const REAL *btc, FN(adi_tables) * T) {
    const int sps = d->lie ? 2 : 3;
    const size_t plane = (size_t)d->C * d->N * d->N;
    T->S = sps * d->steps;
    size_t n = (size_t)(T->S > 0 ? T->S : 1);
    T->sw = (FN(sweep) *)malloc(sizeof(FN(sweep)) * n);
    T->r = (REAL *)malloc(sizeof(REAL) * plane * n);
    T->den = (REAL *)malloc(sizeof(REAL) * plane * n);
    T->cst = (REAL *)malloc(sizeof(REAL) * plane * n);
    T->mask = (unsigned char *)malloc(plane * n);
    if (!T->sw || !T->r || !T->den || !T->cst || !T->mask) return -1;
    FN(build_schedule)(d, T->sw);
    for (int s = 0; s < T->S; ++s) {
        const int ax = T->sw[s].axis;
        FN(build_tables)(d, &T->sw[s], ax == 0 ? ab : bb, ax == 0 ? atc : btc, T->r + plane * s,
                         T->den + plane * s, T->cst + plane * s, T->mask + plane * s);
    }
    return 0;
}

static void FN(adi_tables_free)(FN(adi_tables) * T) {
    free(T->sw); free(T->r); free(T->den); free(T->cst); free(T->mask);
}

/* Runs the forward trajectory of one sample.  cur: [C][N][N] in/out.  If hist != NULL the
 * output of every sweep is stored at hist + s*C*N*N, and if in_state != NULL the state at
 * the start of every step (before the pre-step channel mix) at in_state + step*C*N*N. */
static void FN(adi_run_sample)(const oracle_adi_desc *d, const FN(adi_tables) * T, const REAL *chan,
                               REAL *cur, REAL *tmp, REAL *ds, REAL *hist, REAL *in_state) {
    const int N = d->N, C = d->C, sps = d->lie ? 2 : 3;
    const size_t P = (size_t)N * N, plane = (size_t)C * P;
    for (int step = 0; step < d->steps; ++step) {
        if (in_state) memcpy(in_state + plane * step, cur, sizeof(REAL) * plane);
        if (d->chan_op == 1) {
            FN(chan_apply)(C, P, chan, cur, tmp);
            memcpy(cur, tmp, sizeof(REAL) * plane);
        }
        for (int k = 0; k < sps; ++k) {
            const int s = step * sps + k;
            for (int c = 0; c < C; ++c)
                FN(sweep_plane)(T->sw[s].axis, N, cur + P * c, T->r + plane * s + P * c,
                                T->den + plane * s + P * c, T->cst + plane * s + P * c, ds);
            if (hist) memcpy(hist + plane * s, cur, sizeof(REAL) * plane);
        }
        if (d->chan_op == 2) {
            FN(chan_apply)(C, P, chan, cur, tmp);
            memcpy(cur, tmp, sizeof(REAL) * plane);
        }
    }
}

int FN(oracle_adi_forward)(const oracle_adi_desc *d, const REAL *u, const REAL *ab, const REAL *bb,
                           const REAL *atc, const REAL *btc, const REAL *chan, const REAL *skipw,
                           REAL *out) {
    if (d->N < 2 || d->C < 1 || d->B < 0 || d->steps < 0) return -2;
    if (d->chan_op && !chan) return -3;
    if (d->skip && !skipw) return -3;
    FN(adi_tables) T;
    if (FN(adi_tables_make)(d, ab, bb, atc, btc, &T)) return -1;
    const size_t plane = (size_t)d->C * d->N * d->N;
    REAL sig = R(0);
    if (d->skip) sig = R(1) / (R(1) + FN(r_exp)(-skipw[0]));
    const int nt = FN(pick_threads)(d->nthreads);
    int fail = 0;
#pragma omp parallel num_threads(nt)
    {
        REAL *cur = (REAL *)malloc(sizeof(REAL) * plane);
        REAL *tmp = (REAL *)malloc(sizeof(REAL) * plane);
        REAL *ds = (REAL *)malloc(sizeof(REAL) * (size_t)d->N);
        if (!cur || !tmp || !ds) {
#pragma omp atomic write
            fail = 1;
        } else {
#pragma omp for schedule(static)
            for (int b = 0; b < d->B; ++b) {
                memcpy(cur, u + plane * b, sizeof(REAL) * plane);
                FN(adi_run_sample)(d, &T, chan, cur, tmp, ds, NULL, NULL);
                REAL *o = out + plane * b;
                if (d->skip) {
                    const REAL *u0 = u + plane * b;
                    const REAL om = R(1) - sig;
                    for (size_t q = 0; q < plane; ++q) {
                        REAL x0 = sig * u0[q];
                        REAL x1 = om * cur[q];
                        o[q] = x0 + x1;
                    }
                } else
                    memcpy(o, cur, sizeof(REAL) * plane);
            }
        }
        free(cur); free(tmp); free(ds);
    }
    FN(adi_tables_free)(&T);
    return fail ? -1 : 0;
}

int FN(oracle_adi_backward)(const oracle_adi_desc *d, const REAL *u, const REAL *gout,
                            const REAL *ab, const REAL *bb, const REAL *atc, const REAL *btc,
                            const REAL *chan, const REAL *skipw, REAL *gin, double *g_ab,
                            double *g_bb, double *g_atc, double *g_btc, double *g_chan,
                            double *g_skipw) {
    if (d->N < 2 || d->C < 1 || d->B < 0 || d->steps < 0) return -2;
    if (d->chan_op && (!chan || !g_chan)) return -3;
    if (d->skip && (!skipw || !g_skipw)) return -3;
    FN(adi_tables) T;
    if (FN(adi_tables_make)(d, ab, bb, atc, btc, &T)) return -1;
    const int N = d->N, C = d->C, S = T.S, sps = d->lie ? 2 : 3;
    const size_t P = (size_t)N * N, plane = (size_t)C * P;
    REAL sig = R(0);
    if (d->skip) sig = R(1) / (R(1) + FN(r_exp)(-skipw[0]));
    double *Gr = (double *)calloc(plane * (size_t)(S > 0 ? S : 1), sizeof(double));
    double gM[64];
    double gW = 0.0;
    if (C * C > 64 || !Gr) { FN(adi_tables_free)(&T); free(Gr); return -1; }
    for (int q = 0; q < C * C; ++q) gM[q] = 0.0;
    const int nt = FN(pick_threads)(d->nthreads);
    int fail = 0;
#pragma omp parallel num_threads(nt)
    {
        REAL *cur = (REAL *)malloc(sizeof(REAL) * plane);
        REAL *tmp = (REAL *)malloc(sizeof(REAL) * plane);
        REAL *g = (REAL *)malloc(sizeof(REAL) * plane);
        REAL *ds = (REAL *)malloc(sizeof(REAL) * (size_t)N);
        REAL *hist = (REAL *)malloc(sizeof(REAL) * plane * (size_t)(S > 0 ? S : 1));
        REAL *ins = (REAL *)malloc(sizeof(REAL) * plane * (size_t)(d->steps > 0 ? d->steps : 1));
        double *myGr = (double *)calloc(plane * (size_t)(S > 0 ? S : 1), sizeof(double));
        double mygM[64];
        double mygW = 0.0;
        for (int q = 0; q < 64; ++q) mygM[q] = 0.0;
        if (!cur || !tmp || !g || !ds || !hist || !ins || !myGr) {
#pragma omp atomic write
            fail = 1;
        } else {
#pragma omp for schedule(static)
            for (int b = 0; b < d->B; ++b) {
                const REAL *u0 = u + plane * b, *go = gout + plane * b;
                memcpy(cur, u0, sizeof(REAL) * plane);
                FN(adi_run_sample)(d, &T, chan, cur, tmp, ds, hist, ins);
                if (d->skip) {
                    /* out = sig*u0 + (1-sig)*uF  (SVHN.py:74) */
                    double acc = 0.0;
                    for (size_t q = 0; q < plane; ++q) acc += (double)go[q] * ((double)u0[q] - (double)cur[q]);
                    mygW += acc * (double)sig * (1.0 - (double)sig);
                    const REAL om = R(1) - sig;
                    for (size_t q = 0; q < plane; ++q) g[q] = om * go[q];
                } else
                    memcpy(g, go, sizeof(REAL) * plane);
                for (int step = d->steps - 1; step >= 0; --step) {
                    if (d->chan_op == 2)
                        FN(chan_adjoint)(C, P, chan, hist + plane * (size_t)(step * sps + sps - 1), g, tmp, mygM);
                    for (int k = sps - 1; k >= 0; --k) {
                        const int s = step * sps + k;
                        for (int c = 0; c < C; ++c)
                            FN(sweep_plane_adj)(T.sw[s].axis, N, g + P * c, hist + plane * s + P * c,
                                                T.r + plane * s + P * c, T.den + plane * s + P * c,
                                                T.cst + plane * s + P * c, ds, myGr + plane * s + P * c);
                    }
                    if (d->chan_op == 1) FN(chan_adjoint)(C, P, chan, ins + plane * step, g, tmp, mygM);
                }
                if (gin) {
                    REAL *gi = gin + plane * b;
                    if (d->skip)
                        for (size_t q = 0; q < plane; ++q) {
                            REAL s0 = sig * go[q];
                            gi[q] = g[q] + s0;
                        }
                    else
                        memcpy(gi, g, sizeof(REAL) * plane);
                }
            }
#pragma omp critical
            {
                for (size_t q = 0; q < plane * (size_t)S; ++q) Gr[q] += myGr[q];
                for (int q = 0; q < C * C; ++q) gM[q] += mygM[q];
                gW += mygW;
            }
        }
        free(cur); free(tmp); free(g); free(ds); free(hist); free(ins); free(myGr);
    }
    /* dL/dr -> dL/d(smoothed coef) -> smoothing^T -> clamp mask -> base and time_coeff */
    for (size_t q = 0; q < plane; ++q) g_ab[q] = g_bb[q] = g_atc[q] = g_btc[q] = 0.0;
    double *gk = (double *)malloc(sizeof(double) * (size_t)N * 2);
    for (int s = 0; s < S && gk; ++s) {
        const FN(sweep) *sw = &T.sw[s];
        const double scale = (double)R(sw->dts) / (double)R(sw->h * sw->h);
        const double tt = (double)R(sw->t);
        double *gb = sw->axis == 0 ? g_ab : g_bb, *gt = sw->axis == 0 ? g_atc : g_btc;
        for (int c = 0; c < C; ++c)
            for (int l = 0; l < N; ++l) {
                size_t o = plane * s + ((size_t)c * N + l) * N;
                double *gs = gk, *gc = gk + N;
                for (int i = 0; i < N; ++i) { gs[i] = Gr[o + i] * scale; gc[i] = 0.0; }
                if (d->smooth) {
                    for (int i = 0; i < N; ++i) {
                        gc[i > 0 ? i - 1 : 0] += gs[i] / 3.0;
                        gc[i] += gs[i] / 3.0;
                        gc[i < N - 1 ? i + 1 : N - 1] += gs[i] / 3.0;
                    }
                } else
                    for (int i = 0; i < N; ++i) gc[i] = gs[i];
                for (int i = 0; i < N; ++i) {
                    if (!T.mask[o + i]) continue;
                    size_t q = FN(cell)(sw->axis, N, c, l, i);
                    gb[q] += gc[i];
                    gt[q] += gc[i] * tt;
                }
            }
    }
    if (!gk) fail = 1;
    free(gk);
    if (d->chan_op) for (int q = 0; q < C * C; ++q) g_chan[q] = gM[q];
    if (d->skip) g_skipw[0] = gW;
    free(Gr);
    FN(adi_tables_free)(&T);
    return fail ? -1 : 0;
}

/* ===================================================================================== */
/* emotion_recognition.PDELayer  (emotion_recognition.py:56-97)                          */
/* ===================================================================================== */

/* a_i = 0.5*dt*(w1 + w2 sin(2 pi y_i) + w3 sin(4 pi y_i))/dx^2   emotion_recognition.py:76-77
 * b_j =     dt*(w1 + w2 cos(2 pi x_j) + w3 cos(4 pi x_j))/dy^2   emotion_recognition.py:79-80
 * Also returns the basis values so the adjoint can project onto the six scalars. */
static void FN(emo_profiles)(const oracle_emo_desc *d, const REAL *w, const REAL *xs, const REAL *ys,
                             REAL *a, REAL *b, REAL *s1, REAL *s2, REAL *c1, REAL *c2) {
    const REAL two_pi = R(2.0 * 3.141592653589793), four_pi = R(4.0 * 3.141592653589793);
    const REAL ka = R(0.5 * d->dt), kb = R(d->dt), dx2 = R(d->dx * d->dx), dy2 = R(d->dy * d->dy);
    for (int i = 0; i < d->N; ++i) {
        REAL y2 = two_pi * ys[i], y4 = four_pi * ys[i];
        s1[i] = FN(r_sin)(y2); s2[i] = FN(r_sin)(y4);
        REAL t2 = w[1] * s1[i], t3 = w[2] * s2[i];
        REAL sum = w[0] + t2; sum = sum + t3;
        REAL sc = ka * sum;
        a[i] = sc / dx2;
        REAL x2 = two_pi * xs[i], x4 = four_pi * xs[i];
        c1[i] = FN(r_cos)(x2); c2[i] = FN(r_cos)(x4);
        REAL u2 = w[4] * c1[i], u3 = w[5] * c2[i];
        REAL sb = w[3] + u2; sb = sb + u3;
        REAL scb = kb * sb;
        b[i] = scb / dy2;
    }
}

/* One explicit step on the padded (N+2)x(N+2) buffer: interior from src to dst, ghost ring
 * untouched (it keeps the initial reflected values, emotion_recognition.py:85,91-95). */
static void FN(emo_step)(int N, const REAL *a, const REAL *b, const REAL *src, REAL *dst) {
    const int M = N + 2;
    for (int i = 1; i <= N; ++i)
        for (int j = 1; j <= N; ++j) {
            REAL uc = src[i * M + j];
            REAL tw = R(2) * uc;
            REAL dr = src[(i + 1) * M + j] - tw; dr = dr + src[(i - 1) * M + j];
            REAL dc = src[i * M + j + 1] - tw;   dc = dc + src[i * M + j - 1];
            REAL p1 = a[i - 1] * dr, p2 = b[j - 1] * dc;
            REAL acc = uc + p1;
            dst[i * M + j] = acc + p2;
        }
}

static void FN(emo_pad)(int N, const REAL *u0, REAL *pad) {
    const int M = N + 2;
    for (int i = 0; i < M; ++i) {
        int si = i == 0 ? 1 : (i == M - 1 ? N - 2 : i - 1);
        for (int j = 0; j < M; ++j) {
            int sj = j == 0 ? 1 : (j == M - 1 ? N - 2 : j - 1);
            pad[i * M + j] = u0[si * N + sj];
        }
    }
}

int FN(oracle_emotion_forward)(const oracle_emo_desc *d, const REAL *u0, const REAL *w,
                               const REAL *xs, const REAL *ys, REAL *out) {
    const int N = d->N, M = N + 2;
    if (N < 2 || d->Nt < 0) return -2;
    REAL *prof = (REAL *)malloc(sizeof(REAL) * (size_t)N * 6);
    if (!prof) return -1;
    REAL *a = prof, *b = prof + N;
    FN(emo_profiles)(d, w, xs, ys, a, b, prof + 2 * N, prof + 3 * N, prof + 4 * N, prof + 5 * N);
    const int nt = FN(pick_threads)(d->nthreads);
#pragma omp parallel num_threads(nt)
    {
        REAL *p0 = (REAL *)malloc(sizeof(REAL) * (size_t)M * M * 2);
        REAL *p1 = p0 + (size_t)M * M;
#pragma omp for schedule(static)
        for (int bb = 0; bb < d->B; ++bb) {
            FN(emo_pad)(N, u0 + (size_t)bb * N * N, p0);
            memcpy(p1, p0, sizeof(REAL) * (size_t)M * M);
            REAL *src = p0, *dst = p1;
            for (int k = 0; k < d->Nt; ++k) {
                FN(emo_step)(N, a, b, src, dst);
                REAL *t = src; src = dst; dst = t;
            }
            for (int i = 0; i < N; ++i)
                memcpy(out + (size_t)bb * N * N + (size_t)i * N, src + (i + 1) * M + 1, sizeof(REAL) * (size_t)N);
        }
        free(p0);
    }
    free(prof);
    return 0;
}

int FN(oracle_emotion_backward)(const oracle_emo_desc *d, const REAL *u0, const REAL *gout,
                                const REAL *w, const REAL *xs, const REAL *ys, REAL *gin,
                                double *gw) {
    const int N = d->N, M = N + 2, Nt = d->Nt;
    if (N < 2 || Nt < 0) return -2;
    REAL *prof = (REAL *)malloc(sizeof(REAL) * (size_t)N * 6);
    if (!prof) return -1;
    REAL *a = prof, *b = prof + N, *s1 = prof + 2 * N, *s2 = prof + 3 * N, *c1 = prof + 4 * N,
         *c2 = prof + 5 * N;
    FN(emo_profiles)(d, w, xs, ys, a, b, s1, s2, c1, c2);
    double *dA = (double *)calloc((size_t)N * 2, sizeof(double));
    double *dB = dA + N;
    const int nt = FN(pick_threads)(d->nthreads);
#pragma omp parallel num_threads(nt)
    {
        const size_t PP = (size_t)M * M;
        REAL *hist = (REAL *)malloc(sizeof(REAL) * PP * (size_t)(Nt + 1));
        REAL *lam = (REAL *)malloc(sizeof(REAL) * PP * 2);
        REAL *lam2 = lam + PP;
        double *mA = (double *)calloc((size_t)N * 2, sizeof(double));
        double *mB = mA + N;
#pragma omp for schedule(static)
        for (int bb = 0; bb < d->B; ++bb) {
            FN(emo_pad)(N, u0 + (size_t)bb * N * N, hist);
            for (int k = 0; k < Nt; ++k) {
                memcpy(hist + PP * (k + 1), hist + PP * k, sizeof(REAL) * PP);
                FN(emo_step)(N, a, b, hist + PP * k, hist + PP * (k + 1));
            }
            /* lam lives on the padded grid: interior = dL/du^k, ring = dL/d(ghost), the
             * ghost ring being a constant input of every step. */
            memset(lam, 0, sizeof(REAL) * PP);
            for (int i = 0; i < N; ++i)
                for (int j = 0; j < N; ++j) lam[(i + 1) * M + j + 1] = gout[(size_t)bb * N * N + (size_t)i * N + j];
            for (int k = Nt - 1; k >= 0; --k) {
                const REAL *uk = hist + PP * k;
                /* new interior adjoint; ring entries accumulate (they feed every step) */
                memcpy(lam2, lam, sizeof(REAL) * PP);
                for (int i = 1; i <= N; ++i)
                    for (int j = 1; j <= N; ++j) lam2[i * M + j] = R(0);
                for (int i = 1; i <= N; ++i)
                    for (int j = 1; j <= N; ++j) {
                        REAL l = lam[i * M + j];
                        REAL uc = uk[i * M + j], tw = R(2) * uc;
                        REAL dr = uk[(i + 1) * M + j] - tw; dr = dr + uk[(i - 1) * M + j];
                        REAL dc = uk[i * M + j + 1] - tw;   dc = dc + uk[i * M + j - 1];
                        mA[i - 1] += (double)l * (double)dr;
                        mB[j - 1] += (double)l * (double)dc;
                        REAL al = a[i - 1] * l, bl = b[j - 1] * l;
                        REAL self = (l - R(2) * al) - R(2) * bl;
                        lam2[i * M + j] += self;
                        lam2[(i + 1) * M + j] += al;
                        lam2[(i - 1) * M + j] += al;
                        lam2[i * M + j + 1] += bl;
                        lam2[i * M + j - 1] += bl;
                    }
                memcpy(lam, lam2, sizeof(REAL) * PP);
            }
            if (gin) {
                /* fold the ring back through the reflect padding (corners are never read) */
                REAL *gi = gin + (size_t)bb * N * N;
                for (int i = 0; i < N; ++i)
                    for (int j = 0; j < N; ++j) gi[(size_t)i * N + j] = lam[(i + 1) * M + j + 1];
                for (int j = 0; j < N; ++j) {
                    gi[(size_t)1 * N + j] += lam[0 * M + j + 1];
                    gi[(size_t)(N - 2) * N + j] += lam[(M - 1) * M + j + 1];
                }
                for (int i = 0; i < N; ++i) {
                    gi[(size_t)i * N + 1] += lam[(i + 1) * M + 0];
                    gi[(size_t)i * N + N - 2] += lam[(i + 1) * M + M - 1];
                }
            }
        }
#pragma omp critical
        for (int i = 0; i < N; ++i) { dA[i] += mA[i]; dB[i] += mB[i]; }
        free(hist); free(lam); free(mA);
    }
    const double ka = (double)R(0.5 * d->dt) / (double)R(d->dx * d->dx);
    const double kb = (double)R(d->dt) / (double)R(d->dy * d->dy);
    for (int q = 0; q < 6; ++q) gw[q] = 0.0;
    for (int i = 0; i < N; ++i) {
        gw[0] += ka * dA[i]; gw[1] += ka * dA[i] * (double)s1[i]; gw[2] += ka * dA[i] * (double)s2[i];
        gw[3] += kb * dB[i]; gw[4] += kb * dB[i] * (double)c1[i]; gw[5] += kb * dB[i] * (double)c2[i];
    }
    free(dA); free(prof);
    return 0;
}

/* ===================================================================================== */
/* tiny_imagenet.ImprovedDiffusionLayer, live path (tiny_imagenet.py:34-72)              */
/* ===================================================================================== */

static REAL FN(tiny_alpha)(const oracle_tiny_desc *d, REAL raw, int *m) {
    const REAL lo = R(d->cmin), hi = R(d->cmax);
    *m = raw >= lo && raw <= hi;
    return raw < lo ? lo : (raw > hi ? hi : raw);
}

/* lap0(v)(i,j) with zero ghosts: conv2d(padding=1) with [[0,1,0],[1,-4,1],[0,1,0]] */
static inline REAL FN(lap0)(const REAL *v, int H, int W, int i, int j) {
    REAL up = i > 0 ? v[(size_t)(i - 1) * W + j] : R(0);
    REAL lf = j > 0 ? v[(size_t)i * W + j - 1] : R(0);
    REAL ct = v[(size_t)i * W + j];
    REAL rt = j < W - 1 ? v[(size_t)i * W + j + 1] : R(0);
    REAL dn = i < H - 1 ? v[(size_t)(i + 1) * W + j] : R(0);
    REAL acc = up + lf;
    acc = acc + R(-4) * ct;
    acc = acc + rt;
    return acc + dn;
}

static void FN(tiny_step)(const oracle_tiny_desc *d, REAL adt, REAL sc, const REAL *u, REAL *v, REAL *un) {
    const int H = d->H, W = d->W;
    const REAL bl = R(d->blend);
    for (size_t q = 0; q < (size_t)H * W; ++q) v[q] = u[q] * sc;
    for (int i = 0; i < H; ++i)
        for (int j = 0; j < W; ++j) {
            size_t q = (size_t)i * W + j;
            REAL lp = adt * FN(lap0)(v, H, W, i, j);
            REAL wv = v[q] + lp;
            REAL df = wv - u[q];
            REAL sd = bl * df;
            un[q] = u[q] + sd;
        }
}

int FN(oracle_tiny_forward)(const oracle_tiny_desc *d, const REAL *u, const REAL *alpha_base,
                            const REAL *scaling, REAL *out) {
    const size_t P = (size_t)d->H * d->W;
    if (d->H < 1 || d->W < 1 || d->steps < 0) return -2;
    const int nt = FN(pick_threads)(d->nthreads);
#pragma omp parallel num_threads(nt)
    {
        REAL *buf = (REAL *)malloc(sizeof(REAL) * P * 3);
#pragma omp for schedule(static)
        for (int bc = 0; bc < d->B * d->C; ++bc) {
            const int c = bc % d->C;
            int m;
            REAL al = FN(tiny_alpha)(d, alpha_base[c], &m);
            REAL adt = al * R(d->dt);
            REAL *cur = buf, *v = buf + P, *nx = buf + 2 * P;
            memcpy(cur, u + P * bc, sizeof(REAL) * P);
            for (int k = 0; k < d->steps; ++k) {
                FN(tiny_step)(d, adt, scaling[c], cur, v, nx);
                REAL *t = cur; cur = nx; nx = t;
            }
            memcpy(out + P * bc, cur, sizeof(REAL) * P);
        }
        free(buf);
    }
    return 0;
}

int FN(oracle_tiny_backward)(const oracle_tiny_desc *d, const REAL *u, const REAL *gout,
                             const REAL *alpha_base, const REAL *scaling, REAL *gin,
                             double *g_alpha, double *g_scaling) {
    const int H = d->H, W = d->W, C = d->C, K = d->steps;
    const size_t P = (size_t)H * W;
    if (H < 1 || W < 1 || K < 0 || C > 4096) return -2;
    double *ga = (double *)calloc((size_t)C * 2, sizeof(double));
    double *gs = ga + C;
    const int nt = FN(pick_threads)(d->nthreads);
    const REAL bl = R(d->blend);
#pragma omp parallel num_threads(nt)
    {
        REAL *hist = (REAL *)malloc(sizeof(REAL) * P * (size_t)(K + 1));
        REAL *buf = (REAL *)malloc(sizeof(REAL) * P * 4);
        double *mga = (double *)calloc((size_t)C * 2, sizeof(double));
        double *mgs = mga + C;
#pragma omp for schedule(static)
        for (int bc = 0; bc < d->B * d->C; ++bc) {
            const int c = bc % C;
            int m;
            REAL al = FN(tiny_alpha)(d, alpha_base[c], &m);
            REAL adt = al * R(d->dt);
            REAL sc = scaling[c];
            REAL *v = buf, *g = buf + P, *gwv = buf + 2 * P, *z = buf + 3 * P;
            memcpy(hist, u + P * bc, sizeof(REAL) * P);
            for (int k = 0; k < K; ++k) FN(tiny_step)(d, adt, sc, hist + P * k, v, hist + P * (k + 1));
            memcpy(g, gout + P * bc, sizeof(REAL) * P);
            for (int k = K - 1; k >= 0; --k) {
                const REAL *uk = hist + P * k;
                /* u' = u + bl*(w - u), w = v + adt*lap0(v), v = sc*u */
                for (size_t q = 0; q < P; ++q) { v[q] = uk[q] * sc; gwv[q] = bl * g[q]; }
                double accA = 0.0, accS = 0.0;
                for (int i = 0; i < H; ++i)
                    for (int j = 0; j < W; ++j) {
                        size_t q = (size_t)i * W + j;
                        accA += (double)gwv[q] * (double)FN(lap0)(v, H, W, i, j);
                        REAL lg = adt * FN(lap0)(gwv, H, W, i, j);   /* lap0 is self-adjoint */
                        z[q] = gwv[q] + lg;                           /* dL/dv */
                        accS += (double)z[q] * (double)uk[q];
                    }
                if (m) mga[c] += accA * (double)R(d->dt);
                mgs[c] += accS;
                for (size_t q = 0; q < P; ++q) {
                    REAL keep = g[q] - gwv[q];
                    REAL back = sc * z[q];
                    g[q] = keep + back;
                }
            }
            if (gin) memcpy(gin + P * bc, g, sizeof(REAL) * P);
        }
#pragma omp critical
        for (int c = 0; c < C; ++c) { ga[c] += mga[c]; gs[c] += mgs[c]; }
        free(hist); free(buf); free(mga);
    }
    for (int c = 0; c < C; ++c) { g_alpha[c] = ga[c]; g_scaling[c] = gs[c]; }
    free(ga);
    return 0;
}

/* ===================================================================================== */
/* tiny_imagenet.ImprovedDiffusionLayer, dormant methods (tiny_imagenet.py:88-233):      */
/* scalar-coefficient ADI step with clamp(denom) pivots, and the explicit x / y splits.  */
/* Nobody in the reference calls them (use_implicit is stored at :21 and never read);    */
/* they are restated for whoever does.  All of them are linear in u with a constant      */
/* (Python float) coefficient, so the only gradient is the one with respect to u.        */
/* ===================================================================================== */

/* thomas_algorithm_batch (tiny_imagenet.py:159-190) for one line of n cells with stride st, the
 * constant bands of solve_implicit_x / _y (:107-118, :141-152): a = c = -r, b = 1 + 2r, b_0 =
 * b_{n-1} = 1 + r, every one a Python double rounded to the tensor dtype by torch.full. */
static void FN(tiny_thomas_line)(int n, REAL am, REAL bmid, REAL bend, REAL eps, const REAL *d, size_t st, REAL *x,
                                 REAL *cp, REAL *dp) {
    /* first row (:170-171): no clamp on b_0 */
    cp[0] = am / bend;
    dp[0] = d[0] / bend;
    for (int i = 1; i < n; ++i) {
        const REAL b = i == n - 1 ? bend : bmid;
        REAL prod = am * cp[i - 1];
        REAL denom = b - prod;                       /* :175 */
        denom = denom < eps ? eps : denom;           /* :176 clamp(min=stability_eps) */
        cp[i] = i < n - 1 ? am / denom : R(0);       /* :178-179 */
        REAL pd = am * dp[i - 1];
        REAL num = d[(size_t)i * st] - pd;
        dp[i] = num / denom;                         /* :180 */
    }
    x[(size_t)(n - 1) * st] = dp[n - 1];             /* :184 */
    for (int i = n - 2; i >= 0; --i) {
        REAL pr = cp[i] * x[(size_t)(i + 1) * st];
        x[(size_t)i * st] = dp[i] - pr;              /* :187 */
    }
}

/* diffuse_x_explicit / diffuse_y_explicit (tiny_imagenet.py:199-233) for one line: k = coeff * dt. */
static void FN(tiny_explicit_line)(int n, REAL k, const REAL *u, size_t st, REAL *out) {
    for (int i = 0; i < n; ++i) out[(size_t)i * st] = u[(size_t)i * st];
    if (n > 2)
        for (int i = 1; i < n - 1; ++i) {
            REAL two = R(2) * u[(size_t)i * st];
            REAL acc = u[(size_t)(i - 1) * st] - two;
            acc = acc + u[(size_t)(i + 1) * st];
            REAL sc = k * acc;
            out[(size_t)i * st] = u[(size_t)i * st] + sc;
        }
    if (n > 1) {
        REAL d0 = u[st] - u[0];
        REAL s0 = k * d0;
        out[0] = u[0] + s0;
        REAL d1 = u[(size_t)(n - 2) * st] - u[(size_t)(n - 1) * st];
        REAL s1 = k * d1;
        out[(size_t)(n - 1) * st] = u[(size_t)(n - 1) * st] + s1;
    } else {
        /* W == 1: u[:, :, 1] does not exist in the reference (IndexError); keep the value */
    }
}

/* mode 0: implicit_diffusion_step (x solve with rx, then y solve with ry; :88-102)
 *      1: solve_implicit_x   2: solve_implicit_y   3: diffuse_x_explicit   4: diffuse_y_explicit
 * u, out: [B][H][W] planes.  For modes 0-2 cx / cy = {a, b_mid, b_end} already rounded to REAL;
 * for modes 3-4 cx[0] / cy[0] = coeff * dt rounded to REAL.  All five maps are self-adjoint,
 * so the gradient with respect to u is the same call on the upstream gradient. */
int FN(oracle_tiny_split)(int B, int H, int W, int mode, const REAL *cx, const REAL *cy, REAL eps, int nthreads,
                          const REAL *u, REAL *out) {
    if (H < 1 || W < 1 || H > 4096 || W > 4096 || mode < 0 || mode > 4) return -2;
    const size_t P = (size_t)H * W;
    const int nt = FN(pick_threads)(nthreads);
#pragma omp parallel num_threads(nt)
    {
        const int n = H > W ? H : W;
        REAL *cp = (REAL *)malloc(sizeof(REAL) * (size_t)n * 2);
        REAL *dp = cp + n;
        REAL *tmp = (REAL *)malloc(sizeof(REAL) * P);
#pragma omp for schedule(static)
        for (int b = 0; b < B; ++b) {
            const REAL *src = u + P * b;
            REAL *dst = out + P * b;
            if (mode == 0 || mode == 1) {
                REAL *xo = mode == 0 ? tmp : dst;
                for (int h = 0; h < H; ++h)
                    FN(tiny_thomas_line)(W, cx[0], cx[1], cx[2], eps, src + (size_t)h * W, 1, xo + (size_t)h * W, cp, dp);
                src = xo;
            }
            if (mode == 0 || mode == 2)
                for (int w = 0; w < W; ++w) FN(tiny_thomas_line)(H, cy[0], cy[1], cy[2], eps, src + w, (size_t)W, dst + w, cp, dp);
            if (mode == 3)
                for (int h = 0; h < H; ++h) FN(tiny_explicit_line)(W, cx[0], src + (size_t)h * W, 1, dst + (size_t)h * W);
            if (mode == 4)
                for (int w = 0; w < W; ++w) FN(tiny_explicit_line)(H, cy[0], src + w, (size_t)W, dst + w);
        }
        free(cp);
        free(tmp);
    }
    return 0;
}
