// rmsb_solver.cuh -- one-warp-per-reactor variable-order (1..5) variable-step BDF/NDF integrator.
//
// Role in the reference: `solve(react.ode, CVODE_BDF(); abstol, reltol)` (src/Reactor.jl:76-81; call sites
// src/TestReactors.jl:22,197,305).  CVODE itself is a third-party C library that is NOT in the reference
// tree; what the reference pins is only the converged solution (src/TestReactors.jl golden values), so the
// integrator here is an independent implementation of the quasi-constant-step NDF/BDF family of
// Shampine & Reichelt, "The MATLAB ODE Suite", SIAM J. Sci. Comput. 18 (1997) -- backward differences
// D[0..order+2], modified Newton on (I - c*J) with Jacobian reuse, WRMS error control with
// scale_i = atol + rtol*|y_i| (the same weights CVODE uses), order selection every order+1 equal steps.
// The same algorithm is available on the CPU as scipy.integrate.BDF, which the tests use as an
// independent cross-check of step counts and solution values.
//
// Data placement: state vectors (y, D[0..7], scale, psi, d) are distributed over the warp's registers;
// J and the LU factors of I - c*J live in the warp's shared-memory slice (column-major, odd leading
// dimension); HBM is touched only for the initial state, the saved outputs and the final statistics.
#pragma once
#include "rmsb_device.cuh"

namespace rmsb {

constexpr int BDF_MAX_ORDER = 5;
constexpr int NEWTON_MAXITER = 4;
constexpr double BDF_MIN_FACTOR = 0.2;
constexpr double BDF_MAX_FACTOR = 10.0;

struct SolveArgs {
    int64_t n;
    const double* y0;
    const double* cond;
    int ncond;
    const double* p;
    int64_t p_stride;
    const int* constspc;
    int nconst;
    const double* t_end;
    double rtol, atol;
    int nsave;
    const double* tsave;
    double* ysave;
    double* yend;
    double* idt;
    int* status;
    rmsb_stats* stats;
    int max_steps;
    int max_order;
    int idt_index;
    double first_step;
    unsigned long long* counter;
    int domain_kind;
    int Ny;
};

// ---- distributed vector helpers -----------------------------------------------------------------
template <int R>
__device__ __forceinline__ double rms_norm_scaled(const double (&x)[R], const double (&scale)[R], int Ny, int lane) {
    double a = 0.0;
#pragma unroll
    for (int r = 0; r < R; ++r) {
        int i = r * 32 + lane;
        if (i < Ny) { double q = x[r] / scale[r]; a += q * q; }
    }
    return sqrt(warp_sum(a) / (double)Ny);
}

template <int R>
__device__ __forceinline__ bool all_finite(const double (&x)[R], int Ny, int lane) {
    bool ok = true;
#pragma unroll
    for (int r = 0; r < R; ++r) { int i = r * 32 + lane; if (i < Ny && !isfinite(x[r])) ok = false; }
    return __all_sync(FULL, ok);
}

// ---- dense LU with partial pivoting of A = I - c*J in shared memory ------------------------------
// Rows are owned by lanes (row i -> lane i&31), columns are walked sequentially; row k is broadcast.
// Returns false if a zero pivot is met.
template <int R>
__device__ inline bool lu_factor_warp(WarpWs& w, int Ny, double c, int lane) {
    const int LD = w.LD;
    double* A = w.LU;
    for (int idx = lane; idx < Ny * LD; idx += 32) {
        int col = idx / LD, row = idx - col * LD;
        if (row < Ny) A[idx] = ((row == col) ? 1.0 : 0.0) - c * w.J[idx];
    }
    for (int i = lane; i < Ny; i += 32) w.perm[i] = i;
    bool ok = true;
    for (int k = 0; k < Ny; ++k) {
        __syncwarp();
        // pivot search on the high word of |a_ik| (ordering of non-negative doubles == ordering of their bits;
        // ties within 2^-20 relative are equally good pivots)
        unsigned best = 0; int bi = -1;
#pragma unroll
        for (int r = 0; r < R; ++r) {
            int i = r * 32 + lane;
            if (i >= k && i < Ny) {
                unsigned key = (unsigned)__double2hiint(fabs(A[i + k * LD])) + 1u;
                if (key > best || bi < 0) { best = key; bi = i; }
            }
        }
        unsigned mx = __reduce_max_sync(FULL, best);
        unsigned vote = __ballot_sync(FULL, best == mx && bi >= 0);
        int src = __ffs(vote) - 1;
        int pr = __shfl_sync(FULL, bi, src);
        if (pr != k) {
            for (int j = lane; j < Ny; j += 32) {
                double t = A[k + j * LD]; A[k + j * LD] = A[pr + j * LD]; A[pr + j * LD] = t;
            }
            if (lane == 0) { int t = w.perm[k]; w.perm[k] = w.perm[pr]; w.perm[pr] = t; }
            __syncwarp();
        }
        double piv = A[k + k * LD];
        if (piv == 0.0 || !isfinite(piv)) { ok = false; piv = 1.0; }
        double inv = 1.0 / piv;
        if (lane == 0) w.dinv[k] = inv;
        double l[R];
#pragma unroll
        for (int r = 0; r < R; ++r) {
            int i = r * 32 + lane;
            l[r] = 0.0;
            if (i > k && i < Ny) { l[r] = A[i + k * LD] * inv; A[i + k * LD] = l[r]; }
        }
        for (int j = k + 1; j < Ny; ++j) {
            double akj = A[k + j * LD];
#pragma unroll
            for (int r = 0; r < R; ++r) {
                int i = r * 32 + lane;
                if (i > k && i < Ny) A[i + j * LD] -= l[r] * akj;
            }
        }
    }
    __syncwarp();
    return ok;
}

// Solve (I - cJ) x = b with the factors above; b distributed in registers, overwritten with x.
template <int R>
__device__ inline void lu_solve_warp(const WarpWs& w, int Ny, double (&b)[R], int lane) {
    const int LD = w.LD;
    const double* A = w.LU;
    // apply the row permutation through the warp's scratch vector
#pragma unroll
    for (int r = 0; r < R; ++r) { int i = r * 32 + lane; if (i < Ny) w.xs[i] = b[r]; }
    __syncwarp();
#pragma unroll
    for (int r = 0; r < R; ++r) { int i = r * 32 + lane; if (i < Ny) b[r] = w.xs[w.perm[i]]; }
    // forward substitution (unit lower)
#pragma unroll
    for (int kr = 0; kr < R; ++kr) {
        for (int kl = 0; kl < 32; ++kl) {
            int k = kr * 32 + kl;
            if (k >= Ny - 1) break;
            double xk = __shfl_sync(FULL, b[kr], kl);
#pragma unroll
            for (int r = 0; r < R; ++r) {
                int i = r * 32 + lane;
                if (i > k && i < Ny) b[r] -= A[i + k * LD] * xk;
            }
        }
    }
    // backward substitution
#pragma unroll
    for (int kr = R - 1; kr >= 0; --kr) {
        for (int kl = 31; kl >= 0; --kl) {
            int k = kr * 32 + kl;
            if (k >= Ny) continue;
            if (lane == kl) b[kr] *= w.dinv[k];
            double xk = __shfl_sync(FULL, b[kr], kl);
#pragma unroll
            for (int r = 0; r < R; ++r) {
                int i = r * 32 + lane;
                if (i < k) b[r] -= A[i + k * LD] * xk;
            }
        }
    }
    __syncwarp();
}

// ---- change_D: rescale the backward differences for a step-size ratio `factor` -------------------
// D_new[j] = sum_i (R(factor) U)[i][j] D[i], R[i][k] = prod_{l=1..i} (l-1-factor*k)/l, U = R(1).
template <int R>
__device__ inline void change_D(double (&D)[BDF_MAX_ORDER + 3][R], int order, double factor) {
    double Rm[BDF_MAX_ORDER + 1][BDF_MAX_ORDER + 1], Um[BDF_MAX_ORDER + 1][BDF_MAX_ORDER + 1];
#pragma unroll
    for (int k = 0; k <= BDF_MAX_ORDER; ++k) { Rm[0][k] = 1.0; Um[0][k] = 1.0; }
#pragma unroll
    for (int i = 1; i <= BDF_MAX_ORDER; ++i) {
#pragma unroll
        for (int k = 0; k <= BDF_MAX_ORDER; ++k) {
            if (k == 0) { Rm[i][k] = 0.0; Um[i][k] = 0.0; }
            else {
                Rm[i][k] = Rm[i - 1][k] * (((double)(i - 1) - factor * (double)k) / (double)i);
                Um[i][k] = Um[i - 1][k] * (((double)(i - 1) - (double)k) / (double)i);
            }
        }
    }
    double Dn[BDF_MAX_ORDER + 1][R];
#pragma unroll
    for (int j = 1; j <= BDF_MAX_ORDER; ++j) {
        double ru[BDF_MAX_ORDER + 1];
#pragma unroll
        for (int i = 0; i <= BDF_MAX_ORDER; ++i) {
            double a = 0.0;
#pragma unroll
            for (int k = 0; k <= BDF_MAX_ORDER; ++k) if (k <= j) a += Rm[i][k] * Um[k][j];
            ru[i] = a;
        }
#pragma unroll
        for (int r = 0; r < R; ++r) {
            double a = 0.0;
#pragma unroll
            for (int i = 0; i <= BDF_MAX_ORDER; ++i) if (i <= order) a += ru[i] * D[i][r];
            Dn[j][r] = a;
        }
    }
#pragma unroll
    for (int j = 1; j <= BDF_MAX_ORDER; ++j)
        if (j <= order) {
#pragma unroll
            for (int r = 0; r < R; ++r) D[j][r] = Dn[j][r];
        }
}

// ---- the integrator ---------------------------------------------------------------------------
template <int R>
__device__ void bdf_solve_warp(const DevMech& m, const SolveArgs& a, WarpWs& w, int64_t ir, int lane) {
    const int Ny = a.Ny;
    DomainArgs d;
    d.kind = a.domain_kind; d.Ny = Ny;
    d.p = a.p ? a.p + ir * a.p_stride : nullptr;
    d.constspc = a.constspc; d.nconst = a.nconst;
    const double* cond = a.cond + ir * a.ncond;
    const double rtol = a.rtol, atol = a.atol;
    const double t_end = a.t_end[ir];

    // BDF constants (Shampine-Reichelt NDF kappa; gamma_k = sum_{j<=k} 1/j)
    const double kappa[BDF_MAX_ORDER + 1] = {0.0, -0.1850, -1.0 / 9.0, -0.0823, -0.0415, 0.0};
    double gamma[BDF_MAX_ORDER + 1], alpha[BDF_MAX_ORDER + 1], error_const[BDF_MAX_ORDER + 1];
    gamma[0] = 0.0;
#pragma unroll
    for (int k = 1; k <= BDF_MAX_ORDER; ++k) gamma[k] = gamma[k - 1] + 1.0 / (double)k;
#pragma unroll
    for (int k = 0; k <= BDF_MAX_ORDER; ++k) {
        alpha[k] = (1.0 - kappa[k]) * gamma[k];
        error_const[k] = kappa[k] * gamma[k] + 1.0 / (double)(k + 1);
    }

    double D[BDF_MAX_ORDER + 3][R];
#pragma unroll
    for (int j = 0; j < BDF_MAX_ORDER + 3; ++j)
#pragma unroll
        for (int r = 0; r < R; ++r) D[j][r] = 0.0;
    double y[R], f[R], scale[R];
#pragma unroll
    for (int r = 0; r < R; ++r) {
        int i = r * 32 + lane;
        y[r] = (i < Ny) ? a.y0[ir * Ny + i] : 0.0;
    }
    int n_rhs = 0, n_jac = 0, n_lu = 0, n_steps = 0, n_nfail = 0, n_efail = 0;
    int status = RMSB_SOLVE_OK;

    // constant-k domains: rate coefficients once per reactor
    if (d.kind != RMSB_DOMAIN_CONSTANT_V) {
#pragma unroll
        for (int r = 0; r < R; ++r) { int s = r * 32 + lane; if (s < m.Ns) w.ns[s] = y[r]; }
        __syncwarp();
        double T = cond[0], P, V;
        if (d.kind == RMSB_DOMAIN_CONSTANT_TP) { P = cond[1]; V = bcast<double, R>(y, m.Ns); }
        else { P = (d.kind == RMSB_DOMAIN_CONSTANT_TV) ? 1.0e8 : 0.0; V = cond[1]; }
        setup_constk(m, d, w, T, P, V, lane);
    }

    // output bookkeeping
    int isave = 0;
    const int idt_index = a.idt_index;
    double idt_best = -INFINITY, idt_time = NAN;
    double* ysave = a.ysave ? a.ysave + (size_t)ir * a.nsave * Ny : nullptr;
    while (isave < a.nsave && a.tsave[isave] <= 0.0) {
        if (ysave) {
#pragma unroll
            for (int r = 0; r < R; ++r) { int i = r * 32 + lane; if (i < Ny) ysave[(size_t)isave * Ny + i] = y[r]; }
        }
        ++isave;
    }

    // f0 and initial Jacobian
    rhs_warp<double, R>(m, d, w, cond, y, f, lane, false, nullptr); ++n_rhs;
    double t = 0.0;
    double h_abs;
    bool finite_ok = all_finite<R>(f, Ny, lane);
    if (!finite_ok) status = RMSB_SOLVE_NAN;
    if (a.first_step > 0.0) h_abs = a.first_step;
    else {
        // select_initial_step (Hairer-Norsett-Wanner II.4), order = 1
#pragma unroll
        for (int r = 0; r < R; ++r) scale[r] = atol + fabs(y[r]) * rtol;
        double d0 = rms_norm_scaled<R>(y, scale, Ny, lane);
        double d1 = rms_norm_scaled<R>(f, scale, Ny, lane);
        double h0 = (d0 < 1e-5 || d1 < 1e-5) ? 1e-6 : 0.01 * d0 / d1;
        double y1[R], f1[R], df[R];
#pragma unroll
        for (int r = 0; r < R; ++r) y1[r] = y[r] + h0 * f[r];
        rhs_warp<double, R>(m, d, w, cond, y1, f1, lane, false, nullptr); ++n_rhs;
#pragma unroll
        for (int r = 0; r < R; ++r) df[r] = f1[r] - f[r];
        double d2 = rms_norm_scaled<R>(df, scale, Ny, lane) / h0;
        double h1 = (d1 <= 1e-15 && d2 <= 1e-15) ? fmax(1e-6, h0 * 1e-3) : pow(0.01 / fmax(d1, d2), 0.5);
        h_abs = fmin(fmin(100.0 * h0, h1), t_end);
    }
    jac_warp<R>(m, d, w, cond, y, lane); ++n_jac;
    bool lu_valid = false;
#pragma unroll
    for (int r = 0; r < R; ++r) { D[0][r] = y[r]; D[1][r] = f[r] * h_abs; }
    int order = 1, n_equal = 0;
    const int max_order = a.max_order;
    const double newton_tol = fmax(10.0 * 2.220446049250313e-16 / rtol, fmin(0.03, sqrt(rtol)));

    while (status == RMSB_SOLVE_OK && t < t_end) {
        if (n_steps >= a.max_steps) { status = RMSB_SOLVE_MAX_STEPS; break; }
        const double min_step = 10.0 * (__longlong_as_double(__double_as_longlong(t) + 1) - t);
        if (h_abs < min_step) {
            change_D<R>(D, order, min_step / h_abs);
            h_abs = min_step; n_equal = 0; lu_valid = false;
        }
        bool accepted = false, current_jac = false;
        double t_new = t, h = 0.0, error_norm = 0.0, safety = 0.9;
        double dvec[R], ynew[R];
        while (!accepted) {
            if (h_abs < min_step) { status = RMSB_SOLVE_STEP_TOO_SMALL; break; }
            t_new = t + h_abs;
            if (t_new - t_end > 0.0) {
                t_new = t_end;
                change_D<R>(D, order, fabs(t_new - t) / h_abs);
                n_equal = 0; lu_valid = false;
            }
            h = t_new - t; h_abs = fabs(h);
            double ypred[R], psi[R];
#pragma unroll
            for (int r = 0; r < R; ++r) {
                double s = 0.0, ps = 0.0;
#pragma unroll
                for (int j = 0; j <= BDF_MAX_ORDER; ++j) if (j <= order) s += D[j][r];
#pragma unroll
                for (int j = 1; j <= BDF_MAX_ORDER; ++j) if (j <= order) ps += D[j][r] * gamma[j];
                ypred[r] = s;
                psi[r] = ps / alpha[order];
                scale[r] = atol + rtol * fabs(s);
            }
            const double c = h / alpha[order];
            bool converged = false;
            int n_iter = 0;
            while (!converged) {
                bool lu_ok = true;
                if (!lu_valid) { lu_ok = lu_factor_warp<R>(w, Ny, c, lane); lu_valid = true; ++n_lu; }
                // ---- solve_bdf_system: simplified Newton on y = ypred + d, c*f(y) - psi - d = 0
#pragma unroll
                for (int r = 0; r < R; ++r) { dvec[r] = 0.0; ynew[r] = ypred[r]; }
                double dy_norm_old = -1.0;
                if (lu_ok) {
                    for (int k = 0; k < NEWTON_MAXITER; ++k) {
                        n_iter = k + 1;
                        rhs_warp<double, R>(m, d, w, cond, ynew, f, lane, false, nullptr); ++n_rhs;
                        if (!all_finite<R>(f, Ny, lane)) break;
                        double dy[R];
#pragma unroll
                        for (int r = 0; r < R; ++r) dy[r] = c * f[r] - psi[r] - dvec[r];
                        lu_solve_warp<R>(w, Ny, dy, lane);
                        double dy_norm = rms_norm_scaled<R>(dy, scale, Ny, lane);
                        double rate = (dy_norm_old < 0.0) ? -1.0 : dy_norm / dy_norm_old;
                        if (!(dy_norm == dy_norm)) break;   // NaN
                        if (rate >= 0.0 && (rate >= 1.0 || pow(rate, (double)(NEWTON_MAXITER - k)) / (1.0 - rate) * dy_norm > newton_tol)) break;
#pragma unroll
                        for (int r = 0; r < R; ++r) { ynew[r] += dy[r]; dvec[r] += dy[r]; }
                        if (dy_norm == 0.0 || (rate >= 0.0 && rate / (1.0 - rate) * dy_norm < newton_tol)) { converged = true; break; }
                        dy_norm_old = dy_norm;
                    }
                }
                if (!converged) {
                    if (current_jac) break;
                    jac_warp<R>(m, d, w, cond, ypred, lane); ++n_jac;
                    lu_valid = false; current_jac = true;
                }
            }
            if (!converged) {
                ++n_nfail;
                h_abs *= 0.5;
                change_D<R>(D, order, 0.5);
                n_equal = 0; lu_valid = false;
                continue;
            }
            safety = 0.9 * (2.0 * NEWTON_MAXITER + 1.0) / (2.0 * NEWTON_MAXITER + (double)n_iter);
            double err[R];
#pragma unroll
            for (int r = 0; r < R; ++r) { scale[r] = atol + rtol * fabs(ynew[r]); err[r] = error_const[order] * dvec[r]; }
            error_norm = rms_norm_scaled<R>(err, scale, Ny, lane);
            if (error_norm > 1.0) {
                ++n_efail;
                double factor = fmax(BDF_MIN_FACTOR, safety * pow(error_norm, -1.0 / (double)(order + 1)));
                h_abs *= factor;
                change_D<R>(D, order, factor);
                n_equal = 0; lu_valid = false;
            } else accepted = true;
        }
        if (!accepted) break;
        ++n_equal; ++n_steps;
        const double t_old = t;
        t = t_new;
        // D[order+2] = d - D[order+1]; D[order+1] = d; D[i] += D[i+1] for i = order..0
#pragma unroll
        for (int r = 0; r < R; ++r) {
#pragma unroll
            for (int j = 2; j < BDF_MAX_ORDER + 3; ++j) if (j == order + 2) D[j][r] = dvec[r] - D[j - 1][r];
#pragma unroll
            for (int j = 1; j < BDF_MAX_ORDER + 3; ++j) if (j == order + 1) D[j][r] = dvec[r];
#pragma unroll
            for (int j = BDF_MAX_ORDER; j >= 0; --j) if (j <= order) D[j][r] += D[j + 1][r];
        }
        // ---- dense output on (t_old, t]: y(ts) = D0 + sum_j D_j prod_{m<j} (ts - (t - m h))/(h (1+m))
        while (isave < a.nsave && a.tsave[isave] <= t) {
            double ts = a.tsave[isave];
            if (ysave) {
#pragma unroll
                for (int r = 0; r < R; ++r) {
                    double pj = 1.0, acc = D[0][r];
#pragma unroll
                    for (int j = 1; j <= BDF_MAX_ORDER; ++j) {
                        if (j <= order) {
                            pj *= (ts - (t - (double)(j - 1) * h)) / (h * (double)j);
                            acc += D[j][r] * pj;
                        }
                    }
                    int i = r * 32 + lane;
                    if (i < Ny) ysave[(size_t)isave * Ny + i] = acc;
                }
            }
            ++isave;
        }
        // ---- ignition-delay tracker: max over the step of d/dt of component idt_index, 32 samples per step
        if (idt_index >= 0) {
            double Dk[BDF_MAX_ORDER + 1];
#pragma unroll
            for (int j = 0; j <= BDF_MAX_ORDER; ++j) Dk[j] = bcast<double, R>(D[j], idt_index);
            double ts = t_old + h * (double)(lane + 1) * (1.0 / 32.0);
            double pj = 1.0, dpj = 0.0, der = 0.0;
#pragma unroll
            for (int j = 1; j <= BDF_MAX_ORDER; ++j) {
                if (j <= order) {
                    double den = h * (double)j;
                    double x = (ts - (t - (double)(j - 1) * h)) / den;
                    dpj = dpj * x + pj / den;
                    pj = pj * x;
                    der += Dk[j] * dpj;
                }
            }
            double best = der; double bt = ts;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                double ob = __shfl_xor_sync(FULL, best, o), ot = __shfl_xor_sync(FULL, bt, o);
                if (ob > best || (ob == best && ot < bt)) { best = ob; bt = ot; }
            }
            if (best > idt_best) { idt_best = best; idt_time = bt; }
        }
#pragma unroll
        for (int r = 0; r < R; ++r) y[r] = D[0][r];
        if (n_equal < order + 1) continue;
        // ---- order / step-size selection
        double e_m = INFINITY, e_p = INFINITY;
        double tmp[R];
        if (order > 1) {
#pragma unroll
            for (int r = 0; r < R; ++r) {
                double v = 0.0;
#pragma unroll
                for (int j = 0; j < BDF_MAX_ORDER + 3; ++j) if (j == order) v = D[j][r];
                tmp[r] = error_const[order - 1] * v;
            }
            e_m = rms_norm_scaled<R>(tmp, scale, Ny, lane);
        }
        if (order < max_order) {
#pragma unroll
            for (int r = 0; r < R; ++r) {
                double v = 0.0;
#pragma unroll
                for (int j = 0; j < BDF_MAX_ORDER + 3; ++j) if (j == order + 2) v = D[j][r];
                tmp[r] = error_const[order + 1] * v;
            }
            e_p = rms_norm_scaled<R>(tmp, scale, Ny, lane);
        }
        double fm = pow(e_m, -1.0 / (double)order);
        double f0 = pow(error_norm, -1.0 / (double)(order + 1));
        double fp = pow(e_p, -1.0 / (double)(order + 2));
        int delta; double fmax_;   // np.argmax([fm, f0, fp]) - 1: the first maximum wins
        if (fm >= f0 && fm >= fp) { delta = -1; fmax_ = fm; }
        else if (f0 >= fp) { delta = 0; fmax_ = f0; }
        else { delta = 1; fmax_ = fp; }
        order += delta;
        double factor = fmin(BDF_MAX_FACTOR, safety * fmax_);
        h_abs *= factor;
        change_D<R>(D, order, factor);
        n_equal = 0; lu_valid = false;
    }

    // ---- outputs
    if (a.yend) {
#pragma unroll
        for (int r = 0; r < R; ++r) { int i = r * 32 + lane; if (i < Ny) a.yend[ir * Ny + i] = y[r]; }
    }
    if (ysave) {
        for (; isave < a.nsave; ++isave)
            for (int i = lane; i < Ny; i += 32) ysave[(size_t)isave * Ny + i] = NAN;
    }
    if (lane == 0) {
        if (a.idt) a.idt[ir] = (idt_index >= 0) ? idt_time : NAN;
        a.status[ir] = status;
        if (a.stats) {
            rmsb_stats s;
            s.n_steps = n_steps; s.n_rhs = n_rhs; s.n_jac = n_jac; s.n_lu = n_lu;
            s.n_newton_fail = n_nfail; s.n_err_fail = n_efail; s.last_order = order; s.reserved = 0; s.last_h = h_abs;
            a.stats[ir] = s;
        }
    }
    __syncwarp();
}

}  // namespace rmsb
