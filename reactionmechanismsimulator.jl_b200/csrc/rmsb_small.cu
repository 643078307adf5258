// rmsb_small.cu -- the small-mechanism path: ONE LANE GROUP (16 or 32 lanes) INTEGRATES ONE REACTOR.
//
// Scope: all-Arrhenius ideal-gas mechanisms with Ns <= 32 species and Nr <= 64 reactions in the variable-k domains
// ConstantV (src/Domain.jl:1064-1113,1901-1906) and ConstantP (:1141-1192,1950-1953) -- the headline configuration
// C2 (H2/O2 ignition-delay ensembles) and everything shaped like it.  Everything else stays on the warp-per-reactor
// kernels of rmsb_solver.cuh.
//
// What is different from the warp-per-reactor kernel (round-1 ncu: 5 600 warp instructions per step, 28 % issue
// utilisation, top stall `no_instruction`, every table read an LDG):
//   * the mechanism is re-packed at rmsb_mech_create into ONE lane-major image (NASA rows per species lane, Arrhenius
//     triplets / packed reactant-product bytes / packed Kc rows per reaction lane, packed species -> reaction lists) that a
//     block stages into shared memory once; the RHS issues no global loads at all;
//   * species values (G/RT, concentrations) travel between lanes by warp shuffles, net rates through a [q | 0 | -q]
//     array in shared memory, so the reference's accumulation orders are kept (results are bit-identical to rhs_v);
//   * state vectors live one component per lane; with Ny <= 16 a reactor needs 16 lanes, so a warp carries TWO reactors
//     (G = 16) and every Newton iteration (RHS + triangular solves) serves both;
//   * the integrator is the same NDF/BDF as bdf_solve_warp (same constants, Newton test, hysteresis, unit rows, Jacobian
//     refresh rule) written as a flat state machine: one loop iteration = one Newton iteration; the step logic, the LU
//     factorisation and the Jacobian are regions of that loop entered when EITHER group of the warp needs them
//     (`__any_sync`), executed by all 32 lanes with each group's state updates predicated on its own need.  Every warp
//     primitive therefore runs at warp-uniform control flow with the full mask (a shuffle with a run-time lane mask costs
//     MATCH.ANY + VOTEU + REDUX + BRA.DIV on top of the SHFL -- round-2 ncu, profiles/r02_small_a_*), and the two groups
//     never leave lockstep.  A group that finishes its reactor pulls the next one from the work queue inside the loop.
#include <algorithm>
#include <cmath>
#include <cstring>
#include <map>

#include <cub/device/device_radix_sort.cuh>

#include "rmsb_internal.h"
#include "rmsb_solver.cuh"

using namespace rmsb;

namespace {

struct SmallLayout {          // shared memory of a block, in doubles from smem_ws
    int tab;                  // table image
    int grp0, gstride;        // first group slice, doubles per group
    int o_q, o_J, o_W, o_ru, o_perm, o_dinv;   // inside a group slice
    int qplane;               // 2*NRP + 2
    int LD;                   // leading dimension of J / LU (Ny | 1)
    int NRS;                  // reaction slots per lane = ceil(Nr / G)
};

template <int G> __device__ __forceinline__ double gshfl(unsigned m, double v, int src) { return __shfl_sync(m, v, src, G); }
template <int G> __device__ __forceinline__ Dual2 gshfl(unsigned m, Dual2 v, int src) {
    return Dual2(__shfl_sync(m, v.v, src, G), __shfl_sync(m, v.a, src, G), __shfl_sync(m, v.b, src, G));
}
template <int G> __device__ __forceinline__ double gsum(unsigned m, double x) {
#pragma unroll
    for (int o = G / 2; o > 0; o >>= 1) x += __shfl_xor_sync(m, x, o, G);
    return x;
}
template <int G> __device__ __forceinline__ Dual2 gsum(unsigned m, Dual2 x) { return Dual2(gsum<G>(m, x.v), gsum<G>(m, x.a), gsum<G>(m, x.b)); }

// net rates of the group's reactions: value plane in the group's q array, derivative planes (Dual2) in scratch
template <class N> struct QIO;
template <> struct QIO<double> {
    static __device__ __forceinline__ void st(double* q, double*, int, int NRP, int j, double v, bool) { q[j] = v; q[NRP + 1 + j] = -v; }
    static __device__ __forceinline__ double ld(const double* q, const double*, int, int idx) { return q[idx]; }
};
template <> struct QIO<Dual2> {
    static __device__ __forceinline__ void st(double* q, double* qd, int qp, int NRP, int j, Dual2 v, bool mine) {
        q[j] = v.v; q[NRP + 1 + j] = -v.v;
        if (mine) { qd[j] = v.a; qd[NRP + 1 + j] = -v.a; qd[qp + j] = v.b; qd[qp + NRP + 1 + j] = -v.b; }   // the planes overlay the LU area
    }
    static __device__ __forceinline__ Dual2 ld(const double* q, const double* qd, int qp, int idx) { return Dual2(q[idx], qd[idx], qd[qp + idx]); }
};

struct GCtx {                 // one reactor's lane group: uniform within the group
    const double* tab;        // shared-memory table image
    double* gs;               // the group's shared-memory slice
    const double* p;          // per-reactor parameter row (global) or nullptr
    int mine;                 // this group takes part in the current Jacobian / factorisation region (stores are predicated on it)
    int gl;                   // lane within the group
    int kind;                 // RMSB_DOMAIN_CONSTANT_V / _P
    int Ny;
    double cond0;             // V (ConstantV) or P (ConstantP)
    double invV0;             // 1/V (ConstantV)
    double pG;                // dG of my species (0 without p)
    int is_const;             // my component is a constant species
};

struct RhsAux { double T, P, V, Nt, C, Cv, dotU; };

// calcHSCpdless of my species from the staged image (src/Calculators/Thermovec.jl:76-119); same arithmetic as nasa_hscp
template <class N>
__device__ __forceinline__ void small_nasa(const SmallTables& st, const double* tab, int gl, const TPow<N>& t, double Tv, N& cp, N& h, N& sr) {
    int sel = 0;                                                  // selectPoly (src/Calculators/Thermo.jl:101-111): tmax rows are +inf from a species' last range on
    const double* tm = tab + st.o_tmax + gl;
#pragma unroll 1
    for (int r = 0; r + 1 < st.maxr; ++r) sel += (Tv > tm[r * SM_SP]) ? 1 : 0;
    const double* c = tab + st.o_nasa + sel * st.rows * SM_SP + gl;
    if (st.rows == 9) {                                           // 7-coefficient species are stored with two leading zeros: same sums
        N ct0 = c[0] * t.T5, ct1 = c[SM_SP] * t.T4;
        double ct2 = c[2 * SM_SP];
        N ct3 = c[3 * SM_SP] * t.T0, ct4 = c[4 * SM_SP] * t.T1, ct5 = c[5 * SM_SP] * t.T2, ct6 = c[6 * SM_SP] * t.T3;
        cp = ct0 + ct1 + ct2 + ct3 + ct4 + ct5 + ct6;
        h = -ct0 + t.T6 * ct1 + ct2 + 0.5 * ct3 + 0.33333333333 * ct4 + 0.25 * ct5 + 0.2 * ct6 + c[7 * SM_SP] * t.T4;
        sr = -0.5 * ct0 - ct1 + t.T6 * ct2 + ct3 + 0.5 * ct4 + 0.33333333333 * ct5 + 0.25 * ct6 + c[8 * SM_SP];
    } else {
        double ct2 = c[0];
        N ct3 = c[SM_SP] * t.T0, ct4 = c[2 * SM_SP] * t.T1, ct5 = c[3 * SM_SP] * t.T2, ct6 = c[4 * SM_SP] * t.T3;
        cp = ct2 + ct3 + ct4 + ct5 + ct6;
        h = ct2 + 0.5 * ct3 + 0.33333333333 * ct4 + 0.25 * ct5 + 0.2 * ct6 + c[5 * SM_SP] * t.T4;
        sr = t.T6 * ct2 + ct3 + 0.5 * ct4 + 0.33333333333 * ct5 + 0.25 * ct6 + c[6 * SM_SP];
    }
}

// dydtreactor! (src/Reactor.jl:653-660) of the group's reactor: my state component in `yv`, my component of dy/dt returned.
// The formulas, reciprocals and accumulation orders are those of rhs_v (rmsb_device.cuh).  STORE keeps what the Jacobian
// needs (cs, kf, kr, U, cp/R in the scratch arrays, the scalars in *aux).
template <int G, class N, bool STORE>
__device__ __forceinline__ N small_rhs(const SmallTables& st, const SmallLayout& lay, const GCtx& g, double yv, RhsAux* aux) {
    const int Ns = st.Ns, NRP = st.NRP, gl = g.gl;
    const unsigned gm = FULL;           // both groups of a warp always evaluate together
    const bool isp = (g.kind == RMSB_DOMAIN_CONSTANT_P);
    double* qarr = g.gs + lay.o_q;
    double* W = g.gs + lay.o_W;
    double* qd = W + SM_SP * 3 + 2 * NRP;          // Dual2 planes: behind cs / us / cp / kf / kr (jac scratch), only used when N = Dual2
    const double Tv = gshfl<G>(gm, yv, Ns), second = gshfl<G>(gm, yv, Ns + 1);
    const double V = isp ? second : g.cond0;
    const double invV = isp ? 1.0 / V : g.invV0;
    N T = seed<N>(Tv, 0), P = isp ? N(g.cond0) : seed<N>(second, 1);
    TPow<N> tp(T);
    N RT = RMSB_R * T;
    N invRT = tp.T4 * RMSB_INV_R;
    N Nt = P * V * invRT;
    N C = P * invRT;
    N cp, h, sr;
    small_nasa<N>(st, g.tab, gl, tp, Tv, cp, h, sr);
    h = h + g.pG * invRT;                                          // Domain.jl:1100 (pG = 0 without p: exact)
    N U = isp ? h * RT : (h - 1.0) * RT;
    N gib = h - sr;
    const double cs = yv * invV;
    N cpsum = gsum<G>(gm, cp * yv);                               // species lanes only: the table rows of lanes >= Ns are zero
    if (STORE && g.mine) { W[gl] = cs; W[SM_SP + gl] = val(U); W[2 * SM_SP + gl] = val(cp); }
    N logC0 = RMSB_LOG_1E5_OVER_R - tp.T6;
    const unsigned long long* rxw = reinterpret_cast<const unsigned long long*>(g.tab + st.o_rx);
    const unsigned long long* kcw = reinterpret_cast<const unsigned long long*>(g.tab + st.o_kc);
    const unsigned long long* kcc = reinterpret_cast<const unsigned long long*>(g.tab + st.o_kcc);
    if (sizeof(N) != sizeof(double) && gl == 0 && g.mine) { qd[NRP] = 0.0; qd[lay.qplane + NRP] = 0.0; }
#pragma unroll 1
    for (int q = 0; q < lay.NRS; ++q) {
        const int j = q * G + gl;
        const double* a = g.tab + st.o_arr + j;
        N kf = a[0] * dexp(a[SM_MAXNR] * tp.T6 - a[2 * SM_MAXNR] * invRT);          // Ratevec.jl:24
        const unsigned long long kw = kcw[j], kc = kcc[j];
        const unsigned klo = (unsigned)kw, khi = (unsigned)(kw >> 32), clo = (unsigned)kc, chi = (unsigned)(kc >> 32);
        N ex = 0.0;                                                                   // row j of the stoichiometric matrix times G/RT, species ascending
#pragma unroll
        for (int e = 0; e < 8; ++e) {
            if (e >= st.max_kc) break;
            const int spc = (int)(((e < 4 ? klo : khi) >> (8 * (e & 3))) & 0xffu);
            const int cf = (int)((clo >> (4 * e)) & 15u) - 8;
            ex = ex + (double)cf * gshfl<G>(gm, gib, spc);
        }
        ex = ex + (double)((int)(chi & 15u) - 8) * logC0;
        N kr = kf * dexp(-ex);                                                        // kf/Kc (PhaseState.jl:165-167,230)
        kf = kf * (double)((chi >> 8) & 1u); kr = kr * (double)((chi >> 9) & 1u);
        if (g.p) { const double pk = (j < st.Nr) ? g.p[Ns + j] : 1.0; kf = kf * pk; kr = kr * pk; }   // Domain.jl:1112
        if (STORE && g.mine) { W[3 * SM_SP + j] = val(kf); W[3 * SM_SP + NRP + j] = val(kr); }
        const unsigned long long rw = rxw[j];
        const unsigned rlo = (unsigned)rw, rhi = (unsigned)(rw >> 32);
        N fR = kf * gshfl<G>(gm, cs, (int)(rlo & 0xffu));                              // Reactor.jl:465-483
#pragma unroll
        for (int k = 1; k < 4; ++k) {
            if (k >= st.max_nre) break;
            const unsigned b = (rlo >> (8 * k)) & 0xffu;
            const double cv = gshfl<G>(gm, cs, (int)b);
            fR = fR * ((b == 0xffu) ? 1.0 : cv);
        }
        N rR = kr * gshfl<G>(gm, cs, (int)(rhi & 0xffu));
#pragma unroll
        for (int k = 1; k < 4; ++k) {
            if (k >= st.max_npr) break;
            const unsigned b = (rhi >> (8 * k)) & 0xffu;
            const double cv = gshfl<G>(gm, cs, (int)b);
            rR = rR * ((b == 0xffu) ? 1.0 : cv);
        }
        QIO<N>::st(qarr, qd, lay.qplane, NRP, j, fR - rR, g.mine != 0);
    }
    __syncwarp();
    // my species' reactions in reaction order (Reactor.jl:484-505): entries index [q | 0 | -q]
    N acc = 0.0;
    const unsigned long long* srw = reinterpret_cast<const unsigned long long*>(g.tab + st.o_sr) + gl;
#pragma unroll 1
    for (int w = 0; w < st.sr_words; ++w) {
        const unsigned long long e8 = srw[w * SM_SP];
        const unsigned lo = (unsigned)e8, hi = (unsigned)(e8 >> 32);
#pragma unroll
        for (int b = 0; b < 8; ++b) {
            const int idx = (int)(((b < 4 ? lo : hi) >> (8 * (b & 3))) & 0xffu);
            acc = acc + QIO<N>::ld(qarr, qd, lay.qplane, idx);
        }
    }
    __syncwarp();
    N dn = acc * V;                                                                   // dydt .*= V
    N dotU = gsum<G>(gm, U * dn), sumn = gsum<G>(gm, dn);
    // dT/dt = -dot(Us,dn)/(N*Cv), N*Cv = R*(dot(cp/R,ns) - N)  (Domain.jl:1103-1105,1902); ConstantP: N*Cp = R*dot(cp/R,ns) (:1951-1953)
    N dT = isp ? -dotU / (RMSB_R * cpsum) : -dotU / (RMSB_R * (cpsum - Nt));
    N dP = isp ? sumn * RT / P + dT * V * tp.T4 : sumn * RT * invV + P * tp.T4 * dT;
    N f = (gl == Ns) ? dT : ((gl == Ns + 1) ? dP : dn);
    if (g.is_const) f = N(0.0);                                                       // Domain.jl:1904-1906
    if (STORE) {
        aux->T = val(T); aux->P = val(P); aux->V = V; aux->Nt = val(Nt); aux->C = val(C); aux->dotU = val(dotU);
        aux->Cv = val(cpsum) * (RMSB_R / val(Nt)) - RMSB_R;
    }
    return f;
}

// 1/x to about one ulp for the step heuristics, pivots and weights: MUFU.RCP64H + two Newton steps, no special cases
// (callers guarantee a finite, non-zero x).  The compiler's IEEE division is 12 instructions plus an out-of-line slow path.
__device__ __forceinline__ double rcp_fast(double x) {
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
    double e = fma(-x, r, 1.0);
    r = fma(r, e, r);
    e = fma(-x, r, 1.0);
    return fma(r, e, r);
}

// jacobiany! of the reactors of the warp's groups (src/Domain.jl:2204-2632; steps as in jac_warp): my state component in yv, J
// (column-major, leading dimension LD) in the group's slice.  All lanes execute; only groups with g.mine store.  The scratch
// arrays overlay the LU area: a new Jacobian always invalidates the factors.
template <int G>
__device__ __noinline__ void small_jac(SmallTables st, SmallLayout lay, GCtx g, double yv, int nconst, const int* constspc) {
    const DevMech& m = mech();
    const int Ns = st.Ns, Nr = st.Nr, NRP = st.NRP, Ny = g.Ny, LD = lay.LD, gl = g.gl;
    const bool mine = g.mine != 0;
    RhsAux ax;
    small_rhs<G, double, true>(st, lay, g, yv, &ax);
    double* J = g.gs + lay.o_J;
    double* W = g.gs + lay.o_W;
    const double* cs = W;
    const double* us = W + SM_SP;
    const double* cpb = W + 2 * SM_SP;
    const double* kfb = W + 3 * SM_SP;
    const double* krb = kfb + NRP;
    double* dterm = W + 3 * SM_SP + 2 * NRP;
    __syncwarp();
    if (mine) for (int i = gl; i < Ny * LD; i += G) J[i] = 0.0;
    if (mine) for (int t = gl; t < m.n_dt; t += G) {              // (1) derivative terms (src/Reactor.jl:899-1366)
        int jr = G(m.dt_rxn, t);
        int j = jr & 0x3fffffff;
        double v = (double)G(m.dt_mult, t) * ((jr >> 30) ? krb[j] : kfb[j]);
        for (int k = 0; k < 3; ++k) { int s = G(m.dt_oth, 3 * t + k); if (s < 0) break; v *= cs[s]; }
        dterm[t] = v;
    }
    __syncwarp();
    if (mine) for (int e = gl; e < m.n_je; e += G) {              // (2) entries, in the reference's accumulation order
        double acc = 0.0;
        for (int k = G(m.je_ptr, e); k < G(m.je_ptr, e + 1); ++k) {
            int ti = G(m.jt_idx, k);
            double v = dterm[ti & 0x7fffffff];
            acc = (ti < 0) ? acc - v : acc + v;
        }
        J[G(m.je_row, e) + G(m.je_col, e) * LD] = acc;
    }
    __syncwarp();
    const bool isp = (g.kind == RMSB_DOMAIN_CONSTANT_P);
    const int Vcol = Ns + 1;
    if (isp) {                                                    // (3) V column of the species block (src/Reactor.jl:1368-1413)
        if (mine) for (int s = gl; s < Ns; s += G) {
            double acc = 0.0;
            for (int k = G(m.sr_ptr, s); k < G(m.sr_ptr, s + 1); ++k) {
                int jr = G(m.sr_rxn, k);
                int j = jr & 0x7fffffff;
                double df = 0.0, dr = 0.0;
                int nre = 0, npr = 0;
                double pf = kfb[j], pr = krb[j];
                for (int q = 0; q < 4; ++q) { int sp = G(m.rx, q * Nr + j); if (sp < 0) break; pf *= cs[sp]; ++nre; }
                for (int q = 4; q < 8; ++q) { int sp = G(m.rx, q * Nr + j); if (sp < 0) break; pr *= cs[sp]; ++npr; }
                if (nre > 1) df = -(double)(nre - 1) * pf;
                if (npr > 1) dr = -(double)(npr - 1) * pr;
                if (jr < 0) { acc -= df; acc += dr; } else { acc += df; acc -= dr; }
            }
            J[s + Vcol * LD] = acc;
        }
        __syncwarp();
    }
    // (5) thermodynamic-variable rows (no efficiency reactions on this path, so step (4) of jac_warp is empty)
    const double aT = ax.T, aP = ax.P, aV = ax.V, aN = ax.Nt, aCv = ax.Cv, adotU = ax.dotU;
    if (!isp) {
        if (mine) for (int c = gl; c < Ns; c += G) {              // Domain.jl:2312-2316
            double dCv = cpb[c] * RMSB_R / aN;
            double a = 0.0, sm = 0.0;
            for (int i = 0; i < Ns; ++i) { double v = J[i + c * LD]; a += us[i] * v; sm += v; }
            double jt = -a / (aN * aCv) + adotU / (aN * aCv * aCv) * dCv;
            J[Ns + c * LD] = jt;
            J[Ns + 1 + c * LD] = sm * RMSB_R * aT / aV + aP / aT * jt;
        }
    } else {                                                      // Domain.jl:2480-2490 (us holds the enthalpies)
        const double aCp = aCv + RMSB_R;
        const double dTdt = -adotU / (aN * aCp);
        if (mine) for (int c = gl; c < Ns + 2; c += G) {
            if (c == Ns) continue;
            double a = 0.0, sm = 0.0;
            for (int i = 0; i < Ns; ++i) { double v = J[i + c * LD]; a += us[i] * v; sm += v; }
            double jt = -a / (aN * aCp);
            if (c < Ns) jt -= dTdt * ((cpb[c] * RMSB_R / aN) / aCp);
            J[Ns + c * LD] = jt;
            double jv = sm * RMSB_R * aT / aP + aV / aT * jt;
            if (c == Ns + 1) jv += dTdt / aT;
            J[Ns + 1 + c * LD] = jv;
        }
    }
    __syncwarp();
    if (mine) for (int k = 0; k < nconst; ++k) {                  // (6) constant species rows
        int s = constspc[k] - 1;
        for (int c = gl; c < Ny; c += G) J[s + c * LD] = 0.0;
    }
    __syncwarp();
    // (7) T (and P) columns: derivative of the complete RHS (src/Reactor.jl:1469-1477)
    Dual2 fd = small_rhs<G, Dual2, false>(st, lay, g, yv, nullptr);
    if (mine && gl < Ny) { J[gl + Ns * LD] = fd.a; if (!isp) J[gl + (Ns + 1) * LD] = fd.b; }
    __syncwarp();
}

// LU of A = I - c*J with partial pivoting, rows owned by lanes (the algorithm of lu_factor_warp<1>).  All lanes execute; only
// groups with `mine` store (the others walk through their existing factors and discard the result).
template <int G>
__device__ __noinline__ int small_lu(double* A, const double* J, int* perm, double* dinv, int Ny, int LD, double c, int gl, int gbase, bool mine) {
    const bool act = gl < Ny;
    const bool wr = act && mine;
    if (wr) {
        const double* pj = J + gl;
        double* pa = A + gl;
        for (int col = 0; col < Ny; ++col, pj += LD, pa += LD) *pa = ((gl == col) ? 1.0 : 0.0) - c * *pj;
        perm[gl] = gl;
    }
    int ok = 1;
    double* ck = A + gl;
    for (int k = 0; k < Ny; ++k, ck += LD) {
        __syncwarp();
        // pivot search on the high word of |a_ik| (ordering of non-negative doubles == ordering of their bits)
        unsigned key = (act && gl >= k) ? (unsigned)__double2hiint(fabs(*ck)) + 1u : 0u;
        unsigned mx = key;
#pragma unroll
        for (int o = G / 2; o > 0; o >>= 1) mx = max(mx, __shfl_xor_sync(FULL, mx, o, G));
        const unsigned vote = (__ballot_sync(FULL, key == mx && key != 0u) >> gbase) & ((G == 32) ? 0xffffffffu : 0xffffu);
        const int pr = __ffs(vote) - 1;
        if (__any_sync(FULL, pr != k)) {
            if (wr && pr != k) { double* pa = A + k + gl * LD; double* pb = A + pr + gl * LD; double t = *pa; *pa = *pb; *pb = t; }
            if (mine && gl == 0 && pr != k) { int t = perm[k]; perm[k] = perm[pr]; perm[pr] = t; }
            __syncwarp();
        }
        double piv = A[k + k * LD];
        if (piv == 0.0 || !isfinite(piv)) { ok = 0; piv = 1.0; }
        const double inv = rcp_fast(piv);
        if (mine && gl == 0) dinv[k] = inv;
        const bool below = wr && gl > k;
        double l = 0.0;
        if (below) { l = *ck * inv; *ck = l; }
        const double* pk = A + k + (k + 1) * LD;
        double* pi = ck + LD;
#pragma unroll 2
        for (int j = k + 1; j < Ny; ++j, pk += LD, pi += LD) {
            const double akj = *pk;
            if (below) *pi -= l * akj;
        }
    }
    __syncwarp();
    return ok;
}

// the 6x6 matrix R(factor) U of change_D for each group of the warp (compute_ru of rmsb_solver.cuh), one entry per lane
template <int G>
__device__ __noinline__ void small_ru(double* ru, double factor, int gl) {
    if (gl < 6) {
        const double fk = factor * (double)gl;
        double r = 1.0;
        ru[gl] = r;
#pragma unroll
        for (int i = 1; i < 6; ++i) { r *= ((double)(i - 1) - fk) * c_invl[i]; ru[i * 6 + gl] = r; }
    }
    __syncwarp();
    double acc[36 / G + 1];
#pragma unroll
    for (int pass = 0; pass < 36 / G + 1; ++pass) {
        const int e = gl + G * pass;
        acc[pass] = 0.0;
        if (e < 36) {
            const int i = e / 6, j = e - 6 * i;
            double a = 0.0;
            for (int k = 0; k <= j; ++k) a += ru[i * 6 + k] * c_U[k * 6 + j];
            acc[pass] = a;
        }
    }
    __syncwarp();
#pragma unroll
    for (int pass = 0; pass < 36 / G + 1; ++pass) { const int e = gl + G * pass; if (e < 36) ru[e] = acc[pass]; }
    __syncwarp();
}

static __constant__ double c_inv_alpha[6] = {0.0, 1.0 / ((1.0 - RMSB_K1) * RMSB_G1), 1.0 / ((1.0 - RMSB_K2) * RMSB_G2), 1.0 / ((1.0 - RMSB_K3) * RMSB_G3),
                                             1.0 / ((1.0 - RMSB_K4) * RMSB_G4), 1.0 / RMSB_G5};

enum { PH_FETCH = 0, PH_INIT0 = 1, PH_INIT1 = 2, PH_ATTEMPT = 3, PH_NEWTON = 4, PH_IDLE = 5 };

__device__ __forceinline__ void stage_small(const DevMech& mp, const SmallTables& st, const SmallLayout& lay) {
    const int* src = reinterpret_cast<const int*>(&mp);
    int* dst = reinterpret_cast<int*>(smem_ws);
    for (int i = threadIdx.x; i < (int)(sizeof(DevMech) / sizeof(int)); i += blockDim.x) dst[i] = src[i];
    double* tab = smem_ws + lay.tab;
    for (int i = threadIdx.x; i < st.tab_doubles; i += blockDim.x) tab[i] = __ldg(st.tab + i);
    __syncthreads();
}

#ifndef RMSB_SMALL_MINBLOCKS
#define RMSB_SMALL_MINBLOCKS 4
#endif
#ifndef RMSB_SMALL_THREADS
#define RMSB_SMALL_THREADS 128
#endif
#ifndef RMSB_SMALL_HYST
#define RMSB_SMALL_HYST 1.2        // step-size increases below this factor (at unchanged order) are not taken: no rescale, no refactorisation
#endif
#ifndef RMSB_SMALL_DGMAX
#define RMSB_SMALL_DGMAX 0.0       // > 0: keep the LU factors while c = h/alpha stays within this fraction of the factorised value (CVODE's
#endif                             // dgmax = 0.3, corrections scaled by 2/(1 + c/c_lu)); measured, see DESIGN 5.1
#ifndef RMSB_SMALL_TICKSYNC
#define RMSB_SMALL_TICKSYNC 0      // 1: the warps of a block start every tick together (experiment: shared instruction fetch)
#endif

template <int G>
__global__ void __launch_bounds__(RMSB_SMALL_THREADS, RMSB_SMALL_MINBLOCKS)
small_solve_kernel(const __grid_constant__ DevMech mp, const __grid_constant__ SmallTables st, const __grid_constant__ SmallLayout lay,
                   const __grid_constant__ SolveArgs a) {
    stage_small(mp, st, lay);
    const int lane = threadIdx.x & 31, gl = lane & (G - 1), gbase = lane & ~(G - 1);
    const int grp = threadIdx.x / G;
    const int Ny = a.Ny, Ns = st.Ns, LD = lay.LD;
    GCtx g;
    g.tab = smem_ws + lay.tab;
    g.gs = smem_ws + lay.grp0 + grp * lay.gstride;
    g.mine = 0; g.gl = gl; g.kind = a.domain_kind; g.Ny = Ny;
    g.p = nullptr; g.cond0 = 1.0; g.invV0 = 1.0; g.pG = 0.0; g.is_const = 0;
    double* const qarr = g.gs + lay.o_q;
    double* const Jm = g.gs + lay.o_J;
    double* const LU = g.gs + lay.o_W;
    double* const ru = g.gs + lay.o_ru;
    int* const perm = reinterpret_cast<int*>(g.gs + lay.o_perm);
    double* const dinvs = g.gs + lay.o_dinv;
    if (gl == 0) qarr[st.NRP] = 0.0;
    __syncwarp();
    const bool act = gl < Ny;
    const double rtol = a.rtol, atol = a.atol;
    const int max_order = a.max_order;
    const double newton_tol = fmax(10.0 * 2.220446049250313e-16 / rtol, fmin(0.03, sqrt(rtol)));
    const bool scipy_newton = (a.flags & 1) != 0, hysteresis = (a.flags & 2) == 0;
    const int nls_v = (a.flags >> 8) & 0xff;
    const double nls_coef = ldexp(0.1, -(nls_v == 0 ? 3 : nls_v - 1));
    const int idt_index = a.idt_index;
    const double inv_ny = 1.0 / (double)Ny;

    // ---- per-reactor state (scalars are uniform within a group, not within the warp) ---------------------
    double D[BDF_MAX_ORDER + 3];
    double ypred = 0.0, psi = 0.0, iw = 0.0, dvec = 0.0, ynew = 1.0, dinv = 1.0;
    int myperm = gl;
    bool unit_row = false;
    long long ir = 0;
    double t = 0.0, t_end = 0.0, h_abs = 0.0, t_new = 0.0, h = 1.0, c = 0.0, error_norm = 0.0;
    double crate = 1.0, c_jac = 0.0, dy_norm_old = -1.0, idt_best = -INFINITY, idt_time = NAN, pend = 1.0, h0 = 0.0, d1n = 0.0, s_prev = 0.0;
    int order = 1, n_equal = 0, kiter = 0, isave = 0, status = RMSB_SOLVE_OK;
    int n_rhs = 0, n_jac = 0, n_lu = 0, n_steps = 0, n_nfail = 0, n_efail = 0;
    bool lu_valid = false, current_jac = false, jac_pending = false, jac_from_y = false, lu_ok = true, new_step = true;
    double c_lu = 0.0, nscale = 1.0;      // RMSB_SMALL_DGMAX: c of the factors in place (0: they do not belong to the Jacobian in place)
    int phase = PH_FETCH;
#pragma unroll
    for (int j = 0; j < BDF_MAX_ORDER + 3; ++j) D[j] = 0.0;

#if RMSB_SMALL_TICKSYNC
    bool warp_done = false;
#endif
    for (;;) {
#if RMSB_SMALL_TICKSYNC
        if (__syncthreads_and(warp_done)) break;
        if (warp_done) continue;
#endif
        // ================================================================================ next reactor (cold)
        {
            const bool want = (phase == PH_FETCH);
            if (__any_sync(FULL, want)) {
                unsigned long long nx = 0;
                if (want && gl == 0) nx = atomicAdd(a.counter, 1ULL);
                nx = __shfl_sync(FULL, nx, 0, G);
                if (want) {
                    if (nx >= (unsigned long long)a.n) phase = PH_IDLE;
                    else {
                        ir = a.order ? (long long)a.order[nx] : (long long)nx;
                        g.p = a.p ? a.p + ir * a.p_stride : nullptr;
                        g.cond0 = a.cond[ir * a.ncond];
                        g.invV0 = (a.domain_kind == RMSB_DOMAIN_CONSTANT_V) ? 1.0 / g.cond0 : 0.0;
                        g.pG = (g.p && gl < Ns) ? g.p[gl] : 0.0;
                        t_end = a.t_end[ir];
                        g.is_const = 0;
#pragma unroll 1
                        for (int k = 0; k < a.nconst; ++k) if (a.constspc[k] - 1 == gl) g.is_const = 1;
                        unit_row = (gl < Ns) && (G(mech().inert, gl) != 0 || g.is_const);
                        if (a.flags & 4) unit_row = false;
#pragma unroll
                        for (int j = 0; j < BDF_MAX_ORDER + 3; ++j) D[j] = 0.0;
                        D[0] = act ? a.y0[ir * Ny + gl] : 0.0;
                        ynew = D[0];
                        t = 0.0; order = 1; n_equal = 0; kiter = 0; status = RMSB_SOLVE_OK;
                        n_rhs = n_jac = n_lu = n_steps = n_nfail = n_efail = 0;
                        crate = 1.0; c_jac = 0.0; idt_best = -INFINITY; idt_time = NAN; pend = 1.0;
                        lu_valid = false; current_jac = false; jac_pending = false; jac_from_y = false; lu_ok = true; new_step = true;
                        c_lu = 0.0; nscale = 1.0;
                        isave = 0;
                        if (a.nsave > 0) {
                            double* ysave = a.ysave ? a.ysave + (size_t)ir * a.nsave * Ny : nullptr;
                            while (isave < a.nsave && a.tsave[isave] <= 0.0) {
                                if (ysave && act) ysave[(size_t)isave * Ny + gl] = D[0];
                                ++isave;
                            }
                        }
                        phase = PH_INIT0;
                    }
                }
#if RMSB_SMALL_TICKSYNC
                if (__all_sync(FULL, phase == PH_IDLE)) { warp_done = true; continue; }
#else
                if (__all_sync(FULL, phase == PH_IDLE)) break;
#endif
            }
        }
        bool finish = false;
        // ================================================================================ start / restart an attempt
        {
            const bool att = (phase == PH_ATTEMPT);
            if (__any_sync(FULL, att)) {
                const bool resc = att && pend != 1.0;
                if (__any_sync(FULL, resc)) {
                    // change_D: D_new[j] = sum_i (R(factor) U)[i][j] D[i]
                    small_ru<G>(ru, resc ? pend : 1.0, gl);
                    if (resc) {
                        double Dn[BDF_MAX_ORDER + 1];
#pragma unroll
                        for (int j = 1; j <= BDF_MAX_ORDER; ++j) {
                            Dn[j] = 0.0;
#pragma unroll
                            for (int i = 0; i <= BDF_MAX_ORDER; ++i) if (i <= order) Dn[j] += ru[i * 6 + j] * D[i];
                        }
#pragma unroll
                        for (int j = 1; j <= BDF_MAX_ORDER; ++j) if (j <= order) D[j] = Dn[j];
                        pend = 1.0;
                    }
                    __syncwarp();
                }
                if (att) {
                    const double min_step = 10.0 * (__longlong_as_double(__double_as_longlong(t) + 1) - t);
                    const double hc = t_end - t;
                    if (new_step && n_steps >= a.max_steps) { status = RMSB_SOLVE_MAX_STEPS; finish = true; }
                    else if (h_abs < min_step) {
                        if (new_step) { pend = min_step / h_abs; h_abs = min_step; n_equal = 0; lu_valid = false; }   // once per step: stretch to the minimum
                        else { status = RMSB_SOLVE_STEP_TOO_SMALL; finish = true; }
                    } else if (h_abs != hc && t + h_abs - t_end > 0.0) {              // land on t_end exactly
                        pend = hc / h_abs; h_abs = hc; n_equal = 0; lu_valid = false;
                    }
                    new_step = false;
                    if (!finish && pend == 1.0) {
                        t_new = (h_abs == hc) ? t_end : t + h_abs;
                        h = h_abs;
                        const double ia = c_inv_alpha[order];
                        double s = 0.0, ps = 0.0;
#pragma unroll
                        for (int j = 0; j <= BDF_MAX_ORDER; ++j) if (j <= order) s += D[j];
#pragma unroll
                        for (int j = 1; j <= BDF_MAX_ORDER; ++j) if (j <= order) ps += D[j] * c_gamma[j];
                        ypred = s;
                        psi = ps * ia;
                        iw = rcp_fast(atol + rtol * fabs(s));
                        c = h * ia;
                        dvec = 0.0; ynew = ypred; dy_norm_old = -1.0;
                        if (!lu_valid) {
                            // a Jacobian from a step size 100x away is refreshed before it is factorised again (see bdf_solve_warp)
                            if (c_jac == 0.0) c_jac = c;
                            else if (!current_jac && (c > JAC_REFRESH_RATIO * c_jac || c * JAC_REFRESH_RATIO < c_jac)) jac_pending = true;
                            if (RMSB_SMALL_DGMAX > 0.0 && !jac_pending && lu_ok && c_lu > 0.0 && fabs(c / c_lu - 1.0) <= RMSB_SMALL_DGMAX) lu_valid = true;
                        }
                        nscale = (RMSB_SMALL_DGMAX > 0.0 && lu_valid && c_lu > 0.0 && c != c_lu) ? 2.0 / (1.0 + c / c_lu) : 1.0;
                        kiter = 0;
                        phase = PH_NEWTON;
                    }
                }
            }
        }
        // ================================================================================ Jacobian (about one step in twelve)
        {
            const bool jn = (phase == PH_NEWTON) && jac_pending;
            if (__any_sync(FULL, jn)) {
                g.mine = jn;
                small_jac<G>(st, lay, g, jac_from_y ? D[0] : ypred, a.nconst, a.constspc);
                g.mine = 0;
                if (jn) {
                    ++n_jac;
                    if (!jac_from_y) { current_jac = true; c_jac = c; }
                    jac_pending = false; jac_from_y = false; lu_valid = false; c_lu = 0.0;
                }
            }
        }
        // ================================================================================ LU of I - c*J (about one step in four)
        {
            const bool ln = (phase == PH_NEWTON) && !lu_valid;
            if (__any_sync(FULL, ln)) {
                const int ok = small_lu<G>(LU, Jm, perm, dinvs, Ny, LD, c, gl, gbase, ln);
                if (ln) {
                    lu_ok = ok != 0;
                    myperm = act ? perm[gl] : gl;
                    dinv = act ? dinvs[gl] : 1.0;
                    lu_valid = true; ++n_lu; crate = 1.0; c_lu = c; nscale = 1.0;
                }
            }
        }
        // ================================================================================ RHS: every tick, both groups
        const double f = small_rhs<G, double, false>(st, lay, g, ynew, nullptr);
        const bool fin = __all_sync(FULL, !act || isfinite(f) || phase == PH_IDLE) ||
                         ((__ballot_sync(FULL, act && !isfinite(f)) >> gbase) & ((G == 32) ? 0xffffffffu : 0xffffu)) == 0u;
        // ---- (I - cJ) x = c f - psi - d with the pivoted factors: row gl of L and U is read column by column
        double dy = c * f - psi - dvec;
        {
            double b = __shfl_sync(FULL, dy, myperm, G);
            const double* col = LU + gl;
#pragma unroll 2
            for (int k = 0; k < Ny - 1; ++k, col += LD) {
                const double xk = __shfl_sync(FULL, b, k, G);
                if (act && gl > k) b -= *col * xk;
            }
#pragma unroll 2
            for (int k = Ny - 1; k >= 0; --k, col -= LD) {
                if (gl == k) b *= dinv;
                const double xk = __shfl_sync(FULL, b, k, G);
                if (gl < k) b -= *col * xk;
            }
            dy = unit_row ? dy : (RMSB_SMALL_DGMAX > 0.0 ? b * nscale : b);      // unit rows keep their right-hand side exactly (see bdf_solve_warp)
        }
        const double qn = act ? dy * iw : 0.0;
        const double dy_norm = sqrt(gsum<G>(FULL, qn * qn) * inv_ny);
        int outcome = 0;          // 1 = converged, 2 = failed
        if (phase == PH_NEWTON) {
            ++n_rhs;
            const double nl_tol = scipy_newton ? newton_tol : fmax(newton_tol, nls_coef / c_ec[order]);
            const double rate = (dy_norm_old < 0.0) ? -1.0 : dy_norm * rcp_fast(dy_norm_old);
            if (!lu_ok || !fin || !(dy_norm == dy_norm)) outcome = 2;
            else {
                if (rate >= 0.0) {
                    double rp = rate;
                    for (int q = 1; q < NEWTON_MAXITER - kiter; ++q) rp *= rate;
                    if (rate >= 1.0 || rp * dy_norm > nl_tol * (1.0 - rate)) outcome = 2;
                    else crate = fmax(0.3 * crate, rate);
                }
                if (outcome == 0) {
                    ynew += dy; dvec += dy;
                    if (scipy_newton) { if (dy_norm == 0.0 || (rate >= 0.0 && rate * dy_norm < nl_tol * (1.0 - rate))) outcome = 1; }
                    else if (dy_norm * fmin(1.0, crate) <= nl_tol) outcome = 1;
                    if (outcome == 0) {
                        dy_norm_old = dy_norm;
                        if (++kiter == NEWTON_MAXITER) outcome = 2;
                    }
                }
            }
            if (outcome == 2) {
                if (current_jac) {             // the Jacobian is current: halve the step
                    ++n_nfail; h_abs *= 0.5; pend = 0.5; n_equal = 0; lu_valid = false; c_lu = 0.0; phase = PH_ATTEMPT;
                } else {                        // new Jacobian at the predicted state, same step
                    jac_pending = true; kiter = 0; dvec = 0.0; ynew = ypred; dy_norm_old = -1.0;
                }
            }
        }
        // ================================================================================ first two evaluations of a reactor (cold)
        if (__any_sync(FULL, phase == PH_INIT0 || phase == PH_INIT1)) {
            // select_initial_step (Hairer-Norsett-Wanner II.4): INIT0 has f(y0) in f, INIT1 has f(y0 + h0 f0)
            const bool i0 = (phase == PH_INIT0), i1 = (phase == PH_INIT1);
            if (i0) iw = 1.0 / (atol + fabs(D[0]) * rtol);
            const double qa = act ? (i0 ? D[0] : (f - dvec)) * iw : 0.0, qb = act ? f * iw : 0.0;
            const double na = sqrt(gsum<G>(FULL, qa * qa) * inv_ny), nb = sqrt(gsum<G>(FULL, qb * qb) * inv_ny);
            if (i0) {
                ++n_rhs;
                if (!fin) { status = RMSB_SOLVE_NAN; finish = true; }
                dvec = f;                                   // keeps f0
                if (a.first_step > 0.0) { h_abs = a.first_step; D[1] = dvec * h_abs; jac_pending = true; jac_from_y = true; phase = PH_ATTEMPT; }
                else {
                    d1n = nb;
                    h0 = (na < 1e-5 || d1n < 1e-5) ? 1e-6 : 0.01 * na / d1n;
                    ynew = D[0] + h0 * f;
                    phase = PH_INIT1;
                }
            } else if (i1) {
                ++n_rhs;
                const double d2 = na / h0;
                const double h1 = (d1n <= 1e-15 && d2 <= 1e-15) ? fmax(1e-6, h0 * 1e-3) : sqrt(0.01 / fmax(d1n, d2));
                h_abs = fmin(fmin(100.0 * h0, h1), t_end);
                D[1] = dvec * h_abs;
                jac_pending = true; jac_from_y = true;
                phase = PH_ATTEMPT;
            }
            if (i0 || i1) s_prev = __shfl_sync(FULL, dvec, idt_index < 0 ? 0 : idt_index, G);   // slope of the tracked component at t = 0
            else (void)__shfl_sync(FULL, dvec, 0, G);
        }
        // ================================================================================ converged: error test, accept / reject
        const bool conv = (outcome == 1);
        if (__any_sync(FULL, conv)) {
            const double safety = 0.9 * (2.0 * NEWTON_MAXITER + 1.0) / (2.0 * NEWTON_MAXITER + (double)(kiter + 1));
            const double iwn = rcp_fast(atol + rtol * fabs(ynew));
            const double qe = act ? c_ec[order] * dvec * iwn : 0.0;
            const double en = sqrt(gsum<G>(FULL, qe * qe) * inv_ny);
            bool accepted = false;
            if (conv) {
                iw = iwn; error_norm = en;
                if (error_norm > 1.0) {
                    ++n_efail;
                    const double factor = fmax(BDF_MIN_FACTOR, safety * fast_pow(error_norm, -1.0f / (float)(order + 1)));
                    h_abs *= factor; pend = factor; n_equal = 0; lu_valid = false; phase = PH_ATTEMPT;
                } else accepted = true;
            }
            if (__any_sync(FULL, accepted)) {
                // ---------------------------------------------------------------------- accepted step
                const double t_old = t;
                if (accepted) {
                    ++n_equal; ++n_steps;
                    t = t_new;
#pragma unroll
                    for (int j = 2; j < BDF_MAX_ORDER + 3; ++j) if (j == order + 2) D[j] = dvec - D[j - 1];
#pragma unroll
                    for (int j = 1; j < BDF_MAX_ORDER + 3; ++j) if (j == order + 1) D[j] = dvec;
#pragma unroll
                    for (int j = BDF_MAX_ORDER; j >= 0; --j) if (j <= order) D[j] += D[j + 1];
                    // dense output on (t_old, t]
                    if (isave < a.nsave) {
                        double* ysave = a.ysave ? a.ysave + (size_t)ir * a.nsave * Ny : nullptr;
                        while (isave < a.nsave && a.tsave[isave] <= t) {
                            const double ts = a.tsave[isave];
                            double pj = 1.0, acc = D[0];
#pragma unroll
                            for (int j = 1; j <= BDF_MAX_ORDER; ++j)
                                if (j <= order) { pj *= (ts - (t - (double)(j - 1) * h)) / (h * (double)j); acc += D[j] * pj; }
                            if (ysave && act) ysave[(size_t)isave * Ny + gl] = acc;
                            ++isave;
                        }
                    }
                    current_jac = false; new_step = true;
                    phase = PH_ATTEMPT;
                }
                // ---- ignition-delay tracker: time of the largest d/dt of component idt_index.  The slope at the end of every
                //      step is one dot product; where it has passed through a maximum (it fell over this step, and is within 8x of
                //      the running maximum) the derivative of the interpolant is sampled at 32 points over the last two steps.
                if (idt_index >= 0) {
                    const double invh = rcp_fast(h);
                    double sl = 0.0;
#pragma unroll
                    for (int j = 1; j <= BDF_MAX_ORDER; ++j) if (j <= order) sl += c_gamma[j] * D[j];
                    const double s_end = __shfl_sync(FULL, sl, idt_index, G) * invh;
                    const bool refine = accepted && s_end < s_prev && 8.0 * fmax(s_prev, s_end) > idt_best;
                    if (accepted) {
                        if (s_end > idt_best) { idt_best = s_end; idt_time = t; }
                        s_prev = s_end;
                    }
                    if (__any_sync(FULL, refine)) {
                        double dk[BDF_MAX_ORDER + 1];
#pragma unroll
                        for (int j = 1; j <= BDF_MAX_ORDER; ++j) dk[j] = __shfl_sync(FULL, D[j], idt_index, G);
                        double best = -INFINITY, bt = 0.0;
#pragma unroll
                        for (int pass = 0; pass < 32 / G; ++pass) {
                            const double ts = t - 2.0 * h + h * (double)(gl + pass * G + 1) * (1.0 / 16.0);
                            double pj = 1.0, dpj = 0.0, der = 0.0;
#pragma unroll
                            for (int j = 1; j <= BDF_MAX_ORDER; ++j) {
                                if (j <= order) {
                                    const double iden = invh * c_invl[j];
                                    const double x = (ts - (t - (double)(j - 1) * h)) * iden;
                                    dpj = dpj * x + pj * iden;
                                    pj = pj * x;
                                    der += dk[j] * dpj;
                                }
                            }
                            if (der > best && ts > 0.0) { best = der; bt = ts; }
                        }
#pragma unroll
                        for (int o = G / 2; o > 0; o >>= 1) {
                            const double ob = __shfl_xor_sync(FULL, best, o, G), ot = __shfl_xor_sync(FULL, bt, o, G);
                            if (ob > best || (ob == best && ot < bt)) { best = ob; bt = ot; }
                        }
                        if (refine && best > idt_best) { idt_best = best; idt_time = bt; }
                    }
                    (void)t_old;
                }
                // ---- order / step-size selection every order + 1 equal steps
                const bool sel = accepted && n_equal >= order + 1;
                if (__any_sync(FULL, sel)) {
                    double vm = 0.0, vp = 0.0;
#pragma unroll
                    for (int j = 0; j < BDF_MAX_ORDER + 3; ++j) { if (j == order) vm = D[j]; if (j == order + 2) vp = D[j]; }
                    const double qm = act ? c_ec[order > 1 ? order - 1 : 0] * vm * iw : 0.0, qp = act ? c_ec[order < 5 ? order + 1 : 5] * vp * iw : 0.0;
                    const double nm = sqrt(gsum<G>(FULL, qm * qm) * inv_ny), np_ = sqrt(gsum<G>(FULL, qp * qp) * inv_ny);
                    if (sel) {
                        const double e_m = (order > 1) ? nm : INFINITY, e_p = (order < max_order) ? np_ : INFINITY;
                        const double fm = fast_pow(e_m, -1.0f / (float)order);
                        const double f0 = fast_pow(error_norm, -1.0f / (float)(order + 1));
                        const double fp = fast_pow(e_p, -1.0f / (float)(order + 2));
                        int delta; double fmax_;
                        if (fm >= f0 && fm >= fp) { delta = -1; fmax_ = fm; }
                        else if (f0 >= fp) { delta = 0; fmax_ = f0; }
                        else { delta = 1; fmax_ = fp; }
                        const double factor = fmin(BDF_MAX_FACTOR, safety * fmax_);
                        if (!(hysteresis && delta == 0 && factor >= 1.0 && factor < RMSB_SMALL_HYST)) {
                            order += delta; h_abs *= factor; pend = factor; n_equal = 0; lu_valid = false;
                        }
                    }
                }
                if (accepted && !(t < t_end)) finish = true;
            }
        }
        // ================================================================================ outputs of a finished reactor (cold)
        if (finish) {
            if (a.yend && act) a.yend[ir * Ny + gl] = D[0];
            if (a.ysave) {
                double* ysave = a.ysave + (size_t)ir * a.nsave * Ny;
                for (; isave < a.nsave; ++isave) if (act) ysave[(size_t)isave * Ny + gl] = NAN;
            }
            if (gl == 0) {
                if (a.idt) a.idt[ir] = (idt_index >= 0) ? idt_time : NAN;
                a.status[ir] = status;
                if (a.stats) {
                    rmsb_stats s;
                    s.n_steps = n_steps; s.n_rhs = n_rhs; s.n_jac = n_jac; s.n_lu = n_lu;
                    s.n_newton_fail = n_nfail; s.n_err_fail = n_efail; s.last_order = order; s.reserved = 0; s.last_h = h_abs;
                    a.stats[ir] = s;
                }
            }
            phase = PH_FETCH;
        }
    }
}

// dydt / Jacobian of a batch with the same device routines (parity tests exercise the solver's own RHS and Jacobian)
template <int G, int MODE>
__global__ void __launch_bounds__(128) small_eval_kernel(const __grid_constant__ DevMech mp, const __grid_constant__ SmallTables st,
                                                         const __grid_constant__ SmallLayout lay, const __grid_constant__ BatchArgs a) {
    stage_small(mp, st, lay);
    const int lane = threadIdx.x & 31, gl = lane & (G - 1);
    const int grp = threadIdx.x / G, gpb = blockDim.x / G, wpb = blockDim.x / 32;
    const int Ny = a.Ny, Ns = st.Ns;
    GCtx g;
    g.tab = smem_ws + lay.tab;
    g.gs = smem_ws + lay.grp0 + grp * lay.gstride;
    g.gl = gl; g.kind = a.domain_kind; g.Ny = Ny;
    if (gl == 0) (g.gs + lay.o_q)[st.NRP] = 0.0;
    __syncwarp();
    // warp-uniform trip count: the groups of a warp evaluate together, a group past the end of the batch computes on zeros
    for (int64_t base = ((int64_t)blockIdx.x * wpb + (threadIdx.x >> 5)) * (32 / G); base < a.n; base += (int64_t)gridDim.x * gpb) {
        const int64_t ir = base + (lane / G);
        const bool live = ir < a.n;
        g.mine = live;
        g.p = (live && a.p) ? a.p + ir * a.p_stride : nullptr;
        g.cond0 = live ? a.cond[ir * a.ncond] : 1.0;
        g.invV0 = (a.domain_kind == RMSB_DOMAIN_CONSTANT_V) ? 1.0 / g.cond0 : 0.0;
        g.pG = (g.p && gl < Ns) ? g.p[gl] : 0.0;
        g.is_const = 0;
        for (int k = 0; k < a.nconst; ++k) if (a.constspc[k] - 1 == gl) g.is_const = 1;
        const double yv = (live && gl < Ny) ? a.y[ir * Ny + gl] : 1.0;
        if (MODE == 0) {
            const double f = small_rhs<G, double, false>(st, lay, g, yv, nullptr);
            if (live && gl < Ny) a.out[ir * Ny + gl] = f;
        } else {
            small_jac<G>(st, lay, g, yv, a.nconst, a.constspc);
            const double* Js = g.gs + lay.o_J;
            if (live && gl < Ny) {
                double* J = a.out + (size_t)ir * Ny * Ny;
                for (int col = 0; col < Ny; ++col) J[col * Ny + gl] = Js[gl + col * lay.LD];
            }
            __syncwarp();
        }
    }
}

}  // namespace

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
int rmsb_small_build(rmsb_mech* h, const std::vector<int>& rx, const std::vector<int>& srptr, const std::vector<int>& srrxn,
                     const std::vector<int>& kcptr, const std::vector<int>& kcspc, const std::vector<double>& kccoef,
                     const std::vector<double>& nrp, const std::vector<int>& ktype, const std::vector<int>& koff,
                     const std::vector<double>& kpar, const std::vector<double>& fwd, const std::vector<double>& rev,
                     const std::vector<int>& nasa_nr, const std::vector<int>& nasa_nc, const std::vector<double>& nasa_tmax,
                     const std::vector<double>& nasa_coef, int maxr) {
    SmallTables& st = h->st;
    std::memset(&st, 0, sizeof(st));
    const DevMech& dm = h->dm;
    const int Ns = dm.Ns, Nr = dm.Nr;
    if (Ns > SM_SP - 2 || Nr > SM_MAXNR || dm.coupled || dm.difflim || dm.phase_kind != RMSB_PHASE_GAS || dm.n_effrxn > 0) return RMSB_OK;
    for (int j = 0; j < Nr; ++j) if (ktype[j] != RMSB_KIN_ARRHENIUS_VEC) return RMSB_OK;
    if (dm.eff_ptr == nullptr) return RMSB_OK;
    const int NRP = (Nr + 31) / 32 * 32;
    int rows = 7;
    for (int s = 0; s < Ns; ++s) for (int r = 0; r < nasa_nr[s]; ++r) if (nasa_nc[(size_t)r * Ns + s] == 9) rows = 9;
    int max_nre = 1, max_npr = 1, max_kc = 1, max_sr = 0;
    for (int j = 0; j < Nr; ++j) {
        int nre = 0, npr = 0;
        for (int k = 0; k < 4; ++k) { if (rx[(size_t)k * Nr + j] >= 0) ++nre; if (rx[(size_t)(k + 4) * Nr + j] >= 0) ++npr; }
        max_nre = std::max(max_nre, nre); max_npr = std::max(max_npr, npr);
        max_kc = std::max(max_kc, kcptr[j + 1] - kcptr[j]);
        for (int k = kcptr[j]; k < kcptr[j + 1]; ++k) if (std::fabs(kccoef[k]) > 7.0 || kccoef[k] != std::floor(kccoef[k])) return RMSB_OK;
        if (std::fabs(nrp[j]) > 7.0) return RMSB_OK;
    }
    if (max_kc > 8) return RMSB_OK;
    for (int s = 0; s < Ns; ++s) max_sr = std::max(max_sr, srptr[s + 1] - srptr[s]);
    const int sr_words = (max_sr + 7) / 8;
    st.Ns = Ns; st.Nr = Nr; st.NRP = NRP; st.maxr = maxr; st.rows = rows;
    st.max_nre = max_nre; st.max_npr = max_npr; st.max_kc = max_kc; st.sr_words = sr_words;
    int off = 0;
    st.o_tmax = off; off += std::max(0, maxr - 1) * SM_SP;
    st.o_nasa = off; off += maxr * rows * SM_SP;
    st.o_arr = off; off += 3 * SM_MAXNR;
    st.o_rx = off; off += SM_MAXNR;
    st.o_kc = off; off += SM_MAXNR;
    st.o_kcc = off; off += SM_MAXNR;
    st.o_sr = off; off += sr_words * SM_SP;
    st.tab_doubles = off;
    std::vector<double> tab(off, 0.0);
    for (int r = 0; r + 1 < maxr; ++r)
        for (int s = 0; s < SM_SP; ++s)
            tab[st.o_tmax + r * SM_SP + s] = (s < Ns && r < nasa_nr[s] - 1) ? nasa_tmax[(size_t)r * Ns + s] : INFINITY;
    for (int s = 0; s < Ns; ++s)
        for (int r = 0; r < maxr; ++r) {
            const int rr = std::min(r, nasa_nr[s] - 1);                    // ranges past a species' last one repeat it (never selected)
            const int nc = nasa_nc[(size_t)rr * Ns + s];
            for (int c = 0; c < nc; ++c) {
                const int row = (rows == 9 && nc == 7) ? c + 2 : c;
                tab[st.o_nasa + (r * rows + row) * SM_SP + s] = nasa_coef[((size_t)rr * 9 + c) * Ns + s];
            }
        }
    unsigned long long* rxw = reinterpret_cast<unsigned long long*>(tab.data() + st.o_rx);
    unsigned long long* kcw = reinterpret_cast<unsigned long long*>(tab.data() + st.o_kc);
    unsigned long long* kcc = reinterpret_cast<unsigned long long*>(tab.data() + st.o_kcc);
    for (int j = 0; j < SM_MAXNR; ++j) {
        unsigned long long w = ~0ULL, kw = 0ULL, kc = 0ULL;
        for (int e = 0; e < 8; ++e) kc |= 8ULL << (4 * e);                 // coefficient 0
        kc |= 8ULL << 32;                                                  // nrp 0
        if (j < Nr) {
            const double* a = kpar.data() + koff[j];
            tab[st.o_arr + j] = a[0]; tab[st.o_arr + SM_MAXNR + j] = a[1]; tab[st.o_arr + 2 * SM_MAXNR + j] = a[2];
            w = 0ULL;
            for (int k = 0; k < 8; ++k) {
                const int s = rx[(size_t)k * Nr + j];
                w |= (unsigned long long)(s < 0 ? 0xff : s) << (8 * k);
            }
            kc = 0ULL;
            for (int e = 0; e < 8; ++e) {
                const int k = kcptr[j] + e;
                int spc = 0, cf = 0;
                if (k < kcptr[j + 1]) { spc = kcspc[k]; cf = (int)kccoef[k]; }
                kw |= (unsigned long long)spc << (8 * e);
                kc |= (unsigned long long)(cf + 8) << (4 * e);
            }
            kc |= (unsigned long long)((int)nrp[j] + 8) << 32;
            if (fwd[j] != 0.0) kc |= 1ULL << 40;
            if (rev[j] != 0.0) kc |= 1ULL << 41;
        }
        rxw[j] = w; kcw[j] = kw; kcc[j] = kc;
    }
    unsigned long long* srw = reinterpret_cast<unsigned long long*>(tab.data() + st.o_sr);
    for (int s = 0; s < SM_SP; ++s)
        for (int w = 0; w < sr_words; ++w) {
            unsigned long long v = 0ULL;
            for (int b = 0; b < 8; ++b) {
                int idx = NRP;                                               // the zero slot
                const int k = (s < Ns) ? srptr[s] + w * 8 + b : -1;
                if (s < Ns && k < srptr[s + 1]) {
                    const int jr = srrxn[k];
                    const int j = jr & 0x7fffffff;
                    idx = (jr < 0) ? NRP + 1 + j : j;                        // reactant slot: subtract
                }
                v |= (unsigned long long)idx << (8 * b);
            }
            srw[w * SM_SP + s] = v;
        }
    int rc = rmsb_upload(h, tab, &st.tab);
    if (rc != RMSB_OK) return rc;
    st.ok = 1;
    return RMSB_OK;
}

static int small_group(int Ny) { return Ny <= 16 ? 16 : 32; }

bool rmsb_small_eligible(const rmsb_mech* h, int kind, int Ny, int flags) {
    if (!h->st.ok) return false;
    if (kind != RMSB_DOMAIN_CONSTANT_V && kind != RMSB_DOMAIN_CONSTANT_P) return false;
    if (Ny > 32) return false;
    if (flags & (8 | 32)) return false;       // bit 3: diagnostic step recomputation; bit 5: force the warp-per-reactor kernel
    return true;
}

static SmallLayout small_layout(const rmsb_mech* h, int Ny, int G, int groups, bool with_lu) {
    const SmallTables& st = h->st;
    SmallLayout l;
    l.LD = Ny | 1;
    l.NRS = (st.Nr + G - 1) / G;
    l.qplane = 2 * st.NRP + 2;
    l.tab = MECH_DOUBLES;
    l.grp0 = (l.tab + st.tab_doubles + 1) & ~1;
    int o = 0;
    l.o_q = o; o += l.qplane;
    l.o_J = o; o += Ny * l.LD;
    const int scratch = 3 * SM_SP + 2 * st.NRP + std::max(h->n_dt, 2 * l.qplane);
    l.o_W = o; o += std::max(with_lu ? Ny * l.LD : 0, scratch);
    l.o_ru = o; o += 36;
    l.o_perm = o; o += (Ny + 1) / 2 + 1;
    l.o_dinv = o; o += Ny;
    l.gstride = (o + 1) & ~1;
    (void)groups;
    return l;
}

template <class K>
static int small_config(K kernel, const rmsb_mech* h, const SmallLayout& lay, int groups, size_t* smem, int* bps, int* sms, int threads = 128) {
    cudaDeviceProp prop;
    CU(rmsb_cached_props(h->device, &prop));
    *smem = sizeof(double) * ((size_t)lay.grp0 + (size_t)groups * lay.gstride);
    if (*smem > prop.sharedMemPerBlockOptin) return fail(RMSB_ERR_UNSUPPORTED, "small-mechanism path: shared memory");
    CU(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)*smem));
    int b = 0;
    CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&b, kernel, threads, *smem));
    *bps = b < 1 ? 1 : b;
    *sms = prop.multiProcessorCount;
    return RMSB_OK;
}

// ---- work order -------------------------------------------------------------------------------------------------------------
// The two reactors of a warp advance in lockstep (the step-control, Jacobian and LU regions run when either needs them), so a warp
// wastes least when its reactors behave alike, and the sweep finishes earliest when the expensive reactors start first.  Both are
// served by handing the reactors out in order of descending initial temperature (measured on the C2 sweep: +5 %; sorting by the
// true cost of a pilot solve gives +14 %, scripts/gpu_order_study.py).  Results do not depend on the order: a reactor's integration
// never looks at its neighbour.  Layout of the work buffer behind the 256-byte header:
//   [order int32 n][keys uint64 n][keys_out uint64 n][idx int32 n][cub temporary storage]
constexpr int64_t ORDER_MIN_N = 4096;
__global__ void order_keys_kernel(const double* __restrict__ y0, int Ny, int iT, int64_t n, unsigned long long* keys, int* idx) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const double T = y0[i * Ny + iT];
        keys[i] = (T > 0.0 && T == T) ? (unsigned long long)__double_as_longlong(T) : 0ULL;     // positive doubles order like their bits
        idx[i] = (int)i;
    }
}
static size_t al256(size_t b) { return (b + 255) & ~(size_t)255; }
static size_t order_temp_bytes(int64_t n) {
    size_t tb = 0;
    cub::DeviceRadixSort::SortPairsDescending(nullptr, tb, (const unsigned long long*)nullptr, (unsigned long long*)nullptr, (const int*)nullptr, (int*)nullptr, (int)n);
    return tb;
}
size_t rmsb_small_work_bytes(const rmsb_mech* h, int kind, int Ny, int64_t n) {
    if (!rmsb_small_eligible(h, kind, Ny, 0) || n < ORDER_MIN_N || n > 0x7fffffff) return 0;
    return al256(4 * (size_t)n) * 2 + al256(8 * (size_t)n) * 2 + al256(order_temp_bytes(n));
}

template <int G>
static int small_solve_G(const rmsb_mech* h, const SolveArgs& a_in, cudaStream_t stream) {
    SolveArgs a = a_in;
    // flags bit 6: the caller has ordered the reactors itself (e.g. by the cost a previous sweep measured): keep that order
    if (!(a.flags & 64) && a.n >= ORDER_MIN_N && a.n <= 0x7fffffff && a.work_bytes >= 256 + rmsb_small_work_bytes(h, a.domain_kind, a.Ny, a.n)) {
        char* base = reinterpret_cast<char*>(a.counter) + 256;
        int* order = reinterpret_cast<int*>(base); base += al256(4 * (size_t)a.n);
        unsigned long long* keys = reinterpret_cast<unsigned long long*>(base); base += al256(8 * (size_t)a.n);
        unsigned long long* keys_out = reinterpret_cast<unsigned long long*>(base); base += al256(8 * (size_t)a.n);
        int* idx = reinterpret_cast<int*>(base); base += al256(4 * (size_t)a.n);
        size_t tb = order_temp_bytes(a.n);
        order_keys_kernel<<<(int)std::min<int64_t>((a.n + 255) / 256, 1184), 256, 0, stream>>>(a.y0, a.Ny, h->st.Ns, a.n, keys, idx);
        ++g_launches;
        CU(cub::DeviceRadixSort::SortPairsDescending(base, tb, keys, keys_out, idx, order, (int)a.n, 0, 64, stream));
        a.order = order;
    }
    const int groups = RMSB_SMALL_THREADS / G;
    const SmallLayout lay = small_layout(h, a.Ny, G, groups, true);
    size_t smem; int bps, sms;
    int rc = small_config(small_solve_kernel<G>, h, lay, groups, &smem, &bps, &sms, RMSB_SMALL_THREADS);
    if (rc != RMSB_OK) return rc;
    const int64_t resident = (int64_t)sms * bps, need = (a.n + groups - 1) / groups;
    const int grid = (int)std::max<int64_t>(1, std::min(resident, need));
    CU(cudaMemsetAsync(a.counter, 0, sizeof(unsigned long long), stream));
    small_solve_kernel<G><<<grid, RMSB_SMALL_THREADS, smem, stream>>>(h->dm, h->st, lay, a);
    ++g_launches;
    CU(cudaGetLastError());
    return RMSB_OK;
}

int rmsb_small_launch_solve(const rmsb_mech* h, const SolveArgs& a, cudaStream_t st) {
    return small_group(a.Ny) == 16 ? small_solve_G<16>(h, a, st) : small_solve_G<32>(h, a, st);
}

template <int G, int MODE>
static int small_eval_G(const rmsb_mech* h, const BatchArgs& a, cudaStream_t stream) {
    const int groups = 128 / G;
    const SmallLayout lay = small_layout(h, a.Ny, G, groups, false);
    size_t smem; int bps, sms;
    int rc = small_config(small_eval_kernel<G, MODE>, h, lay, groups, &smem, &bps, &sms);
    if (rc != RMSB_OK) return rc;
    const int64_t cap = (int64_t)sms * bps * 4, need = (a.n + groups - 1) / groups;
    const int grid = (int)std::max<int64_t>(1, std::min(cap, need));
    small_eval_kernel<G, MODE><<<grid, 128, smem, stream>>>(h->dm, h->st, lay, a);
    ++g_launches;
    CU(cudaGetLastError());
    return RMSB_OK;
}

int rmsb_small_launch_eval(const rmsb_mech* h, int mode, const BatchArgs& a, cudaStream_t st) {
    const bool g16 = small_group(a.Ny) == 16;
    if (mode == 0) return g16 ? small_eval_G<16, 0>(h, a, st) : small_eval_G<32, 0>(h, a, st);
    return g16 ? small_eval_G<16, 1>(h, a, st) : small_eval_G<32, 1>(h, a, st);
}
