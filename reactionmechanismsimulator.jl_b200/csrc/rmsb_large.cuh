// rmsb_large.cuh -- large-mechanism path (config C3: ~1000+ species, sparse Jacobian): ONE THREAD BLOCK PER REACTOR.
//
// Constant-k domains with y = ns (ConstantTV, ConstantTAPhi; src/Domain.jl:500-560,660-711): the rate coefficients are
// evaluated once per reactor, the RHS is the mass-action scatter (src/Reactor.jl:465-507), the Jacobian is the analytic
// species block (src/Reactor.jl:899-1366) kept SPARSE on the fixed pattern analysed on the host (rmsb_sparse.h), and the
// Newton matrix I - c*J is factorised by a right-looking sparse LU on that pattern (the reference: GMRES + ILU
// preconditioner, src/Reactor.jl:342-382).  Vectors of length Ns live in a per-block slice of global scratch (L2
// resident); the linear-solve vector, the per-warp column maps and the reductions live in shared memory.
// The integrator is the same NDF/BDF as rmsb_solver.cuh (same constants, Newton test, hysteresis, unit rows,
// Jacobian refresh), written with block-strided loops instead of warp-distributed registers.
#pragma once
#include "rmsb_solver.cuh"
#include "rmsb_tables.h"

namespace rmsb {

// cycle counters of the block-per-reactor solver (thread 0 of every block, summed over blocks):
// 0 rhs, 1 jacobian, 2 factorisation, 3 triangular solves, 4 whole solve, 5 factor: sparse rows, 6 factor: sparse rows + core rows'
// sparse columns, 7 factor: dense core; read back by rmsb_large_profile().  Compile with -DRMSB_LARGE_PROFILE=0 to drop them;
// the routines synchronise the block themselves, the counters add no barrier.
__device__ unsigned long long g_large_prof[8];
#ifndef RMSB_LARGE_PROFILE
#define RMSB_LARGE_PROFILE 1
#endif
#if RMSB_LARGE_PROFILE
#define LPROF_BEGIN() const long long _t0 = clock64()
#define LPROF_END(slot) do { if (threadIdx.x == 0) atomicAdd(&g_large_prof[slot], (unsigned long long)(clock64() - _t0)); } while (0)
#define LPROF_ADD(slot, cycles) do { if (threadIdx.x == 0) atomicAdd(&g_large_prof[slot], (unsigned long long)(cycles)); } while (0)
#else
#define LPROF_BEGIN() do { } while (0)
#define LPROF_END(slot) do { } while (0)
#define LPROF_ADD(slot, cycles) do { } while (0)
#endif

#ifndef RMSB_LARGE_THREADS
#define RMSB_LARGE_THREADS 256
#endif
constexpr int LARGE_THREADS = RMSB_LARGE_THREADS;
constexpr int LARGE_WARPS = LARGE_THREADS / 32;


// per-block scratch layout (doubles)
struct LargeScratch {
    double* base;
    int n, Ns, Nr, n_dt, n_je, nnz, nb, n_ee, vk;      // vk: variable-k domain (ConstantV / ConstantP): thermo and k's inside the RHS
    __device__ __forceinline__ double* kf() const { return base; }
    __device__ __forceinline__ double* kr() const { return kf() + Nr; }
    __device__ __forceinline__ double* q() const { return kr() + Nr; }
    __device__ __forceinline__ double* dterm() const { return q() + Nr; }
    __device__ __forceinline__ double* jv() const { return dterm() + n_dt; }
    __device__ __forceinline__ double* bcolv() const { return jv() + n_je; }            // [nb][n]  border columns d f_i / d y_{Ns+b}
    __device__ __forceinline__ double* browv() const { return bcolv() + nb * n; }       // [nb][Ns] border rows    d f_{Ns+b} / d n_c
    __device__ __forceinline__ double* effv() const { return browv() + nb * Ns; }       // [n_ee]   third-body efficiency entries
    __device__ __forceinline__ double* gb() const { return effv() + n_ee; }             // vk: [3Ns] G/(RT), Dual2-capable
    __device__ __forceinline__ double* ub() const { return gb() + 3 * Ns * vk; }        // vk: [3Ns] U or H
    __device__ __forceinline__ double* qd() const { return ub() + 3 * Ns * vk; }        // vk: [3Nr] net rates
    __device__ __forceinline__ double* fd() const { return qd() + 3 * Nr * vk; }        // vk: [3n]  dy/dt with the T / P partials
    __device__ __forceinline__ double* us() const { return fd() + 3 * n * vk; }         // vk: [Ns]  U_i (H_i) values for the Jacobian
    __device__ __forceinline__ double* cpb() const { return us() + Ns * vk; }           // vk: [Ns]  cp_i/R
    __device__ __forceinline__ double* a() const { return cpb() + Ns * vk; }
    __device__ __forceinline__ double* vec(int k) const { return a() + nnz + (size_t)k * n; }
    // vec: 0..7 D, 8 y, 9 ypred, 10 psi, 11 inverse weights, 12 dvec, 13 ynew, 14 f, 15 dy, 16 cs, 17 G, 18 diffusivities
};
constexpr int LARGE_NVEC = 19;
__host__ __device__ inline size_t large_scratch_doubles(int n, int Ns, int Nr, int n_dt, int n_je, int nnz, int nb, int n_ee, int vk) {
    return (size_t)3 * Nr + n_dt + n_je + (size_t)nb * (n + Ns) + n_ee + (size_t)vk * (8 * Ns + 3 * Nr + 3 * n) + nnz + (size_t)LARGE_NVEC * n + 8;
}
// shared memory: [DevMech][64 reduction doubles][36 RU][n x (linear-solve vector)][LARGE_WARPS * n int column map]
__host__ __device__ inline size_t large_smem_bytes(int n) {
    return (size_t)(MECH_DOUBLES + 64 + 36 + n) * sizeof(double) + (size_t)LARGE_WARPS * n * sizeof(int);
}

__device__ __forceinline__ double block_sum(double v, double* red) {
    v = warp_sum(v);
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    __syncthreads();
    if (lane == 0) red[wid] = v;
    __syncthreads();
    double t = 0.0;
#pragma unroll
    for (int k = 0; k < LARGE_WARPS; ++k) t += red[k];
    return t;
}

__device__ __forceinline__ Dual2 block_sum(Dual2 v, double* red) {
    return Dual2(block_sum(v.v, red), block_sum(v.a, red), block_sum(v.b, red));
}

struct LargeCtx {
    LargeScratch s;
    LargeTables t;
    DomainArgs d;
    double* red;     // [64]: 0..7 block reductions, 48..55 aux (T P V N C Cv dot(U,dn)) of the last variable-k value pass
    double* ru;      // [36]
    double* x;       // [n] linear-solve vector (permuted order)
    int* map;        // this warp's column map [n]
    double T, V, P;   // V: volume (ConstantTV) or area (ConstantTAPhi); ConstantTP reads the volume from the state, P from cond
    __device__ __forceinline__ double vol(const double* y) const { return t.has_v ? y[t.ns + t.bv] : (d.kind == RMSB_DOMAIN_CONSTANT_V ? d.cond[0] : V); }
    __device__ __forceinline__ double* aux() const { return red + 48; }
    __device__ __forceinline__ bool vk() const { return s.vk != 0; }
};

// ---- constant-k set-up: G, diffusivities, kf, krev (setup_constk with block-strided loops) -----------------------------
__device__ void large_setup(const LargeCtx& c, const double* y0) {
    const DevMech& m = mech();
    const int Ns = m.Ns, Nr = m.Nr, tid = threadIdx.x;
    double* Gs = c.s.vec(17);
    double* diffs = c.s.vec(18);
    const double T = c.T, V = c.vol(y0);
    const double* p = c.d.p;
    for (int s = tid; s < Ns; s += LARGE_THREADS) Gs[s] = p ? p[s] : gibbs_scalar(s, T);
    if (m.difflim) {
        double mu = 0.0;
        if (m.visc_kind == RMSB_VISC_RIEDEL) mu = exp(m.visc[0] + m.visc[1] / T + m.visc[2] * log(T) + m.visc[3] * pow(T, m.visc[4]));
        else if (m.visc_kind == RMSB_VISC_CONSTANT) mu = m.visc[0];
        for (int s = tid; s < Ns; s += LARGE_THREADS) {
            double r = m.diff_r ? G(m.diff_r, s) : 0.0;
            diffs[s] = (r > 0.0) ? RMSB_KB * T / (6 * RMSB_PI * mu * r) : INFINITY;
        }
    }
    double a = 0.0;
    for (int s = tid; s < Ns; s += LARGE_THREADS) a += y0[s];
    const double N = block_sum(a, c.red);
    __syncthreads();
    KinEnv<double> e;
    const double P = (c.d.kind == RMSB_DOMAIN_CONSTANT_TV) ? 1.0e8 : (c.d.kind == RMSB_DOMAIN_CONSTANT_TP ? c.P : 0.0);
    e.T = T; e.logT = log(T); e.RT = RMSB_R * T; e.invRT = 1.0 / (RMSB_R * T); e.P = P;
    const double C0 = (m.phase_kind == RMSB_PHASE_SURFACE) ? m.sitedensity : 1.0e5 / (RMSB_R * T);
    const double logC0 = log(C0);
    const double C = (c.d.kind == RMSB_DOMAIN_CONSTANT_TV) ? N / V : (c.d.kind == RMSB_DOMAIN_CONSTANT_TP ? P / (RMSB_R * T) : 0.0);
    for (int j = tid; j < Nr; j += LARGE_THREADS) {
        double kf, kr;
        constk_rates(j, p, c.d.kind == RMSB_DOMAIN_CONSTANT_TP, Gs, diffs, e, C0, logC0, C, y0, V, T, kf, kr);
        c.s.kf()[j] = kf * G(m.fwd, j);
        c.s.kr()[j] = kr * G(m.rev, j);
    }
    __syncthreads();
}

// ---- dy/dt: y -> f ---------------------------------------------------------------------------------------------------
__device__ void large_rhs(const LargeCtx& c, const double* y, double* f) {
    LPROF_BEGIN();
    const DevMech& m = mech();
    const int Ns = m.Ns, Nr = m.Nr, tid = threadIdx.x;
    double* cs = c.s.vec(16);
    double* qb = c.s.q();
    const double* kfb = c.s.kf();
    const double* krb = c.s.kr();
    const double V = c.vol(y), invV = 1.0 / V;
    for (int s = tid; s < Ns; s += LARGE_THREADS) cs[s] = y[s] * invV;
    __syncthreads();
    if (c.d.kind == RMSB_DOMAIN_CONSTANT_TP && m.n_effrxn > 0) {
        // third-body / falloff reactions see the composition through Ceff: re-evaluated on every call (Domain.jl:924,943-945)
        const double T = c.T, P = c.P;
        KinEnv<double> e;
        e.T = T; e.logT = log(T); e.RT = RMSB_R * T; e.invRT = 1.0 / (RMSB_R * T); e.P = P;
        const double C0 = 1.0e5 / (RMSB_R * T), C = P / (RMSB_R * T);
        double* kfw = c.s.kf();
        double* krw = c.s.kr();
        for (int q = tid; q < m.n_effrxn; q += LARGE_THREADS) {
            const int j = G(m.effrxn, q);
            double kf = eval_kin<double>(j, &e, ceff_of(j, C, y, V), false);
            if (c.d.p) kf = kf * c.d.p[Ns + j];
            const double kr = kf / kc_scalar(j, c.s.vec(17), T, C0);
            kfw[j] = kf * G(m.fwd, j); krw[j] = kr * G(m.rev, j);
        }
        __syncthreads();
    }
    for (int j = tid; j < Nr; j += LARGE_THREADS) qb[j] = net_rate<double>(j, kfb[j], krb[j], cs);
    __syncthreads();
    for (int s = tid; s < Ns; s += LARGE_THREADS) {
        double acc = 0.0;
        for (int k = G(m.sr_ptr, s); k < G(m.sr_ptr, s + 1); ++k) {
            int jr = G(m.sr_rxn, k);
            double qv = qb[jr & 0x7fffffff];
            acc = (jr < 0) ? acc - qv : acc + qv;
        }
        f[s] = acc * V;
    }
    __syncthreads();
    if (c.t.has_v) {                                            // dV/dt = sum(dn/dt)*R*T/P   (Domain.jl:1864), before the zeroing below
        double a = 0.0;
        for (int s = tid; s < Ns; s += LARGE_THREADS) a += f[s];
        const double tot = block_sum(a, c.red);
        if (tid == 0) f[Ns] = tot * RMSB_R * c.T / c.P;
        __syncthreads();
    }
    for (int k = tid; k < c.d.nconst; k += LARGE_THREADS) f[c.d.constspc[k] - 1] = 0.0;
    __syncthreads();
    LPROF_END(0);
}

// ---- dy/dt of the variable-k gas domains: thermo, rate coefficients and the energy balance inside the call -----------------
// ISP = false: ConstantV  y = [ns; T; P], cond = [V]  (src/Domain.jl:1064-1113,1901-1906)
// ISP = true : ConstantP  y = [ns; T; V], cond = [P]  (src/Domain.jl:1141-1192,1950-1953)
// The formulas, reciprocals and accumulation orders are those of rhs_v (rmsb_device.cuh) with block-strided loops; N = Dual2
// carries d/dT (and d/dP) through the complete RHS for the Jacobian's T / P columns (src/Reactor.jl:1469-1477).
// fout: N-capable [n].  store_k keeps kf, kr, U (H), cp/R and the scalars the Jacobian needs.
template <class N, bool ISP>
__device__ void large_rhs_v(const LargeCtx& c, const double* y, double* fout, bool store_k) {
    LPROF_BEGIN();
    const DevMech& m = mech();
    const int Ns = m.Ns, Nr = m.Nr, tid = threadIdx.x;
    typedef SmemNum<N> SM;
    const double* p = c.d.p;
    double* cs = c.s.vec(16);
    double* gb = c.s.gb();
    double* ub = c.s.ub();
    double* qb = c.s.qd();
    const double V = ISP ? y[Ns + 1] : c.d.cond[0];
    const double invV = 1.0 / V;
    N T = seed<N>(y[Ns], 0), P = ISP ? N(c.d.cond[0]) : seed<N>(y[Ns + 1], 1);
    TPow<N> tp(T);
    N RT = RMSB_R * T;
    N invRT = tp.T4 * RMSB_INV_R;
    N Nt = P * V * invRT;
    N C = P * invRT;
    N cpns = 0.0;
    for (int s = tid; s < Ns; s += LARGE_THREADS) {
        N cp, h, sr;
        nasa_hscp(s, tp, val(T), cp, h, sr);
        if (p) h = h + p[s] * invRT;
        N U = ISP ? h * RT : (h - 1.0) * RT;
        SM::st(gb, s, h - sr);
        SM::st(ub, s, U);
        const double nsv = y[s];
        cs[s] = nsv * invV;
        cpns = cpns + cp * nsv;
        if (store_k) { c.s.us()[s] = val(U); c.s.cpb()[s] = val(cp); }
    }
    N cpsum = block_sum(cpns, c.red);                   // ends with a barrier: gb / ub / cs are visible
    KinEnv<N> e;
    e.T = T; e.logT = tp.T6; e.RT = RT; e.invRT = invRT; e.P = P;
    N logC0 = RMSB_LOG_1E5_OVER_R - tp.T6;
    for (int j = tid; j < Nr; j += LARGE_THREADS) {
        N kf;
        if (G(m.ktype, j) == RMSB_KIN_ARRHENIUS_VEC) kf = arr_vec(m.kpar + G(m.koff, j), e);
        else kf = eval_kin<N>(j, &e, ceff_of(j, C, y, V), false);
        N ex = 0.0;
        for (int k = G(m.kc_ptr, j); k < G(m.kc_ptr, j + 1); ++k) ex = ex + G(m.kc_coef, k) * SM::ld(gb, G(m.kc_spc, k));
        ex = ex + G(m.nrp, j) * logC0;
        N kr = kf * dexp(-ex);
        kf = kf * G(m.fwd, j); kr = kr * G(m.rev, j);
        if (p) { kf = kf * p[Ns + j]; kr = kr * p[Ns + j]; }
        if (store_k) { c.s.kf()[j] = val(kf); c.s.kr()[j] = val(kr); }
        SM::st(qb, j, net_rate(j, kf, kr, cs));
    }
    __syncthreads();
    N usum = 0.0, nsum = 0.0;
    for (int s = tid; s < Ns; s += LARGE_THREADS) {
        N acc = 0.0;
        for (int k = G(m.sr_ptr, s); k < G(m.sr_ptr, s + 1); ++k) {
            int jr = G(m.sr_rxn, k);
            N qv = SM::ld(qb, jr & 0x7fffffff);
            acc = (jr < 0) ? acc - qv : acc + qv;
        }
        N dn = acc * V;
        SM::st(fout, s, dn);
        usum = usum + SM::ld(ub, s) * dn;
        nsum = nsum + dn;
    }
    N dotU = block_sum(usum, c.red), sumn = block_sum(nsum, c.red);
    N dT = ISP ? -dotU / (RMSB_R * cpsum) : -dotU / (RMSB_R * (cpsum - Nt));
    N dP = ISP ? sumn * RT / P + dT * V * tp.T4 : sumn * RT * invV + P * tp.T4 * dT;
    if (tid == 0) {
        SM::st(fout, Ns, dT);
        SM::st(fout, Ns + 1, dP);
        if (store_k) {
            double* ax = c.aux();
            ax[0] = val(T); ax[1] = val(P); ax[2] = V; ax[3] = val(Nt); ax[4] = val(C); ax[6] = val(dotU);
            ax[5] = val(cpsum) * (RMSB_R / val(Nt)) - RMSB_R;
        }
    }
    __syncthreads();
    for (int k = tid; k < c.d.nconst; k += LARGE_THREADS) SM::st(fout, c.d.constspc[k] - 1, N(0.0));
    __syncthreads();
    LPROF_END(0);
}

__device__ __forceinline__ void large_rhs_any(const LargeCtx& c, const double* y, double* f) {
    if (!c.vk()) large_rhs(c, y, f);
    else if (c.d.kind == RMSB_DOMAIN_CONSTANT_P) large_rhs_v<double, true>(c, y, f, false);
    else large_rhs_v<double, false>(c, y, f, false);
}

// ---- Jacobian: y -> species-block values on the mechanism's entry list (jv), third-body efficiency entries (effv), border
//      columns (bcolv: V, or T and P / V) and border rows (browv) -----------------------------------------------------------------
__device__ void large_jac(const LargeCtx& c, const double* y) {
    const DevMech& m = mech();
    const LargeTables& t = c.t;
    const int Ns = m.Ns, Nr = m.Nr, tid = threadIdx.x, n = t.n;
    const int kind = c.d.kind;
    const bool isp = kind == RMSB_DOMAIN_CONSTANT_P, isv = kind == RMSB_DOMAIN_CONSTANT_V;
    if (isp) large_rhs_v<double, true>(c, y, c.s.fd(), true);          // value pass: kf, kr, U / H, cp/R, cs, scalars
    else if (isv) large_rhs_v<double, false>(c, y, c.s.fd(), true);
    LPROF_BEGIN();
    double* cs = c.s.vec(16);
    double* dterm = c.s.dterm();
    double* jv = c.s.jv();
    const double* kfb = c.s.kf();
    const double* krb = c.s.kr();
    const double aV = c.vol(y), invV = 1.0 / aV;
    const double aT = c.vk() ? c.aux()[0] : c.T, aP = c.vk() ? c.aux()[1] : c.P;
    if (!c.vk()) {
        for (int s = tid; s < Ns; s += LARGE_THREADS) cs[s] = y[s] * invV;
        __syncthreads();
    }
    for (int tt = tid; tt < m.n_dt; tt += LARGE_THREADS) {
        int jr = G(m.dt_rxn, tt);
        int j = jr & 0x3fffffff;
        double v = (double)G(m.dt_mult, tt) * ((jr >> 30) ? krb[j] : kfb[j]);
        for (int k = 0; k < 3; ++k) { int s = G(m.dt_oth, 3 * tt + k); if (s < 0) break; v *= cs[s]; }
        dterm[tt] = v;
    }
    __syncthreads();
    for (int e = tid; e < m.n_je; e += LARGE_THREADS) {
        double acc = 0.0;
        for (int k = G(m.je_ptr, e); k < G(m.je_ptr, e + 1); ++k) {
            int ti = G(m.jt_idx, k);
            double v = dterm[ti & 0x7fffffff];
            acc = (ti < 0) ? acc - v : acc + v;
        }
        jv[e] = acc;
    }
    double* bcol = c.s.bcolv();
    double* brow = c.s.browv();
    double* vcol = bcol + (size_t)t.bv * n;              // the volume's column (has_v)
    if (t.has_v) {
        // V column of the species block (src/Reactor.jl:1368-1413)
        for (int s = tid; s < Ns; s += LARGE_THREADS) {
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
            vcol[s] = acc;
        }
    }
    for (int k = tid; k < t.n_ee; k += LARGE_THREADS) c.s.effv()[k] = 0.0;
    __syncthreads();
    // third-body efficiencies: dk/dn_i = dk/dCeff*(eff_i-1)/V (+ dk/dV for the volume column), src/Domain.jl:2151-2167,
    // src/Reactor.jl:1418-1467.  One warp walks the efficiency reactions in order (lanes over the listed species, the slots of
    // a reaction in sequence): every entry is accumulated in a fixed order, no atomics.
    if (m.n_effrxn > 0 && t.n_ee > 0 && kind != RMSB_DOMAIN_CONSTANT_TAPHI && tid < 32) {
        const int lane = tid;
        double* effv = c.s.effv();
        const double aC = c.vk() ? c.aux()[4] : aP / (RMSB_R * aT);
        KinEnv<double> e;
        e.T = aT; e.logT = log(aT); e.RT = RMSB_R * aT; e.invRT = 1.0 / (RMSB_R * aT); e.P = aP;
        for (int q = 0; q < m.n_effrxn; ++q) {
            const int j = G(m.effrxn, q);
            const double Kc = kfb[j] / krb[j];
            double acc = 0.0;
            for (int k = G(m.eff_ptr, j); k < G(m.eff_ptr, j + 1); ++k) acc += cs[G(m.eff_spc, k)] * G(m.eff_val, k);
            const double Ceff = aC + acc;
            const double dkdC = dk_dCeff(j, &e, Ceff);
            double pa = 1.0, pb = 1.0;
            for (int k = 0; k < 4; ++k) { int sp = G(m.rx, k * Nr + j); if (sp < 0) break; pa *= cs[sp]; }
            for (int k = 4; k < 8; ++k) { int sp = G(m.rx, k * Nr + j); if (sp < 0) break; pb *= cs[sp]; }
            const double pf = pa * aV, pr = pb * aV;
            const int e0 = G(m.eff_ptr, j), ne = G(m.eff_ptr, j + 1) - e0;
            const int ncol = ne + (t.has_v ? 1 : 0);
            const int* idx = t.ee_idx + G(t.ee_ptr, q);
            for (int cc = lane; cc < ncol; cc += 32) {
                double dkf;
                if (cc < ne) dkf = dkdC * (G(m.eff_val, e0 + cc) / aV);
                else dkf = dkdC * (aC - Ceff) / aV;
                const double dvf = dkf * pf, dvr = (dkf / Kc) * pr;
                for (int k = 0; k < 8; ++k) {
                    const int sp = G(m.rx, k * Nr + j);
                    if (sp < 0) continue;
                    const double dv = (k < 4) ? (dvr - dvf) : (dvf - dvr);          // reactant rows lose the forward gain, product rows win it
                    if (cc < ne) effv[G(idx, cc * 8 + k)] += dv;
                    else vcol[sp] += dv;
                }
            }
            __syncwarp();
        }
    }
    __syncthreads();
    if (t.nb > 0) {
        const double* effv = c.s.effv();
        if (kind == RMSB_DOMAIN_CONSTANT_TP) {
            // V row = column sums * R*T/P (Domain.jl:2211-2214)
            const double rtp = RMSB_R * c.T / c.P;
            for (int col = tid; col < Ns; col += LARGE_THREADS) {
                double acc = 0.0;
                for (int e = G(t.je_colptr, col); e < G(t.je_colptr, col + 1); ++e) acc += jv[e];
                for (int k = G(t.ee_colptr, col); k < G(t.ee_colptr, col + 1); ++k) acc += effv[G(t.ee_bycol, k)];
                brow[col] = acc * rtp;
            }
            double a = 0.0;
            for (int s = tid; s < Ns; s += LARGE_THREADS) a += vcol[s];
            const double tot = block_sum(a, c.red);
            if (tid == 0) vcol[Ns] = tot * rtp;                                      // entry (V, V)
        } else if (c.vk()) {
            const double* us = c.s.us();
            const double* cpb = c.s.cpb();
            const double aN = c.aux()[3], aCv = c.aux()[5], adotU = c.aux()[6];
            const double aCp = aCv + RMSB_R;
            const double den = isp ? aN * aCp : aN * aCv;
            const double dTdt = -adotU / den;
            double* trow = brow;                       // border 0: T
            double* prow = brow + Ns;                  // border 1: P (ConstantV) or V (ConstantP)
            for (int col = tid; col < Ns; col += LARGE_THREADS) {                    // Domain.jl:2312-2316 / 2480-2490
                double a = 0.0, sm = 0.0;
                for (int e = G(t.je_colptr, col); e < G(t.je_colptr, col + 1); ++e) { const double v = jv[e]; a += us[G(m.je_row, e)] * v; sm += v; }
                for (int k = G(t.ee_colptr, col); k < G(t.ee_colptr, col + 1); ++k) {
                    const int id = G(t.ee_bycol, k);
                    const double v = effv[id];
                    a += us[G(t.ee_row, id)] * v; sm += v;
                }
                const double dC = cpb[col] * RMSB_R / aN;
                double jt;
                if (isp) jt = -a / den - dTdt * (dC / aCp);
                else jt = -a / den + adotU / (aN * aCv * aCv) * dC;
                trow[col] = jt;
                prow[col] = isp ? sm * RMSB_R * aT / aP + aV / aT * jt : sm * RMSB_R * aT / aV + aP / aT * jt;
            }
            if (isp) {                                                                // the analytic V column's T and V rows
                double a = 0.0, sm = 0.0;
                for (int s = tid; s < Ns; s += LARGE_THREADS) { const double v = vcol[s]; a += us[s] * v; sm += v; }
                a = block_sum(a, c.red); sm = block_sum(sm, c.red);
                if (tid == 0) {
                    const double jt = -a / den;
                    vcol[Ns] = jt;
                    vcol[Ns + 1] = sm * RMSB_R * aT / aP + aV / aT * jt + dTdt / aT;
                }
            }
        }
        __syncthreads();
    }
    // constant-species rows are zero (src/Domain.jl:2216-2218,2318-2320,3053-3055) -- after the thermodynamic-variable rows have been
    // formed from the complete columns, as in the reference
    if (c.d.nconst > 0) {
        for (int e = tid; e < m.n_je; e += LARGE_THREADS) {
            int r = G(m.je_row, e);
            for (int k = 0; k < c.d.nconst; ++k) if (c.d.constspc[k] - 1 == r) jv[e] = 0.0;
        }
        for (int e = tid; e < t.n_ee; e += LARGE_THREADS) {
            int r = G(t.ee_row, e);
            for (int k = 0; k < c.d.nconst; ++k) if (c.d.constspc[k] - 1 == r) c.s.effv()[e] = 0.0;
        }
        if (t.has_v) for (int k = tid; k < c.d.nconst; k += LARGE_THREADS) vcol[c.d.constspc[k] - 1] = 0.0;
        __syncthreads();
    }
    LPROF_END(1);
    // T (and P) columns: the derivative of the complete RHS (src/Reactor.jl:1469-1477), rows of constant species already zero
    if (c.vk()) {
        double* fd = c.s.fd();
        if (isp) large_rhs_v<Dual2, true>(c, y, fd, false);
        else large_rhs_v<Dual2, false>(c, y, fd, false);
        for (int i = tid; i < n; i += LARGE_THREADS) {
            bcol[i] = fd[3 * i + 1];
            if (!isp) bcol[(size_t)n + i] = fd[3 * i + 2];
        }
        __syncthreads();
    }
}

// ---- numeric factorisation of A = I - cc*J on the fixed pattern (no pivoting) ------------------------------------------------
// Returns false on a vanishing / non-finite pivot.
__device__ bool large_factor(const LargeCtx& c, double cc) {
    LPROF_BEGIN();
    const DevMech& m = mech();
    const LargeTables& t = c.t;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5, n = t.n;
    double* a = c.s.a();
    const double* jv = c.s.jv();
    for (int p = tid; p < t.nnz; p += LARGE_THREADS) a[p] = 0.0;
    __syncthreads();
    for (int i = tid; i < n; i += LARGE_THREADS) a[G(t.diag, i)] = 1.0;
    __syncthreads();
    for (int e = tid; e < m.n_je; e += LARGE_THREADS) a[G(t.amap, e)] -= cc * jv[e];     // one entry per (row, col): no races
    for (int b = 0; b < t.nb; ++b) {                                                      // dense border columns and rows
        for (int i = tid; i < n; i += LARGE_THREADS) a[G(t.bcol_pos, b * n + i)] -= cc * c.s.bcolv()[(size_t)b * n + i];
        for (int col = tid; col < t.ns; col += LARGE_THREADS) a[G(t.brow_pos, b * t.ns + col)] -= cc * c.s.browv()[(size_t)b * t.ns + col];
    }
    if (t.n_ee > 0) {
        __syncthreads();                                                                  // efficiency entries may coincide with block entries
        for (int e = tid; e < t.n_ee; e += LARGE_THREADS) a[G(t.ee_pos, e)] -= cc * c.s.effv()[e];
    }
    __syncthreads();
    // (a) sparse rows (i < core), up-looking, level by level: row i -= l_ik * (U part of row k) for its L entries k ascending;
    // (b) the same for the L entries k < core of the core rows (all core rows are independent here);
    // (c) dense right-looking LU of the core: every core row holds all core columns as its last n - core entries.
    const int core = t.core;
    double bad = 0.0;
    const long long _t1 = clock64();
    for (int lv = 0; lv <= t.n_lvl_l; ++lv) {
        if (lv == t.n_lvl_l) LPROF_ADD(5, clock64() - _t1);
        int rb, re;
        if (lv < t.n_lvl_l) { rb = G(t.lvl_l_ptr, lv); re = G(t.lvl_l_ptr, lv + 1); }
        else { rb = 0; re = n - core; }                                  // phase (b): the core rows
        for (int r = rb + wid; r < re; r += LARGE_WARPS) {
            const int i = (lv < t.n_lvl_l) ? G(t.lvl_l_rows, r) : core + r;
            if ((lv < t.n_lvl_l) == (i >= core)) continue;               // core rows wait for phase (b)
            const int pb = G(t.row_ptr, i), pe = G(t.row_ptr, i + 1), pd = G(t.diag, i);
            if (pb == pd) continue;
            for (int p = pb + lane; p < pe; p += 32) c.map[G(t.col, p)] = p;
            __syncwarp();
            // L entries in chunks of 32: their pivot rows' metadata (diagonal position, U length, pivot value) is fetched by
            // the lanes in one go and broadcast per entry; the first 32 U entries of the NEXT pivot row are fetched while the
            // current update is in flight, so the chain per L entry is one read-modify-write of the target row
            bool stop = false;
            for (int p0 = pb; p0 < pd && !stop; p0 += 32) {
                const int pp = p0 + lane;
                int mk = n, mdk = 0, mnk = 0; double mpv = 1.0;
                if (pp < pd) { mk = G(t.col, pp); mdk = G(t.ldiag, pp); mnk = G(t.lnk, pp); if (mk < core) mpv = a[mdk]; }
                const int cnt = (pd - p0 < 32) ? pd - p0 : 32;
                int k = __shfl_sync(FULL, mk, 0), dk = __shfl_sync(FULL, mdk, 0), nk = __shfl_sync(FULL, mnk, 0);
                double pv = __shfl_sync(FULL, mpv, 0);
                int uc = 0; double uv = 0.0;
                if (k < core && lane < nk) { uc = G(t.col, dk + 1 + lane); uv = a[dk + 1 + lane]; }
                for (int q = 0; q < cnt; ++q) {
                    if (k >= core) { stop = true; break; }
                    int kn = n, dkn = 0, nkn = 0, ucn = 0; double pvn = 1.0, uvn = 0.0;
                    if (q + 1 < cnt) {
                        kn = __shfl_sync(FULL, mk, q + 1); dkn = __shfl_sync(FULL, mdk, q + 1); nkn = __shfl_sync(FULL, mnk, q + 1);
                        pvn = __shfl_sync(FULL, mpv, q + 1);
                        if (kn < core && lane < nkn) { ucn = G(t.col, dkn + 1 + lane); uvn = a[dkn + 1 + lane]; }
                    }
                    if (pv == 0.0 || !isfinite(pv)) bad = 1.0;
                    const int p = p0 + q;
                    const double l = a[p] / pv;
                    if (lane < nk) a[c.map[uc]] -= l * uv;
                    for (int u = lane + 32; u < nk; u += 32) a[c.map[G(t.col, dk + 1 + u)]] -= l * a[dk + 1 + u];
                    __syncwarp();
                    if (lane == 0) a[p] = l;
                    __syncwarp();
                    k = kn; dk = dkn; nk = nkn; pv = pvn; uc = ucn; uv = uvn;
                }
            }
        }
        __syncthreads();
    }
    const long long _t2 = clock64();
    LPROF_ADD(6, _t2 - _t1);
    for (int k = core; k < n; ++k) {
        const int dk = G(t.diag, k);
        const double pv = a[dk];
        if (pv == 0.0 || !isfinite(pv)) { bad = 1.0; break; }            // uniform: a[dk] is settled
        const double ipv = 1.0 / pv;
        const int w = n - k - 1;
        if (w == 0) break;
        for (int r = tid; r < w; r += LARGE_THREADS) {
            const int i = k + 1 + r;
            a[G(t.row_ptr, i + 1) - (n - k)] *= ipv;                     // l_ik
        }
        __syncthreads();
        const double* urow = a + dk + 1;                                 // u_kj, j = k+1 .. n-1
        for (int idx = tid; idx < w * w; idx += LARGE_THREADS) {
            const int r = idx / w, q = idx - r * w;
            const int rowend = G(t.row_ptr, k + 2 + r);
            a[rowend - w + q] -= a[rowend - w - 1] * urow[q];
        }
        __syncthreads();
    }
    LPROF_ADD(7, clock64() - _t2);
    const bool ok = block_sum(bad, c.red) == 0.0;
    LPROF_END(2);
    return ok;
}

// ---- solve A x = b with the factors: b (old order) in `b`, x (old order) written to `xout` --------------------------------
__device__ void large_solve(const LargeCtx& c, const double* b, double* xout) {
    LPROF_BEGIN();
    const LargeTables& t = c.t;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5, n = t.n;
    const double* a = c.s.a();
    double* x = c.x;
    for (int i = tid; i < n; i += LARGE_THREADS) x[i] = b[G(t.perm, i)];
    __syncthreads();
    // The sparse rows (i < core) form a shallow dependency graph: level-parallel sweeps, one warp per row.  The core rows
    // are a chain (each depends on all earlier ones): their sparse-column part is reduced for all of them at once, then the
    // dense triangle is swept column by column with the whole block (position arithmetic, no index loads).
    const int core = t.core;
    // ---- forward, unit lower triangle
    for (int lv = 1; lv < t.n_lvl_l; ++lv) {          // level 0 rows have no L entries
        const int rb = G(t.lvl_l_ptr, lv), re = G(t.lvl_l_ptr, lv + 1);
        for (int r = rb + wid; r < re; r += LARGE_WARPS) {
            const int i = G(t.lvl_l_rows, r);
            if (i >= core) continue;
            const int pb = G(t.row_ptr, i), pd = G(t.diag, i);
            double acc = 0.0;
            for (int p = pb + lane; p < pd; p += 32) acc += a[p] * x[G(t.col, p)];
            acc = warp_sum(acc);
            if (lane == 0) x[i] -= acc;
        }
        __syncthreads();
    }
    for (int i = core + wid; i < n; i += LARGE_WARPS) {
        const int pb = G(t.row_ptr, i), pc = G(t.row_ptr, i + 1) - (n - core);   // L entries in sparse columns: [pb, pc)
        double acc = 0.0;
        for (int p = pb + lane; p < pc; p += 32) acc += a[p] * x[G(t.col, p)];
        acc = warp_sum(acc);
        if (lane == 0) x[i] -= acc;
    }
    __syncthreads();
    for (int k = core; k < n - 1; ++k) {
        const double xk = x[k];
        for (int i = k + 1 + tid; i < n; i += LARGE_THREADS) x[i] -= a[G(t.row_ptr, i + 1) - (n - k)] * xk;
        __syncthreads();
    }
    // ---- backward, upper triangle with the diagonal
    for (int j = n - 1; j >= core; --j) {
        const double xj = x[j] / a[G(t.diag, j)];
        __syncthreads();                                   // everyone has read x[j] before it is overwritten
        if (tid == 0) x[j] = xj;
        for (int i = core + tid; i < j; i += LARGE_THREADS) x[i] -= a[G(t.row_ptr, i + 1) - (n - j)] * xj;
        __syncthreads();
    }
    for (int lv = 0; lv < t.n_lvl_u; ++lv) {
        const int rb = G(t.lvl_u_ptr, lv), re = G(t.lvl_u_ptr, lv + 1);
        for (int r = rb + wid; r < re; r += LARGE_WARPS) {
            const int i = G(t.lvl_u_rows, r);
            if (i >= core) continue;
            const int pd = G(t.diag, i), pe = G(t.row_ptr, i + 1);
            double acc = 0.0;
            for (int p = pd + 1 + lane; p < pe; p += 32) acc += a[p] * x[G(t.col, p)];
            acc = warp_sum(acc);
            if (lane == 0) x[i] = (x[i] - acc) / a[pd];
        }
        __syncthreads();
    }
    for (int i = tid; i < n; i += LARGE_THREADS) xout[G(t.perm, i)] = x[i];
    __syncthreads();
    LPROF_END(3);
}

// ---- block-wide vector helpers ------------------------------------------------------------------------------------------
__device__ __forceinline__ double large_norm(const LargeCtx& c, const double* v, const double* iw, double scale_by) {
    double a = 0.0;
    for (int i = threadIdx.x; i < c.t.n; i += LARGE_THREADS) { double q = scale_by * v[i] * iw[i]; a += q * q; }
    return sqrt(block_sum(a, c.red) / (double)c.t.n);
}

__device__ void large_change_D(const LargeCtx& c, int order, double factor) {
    const int tid = threadIdx.x, n = c.t.n;
    __syncthreads();
    if (tid < 36) {
        int i = tid / 6, j = tid - 6 * i;
        double acc = 0.0;
        for (int k = 0; k <= j; ++k) {
            double rik;
            if (k == 0) rik = (i == 0) ? 1.0 : 0.0;
            else { rik = 1.0; for (int l = 1; l <= i; ++l) rik *= ((double)(l - 1) - factor * (double)k) * c_invl[l]; }
            acc += rik * c_U[k * 6 + j];
        }
        c.ru[tid] = acc;
    }
    __syncthreads();
    for (int i = tid; i < n; i += LARGE_THREADS) {
        double dv[BDF_MAX_ORDER + 1], dn[BDF_MAX_ORDER + 1];
#pragma unroll
        for (int k = 0; k <= BDF_MAX_ORDER; ++k) dv[k] = (k <= order) ? c.s.vec(k)[i] : 0.0;
#pragma unroll
        for (int j = 1; j <= BDF_MAX_ORDER; ++j) {
            double acc = 0.0;
#pragma unroll
            for (int k = 0; k <= BDF_MAX_ORDER; ++k) if (k <= order) acc += c.ru[k * 6 + j] * dv[k];
            dn[j] = acc;
        }
#pragma unroll
        for (int j = 1; j <= BDF_MAX_ORDER; ++j) if (j <= order) c.s.vec(j)[i] = dn[j];
    }
    __syncthreads();
}

// ---- the integrator (mirrors bdf_solve_warp; all scalars are block-uniform) -------------------------------------------------
__device__ void large_bdf_solve(const SolveArgs& a, const LargeCtx& c, int64_t ir) {
    LPROF_BEGIN();
    const DevMech& m = mech();
    const int tid = threadIdx.x, n = c.t.n;
    const double rtol = a.rtol, atol = a.atol;
    const double t_end = a.t_end[ir];
    double* const y = c.s.vec(8);
    double* const ypred = c.s.vec(9);
    double* const psi = c.s.vec(10);
    double* const iw = c.s.vec(11);
    double* const dvec = c.s.vec(12);
    double* const ynew = c.s.vec(13);
    double* const f = c.s.vec(14);
    double* const dy = c.s.vec(15);
#define VLOOP(i) for (int i = tid; i < n; i += LARGE_THREADS)
    VLOOP(i) { y[i] = a.y0[ir * n + i]; for (int j = 0; j < BDF_MAX_ORDER + 3; ++j) c.s.vec(j)[i] = 0.0; }
    __syncthreads();
    int n_rhs = 0, n_jac = 0, n_lu = 0, n_steps = 0, n_nfail = 0, n_efail = 0;
    int status = RMSB_SOLVE_OK;
    if (!c.vk()) large_setup(c, y);
    int isave = 0;
    double* ysave = a.ysave ? a.ysave + (size_t)ir * a.nsave * n : nullptr;
    while (isave < a.nsave && a.tsave[isave] <= 0.0) {
        if (ysave) VLOOP(i) ysave[(size_t)isave * n + i] = y[i];
        ++isave;
    }
    large_rhs_any(c, y, f); ++n_rhs;
    double t = 0.0, h_abs;
    {
        double bad = 0.0;
        VLOOP(i) if (!isfinite(f[i])) bad = 1.0;
        if (block_sum(bad, c.red) > 0.0) status = RMSB_SOLVE_NAN;
    }
    if (a.first_step > 0.0) h_abs = a.first_step;
    else {
        VLOOP(i) iw[i] = 1.0 / (atol + fabs(y[i]) * rtol);
        __syncthreads();
        double d0 = large_norm(c, y, iw, 1.0), d1 = large_norm(c, f, iw, 1.0);
        double h0 = (d0 < 1e-5 || d1 < 1e-5) ? 1e-6 : 0.01 * d0 / d1;
        VLOOP(i) ynew[i] = y[i] + h0 * f[i];
        __syncthreads();
        large_rhs_any(c, ynew, dy); ++n_rhs;
        VLOOP(i) dy[i] -= f[i];
        __syncthreads();
        double d2 = large_norm(c, dy, iw, 1.0) / h0;
        double h1 = (d1 <= 1e-15 && d2 <= 1e-15) ? fmax(1e-6, h0 * 1e-3) : sqrt(0.01 / fmax(d1, d2));
        h_abs = fmin(fmin(100.0 * h0, h1), t_end);
    }
    large_jac(c, y); ++n_jac;
    bool lu_valid = false;
    VLOOP(i) { c.s.vec(0)[i] = y[i]; c.s.vec(1)[i] = f[i] * h_abs; }
    __syncthreads();
    int order = 1, n_equal = 0;
    const int max_order = a.max_order;
    const double newton_tol = fmax(10.0 * 2.220446049250313e-16 / rtol, fmin(0.03, sqrt(rtol)));
    const bool scipy_newton = (a.flags & 1) != 0, hysteresis = (a.flags & 2) == 0;
    const int nls_v = (a.flags >> 8) & 0xff;
    const double nls_coef = ldexp(0.1, -(nls_v == 0 ? 3 : nls_v - 1));
    double crate = 1.0, c_jac = 0.0, idt_best = -INFINITY, idt_time = NAN;
    // A factorisation costs ~20 triangular solves here, so the factors are kept while c = h/alpha stays within 30 % of the c they
    // were computed for (CVODE's dgmax rule); the Newton update is scaled by 2/(1 + c/c_lu) as in CVODE to centre the iteration.
    double c_lu = 0.0;
    bool jac_new = true;
    const double dgmax = (a.flags & 16) ? 0.0 : 0.3;           // flags bit 4: refactorise on every change of c (diagnostic)

    while (status == RMSB_SOLVE_OK && t < t_end) {
        if (n_steps >= a.max_steps) { status = RMSB_SOLVE_MAX_STEPS; break; }
        const double min_step = 10.0 * (__longlong_as_double(__double_as_longlong(t) + 1) - t);
        if (h_abs < min_step) { large_change_D(c, order, min_step / h_abs); h_abs = min_step; n_equal = 0; lu_valid = false; }
        bool accepted = false, current_jac = false;
        double t_new = t, h = 0.0, error_norm = 0.0, safety = 0.9;
        while (!accepted) {
            if (h_abs < min_step) { status = RMSB_SOLVE_STEP_TOO_SMALL; break; }
            t_new = t + h_abs;
            if (t_new - t_end > 0.0) {
                t_new = t_end;
                const double hc = t_end - t;
                large_change_D(c, order, hc / h_abs);
                n_equal = 0; lu_valid = false; h_abs = hc;
            }
            h = h_abs;
            const double alpha_o = c_alpha[order];
            VLOOP(i) {
                double s = 0.0, ps = 0.0;
                for (int j = 0; j <= order; ++j) { double v = c.s.vec(j)[i]; s += v; if (j >= 1) ps += v * c_gamma[j]; }
                ypred[i] = s; psi[i] = ps / alpha_o; iw[i] = 1.0 / (atol + rtol * fabs(s));
            }
            __syncthreads();
            const double cc = h / alpha_o;
            bool converged = false;
            int n_iter = 0;
            while (!converged) {
                bool lu_ok = true;
                if (!lu_valid && !jac_new && c_lu != 0.0 && fabs(cc / c_lu - 1.0) <= dgmax) lu_valid = true;   // keep the factors
                if (!lu_valid) {
                    if (c_jac == 0.0) c_jac = cc;
                    else if (!current_jac && (cc > JAC_REFRESH_RATIO * c_jac || cc * JAC_REFRESH_RATIO < c_jac)) {
                        large_jac(c, ypred); ++n_jac; current_jac = true; c_jac = cc;
                    }
                    lu_ok = large_factor(c, cc); lu_valid = true; ++n_lu; crate = 1.0; c_lu = cc; jac_new = false;
                }
                const double gscale = (cc == c_lu) ? 1.0 : 2.0 / (1.0 + cc / c_lu);
                const double nl_tol = scipy_newton ? newton_tol : fmax(newton_tol, nls_coef / c_ec[order]);
                VLOOP(i) { dvec[i] = 0.0; ynew[i] = ypred[i]; }
                __syncthreads();
                double dy_norm_old = -1.0;
                if (lu_ok) {
                    for (int k = 0; k < NEWTON_MAXITER; ++k) {
                        n_iter = k + 1;
                        large_rhs_any(c, ynew, f); ++n_rhs;
                        double bad = 0.0;
                        VLOOP(i) {
                            if (!isfinite(f[i])) bad = 1.0;
                            dy[i] = cc * f[i] - psi[i] - dvec[i];
                        }
                        if (block_sum(bad, c.red) > 0.0) break;
                        __syncthreads();
                        large_solve(c, dy, f);          // f <- solution; unit rows keep their right-hand side (x_i = b_i exactly)
                        VLOOP(i) {
                            bool unit = i < m.Ns && G(m.inert, i) != 0;
                            for (int q = 0; q < c.d.nconst; ++q) if (c.d.constspc[q] - 1 == i) unit = true;
                            if (!unit) dy[i] = f[i] * gscale;
                        }
                        __syncthreads();
                        double dy_norm = large_norm(c, dy, iw, 1.0);
                        double rate = (dy_norm_old < 0.0) ? -1.0 : dy_norm / dy_norm_old;
                        if (!(dy_norm == dy_norm)) break;
                        if (rate >= 0.0) {
                            double rp = rate;
                            for (int q = 1; q < NEWTON_MAXITER - k; ++q) rp *= rate;
                            if (rate >= 1.0 || rp / (1.0 - rate) * dy_norm > nl_tol) break;
                            crate = fmax(0.3 * crate, rate);
                        }
                        VLOOP(i) { ynew[i] += dy[i]; dvec[i] += dy[i]; }
                        __syncthreads();
                        if (scipy_newton) {
                            if (dy_norm == 0.0 || (rate >= 0.0 && rate / (1.0 - rate) * dy_norm < nl_tol)) { converged = true; break; }
                        } else if (dy_norm * fmin(1.0, crate) <= nl_tol) { converged = true; break; }
                        dy_norm_old = dy_norm;
                    }
                }
                if (!converged) {
                    if (current_jac) break;
                    large_jac(c, ypred); ++n_jac; c_jac = cc;
                    lu_valid = false; current_jac = true; jac_new = true;
                }
            }
            if (!converged) {
                ++n_nfail;
                h_abs *= 0.5;
                large_change_D(c, order, 0.5);
                n_equal = 0; lu_valid = false;
                continue;
            }
            safety = 0.9 * (2.0 * NEWTON_MAXITER + 1.0) / (2.0 * NEWTON_MAXITER + (double)n_iter);
            VLOOP(i) iw[i] = 1.0 / (atol + rtol * fabs(ynew[i]));
            __syncthreads();
            error_norm = large_norm(c, dvec, iw, c_ec[order]);
            if (error_norm > 1.0) {
                ++n_efail;
                double factor = fmax(BDF_MIN_FACTOR, safety * fast_pow(error_norm, -1.0f / (float)(order + 1)));
                h_abs *= factor;
                large_change_D(c, order, factor);
                n_equal = 0; lu_valid = false;
            } else accepted = true;
        }
        if (!accepted) break;
        ++n_equal; ++n_steps;
        t = t_new;
        VLOOP(i) {
            const double dv = dvec[i];
            c.s.vec(order + 2)[i] = dv - c.s.vec(order + 1)[i];
            c.s.vec(order + 1)[i] = dv;
            for (int j = order; j >= 0; --j) c.s.vec(j)[i] += c.s.vec(j + 1)[i];
            y[i] = c.s.vec(0)[i];
        }
        __syncthreads();
        if (a.idt_index >= 0) {
            // ignition delay: time of the largest d y[idt_index]/dt at the step ends (slope of the difference polynomial at t)
            double sl = 0.0;
            for (int j = 1; j <= order; ++j) sl += c_gamma[j] * c.s.vec(j)[a.idt_index];
            sl /= h;
            if (sl > idt_best) { idt_best = sl; idt_time = t; }
        }
        while (isave < a.nsave && a.tsave[isave] <= t) {
            const double ts = a.tsave[isave];
            if (ysave) VLOOP(i) {
                double pj = 1.0, acc = c.s.vec(0)[i];
                for (int j = 1; j <= order; ++j) { pj *= (ts - (t - (double)(j - 1) * h)) / (h * (double)j); acc += c.s.vec(j)[i] * pj; }
                ysave[(size_t)isave * n + i] = acc;
            }
            ++isave;
        }
        if (n_equal < order + 1) continue;
        double e_m = INFINITY, e_p = INFINITY;
        if (order > 1) e_m = large_norm(c, c.s.vec(order), iw, c_ec[order - 1]);
        if (order < max_order) e_p = large_norm(c, c.s.vec(order + 2), iw, c_ec[order + 1]);
        double fm = fast_pow(e_m, -1.0f / (float)order);
        double f0 = fast_pow(error_norm, -1.0f / (float)(order + 1));
        double fp = fast_pow(e_p, -1.0f / (float)(order + 2));
        int delta; double fmax_;
        if (fm >= f0 && fm >= fp) { delta = -1; fmax_ = fm; }
        else if (f0 >= fp) { delta = 0; fmax_ = f0; }
        else { delta = 1; fmax_ = fp; }
        double factor = fmin(BDF_MAX_FACTOR, safety * fmax_);
        if (hysteresis && delta == 0 && factor >= 1.0 && factor < 1.2) continue;
        order += delta;
        h_abs *= factor;
        large_change_D(c, order, factor);
        n_equal = 0; lu_valid = false;
    }
    __syncthreads();
    if (a.yend) VLOOP(i) a.yend[ir * n + i] = y[i];
    if (ysave) for (; isave < a.nsave; ++isave) VLOOP(i) ysave[(size_t)isave * n + i] = NAN;
    if (tid == 0) {
        if (a.idt) a.idt[ir] = (a.idt_index >= 0) ? idt_time : NAN;
        a.status[ir] = status;
        if (a.stats) {
            rmsb_stats s;
            s.n_steps = n_steps; s.n_rhs = n_rhs; s.n_jac = n_jac; s.n_lu = n_lu;
            s.n_newton_fail = n_nfail; s.n_err_fail = n_efail; s.last_order = order; s.reserved = 0; s.last_h = h_abs;
            a.stats[ir] = s;
        }
    }
    LPROF_END(4);
#undef VLOOP
}

}  // namespace rmsb
