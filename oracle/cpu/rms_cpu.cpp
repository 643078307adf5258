// rms_cpu.cpp -- C++ CPU restatement of the RMS hot path.  TEST INFRASTRUCTURE / CPU BASELINE ONLY
// (see oracle/__init__.py): it is the fast checker behind tests/ and the CPU comparator of bench.py
// (`cpu_baseline`, `--impl reference`); the product never links, loads or calls it.
//
// The reference (ReactionMechanismSimulator.jl, pure Julia) cannot be built here, so this is a "port":
// every routine cites the reference file:line it follows.  It is written as straightforward sequential
// per-reactor code in the reference's own loop structure (reaction loop with scatter, hand-ordered
// Jacobian scatter), independently of the CUDA implementation (which gathers per species / per entry).
// Validated against the NumPy oracle (oracle/rms_oracle.py, itself pinned to the reference's golden
// values) and, for the integrator, against scipy.integrate.BDF -- tests/test_cpu_oracle.py.
//
// Integrator: the reference calls SUNDIALS CVODE_BDF (third party, not in the reference tree).  The
// stand-in is the variable-order NDF/BDF of Shampine & Reichelt (1997) -- the algorithm of
// scipy.integrate.BDF -- with the reference's analytic Jacobian; same WRMS error control.
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cstring>
#include <limits>
#include <vector>
#ifdef _OPENMP
#include <omp.h>
#endif
#include "../../include/rmsb200.h"

namespace {

constexpr double R = 8.314;           // src/Constants.jl:1
constexpr double Na = 6.022e23;       // :3
constexpr double kB = R / Na;         // :5
constexpr double PI = 3.141592653589793;

// ---- dual numbers (value + d/dT + d/dP) for the ConstantV thermo columns (src/Reactor.jl:1469-1477)
struct D2 {
    double v, a, b;
    D2() : v(0), a(0), b(0) {}
    D2(double x) : v(x), a(0), b(0) {}
    D2(double x, double y, double z) : v(x), a(y), b(z) {}
};
inline D2 operator+(D2 x, D2 y) { return {x.v + y.v, x.a + y.a, x.b + y.b}; }
inline D2 operator-(D2 x, D2 y) { return {x.v - y.v, x.a - y.a, x.b - y.b}; }
inline D2 operator-(D2 x) { return {-x.v, -x.a, -x.b}; }
inline D2 operator*(D2 x, D2 y) { return {x.v * y.v, x.a * y.v + x.v * y.a, x.b * y.v + x.v * y.b}; }
inline D2 operator/(D2 x, D2 y) { double q = x.v / y.v; return {q, (x.a - q * y.a) / y.v, (x.b - q * y.b) / y.v}; }
inline D2 exp_(D2 x) { double e = std::exp(x.v); return {e, e * x.a, e * x.b}; }
inline D2 log_(D2 x) { return {std::log(x.v), x.a / x.v, x.b / x.v}; }
inline D2 log10_(D2 x) { double c = 0.4342944819032518 / x.v; return {std::log10(x.v), x.a * c, x.b * c}; }
inline D2 pow_(D2 x, double n) { double v = std::pow(x.v, n), d = n * v / x.v; return {v, d * x.a, d * x.b}; }
inline D2 exp10_(D2 x) { double v = std::pow(10.0, x.v), d = v * 2.302585092994046; return {v, d * x.a, d * x.b}; }
inline double exp_(double x) { return std::exp(x); }
inline double log_(double x) { return std::log(x); }
inline double log10_(double x) { return std::log10(x); }
inline double pow_(double x, double n) { return std::pow(x, n); }
inline double exp10_(double x) { return std::pow(10.0, x); }
inline double val(double x) { return x; }
inline double val(D2 x) { return x.v; }

struct Mech {
    int Ns, Nr, phase_kind, difflim, maxr, visc_kind;
    double sitedensity, visc[5];
    std::vector<int> nr, nc;            // [Ns], [Ns*maxr]
    std::vector<double> tmax, coef;     // [Ns*maxr], [Ns*maxr*9]
    std::vector<int> rx;                // [Nr*8] 0-based, -1 none
    std::vector<int> ktype, koff, effptr, effspc, effrxn;
    std::vector<double> kpar, effval, rev, fwd, radius, diffr;
    std::vector<std::vector<std::pair<int, double>>> Mrow;  // stoichmatrix rows sorted by species
    std::vector<double> nrp;
    std::vector<char> iseff;
};

// ---- thermo --------------------------------------------------------------------------------------
inline const double* select_poly(const Mech& m, int s, double T, int& ncoef) {  // Thermo.jl:101-111
    int r = m.nr[s] - 1;
    for (int k = 0; k < m.nr[s]; ++k)
        if (T <= m.tmax[s * m.maxr + k]) { r = k; break; }
    ncoef = m.nc[s * m.maxr + r];
    return &m.coef[((size_t)s * m.maxr + r) * 9];
}

template <class N>
inline void hscp(const Mech& m, int s, N T, N& cp, N& h, N& sr) {   // Thermovec.jl:76-119
    int nc;
    const double* c = select_poly(m, s, val(T), nc);
    N T0 = T, T1 = T * T, T2 = T1 * T, T3 = T2 * T, T4 = N(1.0) / T, T6 = log_(T);
    if (nc != 7) {
        N T5 = T4 * T4;
        N ct0 = N(c[0]) * T5, ct1 = N(c[1]) * T4, ct2 = c[2], ct3 = N(c[3]) * T0, ct4 = N(c[4]) * T1, ct5 = N(c[5]) * T2, ct6 = N(c[6]) * T3;
        cp = ct0 + ct1 + ct2 + ct3 + ct4 + ct5 + ct6;
        h = -ct0 + T6 * ct1 + ct2 + N(0.5) * ct3 + N(0.33333333333) * ct4 + N(0.25) * ct5 + N(0.2) * ct6 + N(c[7]) * T4;
        sr = N(-0.5) * ct0 - ct1 + T6 * ct2 + ct3 + N(0.5) * ct4 + N(0.33333333333) * ct5 + N(0.25) * ct6 + N(c[8]);
    } else {
        N ct2 = c[0], ct3 = N(c[1]) * T0, ct4 = N(c[2]) * T1, ct5 = N(c[3]) * T2, ct6 = N(c[4]) * T3;
        cp = ct2 + ct3 + ct4 + ct5 + ct6;
        h = ct2 + N(0.5) * ct3 + N(0.33333333333) * ct4 + N(0.25) * ct5 + N(0.2) * ct6 + N(c[5]) * T4;
        sr = T6 * ct2 + ct3 + N(0.5) * ct4 + N(0.33333333333) * ct5 + N(0.25) * ct6 + N(c[6]);
    }
}

inline double horner(double x, const double* c, int n, int div0) {   // Tools.jl:13-22 with coefs./(1:n)
    double out = 0.0;
    for (int i = n - 1; i >= 0; --i) { out += c[i] / (double)(i + div0); if (i != 0) out *= x; }
    return out;
}

inline double gibbs_scalar(const Mech& m, int s, double T) {   // Thermo.jl:6-8,75-93
    int nc;
    const double* c = select_poly(m, s, T, nc);
    double H, S;
    if (nc == 7) {
        H = horner(T, c, 5, 1) * R * T + c[5] * R;
        S = (c[0] * std::log(T) + T * horner(T, c + 1, 4, 1) + c[6]) * R;
    } else {
        H = ((-c[0] / T + c[1] * std::log(T)) / T + horner(T, c + 2, 5, 1)) * R * T + c[7] * R;
        S = ((-c[0] / (2 * T) - c[1]) / T + c[2] * std::log(T) + T * horner(T, c + 3, 4, 1) + c[8]) * R;
    }
    return H - T * S;
}

// ---- kinetics ------------------------------------------------------------------------------------
template <class N> inline N arr_sca(const double* a, N T) { return N(a[0]) * pow_(T, a[1]) * exp_(N(-a[2]) / (N(R) * T)); }   // Rate.jl:16
template <class N> inline N arr_vec(const double* a, N T) { return N(a[0]) * exp_(N(a[1]) * log_(T) - N(a[2]) * (N(1.0) / (N(R) * T))); }  // Ratevec.jl:24

inline void bounding(double el, const double* x, int n, int& i0, int& i1) {   // Tools.jl:26-40
    if (el <= x[0]) { i0 = 0; i1 = -1; return; }
    for (int i = 0; i < n; ++i)
        if (x[i] >= el) { if (x[i] == el) { i0 = i; i1 = -1; } else { i0 = i - 1; i1 = i; } return; }
    i0 = n - 1; i1 = -1;
}

template <class N> inline N plog_mix(N lowk, N highk, N P, double Plow, double Phigh) {   // Rate.jl:59
    return lowk * exp10_(log10_(P / N(Plow)) / N(std::log10(Phigh / Plow)) * log10_(highk / lowk));
}

template <class N> N plog_block(const double*& c, N T, N P) {
    int nP = (int)c[0];
    const double* Ps = c + 1;
    const double* q = c + 1 + nP;
    int i0, i1;
    bounding(val(P), Ps, nP, i0, i1);
    N lowk = 0.0, highk = 0.0;
    for (int i = 0; i < nP; ++i) {
        int mc = (int)q[0];
        if (i == i0 || i == i1) {
            N acc = 0.0;
            if (mc == 1) acc = arr_sca(q + 1, T);
            else for (int k = 0; k < mc; ++k) acc = acc + arr_sca(q + 1 + 3 * k, T);
            (i == i0 ? lowk : highk) = acc;
        }
        q += 1 + 3 * mc;
    }
    c = q;
    return i1 < 0 ? lowk : plog_mix(lowk, highk, P, Ps[i0], Ps[i1]);
}

template <class N> inline N chebT(int n, N x) {   // Rate.jl:158-176
    if (n == 0) return N(1.0);
    if (n == 1) return x;
    N T0 = 1.0, T1 = x, Tn = 0.0;
    for (int i = 1; i < n; ++i) { Tn = N(2.0) * x * T1 - T0; T0 = T1; T1 = Tn; }
    return Tn;
}

template <class N> N troeF(const double* t, N T, N Pr) {   // Rate.jl:134-146
    if (t[2] == 0.0 && t[1] == 0.0) return N(1.0);
    N Fcent = N(1 - t[0]) * exp_(-T / N(t[1])) + N(t[0]) * exp_(-T / N(t[2]));
    if (t[3] != 0.0) Fcent = Fcent + exp_(N(-t[3]) / T);
    N lF = log10_(Fcent), n = N(0.75) - N(1.27) * lF, c = N(-0.4) - N(0.67) * lF, lPr = log10_(Pr);
    N x = (lPr + c) / (n - N(0.14) * lPr);
    return exp10_(lF / (N(1.0) + x * x));
}

template <class N> N eval_kin(const Mech& m, int j, N T, N P, N Ceff, bool scalar_forms) {
    const double* a = &m.kpar[m.koff[j]];
    switch (m.ktype[j]) {
    case RMSB_KIN_ARRHENIUS_VEC: return scalar_forms ? arr_sca(a, T) : arr_vec(a, T);
    case RMSB_KIN_ARRHENIUS: return arr_sca(a, T);
    case RMSB_KIN_MULTIARRHENIUS: { N o = 0.0; for (int k = 0; k < (int)a[0]; ++k) o = o + arr_sca(a + 1 + 3 * k, T); return o; }   // Rate.jl:68-74
    case RMSB_KIN_PLOG_VEC: {                                                                                                  // Ratevec.jl:204-215
        int nP = (int)a[0], i0, i1;
        const double *Ps = a + 1, *ar = a + 1 + nP;
        bounding(val(P), Ps, nP, i0, i1);
        N lowk = scalar_forms ? arr_sca(ar + 3 * i0, T) : arr_vec(ar + 3 * i0, T);
        if (i1 < 0) return lowk;
        N highk = scalar_forms ? arr_sca(ar + 3 * i1, T) : arr_vec(ar + 3 * i1, T);
        return plog_mix(lowk, highk, P, Ps[i0], Ps[i1]);
    }
    case RMSB_KIN_PLOG: { const double* c = a; return plog_block(c, T, P); }                                                    // Rate.jl:48-60
    case RMSB_KIN_MULTIPLOG: { const double* c = a + 1; N o = 0.0; for (int k = 0; k < (int)a[0]; ++k) o = o + plog_block(c, T, P); return o; }  // Rate.jl:82-88
    case RMSB_KIN_THIRDBODY: return Ceff * arr_sca(a, T);                                                                      // Rate.jl:98
    case RMSB_KIN_LINDEMANN: { N k0 = arr_sca(a + 3, T), kinf = arr_sca(a, T), Pr = k0 * Ceff / kinf; return kinf * Pr / (N(1.0) + Pr); }  // :109-114
    case RMSB_KIN_TROE: { N k0 = arr_sca(a + 3, T), kinf = arr_sca(a, T), Pr = k0 * Ceff / kinf; return kinf * (Pr / (N(1.0) + Pr)) * troeF(a + 6, T, Pr); }  // :129-147
    case RMSB_KIN_CHEBYSHEV_VEC: {                                                                                             // Ratevec.jl:89-100 / Rate.jl:194-209
        double Tmin = a[0], Tmax = a[1], Pmin = a[2], Pmax = a[3];
        int nT = (int)a[4], nP = (int)a[5];
        const double* cf = a + 6;
        N Tred = (N(2.0) / T - N(1.0 / Tmin) - N(1.0 / Tmax)) / N(1.0 / Tmax - 1.0 / Tmin);
        N Pred = (N(2.0) * log10_(P) - N(std::log10(Pmin)) - N(std::log10(Pmax))) / N(std::log10(Pmax) - std::log10(Pmin));
        N k = 0.0;
        if (!scalar_forms) { for (int jp = 0; jp < nP; ++jp) for (int it = 0; it < nT; ++it) k = k + N(cf[it * nP + jp]) * (chebT(it, Tred) * chebT(jp, Pred)); }
        else { for (int it = 0; it < nT; ++it) for (int jp = 0; jp < nP; ++jp) k = k + N(cf[it * nP + jp]) * chebT(it, Tred) * chebT(jp, Pred); }
        return exp10_(k);
    }
    }
    return N(0.0);
}

double dk_dCeff(const Mech& m, int j, double T, double Ceff) {   // Rate.jl:232-262
    const double* a = &m.kpar[m.koff[j]];
    if (m.ktype[j] == RMSB_KIN_THIRDBODY) return arr_sca(a, T);
    double k0 = arr_sca(a + 3, T), kinf = arr_sca(a, T), Pr = k0 * Ceff / kinf;
    if (m.ktype[j] == RMSB_KIN_LINDEMANN || (a[8] == 0.0 && a[7] == 0.0)) return k0 / ((1.0 + Pr) * (1.0 + Pr));
    const double* t = a + 6;
    double Fcent = (1 - t[0]) * std::exp(-T / t[1]) + t[0] * std::exp(-T / t[2]);
    if (t[3] != 0.0) Fcent += std::exp(-t[3] / T);
    double lF = std::log10(Fcent), n = 0.75 - 1.27 * lF, c = -0.4 - 0.67 * lF, lPr = std::log10(Pr), d = 0.14;
    double x = (lPr + c) / (n - d * lPr);
    double F = std::pow(10.0, lF / (1.0 + x * x));
    double dF = -std::log(10.0) * F * lF / ((1.0 + x * x) * (1.0 + x * x)) * 2.0 * x * 0.4342944819032518 / Pr * k0 / kinf / (n - d * lPr) * (1.0 + d * x);
    return k0 / ((1.0 + Pr) * (1.0 + Pr)) * F + kinf * Pr / (1.0 + Pr) * dF;
}

template <class N> inline N ceff_of(const Mech& m, int j, N C, const double* ns, double V) {   // PhaseState.jl:53-58
    if (m.effptr[j + 1] == m.effptr[j]) return C;
    double acc = 0.0;
    for (int k = m.effptr[j]; k < m.effptr[j + 1]; ++k) acc += ns[m.effspc[k]] * m.effval[k];
    return C + N(acc / V);
}

double kc_scalar(const Mech& m, int j, const double* G, double T, double C0) {   // PhaseState.jl:106-129
    const int* r = &m.rx[j * 8];
    double acc = G[r[0]]; int nre = 1;
    for (int k = 1; k < 4 && r[k] >= 0; ++k) { acc += G[r[k]]; ++nre; }
    double dG = 0.0 - acc;
    acc = G[r[4]]; int npr = 1;
    for (int k = 5; k < 8 && r[k] >= 0; ++k) { acc += G[r[k]]; ++npr; }
    dG += acc;
    return std::exp(-(dG) / (R * T)) * std::pow(C0, (double)(npr - nre));
}

double diffusive_rate(const Mech& m, int j, int side, const double* diffs) {   // PhaseState.jl:87-93
    const int* r = &m.rx[j * 8 + side * 4];
    if (r[1] < 0) return std::numeric_limits<double>::infinity();
    return 4.0 * PI * (diffs[r[0]] + diffs[r[1]]) * (m.radius[r[0]] + m.radius[r[1]]) * Na;
}

// ---- per-reactor workspace -----------------------------------------------------------------------
struct Work {
    std::vector<double> cs, G, kf, kr, us, cp, diffs, jac, lu, f, tmp, tmpns;
    std::vector<double> bufd;
    std::vector<D2> bufD;
    std::vector<int> piv;
    template <class N> N* buf();
    void init(const Mech& m, int Ny) {
        bufd.resize(3 * m.Ns + 2 * m.Nr); bufD.resize(3 * m.Ns + 2 * m.Nr);
        cs.resize(m.Ns); G.resize(m.Ns); kf.resize(m.Nr); kr.resize(m.Nr); us.resize(m.Ns); cp.resize(m.Ns); diffs.resize(m.Ns); tmpns.resize(m.Ns);
        jac.assign((size_t)Ny * Ny, 0.0); lu.resize((size_t)Ny * Ny); f.resize(Ny); tmp.resize(Ny); piv.resize(Ny);
    }
};

template <> inline double* Work::buf<double>() { return bufd.data(); }
template <> inline D2* Work::buf<D2>() { return bufD.data(); }

struct Dom {
    int kind, Ny;
    const double* p;
    const int* constspc;   // 1-based
    int nconst;
    const double* cond;
};

// addreactionratecontributions!  src/Reactor.jl:465-507
template <class N>
inline void add_rates(const Mech& m, const double* cs, const N* kf, const N* kr, N* dydt) {
    for (int i = 0; i < m.Nr; ++i) {
        const int* r = &m.rx[i * 8];
        N fR = kf[i] * N(cs[r[0]]);
        for (int k = 1; k < 4 && r[k] >= 0; ++k) fR = fR * N(cs[r[k]]);
        N rR = kr[i] * N(cs[r[4]]);
        for (int k = 5; k < 8 && r[k] >= 0; ++k) rR = rR * N(cs[r[k]]);
        N Rr = fR - rR;
        for (int k = 0; k < 4 && r[k] >= 0; ++k) dydt[r[k]] = dydt[r[k]] - Rr;
        for (int k = 4; k < 8 && r[k] >= 0; ++k) dydt[r[k]] = dydt[r[k]] + Rr;
    }
}

// constant-k setup: src/Domain.jl:42-102,500-560,660-711 + src/PhaseState.jl:184-257
void setup_constk(const Mech& m, const Dom& d, Work& w, const double* ns, double T, double P, double V) {
    for (int s = 0; s < m.Ns; ++s) w.G[s] = gibbs_scalar(m, s, T);
    if (m.difflim) {
        double mu = 0.0;
        if (m.visc_kind == RMSB_VISC_RIEDEL) mu = std::exp(m.visc[0] + m.visc[1] / T + m.visc[2] * std::log(T) + m.visc[3] * std::pow(T, m.visc[4]));
        else if (m.visc_kind == RMSB_VISC_CONSTANT) mu = m.visc[0];
        for (int s = 0; s < m.Ns; ++s) w.diffs[s] = m.diffr[s] > 0.0 ? kB * T / (6 * PI * mu * m.diffr[s]) : std::numeric_limits<double>::infinity();
    }
    double C0 = m.phase_kind == RMSB_PHASE_SURFACE ? m.sitedensity : 1.0e5 / (R * T);
    double N = 0.0;
    for (int s = 0; s < m.Ns; ++s) N += ns[s];
    double C = d.kind == RMSB_DOMAIN_CONSTANT_TP ? P / (R * T) : (d.kind == RMSB_DOMAIN_CONSTANT_TV ? N / V : 0.0);
    for (int j = 0; j < m.Nr; ++j) {
        double kf, kr;
        double Ceff = ceff_of(m, j, C, ns, V);
        if (!m.difflim) {
            kf = eval_kin<double>(m, j, T, P, Ceff, false);
            double ex = 0.0;
            for (auto& kv : m.Mrow[j]) ex += kv.second * (w.G[kv.first] / (R * T));
            ex += m.nrp[j] * std::log(C0);
            kr = kf / std::exp(ex);
        } else {
            kf = eval_kin<double>(m, j, T, P, Ceff, true);
            double Kc = kc_scalar(m, j, w.G.data(), T, C0);
            kr = kf / Kc;
            const int* r = &m.rx[j * 8];
            int nre = 0, npr = 0;
            for (int k = 0; k < 4; ++k) { nre += r[k] >= 0; npr += r[k + 4] >= 0; }
            if (nre == 1) {
                if (npr > 1) { double krd = diffusive_rate(m, j, 1, w.diffs.data()); kr = kr * krd / (kr + krd); kf = Kc * kr; }
            } else if (npr == 1) {
                double kfd = diffusive_rate(m, j, 0, w.diffs.data()); kf = kf * kfd / (kf + kfd); kr = kf / Kc;
            } else if (npr == nre) {
                double kfd = diffusive_rate(m, j, 0, w.diffs.data()), krd = diffusive_rate(m, j, 1, w.diffs.data());
                double kff = kf * kfd / (kf + kfd), krevr = kr * krd / (kr + krd), kfr = Kc * krevr;
                if (kff > kfr) { kf = kfr; kr = krevr; } else { kf = kff; kr = kf / Kc; }
            }
        }
        w.kf[j] = kf * m.fwd[j]; w.kr[j] = kr * m.rev[j];
    }
}

struct Aux { double T, P, V, N, C, Cv, dotU; };

// dydtreactor!  src/Reactor.jl:653-660 with calcthermo (Domain.jl) and calcdomainderivatives! inlined
template <class N>
void rhs(const Mech& m, const Dom& d, Work& w, const N* y, N* dy, bool keep, Aux* ax) {
    const int Ns = m.Ns, Nr = m.Nr;
    for (int i = 0; i < d.Ny; ++i) dy[i] = N(0.0);
    if (d.kind == RMSB_DOMAIN_CONSTANT_V || d.kind == RMSB_DOMAIN_CONSTANT_P) {
        // ConstantV: y = [ns; T; P], cond = [V].  ConstantP: y = [ns; T; V], cond = [P] (Domain.jl:1141-1192,1950-1953):
        // enthalpies instead of internal energies, Cp instead of Cv, dV/dt instead of dP/dt; V carries no derivative seed
        const bool isP = d.kind == RMSB_DOMAIN_CONSTANT_P;
        N T = y[Ns], P = isP ? N(d.cond[0]) : y[Ns + 1];
        double V = isP ? val(y[Ns + 1]) : d.cond[0];
        N Nt = P * N(V) / (N(R) * T), C = Nt / N(V);           // Domain.jl:1068-1070
        N* g = w.buf<N>(); N* us = g + Ns; N* kf = us + Ns; N* kr = kf + Nr;
        double* ns = w.tmpns.data();
        N cpns = 0.0;
        for (int s = 0; s < Ns; ++s) {
            N cp, h, sr;
            hscp(m, s, T, cp, h, sr);
            if (d.p) h = h + N(d.p[s]) / (N(R) * T);
            N G = (h - sr) * (N(R) * T);
            us[s] = isP ? h * (N(R) * T) : (h - N(1.0)) * (N(R) * T);
            g[s] = G / (N(R) * T);
            ns[s] = val(y[s]);
            w.cs[s] = ns[s] / V;
            cpns = cpns + cp * y[s];
            if (keep) { w.us[s] = val(us[s]); w.cp[s] = val(cp); }
        }
        N Cv = cpns * (N(R) / Nt) - N(R);                       // :1103-1105
        N logC0 = log_(N(1.0e5) / (N(R) * T));
        for (int j = 0; j < Nr; ++j) {
            N Ceff = ceff_of(m, j, C, ns, V);
            N k = eval_kin<N>(m, j, T, P, Ceff, false);
            N ex = 0.0;
            for (auto& kv : m.Mrow[j]) ex = ex + N(kv.second) * g[kv.first];
            ex = ex + N(m.nrp[j]) * logC0;
            N krev = k / exp_(ex);
            k = k * N(m.fwd[j]); krev = krev * N(m.rev[j]);
            if (d.p) { k = k * N(d.p[Ns + j]); krev = krev * N(d.p[Ns + j]); }
            kf[j] = k; kr[j] = krev;
            if (keep) { w.kf[j] = val(k); w.kr[j] = val(krev); }
        }
        add_rates<N>(m, w.cs.data(), kf, kr, dy);
        N dotU = 0.0, sumn = 0.0;
        for (int s = 0; s < Ns; ++s) { dy[s] = dy[s] * N(V); dotU = dotU + us[s] * dy[s]; sumn = sumn + dy[s]; }
        if (isP) {
            dy[Ns] = -dotU / (Nt * (Cv + N(R)));                 // Domain.jl:1951-1952
            dy[Ns + 1] = sumn * N(R) * T / P + dy[Ns] * N(V) / T;  // :1953
        } else {
            dy[Ns] = -dotU / (Nt * Cv);                          // Domain.jl:1902
            dy[Ns + 1] = sumn * N(R) * T / N(V) + P / T * dy[Ns];  // :1903
        }
        if (ax) *ax = {val(T), val(P), V, val(Nt), val(C), val(Cv), val(dotU)};
    } else {
        double T = d.cond[0], P, V;
        if (d.kind == RMSB_DOMAIN_CONSTANT_TP) { P = d.cond[1]; V = val(y[Ns]); }
        else { P = d.kind == RMSB_DOMAIN_CONSTANT_TV ? 1.0e8 : 0.0; V = d.cond[1]; }
        double* ns = w.tmpns.data();
        double Nt = 0.0;
        for (int s = 0; s < Ns; ++s) { ns[s] = val(y[s]); w.cs[s] = ns[s] / V; Nt += ns[s]; }
        if (d.kind == RMSB_DOMAIN_CONSTANT_TP) Nt = P * V / (R * T);
        double C = Nt / V;
        if (d.kind == RMSB_DOMAIN_CONSTANT_TP)
            for (int j : m.effrxn) {                              // Domain.jl:924,943-945
                double Ceff = ceff_of(m, j, C, ns, V);
                double kf = eval_kin<double>(m, j, T, P, Ceff, false);
                double kr = kf / kc_scalar(m, j, w.G.data(), T, 1.0e5 / (R * T));
                w.kf[j] = kf * m.fwd[j]; w.kr[j] = kr * m.rev[j];
            }
        N* kf = w.buf<N>(); N* kr = kf + Nr;
        for (int j = 0; j < Nr; ++j) { kf[j] = N(w.kf[j]); kr[j] = N(w.kr[j]); }
        add_rates<N>(m, w.cs.data(), kf, kr, dy);
        N sumn = 0.0;
        for (int s = 0; s < Ns; ++s) { dy[s] = dy[s] * N(V); sumn = sumn + dy[s]; }
        if (d.kind == RMSB_DOMAIN_CONSTANT_TP) dy[Ns] = sumn * N(R * T / P);   // Domain.jl:1864
        if (ax) *ax = {T, P, V, Nt, C, 0.0, 0.0};
    }
    for (int k = 0; k < d.nconst; ++k) dy[d.constspc[k] - 1] = N(0.0);
}

// _jacobianynswrtns!, _jacobianynswrtV!, _jacobianykderiv!  (src/Reactor.jl:874-1467), multiplicity-aware
void jac_side(const Mech& m, int i, int side, double k, const double* cs, double* J, int Ny, int Vind, double V,
              bool do_ns, bool do_V, double dkdx, int xind) {
    const int* a = &m.rx[i * 8 + side * 4];
    const int* b = &m.rx[i * 8 + (1 - side) * 4];
    int na = 0, nb = 0;
    while (na < 4 && a[na] >= 0) ++na;
    while (nb < 4 && b[nb] >= 0) ++nb;
    if (do_ns) {
        for (int q = 0; q < na; ++q) {
            int s = a[q];
            bool seen = false;
            for (int z = 0; z < q; ++z) if (a[z] == s) seen = true;
            if (seen) continue;
            int mult = 0;
            for (int z = 0; z < na; ++z) if (a[z] == s) ++mult;
            double deriv = (double)mult * k;
            bool removed = false;
            for (int z = 0; z < na; ++z) { if (a[z] == s && !removed) { removed = true; continue; } deriv *= cs[a[z]]; }
            for (int z = 0; z < na; ++z) J[a[z] + (size_t)s * Ny] -= deriv;
            for (int z = 0; z < nb; ++z) J[b[z] + (size_t)s * Ny] += deriv;
        }
    }
    if (do_V && na > 1) {
        double deriv = -(double)(na - 1) * k;
        for (int z = 0; z < na; ++z) deriv *= cs[a[z]];
        for (int z = 0; z < na; ++z) J[a[z] + (size_t)Vind * Ny] -= deriv;
        for (int z = 0; z < nb; ++z) J[b[z] + (size_t)Vind * Ny] += deriv;
    }
    if (xind >= 0) {
        double deriv = dkdx;
        for (int z = 0; z < na; ++z) deriv *= cs[a[z]];
        deriv *= V;
        for (int z = 0; z < na; ++z) J[a[z] + (size_t)xind * Ny] -= deriv;
        for (int z = 0; z < nb; ++z) J[b[z] + (size_t)xind * Ny] += deriv;
    }
}

// jacobiany!  src/Domain.jl:2204-2454,3049-3055 -> column-major Ny x Ny in w.jac
void jacobian(const Mech& m, const Dom& d, Work& w, const double* y) {
    const int Ns = m.Ns, Ny = d.Ny;
    std::vector<double> dy(Ny);
    Aux ax;
    rhs<double>(m, d, w, y, dy.data(), true, &ax);
    std::fill(w.jac.begin(), w.jac.end(), 0.0);
    double* J = w.jac.data();
    const bool hasV = d.kind == RMSB_DOMAIN_CONSTANT_TP || d.kind == RMSB_DOMAIN_CONSTANT_P;
    const int Vind = d.kind == RMSB_DOMAIN_CONSTANT_P ? Ns + 1 : Ns;
    for (int i = 0; i < m.Nr; ++i) {
        jac_side(m, i, 0, w.kf[i], w.cs.data(), J, Ny, Vind, ax.V, true, false, 0.0, -1);
        jac_side(m, i, 1, w.kr[i], w.cs.data(), J, Ny, Vind, ax.V, true, false, 0.0, -1);
        if (hasV) {
            jac_side(m, i, 0, w.kf[i], w.cs.data(), J, Ny, Vind, ax.V, false, true, 0.0, -1);
            jac_side(m, i, 1, w.kr[i], w.cs.data(), J, Ny, Vind, ax.V, false, true, 0.0, -1);
        }
        if (m.iseff[i] && d.kind != RMSB_DOMAIN_CONSTANT_TAPHI) {   // Domain.jl:2151-2167
            double Kc = w.kf[i] / w.kr[i];
            double acc = 0.0;
            for (int k = m.effptr[i]; k < m.effptr[i + 1]; ++k) acc += w.cs[m.effspc[k]] * m.effval[k];
            double Ceff = ax.C + acc;
            double dkdC = dk_dCeff(m, i, ax.T, Ceff);
            for (int k = m.effptr[i]; k < m.effptr[i + 1]; ++k) {
                double dkdni = dkdC * (m.effval[k] / ax.V);
                jac_side(m, i, 0, 0.0, w.cs.data(), J, Ny, Vind, ax.V, false, false, dkdni, m.effspc[k]);
                jac_side(m, i, 1, 0.0, w.cs.data(), J, Ny, Vind, ax.V, false, false, dkdni / Kc, m.effspc[k]);
            }
            if (hasV) {
                double dkdV = dkdC * (ax.C - Ceff) / ax.V;
                jac_side(m, i, 0, 0.0, w.cs.data(), J, Ny, Vind, ax.V, false, false, dkdV, Vind);
                jac_side(m, i, 1, 0.0, w.cs.data(), J, Ny, Vind, ax.V, false, false, dkdV / Kc, Vind);
            }
        }
    }
    if (d.kind == RMSB_DOMAIN_CONSTANT_TP) {
        for (int c = 0; c <= Ns; ++c) {
            double a = 0.0;
            for (int i = 0; i < Ns; ++i) a += J[i + (size_t)c * Ny];
            J[Ns + (size_t)c * Ny] = a * R * ax.T / ax.P;
        }
    } else if (d.kind == RMSB_DOMAIN_CONSTANT_V) {
        for (int c = 0; c < Ns; ++c) {
            double dCv = w.cp[c] * R / ax.N, a = 0.0, sm = 0.0;
            for (int i = 0; i < Ns; ++i) { a += w.us[i] * J[i + (size_t)c * Ny]; sm += J[i + (size_t)c * Ny]; }
            double jt = -a / (ax.N * ax.Cv) + ax.dotU / (ax.N * ax.Cv * ax.Cv) * dCv;
            J[Ns + (size_t)c * Ny] = jt;
            J[Ns + 1 + (size_t)c * Ny] = sm * R * ax.T / ax.V + ax.P / ax.T * jt;
        }
    }
    else if (d.kind == RMSB_DOMAIN_CONSTANT_P) {                 // Domain.jl:2480-2490 (w.us holds the enthalpies)
        const double Cp = ax.Cv + R, dTdt = -ax.dotU / (ax.N * Cp);
        for (int c = 0; c < Ns + 2; ++c) {
            if (c == Ns) continue;
            double a = 0.0, sm = 0.0;
            for (int i = 0; i < Ns; ++i) { a += w.us[i] * J[i + (size_t)c * Ny]; sm += J[i + (size_t)c * Ny]; }
            double jt = -a / (ax.N * Cp);
            if (c < Ns) jt -= dTdt * ((w.cp[c] * R / ax.N) / Cp);
            J[Ns + (size_t)c * Ny] = jt;
            J[Ns + 1 + (size_t)c * Ny] = sm * R * ax.T / ax.P + ax.V / ax.T * jt + (c == Ns + 1 ? dTdt / ax.T : 0.0);
        }
    }
    for (int k = 0; k < d.nconst; ++k) for (int c = 0; c < Ny; ++c) J[d.constspc[k] - 1 + (size_t)c * Ny] = 0.0;
    if (d.kind == RMSB_DOMAIN_CONSTANT_P) {   // T column: derivative of the full RHS (Domain.jl:2630)
        std::vector<D2> yd(Ny), dyd(Ny);
        for (int i = 0; i < Ny; ++i) yd[i] = D2(y[i], i == Ns ? 1.0 : 0.0, 0.0);
        rhs<D2>(m, d, w, yd.data(), dyd.data(), false, nullptr);
        for (int i = 0; i < Ny; ++i) J[i + (size_t)Ns * Ny] = dyd[i].a;
    }
    if (d.kind == RMSB_DOMAIN_CONSTANT_V) {   // T and P columns: derivative of the full RHS (Reactor.jl:1469-1477)
        std::vector<D2> yd(Ny), dyd(Ny);
        for (int i = 0; i < Ny; ++i) yd[i] = D2(y[i], i == Ns ? 1.0 : 0.0, i == Ns + 1 ? 1.0 : 0.0);
        rhs<D2>(m, d, w, yd.data(), dyd.data(), false, nullptr);
        for (int i = 0; i < Ny; ++i) { J[i + (size_t)Ns * Ny] = dyd[i].a; J[i + (size_t)(Ns + 1) * Ny] = dyd[i].b; }
    }
}

// ---- dense LU (partial pivoting), column-major -----------------------------------------------------
int g_flags = 0;   // integrator flags, same meaning as rmsb_solve_opts.reserved (0 = the CUDA path's defaults)

bool lu_factor(double* A, int n, int* piv) {
    for (int k = 0; k < n; ++k) {
        int p = k; double best = std::fabs(A[k + (size_t)k * n]);
        for (int i = k + 1; i < n; ++i) { double v = std::fabs(A[i + (size_t)k * n]); if (v > best) { best = v; p = i; } }
        piv[k] = p;
        if (p != k) for (int j = 0; j < n; ++j) std::swap(A[k + (size_t)j * n], A[p + (size_t)j * n]);
        double pv = A[k + (size_t)k * n];
        if (pv == 0.0 || !std::isfinite(pv)) return false;
        double inv = 1.0 / pv;
        for (int i = k + 1; i < n; ++i) A[i + (size_t)k * n] *= inv;
        for (int j = k + 1; j < n; ++j) {
            double akj = A[k + (size_t)j * n];
            if (akj != 0.0) for (int i = k + 1; i < n; ++i) A[i + (size_t)j * n] -= A[i + (size_t)k * n] * akj;
        }
    }
    return true;
}
void lu_solve(const double* A, int n, const int* piv, double* b) {
    for (int k = 0; k < n; ++k) std::swap(b[k], b[piv[k]]);
    for (int k = 0; k < n; ++k) { double xk = b[k]; if (xk != 0.0) for (int i = k + 1; i < n; ++i) b[i] -= A[i + (size_t)k * n] * xk; }
    for (int k = n - 1; k >= 0; --k) { b[k] /= A[k + (size_t)k * n]; double xk = b[k]; for (int i = 0; i < k; ++i) b[i] -= A[i + (size_t)k * n] * xk; }
}

// ---- NDF/BDF (Shampine & Reichelt 1997; the algorithm of scipy.integrate.BDF) ----------------------
constexpr int MAXO = 5, NEWTON_MAXITER = 4;

double rms(const double* x, const double* scale, int n) {
    double a = 0.0;
    for (int i = 0; i < n; ++i) { double q = x[i] / scale[i]; a += q * q; }
    return std::sqrt(a / n);
}

void change_D(std::vector<double>& D, int n, int order, double factor) {
    double Rm[MAXO + 1][MAXO + 1], Um[MAXO + 1][MAXO + 1], RU[MAXO + 1][MAXO + 1];
    for (int k = 0; k <= order; ++k) { Rm[0][k] = 1.0; Um[0][k] = 1.0; }
    for (int i = 1; i <= order; ++i)
        for (int k = 0; k <= order; ++k) {
            if (k == 0) { Rm[i][k] = 0.0; Um[i][k] = 0.0; }
            else { Rm[i][k] = Rm[i - 1][k] * ((i - 1 - factor * k) / i); Um[i][k] = Um[i - 1][k] * ((double)(i - 1 - k) / i); }
        }
    for (int i = 0; i <= order; ++i)
        for (int j = 0; j <= order; ++j) { double a = 0.0; for (int k = 0; k <= order; ++k) a += Rm[i][k] * Um[k][j]; RU[i][j] = a; }
    std::vector<double> Dn((size_t)(order + 1) * n);
    for (int j = 0; j <= order; ++j)
        for (int c = 0; c < n; ++c) { double a = 0.0; for (int i = 0; i <= order; ++i) a += RU[i][j] * D[(size_t)i * n + c]; Dn[(size_t)j * n + c] = a; }
    std::copy(Dn.begin(), Dn.end(), D.begin());
}

struct SolveOut { int status; rmsb_stats st; double idt; };

SolveOut bdf_solve(const Mech& m, const Dom& d, Work& w, const double* y0, double t_end, double rtol, double atol,
                   int nsave, const double* tsave, double* ysave, double* yend, int idt_index, int max_steps) {
    const int n = d.Ny;
    const double kappa[MAXO + 1] = {0.0, -0.1850, -1.0 / 9.0, -0.0823, -0.0415, 0.0};
    double gamma[MAXO + 1], alpha[MAXO + 1], ec[MAXO + 1];
    gamma[0] = 0.0;
    for (int k = 1; k <= MAXO; ++k) gamma[k] = gamma[k - 1] + 1.0 / k;
    for (int k = 0; k <= MAXO; ++k) { alpha[k] = (1.0 - kappa[k]) * gamma[k]; ec[k] = kappa[k] * gamma[k] + 1.0 / (k + 1); }
    SolveOut out{};
    rmsb_stats& st = out.st;
    std::vector<double> D((size_t)(MAXO + 3) * n, 0.0), y(y0, y0 + n), f(n), scale(n), ypred(n), psi(n), dvec(n), ynew(n), dy(n), err(n);
    if (d.kind != RMSB_DOMAIN_CONSTANT_V && d.kind != RMSB_DOMAIN_CONSTANT_P) {
        double T = d.cond[0], P, V;
        if (d.kind == RMSB_DOMAIN_CONSTANT_TP) { P = d.cond[1]; V = y[m.Ns]; } else { P = d.kind == RMSB_DOMAIN_CONSTANT_TV ? 1.0e8 : 0.0; V = d.cond[1]; }
        setup_constk(m, d, w, y.data(), T, P, V);
    }
    int isave = 0;
    while (isave < nsave && tsave[isave] <= 0.0) { if (ysave) std::copy(y.begin(), y.end(), ysave + (size_t)isave * n); ++isave; }
    rhs<double>(m, d, w, y.data(), f.data(), false, nullptr); ++st.n_rhs;
    double t = 0.0, h_abs;
    {   // select_initial_step
        for (int i = 0; i < n; ++i) scale[i] = atol + std::fabs(y[i]) * rtol;
        double d0 = rms(y.data(), scale.data(), n), d1 = rms(f.data(), scale.data(), n);
        double h0 = (d0 < 1e-5 || d1 < 1e-5) ? 1e-6 : 0.01 * d0 / d1;
        std::vector<double> y1(n), f1(n);
        for (int i = 0; i < n; ++i) y1[i] = y[i] + h0 * f[i];
        rhs<double>(m, d, w, y1.data(), f1.data(), false, nullptr); ++st.n_rhs;
        for (int i = 0; i < n; ++i) f1[i] -= f[i];
        double d2 = rms(f1.data(), scale.data(), n) / h0;
        double h1 = (d1 <= 1e-15 && d2 <= 1e-15) ? std::max(1e-6, h0 * 1e-3) : std::pow(0.01 / std::max(d1, d2), 0.5);
        h_abs = std::min(std::min(100.0 * h0, h1), t_end);
    }
    jacobian(m, d, w, y.data()); ++st.n_jac;
    bool lu_valid = false;
    double c_jac = 0.0;   // c = h/alpha at the last Jacobian evaluation (0: not yet used)
    // A Jacobian evaluated at a step size 100x smaller (or larger) is refreshed before the next factorisation: with a stale J and
    // c -> large the modified Newton iteration looks converged in the weighted norm while the state coasts along the
    // equilibrium manifold (its energy-conservation null vector is the stale one); seen as a T drift of 1e-4 in the
    // post-ignition tail of 2.6 % of the C2 sweep.  Costs +4 Jacobians per solve.
    const double jac_ratio = getenv("RMSO_JRATIO") ? atof(getenv("RMSO_JRATIO")) : 100.0;
    for (int i = 0; i < n; ++i) { D[i] = y[i]; D[n + i] = f[i] * h_abs; }
    int order = 1, n_equal = 0;
    const double newton_tol = std::max(10.0 * 2.220446049250313e-16 / rtol, std::min(0.03, std::sqrt(rtol)));
    const bool scipy_newton = (g_flags & 1) != 0, hysteresis = (g_flags & 2) == 0;
    const int nls_v = (g_flags >> 8) & 0xff;
    const double nls_coef = std::ldexp(0.1, -(nls_v == 0 ? 3 : nls_v - 1));
    double crate = 1.0;
    // species no reaction changes (net stoichiometry zero everywhere) and constant species: dy_i/dt == 0, zero Jacobian row
    std::vector<char> unit_row(n, 0);
    std::vector<double> rhs_keep(n);
    for (int s = 0; s < m.Ns; ++s) unit_row[s] = 1;
    for (int j = 0; j < m.Nr; ++j) {
        const int* r = &m.rx[j * 8];
        for (int a = 0; a < 8; ++a) {
            if (r[a] < 0) continue;
            int net = 0;
            for (int b = 0; b < 8; ++b) if (r[b] == r[a]) net += (b < 4) ? -1 : 1;
            if (net != 0) unit_row[r[a]] = 0;
        }
    }
    for (int k = 0; k < d.nconst; ++k) unit_row[d.constspc[k] - 1] = 1;
    double idt_best = -std::numeric_limits<double>::infinity(), idt_time = std::numeric_limits<double>::quiet_NaN();
    int status = RMSB_SOLVE_OK;
    for (int i = 0; i < n; ++i) if (!std::isfinite(f[i])) status = RMSB_SOLVE_NAN;
    while (status == RMSB_SOLVE_OK && t < t_end) {
        if (st.n_steps >= max_steps) { status = RMSB_SOLVE_MAX_STEPS; break; }
        double min_step = 10.0 * (std::nextafter(t, std::numeric_limits<double>::infinity()) - t);
        if (h_abs < min_step) { change_D(D, n, order, min_step / h_abs); h_abs = min_step; n_equal = 0; lu_valid = false; }
        bool accepted = false, current_jac = false;
        double t_new = t, h = 0.0, error_norm = 0.0, safety = 0.9;
        while (!accepted) {
            if (h_abs < min_step) { status = RMSB_SOLVE_STEP_TOO_SMALL; break; }
            t_new = t + h_abs;
            // the step the differences D are scaled for is h_abs itself; recomputing it as fl(t + h_abs) - t (scipy) makes the
            // formula inconsistent by ulp(t)*|dy/dt|, an error floor that rtol = 1e-12 hits during ignition (autonomous RHS:
            // the rounding of t itself is harmless)
            if (t_new - t_end > 0.0) {
                t_new = t_end;
                double hc = t_end - t;
                change_D(D, n, order, hc / h_abs); n_equal = 0; lu_valid = false;
                h_abs = hc;
            }
            h = h_abs;
            for (int i = 0; i < n; ++i) {
                double s = 0.0, ps = 0.0;
                for (int j = 0; j <= order; ++j) s += D[(size_t)j * n + i];
                for (int j = 1; j <= order; ++j) ps += D[(size_t)j * n + i] * gamma[j];
                ypred[i] = s; psi[i] = ps / alpha[order]; scale[i] = atol + rtol * std::fabs(s);
            }
            double c = h / alpha[order];
            bool converged = false;
            int n_iter = 0;
            while (!converged) {
                bool lu_ok = true;
                if (!lu_valid) {
                    if (c_jac == 0.0) c_jac = c;
                    else if (jac_ratio > 0.0 && !current_jac && (c > jac_ratio * c_jac || c * jac_ratio < c_jac)) {
                        jacobian(m, d, w, ypred.data()); ++st.n_jac; current_jac = true; c_jac = c;
                    }
                    for (int cc = 0; cc < n; ++cc) for (int r = 0; r < n; ++r) w.lu[r + (size_t)cc * n] = (r == cc ? 1.0 : 0.0) - c * w.jac[r + (size_t)cc * n];
                    lu_ok = lu_factor(w.lu.data(), n, w.piv.data()); lu_valid = true; ++st.n_lu; crate = 1.0;
                }
                const double nl_tol = scipy_newton ? newton_tol : std::max(newton_tol, nls_coef / ec[order]);
                std::fill(dvec.begin(), dvec.end(), 0.0);
                ynew = ypred;
                double dy_norm_old = -1.0;
                if (lu_ok)
                    for (int k = 0; k < NEWTON_MAXITER; ++k) {
                        n_iter = k + 1;
                        rhs<double>(m, d, w, ynew.data(), f.data(), false, nullptr); ++st.n_rhs;
                        bool fin = true;
                        for (int i = 0; i < n; ++i) if (!std::isfinite(f[i])) fin = false;
                        if (!fin) break;
                        for (int i = 0; i < n; ++i) dy[i] = c * f[i] - psi[i] - dvec[i];
                        for (int i = 0; i < n; ++i) rhs_keep[i] = dy[i];
                        lu_solve(w.lu.data(), n, w.piv.data(), dy.data());
                        for (int i = 0; i < n; ++i) if (unit_row[i]) dy[i] = rhs_keep[i];   // row i of I - c*J is a unit row: x_i = b_i exactly
                        double dy_norm = rms(dy.data(), scale.data(), n);
                        double rate = dy_norm_old < 0.0 ? -1.0 : dy_norm / dy_norm_old;
                        if (!(dy_norm == dy_norm)) break;
                        if (current_jac && getenv("RMSO_DEBUG") && rate >= 0.0 && (rate >= 1.0 || std::pow(rate, NEWTON_MAXITER - k) / (1.0 - rate) * dy_norm > newton_tol)) {
                            int im = 0; double em = 0; for (int i = 0; i < n; ++i) { double e = std::fabs(dy[i] / scale[i]); if (e > em) { em = e; im = i; } }
                            fprintf(stderr, "nfail t=%.10g h=%.3e order=%d k=%d dy_norm=%.3e rate=%.3e worst %d (y=%.4e dy=%.3e scale=%.3e) T=%.2f\n", t, h_abs, order, k, dy_norm, rate, im, ynew[im], dy[im], scale[im], ynew[n-2]);
                        }
                        if (rate >= 0.0) {
                            if (rate >= 1.0 || std::pow(rate, NEWTON_MAXITER - k) / (1.0 - rate) * dy_norm > nl_tol) break;
                            crate = std::max(0.3 * crate, rate);
                        }
                        for (int i = 0; i < n; ++i) { ynew[i] += dy[i]; dvec[i] += dy[i]; }
                        if (getenv("RMSO_TRACE")) fprintf(stderr, "   it t=%.6e h=%.3e o=%d k=%d dy_norm=%.3e rate=%.3e crate=%.3e\n", t, h_abs, order, k, dy_norm, rate, crate);
                        if (scipy_newton) {
                            if (dy_norm == 0.0 || (rate >= 0.0 && rate / (1.0 - rate) * dy_norm < nl_tol)) { converged = true; break; }
                        } else if (dy_norm * std::min(1.0, crate) <= nl_tol) { converged = true; break; }
                        dy_norm_old = dy_norm;
                    }
                if (!converged) {
                    if (current_jac) break;
                    jacobian(m, d, w, ypred.data()); ++st.n_jac; c_jac = c;
                    lu_valid = false; current_jac = true;
                }
            }
            if (!converged) { ++st.n_newton_fail; h_abs *= 0.5; change_D(D, n, order, 0.5); n_equal = 0; lu_valid = false; continue; }
            safety = 0.9 * (2.0 * NEWTON_MAXITER + 1.0) / (2.0 * NEWTON_MAXITER + n_iter);
            for (int i = 0; i < n; ++i) { scale[i] = atol + rtol * std::fabs(ynew[i]); err[i] = ec[order] * dvec[i]; }
            error_norm = rms(err.data(), scale.data(), n);
            if (error_norm > 1.0) {
                if (getenv("RMSO_DEBUG")) {
                    int im = 0; double em = 0; for (int i = 0; i < n; ++i) { double e = std::fabs(err[i] / scale[i]); if (e > em) { em = e; im = i; } }
                    fprintf(stderr, "efail t=%.17g h=%.3e order=%d err=%.3e worst comp %d (y=%.6e d=%.3e D1=%.3e) iters=%d\n", t, h_abs, order, error_norm, im, ynew[im], dvec[im], D[1 * n + im], n_iter);
                }
                ++st.n_err_fail;
                double factor = std::max(0.2, safety * std::pow(error_norm, -1.0 / (order + 1)));
                h_abs *= factor; change_D(D, n, order, factor); n_equal = 0; lu_valid = false;
            } else accepted = true;
        }
        if (!accepted) break;
        ++n_equal; ++st.n_steps;
        if (getenv("RMSO_TRACE")) {
            double sn = 0; for (int i = 0; i < m.Ns; ++i) sn += ynew[i];
            fprintf(stderr, "ACC t=%.6e h=%.3e o=%d err=%.3e T=%.6f inv=%.3e lu_valid=%d\n", t_new, h_abs, order, error_norm, ynew[m.Ns], ynew[m.Ns + 1] * d.cond[0] / (8.314 * ynew[m.Ns]) / sn - 1.0, (int)lu_valid);
        }
        double t_old = t;
        t = t_new;
        for (int i = 0; i < n; ++i) {
            D[(size_t)(order + 2) * n + i] = dvec[i] - D[(size_t)(order + 1) * n + i];
            D[(size_t)(order + 1) * n + i] = dvec[i];
            for (int j = order; j >= 0; --j) D[(size_t)j * n + i] += D[(size_t)(j + 1) * n + i];
        }
        while (isave < nsave && tsave[isave] <= t) {   // dense output
            double ts = tsave[isave];
            if (ysave)
                for (int i = 0; i < n; ++i) {
                    double pj = 1.0, acc = D[i];
                    for (int j = 1; j <= order; ++j) { pj *= (ts - (t - (j - 1) * h)) / (h * j); acc += D[(size_t)j * n + i] * pj; }
                    ysave[(size_t)isave * n + i] = acc;
                }
            ++isave;
        }
        if (idt_index >= 0)
            for (int q = 1; q <= 32; ++q) {   // max of d/dt of the interpolant over the step, 32 samples
                double ts = t_old + h * q * (1.0 / 32.0), pj = 1.0, dpj = 0.0, der = 0.0;
                for (int j = 1; j <= order; ++j) {
                    double den = h * j, x = (ts - (t - (j - 1) * h)) / den;
                    dpj = dpj * x + pj / den; pj = pj * x;
                    der += D[(size_t)j * n + idt_index] * dpj;
                }
                if (der > idt_best) { idt_best = der; idt_time = ts; }
            }
        for (int i = 0; i < n; ++i) y[i] = D[i];
        if (n_equal < order + 1) continue;
        double e_m = std::numeric_limits<double>::infinity(), e_p = e_m;
        if (order > 1) { for (int i = 0; i < n; ++i) err[i] = ec[order - 1] * D[(size_t)order * n + i]; e_m = rms(err.data(), scale.data(), n); }
        if (order < MAXO) { for (int i = 0; i < n; ++i) err[i] = ec[order + 1] * D[(size_t)(order + 2) * n + i]; e_p = rms(err.data(), scale.data(), n); }
        double fm = std::pow(e_m, -1.0 / order), f0 = std::pow(error_norm, -1.0 / (order + 1)), fp = std::pow(e_p, -1.0 / (order + 2));
        int delta; double fmx;
        if (fm >= f0 && fm >= fp) { delta = -1; fmx = fm; } else if (f0 >= fp) { delta = 0; fmx = f0; } else { delta = 1; fmx = fp; }
        double factor = std::min(10.0, safety * fmx);
        if (hysteresis && delta == 0 && factor >= 1.0 && factor < 1.2) continue;
        order += delta;
        h_abs *= factor; change_D(D, n, order, factor); n_equal = 0; lu_valid = false;
    }
    if (yend) std::copy(y.begin(), y.end(), yend);
    if (ysave) for (; isave < nsave; ++isave) for (int i = 0; i < n; ++i) ysave[(size_t)isave * n + i] = std::numeric_limits<double>::quiet_NaN();
    st.last_order = order; st.last_h = h_abs;
    out.status = status;
    out.idt = idt_index >= 0 ? idt_time : std::numeric_limits<double>::quiet_NaN();
    return out;
}

int domain_ny(const Mech& m, int kind) {
    switch (kind) {
    case RMSB_DOMAIN_CONSTANT_V:
    case RMSB_DOMAIN_CONSTANT_P: return m.Ns + 2;
    case RMSB_DOMAIN_CONSTANT_TP: return m.Ns + 1;
    case RMSB_DOMAIN_CONSTANT_TV: case RMSB_DOMAIN_CONSTANT_TAPHI: return m.Ns;
    }
    return -1;
}
int domain_ncond(int kind) { return (kind == RMSB_DOMAIN_CONSTANT_V || kind == RMSB_DOMAIN_CONSTANT_P) ? 1 : 2; }

}  // namespace

extern "C" {

void* rmso_create(const rmsb_mech_desc* d) {
    Mech* m = new Mech();
    m->Ns = d->n_species; m->Nr = d->n_reactions; m->phase_kind = d->phase_kind; m->difflim = d->diffusion_limited;
    m->maxr = d->nasa_max_ranges; m->sitedensity = d->sitedensity; m->visc_kind = d->viscosity_kind;
    for (int k = 0; k < 5; ++k) m->visc[k] = d->viscosity_params[k];
    const int Ns = m->Ns, Nr = m->Nr, maxr = m->maxr;
    m->nr.assign(d->nasa_nranges, d->nasa_nranges + Ns);
    m->nc.assign(d->nasa_ncoef, d->nasa_ncoef + Ns * maxr);
    m->tmax.assign(d->nasa_tmax, d->nasa_tmax + Ns * maxr);
    m->coef.assign(d->nasa_coefs, d->nasa_coefs + (size_t)Ns * maxr * 9);
    m->rx.resize((size_t)Nr * 8);
    for (size_t i = 0; i < (size_t)Nr * 8; ++i) m->rx[i] = d->rxnarray[i] - 1;
    m->ktype.assign(d->kin_type, d->kin_type + Nr);
    m->koff.assign(d->kin_offset, d->kin_offset + Nr + 1);
    m->kpar.assign(d->kin_params, d->kin_params + m->koff[Nr]);
    m->effptr.assign(d->eff_ptr, d->eff_ptr + Nr + 1);
    for (int k = 0; k < m->effptr[Nr]; ++k) { m->effspc.push_back(d->eff_species[k] - 1); m->effval.push_back(d->eff_value[k]); }
    m->rev.resize(Nr); m->fwd.resize(Nr); m->iseff.assign(Nr, 0);
    for (int j = 0; j < Nr; ++j) {
        m->rev[j] = d->reversible[j] ? 1.0 : 0.0; m->fwd[j] = d->forwardable[j] ? 1.0 : 0.0;
        if (m->effptr[j + 1] > m->effptr[j]) { m->iseff[j] = 1; m->effrxn.push_back(j); }
    }
    if (d->radius) m->radius.assign(d->radius, d->radius + Ns); else m->radius.assign(Ns, 0.0);
    if (d->diff_radius) m->diffr.assign(d->diff_radius, d->diff_radius + Ns); else m->diffr.assign(Ns, -1.0);
    m->Mrow.resize(Nr); m->nrp.resize(Nr);
    for (int j = 0; j < Nr; ++j) {   // getstoichmatrix, src/Phase.jl:203-222 (slot > 3 counts as product)
        std::vector<std::pair<int, double>> row;
        double n = 0.0;
        for (int k = 0; k < 8; ++k) {
            int s = m->rx[j * 8 + k];
            if (s < 0) continue;
            double v = (k + 1 > 3) ? -1.0 : 1.0;
            n -= v;
            bool found = false;
            for (auto& kv : row) if (kv.first == s) { kv.second += v; found = true; }
            if (!found) row.push_back({s, v});
        }
        std::sort(row.begin(), row.end());
        m->Mrow[j] = row; m->nrp[j] = n;
    }
    return m;
}

void rmso_destroy(void* h) { delete static_cast<Mech*>(h); }

static Dom make_dom(const Mech& m, int kind, int64_t r, const double* cond, const double* p, int64_t p_stride, const int32_t* cs, int ncs) {
    Dom d;
    d.kind = kind; d.Ny = domain_ny(m, kind); d.p = p ? p + r * p_stride : nullptr; d.constspc = cs; d.nconst = cs ? ncs : 0;
    d.cond = cond + r * domain_ncond(kind);
    return d;
}

// mode 0: dydt [n][Ny], 1: Jacobian [n][Ny*Ny] column-major
int rmso_eval(void* h, int mode, int kind, int64_t n, const double* y, const double* cond, const double* p, int64_t p_stride,
              const int32_t* cs, int32_t ncs, double* out, int nthreads) {
    const Mech& m = *static_cast<Mech*>(h);
    const int Ny = domain_ny(m, kind);
    if (Ny < 0) return -1;
#ifdef _OPENMP
    if (nthreads > 0) omp_set_num_threads(nthreads);
#endif
#pragma omp parallel
    {
        Work w; w.init(m, Ny);
#pragma omp for schedule(dynamic, 16)
        for (int64_t r = 0; r < n; ++r) {
            Dom d = make_dom(m, kind, r, cond, p, p_stride, cs, ncs);
            const double* yr = y + r * Ny;
            if (kind != RMSB_DOMAIN_CONSTANT_V && kind != RMSB_DOMAIN_CONSTANT_P) {
                double T = d.cond[0], P, V;
                if (kind == RMSB_DOMAIN_CONSTANT_TP) { P = d.cond[1]; V = yr[m.Ns]; } else { P = kind == RMSB_DOMAIN_CONSTANT_TV ? 1.0e8 : 0.0; V = d.cond[1]; }
                setup_constk(m, d, w, yr, T, P, V);
            }
            if (mode == 0) rhs<double>(m, d, w, yr, out + r * Ny, false, nullptr);
            else { jacobian(m, d, w, yr); std::copy(w.jac.begin(), w.jac.end(), out + (size_t)r * Ny * Ny); }
        }
    }
    return 0;
}

void rmso_set_flags(int flags) { g_flags = flags; }

int rmso_solve(void* h, int kind, int64_t n, const double* y0, const double* cond, const double* p, int64_t p_stride,
               const int32_t* cs, int32_t ncs, const double* t_end, double rtol, double atol, int32_t nsave, const double* tsave,
               double* ysave, double* yend, double* idt, int32_t* status, rmsb_stats* stats, int32_t max_steps, int32_t idt_index,
               int nthreads) {
    const Mech& m = *static_cast<Mech*>(h);
    const int Ny = domain_ny(m, kind);
    if (Ny < 0) return -1;
    if (idt_index < 0) idt_index = (kind == RMSB_DOMAIN_CONSTANT_V || kind == RMSB_DOMAIN_CONSTANT_P) ? m.Ns : -1;
    if (max_steps <= 0) max_steps = 500000;
#ifdef _OPENMP
    if (nthreads > 0) omp_set_num_threads(nthreads);
#endif
#pragma omp parallel
    {
        Work w; w.init(m, Ny);
#pragma omp for schedule(dynamic, 1)
        for (int64_t r = 0; r < n; ++r) {
            Dom d = make_dom(m, kind, r, cond, p, p_stride, cs, ncs);
            SolveOut o = bdf_solve(m, d, w, y0 + r * Ny, t_end[r], rtol, atol, nsave, tsave, ysave ? ysave + (size_t)r * nsave * Ny : nullptr,
                                   yend ? yend + r * Ny : nullptr, idt_index, max_steps);
            status[r] = o.status;
            if (idt) idt[r] = o.idt;
            if (stats) stats[r] = o.st;
        }
    }
    return 0;
}

int rmso_max_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

}  // extern "C"
