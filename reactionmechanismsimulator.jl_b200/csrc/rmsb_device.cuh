// rmsb_device.cuh -- device-side mechanism tables and the warp-cooperative RHS / Jacobian.
//
// Execution model: ONE WARP INTEGRATES ONE REACTOR.  Species-indexed work (thermo, gather of reaction
// rates) is spread over lanes by species, reaction-indexed work (rate coefficients, equilibrium
// constants, mass-action rates) by reaction; per-warp scratch lives in the warp's private slice of
// shared memory.  Every warp follows its own control flow, so per-reactor adaptive stepping costs no
// intra-warp divergence.
//
// Code-size discipline (round-1 ncu finding: a fully inlined kernel was 95k SASS instructions and stalled
// on instruction fetch): the RHS, Jacobian, LU and setup routines are NOT inlined; they exchange vectors
// through the shared-memory slice (ybuf -> fbuf) so no register array crosses a call boundary.
//
// What is restated here (reference = ReactionMechanismSimulator.jl, paths relative to its root):
//   calcHSCpdless                 src/Calculators/Thermovec.jl:76-119      (vectorised thermo path)
//   getGibbs/getEnthalpy/Entropy  src/Calculators/Thermo.jl:6-8,75-93      (scalar thermo path)
//   rate calculators              src/Calculators/Rate.jl:10-210, Ratevec.jl:24,89-100,204-215
//   getkf/getkfs/getKc/getKcs/getkfkrev(s)   src/PhaseState.jl:53-76,106-129,165-171,184-257
//   calcthermo                    src/Domain.jl:916-966,1064-1113,1514-1561,1597-1633
//   addreactionratecontributions! src/Reactor.jl:465-507
//   calcdomainderivatives!        src/Domain.jl:1839-1842,1863-1867,1901-1906
//   jacobiany! (+helpers)         src/Domain.jl:2151-2454,3049-3055 ; src/Reactor.jl:874-1477
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <math.h>
#include "../../include/rmsb200.h"

namespace rmsb {

// src/Constants.jl:1-7 (R is 8.314 exactly)
#define RMSB_R   8.314
#define RMSB_NA  6.022e23
#define RMSB_KB  (8.314 / 6.022e23)
#define RMSB_PI  3.141592653589793

constexpr unsigned FULL = 0xffffffffu;

struct DevMech {
    int Ns, Nr;
    int phase_kind;
    int difflim;
    double sitedensity;
    // NASA polynomials, SoA over species: nasa_tmax[r*Ns+s], nasa_nc[r*Ns+s], nasa_c[(r*9+c)*Ns+s]
    int maxr;
    const int* nasa_nr;
    const int* nasa_nc;
    const double* nasa_tmax;
    const double* nasa_c;
    // reactions: rx[slot*Nr+j], 0-based, -1 = none
    const int* rx;
    const int* ktype;
    const int* koff;
    const double* kpar;
    const int* eff_ptr;      // [Nr+1]
    const int* eff_spc;      // 0-based
    const double* eff_val;
    int n_effrxn;            // reactions with >=1 efficiency AND falloff kinetics (Domain.jl:74 efficiencyinds)
    const int* effrxn;       // their indices
    const double* rev;       // 0/1
    const double* fwd;       // 0/1
    // stoichiometric matrix rows (getstoichmatrix, Phase.jl:203-222): CSR sorted by species
    const int* kc_ptr;
    const int* kc_spc;
    const double* kc_coef;
    const double* nrp;
    // species -> (reaction, sign) in reaction order, one entry per slot occurrence
    const int* sr_ptr;
    const int* sr_rxn;       // rxn index; bit 31 set => reactant slot (subtract)
    // Jacobian ns-block: derivative terms and per-entry gather lists (built on host in reference order)
    int n_dt;
    const int* dt_rxn;       // reaction | dir<<30 (dir 1 = reverse)
    const int* dt_mult;      // multiplicity m
    const int* dt_oth;       // [3*n_dt] other species of that side (-1 none)
    const int* dt_spc;       // species differentiated against (column)
    int n_je;
    const int* je_ptr;       // [n_je+1]
    const int* je_row;
    const int* je_col;
    const int* jt_idx;       // term index; bit 31 set => subtract
    // liquid phase
    const double* radius;
    const double* diff_r;
    int visc_kind;
    double visc[5];
    // two-phase (bulk + surface + interface) systems: species [0,n_bulk) live in the bulk phase
    int coupled, n_bulk;
    const int* rx_dom;       // [Nr] 0 bulk reaction, 1 surface reaction, 2 interface reaction (sorted ascending)
    const double* nrp2;      // [Nr] surface-phase exponent of C0 in Kc (nrp is the bulk-phase exponent)
    const int* sr_mid;       // [Ns] first interface entry of the species' slot list
    const int* inert;        // [Ns] 1 = no reaction changes the species (absent, or written on both sides everywhere): dn/dt == 0 exactly
    int heavy_sr, heavy_je;  // 1 = some species row / Jacobian entry list is longer than COOP_ROW: the warp-cooperative sums are compiled in
                             // for those mechanisms only (their ballots cost the 12-species mechanisms 10 %)
};

// ------------------------------------------------------------------------------------------------
// Dual numbers with two partials (d/dT, d/dP) -- the in-kernel analogue of the reference's
// ForwardDiff pass over the complete RHS for the ConstantV T and P columns (src/Reactor.jl:1469-1477).
// ------------------------------------------------------------------------------------------------
struct Dual2 {
    double v, a, b;
    __device__ __forceinline__ Dual2() {}
    __device__ __forceinline__ Dual2(double v_) : v(v_), a(0.0), b(0.0) {}
    __device__ __forceinline__ Dual2(double v_, double a_, double b_) : v(v_), a(a_), b(b_) {}
};
__device__ __forceinline__ Dual2 operator+(Dual2 x, Dual2 y) { return Dual2(x.v + y.v, x.a + y.a, x.b + y.b); }
__device__ __forceinline__ Dual2 operator+(Dual2 x, double y) { return Dual2(x.v + y, x.a, x.b); }
__device__ __forceinline__ Dual2 operator+(double x, Dual2 y) { return Dual2(x + y.v, y.a, y.b); }
__device__ __forceinline__ Dual2 operator-(Dual2 x, Dual2 y) { return Dual2(x.v - y.v, x.a - y.a, x.b - y.b); }
__device__ __forceinline__ Dual2 operator-(Dual2 x, double y) { return Dual2(x.v - y, x.a, x.b); }
__device__ __forceinline__ Dual2 operator-(double x, Dual2 y) { return Dual2(x - y.v, -y.a, -y.b); }
__device__ __forceinline__ Dual2 operator-(Dual2 x) { return Dual2(-x.v, -x.a, -x.b); }
__device__ __forceinline__ Dual2 operator*(Dual2 x, Dual2 y) { return Dual2(x.v * y.v, x.a * y.v + x.v * y.a, x.b * y.v + x.v * y.b); }
__device__ __forceinline__ Dual2 operator*(Dual2 x, double y) { return Dual2(x.v * y, x.a * y, x.b * y); }
__device__ __forceinline__ Dual2 operator*(double x, Dual2 y) { return Dual2(x * y.v, x * y.a, x * y.b); }
__device__ __forceinline__ Dual2 operator/(Dual2 x, Dual2 y) {
    double q = x.v / y.v, iy = 1.0 / y.v;
    return Dual2(q, (x.a - q * y.a) * iy, (x.b - q * y.b) * iy);
}
__device__ __forceinline__ Dual2 operator/(Dual2 x, double y) { return Dual2(x.v / y, x.a / y, x.b / y); }
__device__ __forceinline__ Dual2 operator/(double x, Dual2 y) {
    double q = x / y.v, s = -q / y.v;
    return Dual2(q, s * y.a, s * y.b);
}
__device__ __forceinline__ Dual2 dexp(Dual2 x) { double e = exp(x.v); return Dual2(e, e * x.a, e * x.b); }
__device__ __forceinline__ double dexp(double x) { return exp(x); }
__device__ __forceinline__ Dual2 dlog(Dual2 x) { double i = 1.0 / x.v; return Dual2(log(x.v), x.a * i, x.b * i); }
__device__ __forceinline__ double dlog(double x) { return log(x); }
__device__ __forceinline__ Dual2 dlog10(Dual2 x) {
    double i = 0.4342944819032518 / x.v;  // log10(e)/x
    return Dual2(log10(x.v), x.a * i, x.b * i);
}
__device__ __forceinline__ double dlog10(double x) { return log10(x); }
__device__ __forceinline__ Dual2 dpow(Dual2 x, double n) {  // x^n, real exponent
    double v = pow(x.v, n), d = n * v / x.v;
    return Dual2(v, d * x.a, d * x.b);
}
__device__ __forceinline__ double dpow(double x, double n) { return pow(x, n); }
__device__ __forceinline__ Dual2 dexp10(Dual2 x) {  // 10^x
    double v = pow(10.0, x.v), d = v * 2.302585092994046;
    return Dual2(v, d * x.a, d * x.b);
}
__device__ __forceinline__ double dexp10(double x) { return pow(10.0, x); }
__device__ __forceinline__ double val(double x) { return x; }
__device__ __forceinline__ double val(Dual2 x) { return x.v; }
// seed the thermodynamic state variables: which = 0 -> d/dT = 1, which = 1 -> d/dP = 1
template <class N> __device__ __forceinline__ N seed(double v, int which);
template <> __device__ __forceinline__ double seed<double>(double v, int) { return v; }
template <> __device__ __forceinline__ Dual2 seed<Dual2>(double v, int which) { return Dual2(v, which == 0 ? 1.0 : 0.0, which == 1 ? 1.0 : 0.0); }

// ---- warp helpers --------------------------------------------------------------------------------
__device__ __forceinline__ double warp_sum(double x) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(FULL, x, o);
    return x;
}
__device__ __forceinline__ Dual2 warp_sum(Dual2 x) { return Dual2(warp_sum(x.v), warp_sum(x.a), warp_sum(x.b)); }

// ------------------------------------------------------------------------------------------------
// Per-warp shared-memory workspace.  A small by-value view (base pointer + sizes); sub-arrays are
// located by integer offsets so that the struct travels in registers across non-inlined calls.
// ------------------------------------------------------------------------------------------------
extern __shared__ double smem_ws[];     // [DevMech copy | warp 0 slice | warp 1 slice | ...]

struct Ws {
    int off;                                  // offset (in doubles) of this warp's slice inside smem_ws
    int Ns, Nr, Ny, ndt, LD;
    int jpo;                                  // offset of the parameter-Jacobian tile inside the slice (mode 2: after LU; mode 3: on J)
    int tri;                                  // 3: g / un / q / fbuf hold dual numbers (ConstantV / ConstantP run the dual pass), else 1
    __device__ __forceinline__ double* at(int o) const { return smem_ws + off + o; }
    __device__ __forceinline__ double* cs() const { return at(0); }                                 // [Ns]   concentrations
    __device__ __forceinline__ double* g() const { return at(Ns); }                                 // [tri*Ns]  G/(RT) (Dual2-capable) | absolute G (constant-k)
    __device__ __forceinline__ double* un() const { return at((1 + tri) * Ns); }                    // [tri*Ns]  internal energies U_i (Dual2-capable) | diffusivities (liquid setup)
    __device__ __forceinline__ double* q() const { return at((1 + 2 * tri) * Ns); }                 // [tri*Nr]  net rates (Dual2-capable)
    __device__ __forceinline__ double* kf() const { return at((1 + 2 * tri) * Ns + tri * Nr); }     // [Nr]
    __device__ __forceinline__ double* kr() const { return kf() + Nr; }                             // [Nr]
    __device__ __forceinline__ double* us() const { return kf() + 2 * Nr; }                         // [Ns]   U_i values (Jacobian)
    __device__ __forceinline__ double* cp() const { return kf() + 2 * Nr + Ns; }                    // [Ns]   cp_i/R values (Jacobian)
    // Jacobian derivative terms [n_dt]: scratch of jac_warp / jac_coupled only.  In the solver slices (ndt == 0 here) they live on the
    // LU region, which every Jacobian evaluation invalidates anyway (the factorisation follows it); evaluation slices reserve ndt doubles.
    __device__ __forceinline__ double* dterm() const { return ndt ? kf() + 2 * Nr + 2 * Ns : LU(); }
    __device__ __forceinline__ double* ybuf() const { return kf() + 2 * Nr + 2 * Ns + ndt; }        // [Ny]   RHS / Jacobian input state
    __device__ __forceinline__ double* fbuf() const { return ybuf() + Ny; }                         // [tri*Ny]  RHS output (Dual2-capable)
    __device__ __forceinline__ double* aux() const { return fbuf() + tri * Ny; }                    // [8]    T P V N C Cv dot(U,dn) 1/V
    __device__ __forceinline__ double* ru() const { return aux() + 8; }                             // [36]   change_D matrix
    __device__ __forceinline__ double* dinv() const { return ru() + 36; }                           // [Ny]   spare
    __device__ __forceinline__ double* xs() const { return dinv() + Ny; }                           // [Ny]   linear-solve vector
    __device__ __forceinline__ int* perm() const { return reinterpret_cast<int*>(xs() + Ny); }      // [Ny]   pivot rows
    __device__ __forceinline__ double* J() const { return xs() + Ny + (Ny + 1) / 2 + 1; }           // [Ny*LD] Jacobian, column-major
    __device__ __forceinline__ double* LU() const { return J() + Ny * LD; }                         // [Ny*LD] LU factors of I - c*J (solver only)
    __device__ __forceinline__ double* jp() const { return at(jpo); }                        // [Ny*32] 32 columns of d(dy/dt)/dp, [row][lane] (sensitivities only)
    __device__ __forceinline__ double* xm() const { return jp() + Ny * 32; }                        // [Ny*32] per-lane sensitivity work vectors, [row][lane]
};

// The mechanism descriptor lives at the start of the block's dynamic shared memory: every routine reads it
// through a pointer of known address space (LDS with immediate offsets), never through a generic reference.
constexpr int COOP_ROW = 48;   // species rows with more reaction entries than this are summed by the whole warp (rhs_coupled)
constexpr int MECH_DOUBLES = (int)((sizeof(DevMech) + 15) / 16 * 2);
__device__ __forceinline__ const DevMech& mech() { return *reinterpret_cast<const DevMech*>(smem_ws); }
#define G(ptr, idx) __ldg((ptr) + (idx))   /* mechanism tables are immutable global memory: read-only path */

// mode 0: RHS / Jacobian evaluation, 1: solver (+ LU), 2: forward sensitivities (+ LU, Jp tile, work tile),
// 3: parameter-Jacobian evaluation (the Jp tile takes the place of J)
// Dual2 capacity of the g / un / q / fbuf arrays: only the adiabatic gas domains run the dual-number pass (rhs_v<Dual2>); every other
// domain gets the slimmer slice, i.e. more resident warps per SM for the 33..128-variable mechanisms whose slice is tens of KB
__host__ __device__ inline int ws_tri(int kind) { return (kind == RMSB_DOMAIN_CONSTANT_V || kind == RMSB_DOMAIN_CONSTANT_P) ? 3 : 1; }
__host__ __device__ inline int ws_fixed_doubles(int Ns, int Nr, int Ny, int n_dt, int tri) {
    return (3 + 2 * tri) * Ns + (2 + tri) * Nr + n_dt + Ny + tri * Ny + 8 + 36 + Ny + Ny + (Ny + 1) / 2 + 1;
}
// doubles of the derivative-term array reserved in the fixed part of the slice: none in the solver modes (see Ws::dterm)
__host__ __device__ inline int ws_ndt_reserved(int n_dt, int mode) { return (mode == 1 || mode == 2) ? 0 : n_dt; }
__host__ __device__ inline int ws_lu_doubles(int Ny, int n_dt) { int LD = Ny | 1; return Ny * LD > n_dt ? Ny * LD : n_dt; }
__host__ __device__ inline int ws_doubles(int Ns, int Nr, int Ny, int n_dt, int mode, int tri) {
    int LD = Ny | 1;
    int n = ws_fixed_doubles(Ns, Nr, Ny, ws_ndt_reserved(n_dt, mode), tri);
    if (mode == 3) n += (Ny * LD > Ny * 32) ? Ny * LD : Ny * 32;
    else n += Ny * LD;
    if (mode == 1 || mode == 2) n += ws_lu_doubles(Ny, n_dt);
    if (mode == 2) n += 2 * Ny * 32;
    return (n + 1) & ~1;
}
__host__ __device__ inline int ws_jp_offset(int Ns, int Nr, int Ny, int n_dt, int mode, int tri) {
    int LD = Ny | 1;
    return ws_fixed_doubles(Ns, Nr, Ny, ws_ndt_reserved(n_dt, mode), tri) + (mode == 2 ? Ny * LD + ws_lu_doubles(Ny, n_dt) : 0);
}

// smem load/store of a possibly-dual value at logical index i of an array with Dual2 capacity
template <class N> struct SmemNum;
template <> struct SmemNum<double> {
    static __device__ __forceinline__ void st(double* a, int i, double x) { a[i] = x; }
    static __device__ __forceinline__ double ld(const double* a, int i) { return a[i]; }
};
template <> struct SmemNum<Dual2> {
    static __device__ __forceinline__ void st(double* a, int i, Dual2 x) { a[3 * i] = x.v; a[3 * i + 1] = x.a; a[3 * i + 2] = x.b; }
    static __device__ __forceinline__ Dual2 ld(const double* a, int i) { return Dual2(a[3 * i], a[3 * i + 1], a[3 * i + 2]); }
};

// ------------------------------------------------------------------------------------------------
// Thermo
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ int select_poly(int s, double T) {
    const DevMech& m = mech();
    // selectPoly: first range with T <= Tmax, else the last (src/Calculators/Thermo.jl:101-111)
    int nr = G(m.nasa_nr, s);
    for (int r = 0; r < nr; ++r)
        if (T <= G(m.nasa_tmax, r * m.Ns + s)) return r;
    return nr - 1;
}

template <class N>
struct TPow {  // powers of T built exactly as Thermovec.jl:78-84 / :100-105
    N T0, T1, T2, T3, T4, T5, T6;
    __device__ __forceinline__ explicit TPow(N T) {
        T0 = T; T1 = T * T; T2 = T1 * T; T3 = T2 * T; T4 = 1.0 / T; T5 = T4 * T4; T6 = dlog(T);
    }
};

// calcHSCpdless for species s (src/Calculators/Thermovec.jl:76-119; literal 0.33333333333 kept)
template <class N>
__device__ __forceinline__ void nasa_hscp(int s, const TPow<N>& t, double Tv, N& cp, N& h, N& sr) {
    const DevMech& m = mech();
    int r = select_poly(s, Tv);
    const double* c = m.nasa_c + (size_t)(r * 9) * m.Ns + s;
    const int Ns = m.Ns;
    if (G(m.nasa_nc, r * Ns + s) != 7) {
        N ct0 = G(c, 0) * t.T5, ct1 = G(c, Ns) * t.T4;
        double ct2 = G(c, 2 * Ns);
        N ct3 = G(c, 3 * Ns) * t.T0, ct4 = G(c, 4 * Ns) * t.T1, ct5 = G(c, 5 * Ns) * t.T2, ct6 = G(c, 6 * Ns) * t.T3;
        cp = ct0 + ct1 + ct2 + ct3 + ct4 + ct5 + ct6;
        h = -ct0 + t.T6 * ct1 + ct2 + 0.5 * ct3 + 0.33333333333 * ct4 + 0.25 * ct5 + 0.2 * ct6 + G(c, 7 * Ns) * t.T4;
        sr = -0.5 * ct0 - ct1 + t.T6 * ct2 + ct3 + 0.5 * ct4 + 0.33333333333 * ct5 + 0.25 * ct6 + G(c, 8 * Ns);
    } else {
        double ct2 = G(c, 0);
        N ct3 = G(c, Ns) * t.T0, ct4 = G(c, 2 * Ns) * t.T1, ct5 = G(c, 3 * Ns) * t.T2, ct6 = G(c, 4 * Ns) * t.T3;
        cp = ct2 + ct3 + ct4 + ct5 + ct6;
        h = ct2 + 0.5 * ct3 + 0.33333333333 * ct4 + 0.25 * ct5 + 0.2 * ct6 + G(c, 5 * Ns) * t.T4;
        sr = t.T6 * ct2 + ct3 + 0.5 * ct4 + 0.33333333333 * ct5 + 0.25 * ct6 + G(c, 6 * Ns);
    }
}

// getGibbs through the scalar path: evalpoly Horner (src/Tools.jl:13-22) with exact c_k/k
// (src/Calculators/Thermo.jl:6-8,75-93).  Used by the constant-T domains at setup.
static __device__ __noinline__ double gibbs_scalar(int s, double T) {
    const DevMech& m = mech();
    int r = select_poly(s, T);
    const double* c = m.nasa_c + (size_t)(r * 9) * m.Ns + s;
    const int Ns = m.Ns;
    double H, S;
    if (G(m.nasa_nc, r * Ns + s) == 7) {
        double out = 0.0;
        for (int i = 4; i >= 0; --i) { out += G(c, i * Ns) / (double)(i + 1); if (i != 0) out *= T; }
        H = out * RMSB_R * T + G(c, 5 * Ns) * RMSB_R;
        out = 0.0;
        for (int i = 3; i >= 0; --i) { out += G(c, (i + 1) * Ns) / (double)(i + 1); if (i != 0) out *= T; }
        S = (G(c, 0) * log(T) + T * out + G(c, 6 * Ns)) * RMSB_R;
    } else {
        double out = 0.0;
        for (int i = 4; i >= 0; --i) { out += G(c, (i + 2) * Ns) / (double)(i + 1); if (i != 0) out *= T; }
        H = ((-G(c, 0) / T + G(c, Ns) * log(T)) / T + out) * RMSB_R * T + G(c, 7 * Ns) * RMSB_R;
        out = 0.0;
        for (int i = 3; i >= 0; --i) { out += G(c, (i + 3) * Ns) / (double)(i + 1); if (i != 0) out *= T; }
        S = ((-G(c, 0) / (2 * T) - G(c, Ns)) / T + G(c, 2 * Ns) * log(T) + T * out + G(c, 8 * Ns)) * RMSB_R;
    }
    return H - T * S;
}

// ------------------------------------------------------------------------------------------------
// Rate calculators
// ------------------------------------------------------------------------------------------------
template <class N>
struct KinEnv {
    N T, logT, invRT;   // invRT = 1.0/(R*T)  (Ratevec.jl:24)
    N RT;               // R*T
    N P;
};

template <class N> __device__ __forceinline__ N arr_vec(const double* a, const KinEnv<N>& e) {
    return G(a, 0) * dexp(G(a, 1) * e.logT - G(a, 2) * e.invRT);            // Ratevec.jl:24
}
template <class N> __device__ __forceinline__ N arr_sca(const double* a, const KinEnv<N>& e) {
    return G(a, 0) * dpow(e.T, G(a, 1)) * dexp(-G(a, 2) / e.RT);            // Rate.jl:16
}

__device__ __forceinline__ void bounding_inds(double el, const double* x, int n, int& i0, int& i1) {
    // getBoundingIndsSorted, src/Tools.jl:26-40
    if (el <= x[0]) { i0 = 0; i1 = -1; return; }
    for (int i = 0; i < n; ++i) {
        if (x[i] >= el) {
            if (x[i] == el) { i0 = i; i1 = -1; } else { i0 = i - 1; i1 = i; }
            return;
        }
    }
    i0 = n - 1; i1 = -1;
}

template <class N> __device__ __forceinline__ N plog_interp(N lowk, N highk, N P, double Plow, double Phigh) {
    // Rate.jl:59 : lowk*10^(log10(P/Plow)/log10(Phigh/Plow)*log10(highk/lowk))
    return lowk * dexp10(dlog10(P / Plow) / log10(Phigh / Plow) * dlog10(highk / lowk));
}

// one PLOG block with per-pressure {m, m x Arrhenius}; returns value and advances the cursor
template <class N> __device__ inline N plog_block(const double*& c, const KinEnv<N>& e) {
    int nP = (int)c[0];
    const double* Ps = c + 1;
    const double* ent = c + 1 + nP;
    int i0, i1;
    bounding_inds(val(e.P), Ps, nP, i0, i1);
    N lowk = 0.0, highk = 0.0;
    const double* q = ent;
    for (int i = 0; i < nP; ++i) {
        int mcount = (int)q[0];
        if (i == i0 || i == i1) {
            N acc = 0.0;
            if (mcount == 1) acc = arr_sca(q + 1, e);
            else for (int k = 0; k < mcount; ++k) acc = acc + arr_sca(q + 1 + 3 * k, e);
            if (i == i0) lowk = acc; else highk = acc;
        }
        q += 1 + 3 * mcount;
    }
    c = q;
    if (i1 < 0) return lowk;
    return plog_interp(lowk, highk, e.P, Ps[i0], Ps[i1]);
}

template <class N> __device__ __forceinline__ N cheb_T(int n, N x) {  // Rate.jl:158-176
    if (n == 0) return N(1.0);
    if (n == 1) return x;
    N T0 = 1.0, T1 = x, Tn = 0.0;
    for (int i = 1; i < n; ++i) { Tn = 2.0 * x * T1 - T0; T0 = T1; T1 = Tn; }
    return Tn;
}

template <class N> __device__ inline N troe_F(const double* a, N T, N Pr) {
    // a = [a, T3, T1, T2]; Rate.jl:134-146
    if (a[2] == 0.0 && a[1] == 0.0) return N(1.0);
    N Fcent = (1 - a[0]) * dexp(-T / a[1]) + a[0] * dexp(-T / a[2]);
    if (a[3] != 0.0) Fcent = Fcent + dexp(-a[3] / T);
    const double d = 0.14;
    N lF = dlog10(Fcent);
    N n = 0.75 - 1.27 * lF;
    N c = -0.4 - 0.67 * lF;
    N lPr = dlog10(Pr);
    N x = (lPr + c) / (n - d * lPr);
    return dexp10(lF / (1 + x * x));
}

// kinetics(T=T,P=P,C=C) for every type except the vectorised-Arrhenius fast path handled by callers.
// `scalar_forms` forces the scalar expressions for the *_VEC types (the diffusion-limited path goes through
// getkf for every reaction, src/PhaseState.jl:184-223,249-252).  Not inlined: keeps the hot RHS loop small.
template <class N>
__device__ __noinline__ N eval_kin(int j, const KinEnv<N>* ep, N Ceff, bool scalar_forms) {
    const DevMech& m = mech();
    const KinEnv<N>& e = *ep;
    const double* a = m.kpar + G(m.koff, j);
    switch (G(m.ktype, j)) {
    case RMSB_KIN_ARRHENIUS_VEC:
        return scalar_forms ? arr_sca(a, e) : arr_vec(a, e);
    case RMSB_KIN_ARRHENIUS:
        return arr_sca(a, e);
    case RMSB_KIN_MULTIARRHENIUS: {  // Rate.jl:68-74
        int mc = (int)a[0];
        N out = 0.0;
        for (int k = 0; k < mc; ++k) out = out + arr_sca(a + 1 + 3 * k, e);
        return out;
    }
    case RMSB_KIN_PLOG_VEC: {        // Ratevec.jl:204-215 (vector form leaves) / Rate.jl:48-60
        int nP = (int)a[0];
        const double* Ps = a + 1;
        const double* ar = a + 1 + nP;
        int i0, i1;
        bounding_inds(val(e.P), Ps, nP, i0, i1);
        N lowk = scalar_forms ? arr_sca(ar + 3 * i0, e) : arr_vec(ar + 3 * i0, e);
        if (i1 < 0) return lowk;
        N highk = scalar_forms ? arr_sca(ar + 3 * i1, e) : arr_vec(ar + 3 * i1, e);
        return plog_interp(lowk, highk, e.P, Ps[i0], Ps[i1]);
    }
    case RMSB_KIN_PLOG: {
        const double* c = a;
        return plog_block(c, e);
    }
    case RMSB_KIN_MULTIPLOG: {       // Rate.jl:82-88
        int nk = (int)a[0];
        const double* c = a + 1;
        N out = 0.0;
        for (int k = 0; k < nk; ++k) out = out + plog_block(c, e);
        return out;
    }
    case RMSB_KIN_THIRDBODY:         // Rate.jl:98
        return Ceff * arr_sca(a, e);
    case RMSB_KIN_LINDEMANN: {       // Rate.jl:109-114
        N k0 = arr_sca(a + 3, e), kinf = arr_sca(a, e);
        N Pr = k0 * Ceff / kinf;
        return kinf * Pr / (1.0 + Pr);
    }
    case RMSB_KIN_TROE: {            // Rate.jl:129-147
        N k0 = arr_sca(a + 3, e), kinf = arr_sca(a, e);
        N Pr = k0 * Ceff / kinf;
        N Fv = troe_F(a + 6, e.T, Pr);
        return kinf * (Pr / (1 + Pr)) * Fv;
    }
    case RMSB_KIN_CHEBYSHEV_VEC: {   // Ratevec.jl:89-100 (P outer, T inner) / Rate.jl:194-209 (T outer)
        double Tmin = a[0], Tmax = a[1], Pmin = a[2], Pmax = a[3];
        int nT = (int)a[4], nP = (int)a[5];
        const double* cf = a + 6;
        N Tred = (2.0 / e.T - 1.0 / Tmin - 1.0 / Tmax) / (1.0 / Tmax - 1.0 / Tmin);
        N Pred = (2.0 * dlog10(e.P) - log10(Pmin) - log10(Pmax)) / (log10(Pmax) - log10(Pmin));
        N k = 0.0;
        if (!scalar_forms) {
            for (int jp = 0; jp < nP; ++jp)
                for (int it = 0; it < nT; ++it)
                    k = k + cf[it * nP + jp] * (cheb_T(it, Tred) * cheb_T(jp, Pred));
        } else {
            for (int it = 0; it < nT; ++it)
                for (int jp = 0; jp < nP; ++jp)
                    k = k + cf[it * nP + jp] * cheb_T(it, Tred) * cheb_T(jp, Pred);
        }
        return dexp10(k);
    }
    }
    return N(0.0);
}

// _calcdkdCeff, src/Calculators/Rate.jl:232-262
static __device__ __noinline__ double dk_dCeff(int j, const KinEnv<double>* ep, double Ceff) {
    const DevMech& m = mech();
    const KinEnv<double>& e = *ep;
    const double* a = m.kpar + G(m.koff, j);
    int kt = G(m.ktype, j);
    if (kt == RMSB_KIN_THIRDBODY) return arr_sca(a, e);
    double k0 = arr_sca(a + 3, e), kinf = arr_sca(a, e);
    double Pr = k0 * Ceff / kinf;
    if (kt == RMSB_KIN_LINDEMANN || (a[8] == 0.0 && a[7] == 0.0)) return k0 / ((1.0 + Pr) * (1.0 + Pr));
    const double* t = a + 6;
    double T = e.T;
    double Fcent = (1 - t[0]) * exp(-T / t[1]) + t[0] * exp(-T / t[2]);
    if (t[3] != 0.0) Fcent += exp(-t[3] / T);
    const double d = 0.14;
    double lF = log10(Fcent);
    double n = 0.75 - 1.27 * lF;
    double c = -0.4 - 0.67 * lF;
    double lPr = log10(Pr);
    double x = (lPr + c) / (n - d * lPr);
    double F = pow(10.0, lF / (1.0 + x * x));
    double dFdCeff = -log(10.0) * F * lF / ((1.0 + x * x) * (1.0 + x * x)) * 2.0 * x * 0.4342944819032518 / Pr * k0 / kinf / (n - d * lPr) * (1.0 + d * x);
    return k0 / ((1.0 + Pr) * (1.0 + Pr)) * F + kinf * Pr / (1.0 + Pr) * dFdCeff;
}

// Ceff = C + sum_i ns[i]*(eff_i-1)/V   (src/PhaseState.jl:53-58)
template <class N>
__device__ __forceinline__ N ceff_of(int j, N C, const double* ns, double V) {
    const DevMech& m = mech();
    int b = G(m.eff_ptr, j), en = G(m.eff_ptr, j + 1);
    if (b == en) return C;
    double acc = 0.0;
    for (int k = b; k < en; ++k) acc += ns[G(m.eff_spc, k)] * G(m.eff_val, k);
    return C + acc / V;
}

// mass-action net rate of reaction j: src/Reactor.jl:465-483
template <class N>
__device__ __forceinline__ N net_rate(int j, N kf, N kr, const double* cs) {
    const DevMech& m = mech();
    const int Nr = m.Nr;
    N fR = kf * cs[G(m.rx, j)];
    for (int k = 1; k < 4; ++k) { int s = G(m.rx, k * Nr + j); if (s < 0) break; fR = fR * cs[s]; }
    N rR = kr * cs[G(m.rx, 4 * Nr + j)];
    for (int k = 5; k < 8; ++k) { int s = G(m.rx, k * Nr + j); if (s < 0) break; rR = rR * cs[s]; }
    return fR - rR;
}

// scalar getKc (src/PhaseState.jl:106-129, phi = 0) from absolute Gibbs energies G[]
static __device__ __noinline__ double kc_scalar(int j, const double* G, double T, double C0) {
    const DevMech& m = mech();
    const int Nr = m.Nr;
    double acc = G[G(m.rx, j)];
    int nre = 1;
    for (int k = 1; k < 4; ++k) { int s = G(m.rx, k * Nr + j); if (s < 0) break; acc += G[s]; ++nre; }
    double dG = 0.0 - acc;
    acc = G[G(m.rx, 4 * Nr + j)];
    int npr = 1;
    for (int k = 5; k < 8; ++k) { int s = G(m.rx, k * Nr + j); if (s < 0) break; acc += G[s]; ++npr; }
    dG += acc;
    return exp(-(dG) / (RMSB_R * T)) * pow(C0, (double)(npr - nre));
}

// getDiffusiveRate (src/PhaseState.jl:87-103): Smoluchowski for 2 bodies, its generalisation for 3 and 4 (:92-101)
__device__ __forceinline__ double diffusive_rate(int j, int side, const double* diffs) {
    const DevMech& m = mech();
    const int Nr = m.Nr;
    int sp[4];
    int n = 0;
    for (int k = 0; k < 4; ++k) { sp[k] = G(m.rx, (side * 4 + k) * Nr + j); if (sp[k] >= 0) ++n; else break; }
    if (n < 2) return INFINITY;
    if (n == 2) return 4.0 * RMSB_PI * (diffs[sp[0]] + diffs[sp[1]]) * (G(m.radius, sp[0]) + G(m.radius, sp[1])) * RMSB_NA;
    double D[4], Dinv[4], csum[4];
    double sinv = 0.0, rsum = 0.0;
    for (int k = 0; k < n; ++k) { D[k] = diffs[sp[k]]; Dinv[k] = 1.0 / D[k]; sinv += Dinv[k]; csum[k] = sinv; rsum += G(m.radius, sp[k]); }
    const double a = (3.0 * n - 5.0) / 2.0;
    // Dbar = 1 ./ reverse(cumsum(Dinv)); Dhat = D .+ Dbar; kf ~ prod(Dhat[2:end].^1.5)
    double prod = 1.0;
    for (int k = 1; k < n; ++k) { const double Dhat = D[k] + 1.0 / csum[n - 1 - k]; prod *= Dhat * sqrt(Dhat); }
    double pairs = 0.0;                                    // sum_{i=2..N} sum_{m<i, m<=N-1} 1/(D_i D_m)
    for (int i = 1; i < n; ++i) for (int q = 0; q < i && q < n - 1; ++q) pairs += 1.0 / (D[i] * D[q]);
    const double deltaN = sinv / pairs;
    return prod * 4.0 * pow(RMSB_PI, a + 1.0) / tgamma(a) * pow(rsum / sqrt(deltaN), 2.0 * a) * pow(RMSB_NA, (double)(n - 1));
}

struct DomainArgs {
    int kind;             // RMSB_DOMAIN_*
    int nconst;
    const double* p;      // nullptr or per-reactor parameter row
    const int* constspc;  // 1-based (reference convention)
    const double* cond;   // per-reactor constants
    const double* flow = nullptr;   // boundary interfaces of this reactor (IF_* layout below) or nullptr
    const double* ptab = nullptr;   // time profile of the domain's "constant" (ParametrizedV: V(t), ParametrizedP: P(t)) or nullptr
    double t = 0.0;                 // time of the evaluation (only the profile looks at it: the chemistry is autonomous)
};

// Time profile: a natural cubic spline through (ts, vs) -- the role of getspline (src/Tools.jl:46-85; the reference fits a
// third-party smoothing spline with lambda = 1e-10, which is this interpolating spline to ~1e-10) -- stored as
// [nt, ts[nt], vs[nt], gamma[nt]] with gamma the second derivatives (gamma[0] = gamma[nt-1] = 0); linear extrapolation outside.
// Returns the value; *dv = the ANALYTIC derivative (the reference differentiates numerically, Calculus.derivative, Domain.jl:2003,2047).
__device__ __forceinline__ double profile_eval(const double* tab, double x, double* dv) {
    const int nt = (int)tab[0];
    const double* ts = tab + 1;
    const double* vs = ts + nt;
    const double* gm = vs + nt;
    if (x <= ts[0]) {
        const double h = ts[1] - ts[0];
        const double gp = (vs[1] - vs[0]) / h - h * gm[1] / 6.0;
        *dv = gp;
        return vs[0] - (ts[0] - x) * gp;
    }
    if (x >= ts[nt - 1]) {
        const double h = ts[nt - 1] - ts[nt - 2];
        const double gp = (vs[nt - 1] - vs[nt - 2]) / h + h * gm[nt - 2] / 6.0;
        *dv = gp;
        return vs[nt - 1] + (x - ts[nt - 1]) * gp;
    }
    int lo = 0, hi = nt - 1;
    while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if (ts[mid] <= x) lo = mid; else hi = mid; }
    const double xl = ts[lo], xr = ts[hi], h = xr - xl, gl = vs[lo], gr = vs[hi], cl = gm[lo], cr = gm[hi];
    const double a = x - xl, b = xr - x;
    // src/Tools.jl:76-83: val = ((x-xl)*gr + (xr-x)*gl)/h - 1/6*(x-xl)*(xr-x)*((1+(x-xl)/h)*gamma_r + (1+(xr-x)/h)*gamma_l)
    const double w = (1.0 + a / h) * cr + (1.0 + b / h) * cl;
    *dv = (gr - gl) / h - ((b - a) * w + a * b * (cr - cl) / h) / 6.0;
    return (a * gr + b * gl) / h - a * b * w / 6.0;
}

// Boundary interfaces with constant flows (src/Interface.jl:234-365), one row of 7 + 2*Ns doubles per reactor.  Several interfaces
// of one type are merged by the host (the terms are linear in flow*composition and flow*enthalpy).
//   [0] Inlet molar flow F   [1] Outlet molar flow   [2] VolumetricFlowRateInlet Vin   [3] VolumetricFlowRateOutlet Vout
//   [4] VolumeMaintainingOutlet present (0 / 1)   [5] Inlet enthalpy H (J/mol)   [6] VolumetricFlowRateInlet Hpervolume (J/m^3)
//   [7 .. 7+Ns) Inlet mole fractions   [7+Ns .. 7+2Ns) VolumetricFlowRateInlet concentrations
constexpr int IF_FIN = 0, IF_FOUT = 1, IF_VIN = 2, IF_VOUT = 3, IF_VM = 4, IF_HIN = 5, IF_HPV = 6, IF_Y = 7;

// kf / krev of reaction j for the constant-k domains (getkfkrevs, src/PhaseState.jl:184-257): vector path without diffusion
// limits, per-reaction scalar path with them.  G = absolute Gibbs energies, diffs = Stokes diffusivities (liquid only).
__device__ __forceinline__ void constk_rates(int j, const double* p, bool is_eff_dom, const double* G, const double* diffs,
                                             const KinEnv<double>& e, double C0, double logC0, double C, const double* ns,
                                             double V, double T, double& kf_out, double& kr_out) {
    const DevMech& m = mech();
    const int Ns = m.Ns, Nr = m.Nr;
    double kf, kr;
    bool eff = (G(m.eff_ptr, j + 1) > G(m.eff_ptr, j));
    if (!m.difflim) {
        // getkfkrevs vector path: getkfs + getKcs (src/PhaseState.jl:226-233)
        if (p && !(is_eff_dom && eff)) kf = p[Ns + j];
        else kf = eval_kin<double>(j, &e, ceff_of(j, C, ns, V), false);
        double ex = 0.0;
        for (int k = G(m.kc_ptr, j); k < G(m.kc_ptr, j + 1); ++k) ex += G(m.kc_coef, k) * (G[G(m.kc_spc, k)] / (RMSB_R * T));
        ex += G(m.nrp, j) * logC0;
        kr = kf / exp(ex);
    } else {
        // getkfkrev per reaction (src/PhaseState.jl:184-223): scalar forms, scalar Kc, diffusion limits
        if (p) kf = p[Ns + j];
        else kf = eval_kin<double>(j, &e, ceff_of(j, C, ns, V), true);
        double Kc = kc_scalar(j, G, T, C0);
        kr = kf / Kc;
        int nre = 0, npr = 0;
        for (int k = 0; k < 4; ++k) { if (G(m.rx, k * Nr + j) >= 0) ++nre; if (G(m.rx, (k + 4) * Nr + j) >= 0) ++npr; }
        if (nre == 1) {
            if (npr > 1) {
                double krd = diffusive_rate(j, 1, diffs);
                kr = kr * krd / (kr + krd);
                kf = Kc * kr;
            }
        } else if (npr == 1) {
            double kfd = diffusive_rate(j, 0, diffs);
            kf = kf * kfd / (kf + kfd);
            kr = kf / Kc;
        } else if (npr == nre) {
            double kfd = diffusive_rate(j, 0, diffs);
            double krd = diffusive_rate(j, 1, diffs);
            double kff = kf * kfd / (kf + kfd);
            double krevr = kr * krd / (kr + krd);
            double kfr = Kc * krevr;
            if (kff > kfr) { kf = kfr; kr = krevr; }
            else { kf = kff; kr = kf / Kc; }
        }
    }
    kf_out = kf; kr_out = kr;
}

// ------------------------------------------------------------------------------------------------
// Constant-k domain setup (TP / TV / TAPhi ctors, src/Domain.jl:42-102,500-560,660-711):
// Gs via the scalar thermo path, kfs/krevs once per reactor.  w.g <- absolute G, w.kf/w.kr <- k's.
// The initial amounts are read from w.ybuf (only efficiency reactions look at them, and those are
// re-evaluated on every RHS call for TP).
// ------------------------------------------------------------------------------------------------
static __device__ __noinline__ void setup_constk(DomainArgs d, Ws w, int lane) {
    const DevMech& m = mech();
    const int Ns = m.Ns, Nr = m.Nr;
    const double* p = d.p;
    const double* ns = w.ybuf();
    double* G = w.g();
    double* diffs = w.un();
    // ParametrizedTConstantV (ConstantTV with a time profile, src/Domain.jl:1372-1434): T = T(t), Gibbs energies through the vector
    // thermo path with p = [dG; f] (enthalpy shifts and rate factors) instead of the constant-k convention p = [Gs; kfs]
    const bool param = d.ptab != nullptr;
    double dTdt_unused;
    const double T = param ? profile_eval(d.ptab, d.t, &dTdt_unused) : d.cond[0];
    double P, V;
    if (d.kind == RMSB_DOMAIN_CONSTANT_TP) { P = d.cond[1]; V = ns[Ns]; }
    else { P = (d.kind == RMSB_DOMAIN_CONSTANT_TV) ? 1.0e8 : 0.0; V = d.cond[1]; }
    if (param) {
        const TPow<double> tp(T);
        for (int s = lane; s < Ns; s += 32) {
            double cp, h, sr;
            nasa_hscp(s, tp, T, cp, h, sr);
            if (p) h = h + p[s] / (RMSB_R * T);
            G[s] = (h - sr) * (RMSB_R * T);
        }
    } else {
        for (int s = lane; s < Ns; s += 32) G[s] = p ? p[s] : gibbs_scalar(s, T);
    }
    if (m.difflim) {   // Stokes diffusivities, src/Calculators/Diffusion.jl:19, Viscosity.jl:24
        double mu = 0.0;
        if (m.visc_kind == RMSB_VISC_RIEDEL)
            mu = exp(m.visc[0] + m.visc[1] / T + m.visc[2] * log(T) + m.visc[3] * pow(T, m.visc[4]));
        else if (m.visc_kind == RMSB_VISC_CONSTANT)
            mu = m.visc[0];
        for (int s = lane; s < Ns; s += 32) {
            double r = m.diff_r ? G(m.diff_r, s) : 0.0;
            diffs[s] = (r > 0.0) ? RMSB_KB * T / (6 * RMSB_PI * mu * r) : INFINITY;
        }
    }
    __syncwarp();
    KinEnv<double> e;
    e.T = T; e.logT = log(T); e.RT = RMSB_R * T; e.invRT = 1.0 / (RMSB_R * T); e.P = P;
    double C0 = (m.phase_kind == RMSB_PHASE_SURFACE) ? m.sitedensity : 1.0e5 / (RMSB_R * T);
    double logC0 = log(C0);
    double N = 0.0;
    for (int s = 0; s < Ns; ++s) N += ns[s];
    double C = (d.kind == RMSB_DOMAIN_CONSTANT_TP) ? P / (RMSB_R * T) : (d.kind == RMSB_DOMAIN_CONSTANT_TV ? N / V : 0.0);
    bool is_eff_dom = (d.kind == RMSB_DOMAIN_CONSTANT_TP);
    for (int j = lane; j < Nr; j += 32) {
        double kf, kr;
        constk_rates(j, param ? nullptr : p, is_eff_dom, G, diffs, e, C0, logC0, C, ns, V, T, kf, kr);
        if (param && p) { kf *= p[Ns + j]; kr *= p[Ns + j]; }
        w.kf()[j] = kf * G(m.fwd, j);          // forwardability / reversibility masks (:220-221, :254-255)
        w.kr()[j] = kr * G(m.rev, j);
    }
    __syncwarp();
}

// ------------------------------------------------------------------------------------------------
// dydtreactor! for one reactor (src/Reactor.jl:653-660).
//   in : w.ybuf   state values (the Dual2 instantiation seeds dT = 1 on T and dP = 1 on P itself)
//        w.aux[7] 1/V (ConstantV) -- per-reactor constant, set once by the caller
//   out: w.fbuf   dy/dt (N-capable); w.aux[0..6] <- T P V N C Cv dot(U,dn) (Cv only with store_k);
//        with store_k also kf/kr/us/cp for the Jacobian
// Arithmetic note (DESIGN.md "floating-point contract"): the reference is compiled under @fastmath and is
// itself reproducible only to a few ulp; this kernel keeps the reference's formulas but replaces divisions by
// a shared reciprocal where the reference divides by the same quantity (1/T, 1/(R*T), 1/V) and evaluates
// kf/Kc as kf*exp(-x).  Measured agreement with the CPU restatement: <= 1e-13 relative (tests/test_gpu_parity.py).
// ------------------------------------------------------------------------------------------------
#define RMSB_INV_R (1.0 / 8.314)
#define RMSB_LOG_1E5_OVER_R 9.394984624139783   /* log(1.0e5/8.314) */

// ISP = false: ConstantV  y = [ns; T; P], cond = [V]   (src/Domain.jl:1064-1113,1901-1906)
// ISP = true : ConstantP  y = [ns; T; V], cond = [P]   (src/Domain.jl:1141-1192,1950-1953): enthalpies instead of internal
//              energies, dT/dt = -dot(Hs,dn)/(N*Cp), dV/dt = sum(dn)*R*T/P + dT/dt*V/T; only T carries a derivative seed
template <class N, bool ISP = false>
__device__ __noinline__ void rhs_v(DomainArgs d, Ws w, int lane, bool store_k) {
    const DevMech& m = mech();
    const int Ns = m.Ns, Nr = m.Nr;
    const double* p = d.p;
    typedef SmemNum<N> SM;
    const double* yb = w.ybuf();
    double* cs = w.cs();
    double* fb = w.fbuf();
    double* qb = w.q();
    double* gb = w.g();
    double* ub = w.un();
    // ParametrizedVDomain / ParametrizedPDomain (src/Domain.jl:1219-1370,2003-2090): the domain's constant follows a time profile
    double dprof = 0.0;
    const double prof = d.ptab ? profile_eval(d.ptab, d.t, &dprof) : d.cond[0];
    const double V = ISP ? yb[Ns + 1] : prof;
    const double invV = (ISP || d.ptab) ? 1.0 / V : w.aux()[7];
    N T = seed<N>(yb[Ns], 0), P = ISP ? N(prof) : seed<N>(yb[Ns + 1], 1);
    TPow<N> tp(T);                       // T4 = 1/T, T6 = log(T)
    N RT = RMSB_R * T;
    N invRT = tp.T4 * RMSB_INV_R;        // 1/(R*T)
    N Nt = P * V * invRT;                // N = P*V/(R*T)  (Domain.jl:1068)
    N C = P * invRT;                     // C = N/V
    N cpns = 0.0;
    for (int s = lane; s < Ns; s += 32) {
        N cp, h, sr;
        nasa_hscp(s, tp, val(T), cp, h, sr);
        if (p) h = h + p[s] * invRT;                    // Domain.jl:1100
        N U = ISP ? h * RT : (h - 1.0) * RT;            // :1102 (internal energy) / :1152 (enthalpy)
        SM::st(gb, s, h - sr);                          // G/(R*T) with G = (h-s)*R*T (:1101, PhaseState.jl:166)
        SM::st(ub, s, U);
        double nsv = yb[s];
        cs[s] = nsv * invV;
        cpns = cpns + cp * nsv;
        if (store_k) { w.us()[s] = val(U); w.cp()[s] = val(cp); }
    }
    N cpsum = warp_sum(cpns);                           // dot(cpdivR, ns)
    __syncwarp();
    KinEnv<N> e;
    e.T = T; e.logT = tp.T6; e.RT = RT; e.invRT = invRT; e.P = P;
    N logC0 = RMSB_LOG_1E5_OVER_R - tp.T6;              // log(1e5/(R*T))
    for (int j = lane; j < Nr; j += 32) {
        N kf;
        if (G(m.ktype, j) == RMSB_KIN_ARRHENIUS_VEC) kf = arr_vec(m.kpar + G(m.koff, j), e);
        else kf = eval_kin<N>(j, &e, ceff_of(j, C, yb, V), false);
        N ex = 0.0;
        for (int k = G(m.kc_ptr, j); k < G(m.kc_ptr, j + 1); ++k) ex = ex + G(m.kc_coef, k) * SM::ld(gb, G(m.kc_spc, k));
        ex = ex + G(m.nrp, j) * logC0;
        N kr = kf * dexp(-ex);                          // kf/Kc, Kc = exp(ex)  (PhaseState.jl:165-167,230)
        kf = kf * G(m.fwd, j); kr = kr * G(m.rev, j);
        if (p) { kf = kf * p[Ns + j]; kr = kr * p[Ns + j]; }   // Domain.jl:1112
        if (store_k) { w.kf()[j] = val(kf); w.kr()[j] = val(kr); }
        SM::st(qb, j, net_rate(j, kf, kr, cs));
    }
    __syncwarp();
    N usum = 0.0, nsum = 0.0;
    for (int s = lane; s < Ns; s += 32) {
        N acc = 0.0;
        for (int k = G(m.sr_ptr, s); k < G(m.sr_ptr, s + 1); ++k) {
            int jr = G(m.sr_rxn, k);
            N qv = SM::ld(qb, jr & 0x7fffffff);
            acc = (jr < 0) ? acc - qv : acc + qv;
        }
        N dn = acc * V;                                  // dydt .*= V
        SM::st(fb, s, dn);
        usum = usum + SM::ld(ub, s) * dn;
        nsum = nsum + dn;
    }
    N dotU = warp_sum(usum), sumn = warp_sum(nsum);
    // dT/dt = -dot(Us,dn)/(N*Cv) with N*Cv = R*(dot(cp/R,ns) - N)     (Domain.jl:1103-1105,1902)
    // ConstantP: N*Cp = N*(Cv + R) = R*dot(cp/R,ns)                                    (Domain.jl:1951-1953)
    // with a profile: dT/dt = (-dot(Us,dn) - P*dV/dt)/(N*Cv), dP/dt = ... - P/V*dV/dt   (ParametrizedV, Domain.jl:2003-2006)
    //                 dT/dt = (-dot(Hs,dn) + V*dP/dt)/(N*Cp), dV/dt = ... - dP/dt*V/P   (ParametrizedP, Domain.jl:2047-2051)
    N dotE = ISP ? dotU - V * dprof : dotU + P * dprof;                               // dprof = 0 without a profile: exact
    N dT = ISP ? -dotE / (RMSB_R * cpsum) : -dotE / (RMSB_R * (cpsum - Nt));
    N dP = ISP ? sumn * RT / P + dT * V * tp.T4 : sumn * RT * invV + P * tp.T4 * dT;   // :1903 / :1953 (dV/dt for ConstantP)
    if (d.ptab) dP = ISP ? dP - dprof * V / P : dP - P * invV * dprof;
    if (lane == 0) {
        SM::st(fb, Ns, dT);
        SM::st(fb, Ns + 1, dP);
        double* ax = w.aux();
        // ax[6]: the numerator of -dT/dt*(N*C) -- what the Jacobian's d/dn_i(dT/dt) rows multiply by dC/C (Domain.jl:2743-2746,2892)
        ax[0] = val(T); ax[1] = val(P); ax[2] = V; ax[3] = val(Nt); ax[4] = val(C); ax[6] = val(dotE);
        if (ISP) ax[7] = dprof;                                                       // dP/dt of the profile for the volume column
        if (store_k) ax[5] = val(cpsum) * (RMSB_R / val(Nt)) - RMSB_R;   // Cvave
    }
    __syncwarp();
    for (int k = lane; k < d.nconst; k += 32) SM::st(fb, d.constspc[k] - 1, N(0.0));   // Domain.jl:1904-1906
    __syncwarp();
    if (d.flow) {
        // boundary interfaces (src/Domain.jl:1907-1947 ConstantV, 1957-1997 ConstantP), in N arithmetic: the T / P columns of the
        // Jacobian are the derivative of the COMPLETE right-hand side (jacobianytherm!, src/Reactor.jl:1469-1477)
        const double* fl = d.flow;
        const double Fin = fl[IF_FIN], Fout = fl[IF_FOUT], Vin = fl[IF_VIN], Vout = fl[IF_VOUT];
        N un = 0.0;
        for (int s = lane; s < Ns; s += 32) un = un + SM::ld(ub, s) * yb[s];
        N dotUn = warp_sum(un);                                   // dot(Us, ns) (ConstantP: dot(Hs, ns))
        N NC = ISP ? RMSB_R * cpsum : RMSB_R * (cpsum - Nt);      // N*Cv (N*Cp)
        N iNt = 1.0 / Nt;
        N dTi = 0.0, dPi = 0.0;
        if (!ISP) {
            N a1 = Fin * (fl[IF_HIN] - dotUn * iNt) / NC;          dTi = dTi + a1; dPi = dPi + (Fin * RT * invV + P * tp.T4 * a1);
            N a2 = (P * V * iNt * Fout) / NC;                      dTi = dTi - a2; dPi = dPi - (Fout * RT * invV + P * tp.T4 * a2);
            N a3 = Vin * (fl[IF_HPV] - dotUn * invV) / NC;         dTi = dTi + a3; dPi = dPi + (Vin * P * invV + P * tp.T4 * a3);
            N a4 = (P * Vout) / NC;                                dTi = dTi - a4; dPi = dPi - (Vout * P * invV + P * tp.T4 * a4);
        } else {
            N a1 = Fin * (fl[IF_HIN] - dotUn * iNt) / NC;          dTi = dTi + a1; dPi = dPi + (Fin * RT / P + a1 * V * tp.T4);
            dPi = dPi - Fout * RT / P;
            N a3 = Vin * (fl[IF_HPV] - dotUn * invV) / NC;         dTi = dTi + a3; dPi = dPi + Vin;
            dPi = dPi - Vout;
        }
        for (int s = lane; s < Ns; s += 32) {
            N v = SM::ld(fb, s);
            v = v + fl[IF_Y + s] * Fin;
            v = v - Fout * yb[s] * iNt;
            v = v + Vin * fl[IF_Y + Ns + s];
            v = v - Vout * yb[s] * invV;
            SM::st(fb, s, v);
        }
        __syncwarp();
        N fT = SM::ld(fb, Ns) + dTi, fP = SM::ld(fb, Ns + 1) + dPi;
        __syncwarp();
        if (ISP && fl[IF_VM] != 0.0) {                            // VolumeMaintainingOutlet: after everything else (:1990-1997)
            N flowm = P * fP / RT;
            for (int s = lane; s < Ns; s += 32) SM::st(fb, s, SM::ld(fb, s) - flowm * yb[s] * iNt);
            fP = fP - fP;
        }
        if (lane == 0) { SM::st(fb, Ns, fT); SM::st(fb, Ns + 1, fP); }
        __syncwarp();
    }
}

// constant-k domains: TP (y = [ns; V]), TV / TAPhi (y = ns).  setup_constk() must have filled w.g / w.kf / w.kr.
static __device__ __noinline__ void rhs_ck(DomainArgs d, Ws w, int lane) {
    const DevMech& m = mech();
    const int Ns = m.Ns, Nr = m.Nr;
    const double* p = d.p;
    const double* yb = w.ybuf();
    double* cs = w.cs();
    double* fb = w.fbuf();
    double* qb = w.q();
    double dTdt_unused;
    const double T = (d.kind == RMSB_DOMAIN_CONSTANT_TV && d.ptab) ? profile_eval(d.ptab, d.t, &dTdt_unused) : d.cond[0];
    double P, V;
    if (d.kind == RMSB_DOMAIN_CONSTANT_TP) { P = d.cond[1]; V = yb[Ns]; }
    else { P = (d.kind == RMSB_DOMAIN_CONSTANT_TV) ? 1.0e8 : 0.0; V = d.cond[1]; }
    const double invV = 1.0 / V;
    double a = 0.0;
    for (int s = lane; s < Ns; s += 32) { double nsv = yb[s]; cs[s] = nsv * invV; a += nsv; }
    double Nt, C;
    if (d.kind == RMSB_DOMAIN_CONSTANT_TP) { Nt = P * V / (RMSB_R * T); C = Nt * invV; }
    else { Nt = warp_sum(a); C = Nt * invV; }
    __syncwarp();
    double* kfb = w.kf();
    double* krb = w.kr();
    if (d.kind == RMSB_DOMAIN_CONSTANT_TP && m.n_effrxn > 0) {
        // efficiency reactions are re-evaluated each call through getkfkrev (Domain.jl:924,943-945)
        KinEnv<double> e;
        e.T = T; e.logT = log(T); e.RT = RMSB_R * T; e.invRT = 1.0 / (RMSB_R * T); e.P = P;
        double C0 = 1.0e5 / (RMSB_R * T);
        for (int q = lane; q < m.n_effrxn; q += 32) {
            int j = G(m.effrxn, q);
            double kf = eval_kin<double>(j, &e, ceff_of(j, C, yb, V), false);
            if (p) kf = kf * p[Ns + j];
            double kr = kf / kc_scalar(j, w.g(), T, C0);
            kfb[j] = kf * G(m.fwd, j); krb[j] = kr * G(m.rev, j);
        }
        __syncwarp();
    }
    for (int j = lane; j < Nr; j += 32) qb[j] = net_rate<double>(j, kfb[j], krb[j], cs);
    __syncwarp();
    // dn_s/dt = V * sum over the species' reaction slots of +-q_j.  A species that takes part in most reactions (O2, ROO. in the
    // liquid oxidation mechanisms: hundreds of slots against an average of 40) would serialise the warp behind its lane: rows longer
    // than COOP_ROW are summed by the whole warp (32 contiguous pieces, then a tree; the short rows keep the reference's sequential order) -- ncu on C3_35:
    // this gather was 35 % of the kernel's instructions
    double nsum = 0.0;
    if (!m.heavy_sr) {
        for (int s = lane; s < Ns; s += 32) {
            double acc = 0.0;
            for (int k = G(m.sr_ptr, s); k < G(m.sr_ptr, s + 1); ++k) {
                int jr = G(m.sr_rxn, k);
                double qv = qb[jr & 0x7fffffff];
                acc = (jr < 0) ? acc - qv : acc + qv;
            }
            double dn = acc * V;
            fb[s] = dn;
            nsum += dn;
        }
    } else
    for (int s0 = 0; s0 < Ns; s0 += 32) {
        const int s = s0 + lane;
        int k0 = 0, k2 = 0;
        if (s < Ns) { k0 = G(m.sr_ptr, s); k2 = G(m.sr_ptr, s + 1); }
        const bool heavy = (k2 - k0) > COOP_ROW;
        if (s < Ns && !heavy) {
            double acc = 0.0;
            for (int k = k0; k < k2; ++k) {
                int jr = G(m.sr_rxn, k);
                double qv = qb[jr & 0x7fffffff];
                acc = (jr < 0) ? acc - qv : acc + qv;
            }
            double dn = acc * V;
            fb[s] = dn;
            nsum += dn;
        }
        unsigned hv = __ballot_sync(FULL, heavy);
        while (hv) {
            const int src = __ffs(hv) - 1;
            hv &= hv - 1;
            const int h0 = __shfl_sync(FULL, k0, src), h2 = __shfl_sync(FULL, k2, src);
            const int chunk = (h2 - h0 + 31) >> 5;                  // contiguous pieces of the list, each in the reference's order
            const int kb = h0 + lane * chunk, ke = min(h2, kb + chunk);
            double a = 0.0;
            for (int k = kb; k < ke; ++k) { int jr = G(m.sr_rxn, k); double qv = qb[jr & 0x7fffffff]; a = (jr < 0) ? a - qv : a + qv; }
            const double dn = warp_sum(a) * V;
            if (lane == 0) { fb[s0 + src] = dn; nsum += dn; }
        }
    }
    if (d.kind == RMSB_DOMAIN_CONSTANT_TP) {
        double dV = warp_sum(nsum) * RMSB_R * T / P;     // Domain.jl:1864
        if (lane == 0) fb[Ns] = dV;
    }
    if (lane == 0) {
        double* ax = w.aux();
        ax[0] = T; ax[1] = P; ax[2] = V; ax[3] = Nt; ax[4] = C; ax[5] = 0.0; ax[6] = 0.0;
    }
    __syncwarp();
    for (int k = lane; k < d.nconst; k += 32) fb[d.constspc[k] - 1] = 0.0;   // Domain.jl:1840-1842,1865-1867
    __syncwarp();
    if (d.flow) {
        // boundary interfaces: src/Domain.jl:1843-1860 (ConstantTV / ConstantTAPhi: species only), 1868-1898 (ConstantTP)
        const double* fl = d.flow;
        const double Fin = fl[IF_FIN], Fout = fl[IF_FOUT], Vin = fl[IF_VIN], Vout = fl[IF_VOUT];
        const double iNt = 1.0 / Nt;
        for (int s = lane; s < Ns; s += 32) {
            double v = fb[s];
            v += fl[IF_Y + s] * Fin;
            v -= Fout * yb[s] * iNt;
            v += Vin * fl[IF_Y + Ns + s];
            v -= Vout * yb[s] * invV;
            fb[s] = v;
        }
        __syncwarp();
        if (d.kind == RMSB_DOMAIN_CONSTANT_TP) {
            const double rtp = RMSB_R * T / P;
            double dV = fb[Ns] + Fin * rtp - Fout * rtp + Vin - Vout;
            __syncwarp();
            if (fl[IF_VM] != 0.0) {                                // VolumeMaintainingOutlet (:1890-1897)
                const double flowm = P * dV / (RMSB_R * T);
                for (int s = lane; s < Ns; s += 32) fb[s] -= flowm * yb[s] * iNt;
                dV -= dV;
            }
            if (lane == 0) fb[Ns] = dV;
        }
        __syncwarp();
    }
}

// ------------------------------------------------------------------------------------------------
// Coupled ConstantTP gas + ConstantTAPhi surface + ReactiveInternalInterfaceConstantTPhi (config C4):
// y = [ns_bulk; ns_surf; V], cond = [T, P, A].  Multi-domain dydtreactor! src/Reactor.jl:662-725; interface
// construction src/Interface.jl:89-108; two-phase getKc src/PhaseState.jl:131-162.  All rate coefficients are
// constants of the reactor (T fixed) and are computed once by setup_coupled.
// ------------------------------------------------------------------------------------------------
static __device__ __noinline__ void setup_coupled(DomainArgs d, Ws w, int lane) {
    const DevMech& m = mech();
    const int Ns = m.Ns, Nr = m.Nr;
    double* Gb = w.g();
    const double T = d.cond[0], P = d.cond[1];
    for (int s = lane; s < Ns; s += 32) Gb[s] = gibbs_scalar(s, T);       // calcgibbs of both constant-T domains
    __syncwarp();
    KinEnv<double> e;
    e.T = T; e.logT = log(T); e.RT = RMSB_R * T; e.invRT = 1.0 / (RMSB_R * T); e.P = P;
    const double C0g = 1.0e5 / (RMSB_R * T), sig = m.sitedensity;
    const double logC0g = log(C0g), logsig = log(sig);
    for (int j = lane; j < Nr; j += 32) {
        const int dom = G(m.rx_dom, j);
        double kf, kr;
        if (dom == 2) {
            // interface: kfs = getkf(rxn, nothing, T, ...), Kc = exp(-dG/RT)*C0_1^dN1*C0_2^dN2, krevs = kfs./Kc
            kf = eval_kin<double>(j, &e, 0.0, true);
            double dG = 0.0;
            for (int k = 0; k < 4; ++k) { int sp = G(m.rx, k * Nr + j); if (sp < 0) break; dG -= Gb[sp]; }
            for (int k = 4; k < 8; ++k) { int sp = G(m.rx, k * Nr + j); if (sp < 0) break; dG += Gb[sp]; }
            double Kc = exp(-(dG) / (RMSB_R * T)) * pow(C0g, G(m.nrp, j)) * pow(sig, G(m.nrp2, j));
            kr = kf / Kc;
        } else {
            // single-phase reactions of the TP / TAPhi domains: getkfs + getKcs (src/PhaseState.jl:226-233)
            kf = eval_kin<double>(j, &e, dom == 0 ? P / (RMSB_R * T) : 0.0, false);
            double ex = 0.0;
            for (int k = G(m.kc_ptr, j); k < G(m.kc_ptr, j + 1); ++k) ex += G(m.kc_coef, k) * (Gb[G(m.kc_spc, k)] / (RMSB_R * T));
            ex += (dom == 0) ? G(m.nrp, j) * logC0g : G(m.nrp2, j) * logsig;
            kr = kf / exp(ex);
        }
        w.kf()[j] = kf * G(m.fwd, j);
        w.kr()[j] = kr * G(m.rev, j);
    }
    __syncwarp();
}

// out[s] = sum over the reactions of species s of +-q_j.  RHS: the entries before sr_mid (reactions of the species' own domain)
// are multiplied by the domain size V or A (dydt[block] .*= V, src/Reactor.jl:662-725), interface reactions (already * A) are
// not; returns this lane's part of sum_{bulk} out[s].  A species that takes part in most reactions (the vacant site) would
// serialise the warp behind one lane: rows longer than COOP_ROW are summed by the whole warp.
template <bool RHS>
__device__ __forceinline__ double coupled_rows(const double* qb, double* out, int Ns, int Nb, double V, double A, int lane) {
    const DevMech& m = mech();
    double nsum = 0.0;
    for (int s0 = 0; s0 < Ns; s0 += 32) {
        const int s = s0 + lane;
        int k0 = 0, k1 = 0, k2 = 0;
        if (s < Ns) { k0 = G(m.sr_ptr, s); k1 = RHS ? G(m.sr_mid, s) : k0; k2 = G(m.sr_ptr, s + 1); }
        const bool heavy = (k2 - k0) > COOP_ROW;
        if (s < Ns && !heavy) {
            double acc = 0.0;
            for (int k = k0; k < k1; ++k) { int jr = G(m.sr_rxn, k); double qv = qb[jr & 0x7fffffff]; acc = (jr < 0) ? acc - qv : acc + qv; }
            acc *= (s < Nb ? V : A);
            for (int k = k1; k < k2; ++k) { int jr = G(m.sr_rxn, k); double qv = qb[jr & 0x7fffffff]; acc = (jr < 0) ? acc - qv : acc + qv; }
            out[s] = acc;
            if (s < Nb) nsum += acc;
        }
        unsigned hv = __ballot_sync(FULL, heavy);
        while (hv) {
            const int src = __ffs(hv) - 1;
            hv &= hv - 1;
            const int sh = s0 + src;
            const int h0 = __shfl_sync(FULL, k0, src), h1 = __shfl_sync(FULL, k1, src), h2 = __shfl_sync(FULL, k2, src);
            double a1 = 0.0, a2 = 0.0;
            for (int k = h0 + lane; k < h1; k += 32) { int jr = G(m.sr_rxn, k); double qv = qb[jr & 0x7fffffff]; a1 = (jr < 0) ? a1 - qv : a1 + qv; }
            for (int k = h1 + lane; k < h2; k += 32) { int jr = G(m.sr_rxn, k); double qv = qb[jr & 0x7fffffff]; a2 = (jr < 0) ? a2 - qv : a2 + qv; }
            const double acc = warp_sum(a1) * (sh < Nb ? V : A) + warp_sum(a2);
            if (lane == 0) { out[sh] = acc; if (sh < Nb) nsum += acc; }
        }
    }
    return nsum;
}

static __device__ __noinline__ void rhs_coupled(DomainArgs d, Ws w, int lane) {
    const DevMech& m = mech();
    const int Ns = m.Ns, Nr = m.Nr, Nb = m.n_bulk;
    const double* yb = w.ybuf();
    double* cs = w.cs();
    double* fb = w.fbuf();
    double* qb = w.q();
    const double T = d.cond[0], P = d.cond[1], A = d.cond[2];
    const double V = yb[Ns];
    for (int s = lane; s < Ns; s += 32) cs[s] = yb[s] / (s < Nb ? V : A);
    __syncwarp();
    const double* kfb = w.kf();
    const double* krb = w.kr();
    for (int j = lane; j < Nr; j += 32) {
        double q = net_rate<double>(j, kfb[j], krb[j], cs);
        if (G(m.rx_dom, j) == 2) q *= A;                                   // evaluate(): R = (fR - rR)*A
        qb[j] = q;
    }
    __syncwarp();
    double nsum = coupled_rows<true>(qb, fb, Ns, Nb, V, A, lane);
    double dV = warp_sum(nsum) * RMSB_R * T / P;                           // Domain.jl:1864 on the gas block
    if (lane == 0) {
        fb[Ns] = dV;
        double* ax = w.aux();
        ax[0] = T; ax[1] = P; ax[2] = V; ax[3] = P * V / (RMSB_R * T); ax[4] = P / (RMSB_R * T); ax[5] = A; ax[6] = 0.0;
    }
    __syncwarp();
    for (int k = lane; k < d.nconst; k += 32) fb[d.constspc[k] - 1] = 0.0;
    __syncwarp();
}

// Analytic Jacobian of the coupled system (the reference differentiates the whole RHS with ForwardDiff, src/Reactor.jl:203).
static __device__ __noinline__ void jac_coupled(DomainArgs d, Ws w, int lane) {
    const DevMech& m = mech();
    const int Ns = m.Ns, Nr = m.Nr, Nb = m.n_bulk, Ny = w.Ny, LD = w.LD;
    rhs_coupled(d, w, lane);
    double* J = w.J();
    const double* cs = w.cs();
    const double* kfb = w.kf();
    const double* krb = w.kr();
    double* dterm = w.dterm();
    const double T = d.cond[0], P = d.cond[1], A = d.cond[2];
    const double V = w.aux()[2];
    for (int i = lane; i < Ny * LD; i += 32) J[i] = 0.0;
    for (int t = lane; t < m.n_dt; t += 32) {
        int jr = G(m.dt_rxn, t);
        int j = jr & 0x3fffffff;
        double v = (double)G(m.dt_mult, t) * ((jr >> 30) ? krb[j] : kfb[j]);
        for (int k = 0; k < 3; ++k) { int s = G(m.dt_oth, 3 * t + k); if (s < 0) break; v *= cs[s]; }
        const double scale = (G(m.rx_dom, j) == 0) ? V : A;               // rate * domain size ...
        const double size = (G(m.dt_spc, t) < Nb) ? V : A;                // ... d(c_k)/d(n_k) = 1/size_k
        dterm[t] = v * scale / size;
    }
    __syncwarp();
    for (int e0 = 0; e0 < m.n_je; e0 += 32) {
        const int e = e0 + lane;
        int k0 = 0, k1 = 0, pos = 0;
        if (e < m.n_je) { k0 = G(m.je_ptr, e); k1 = G(m.je_ptr, e + 1); pos = G(m.je_row, e) + G(m.je_col, e) * LD; }
        const bool heavy = (k1 - k0) > COOP_ROW;
        if (e < m.n_je && !heavy) {
            double acc = 0.0;
            for (int k = k0; k < k1; ++k) { int ti = G(m.jt_idx, k); double v = dterm[ti & 0x7fffffff]; acc = (ti < 0) ? acc - v : acc + v; }
            J[pos] = acc;
        }
        unsigned hv = __ballot_sync(FULL, heavy);
        while (hv) {                                                        // (vacant site, vacant site) and the like
            const int src = __ffs(hv) - 1;
            hv &= hv - 1;
            const int h0 = __shfl_sync(FULL, k0, src), h1 = __shfl_sync(FULL, k1, src), hp = __shfl_sync(FULL, pos, src);
            const int chunk = (h1 - h0 + 31) >> 5;                  // contiguous pieces in the reference's order (see jac_warp)
            const int kb = h0 + lane * chunk, ke = min(h1, kb + chunk);
            double acc = 0.0;
            for (int k = kb; k < ke; ++k) { int ti = G(m.jt_idx, k); double v = dterm[ti & 0x7fffffff]; acc = (ti < 0) ? acc - v : acc + v; }
            acc = warp_sum(acc);
            if (lane == 0) J[hp] = acc;
        }
    }
    __syncwarp();
    // V column: d(scale_j * k prod c)/dV = (k prod c) * ([bulk reaction] - scale_j * (#bulk species in the side)/V)
    double* qb = w.q();                                                     // the rates of rhs_coupled are not needed any more
    for (int j = lane; j < Nr; j += 32) {
        const int dom = G(m.rx_dom, j);
        const double scale = (dom == 0) ? V : A;
        double pf = kfb[j], pr = krb[j];
        int gf = 0, gr = 0;
        for (int q = 0; q < 4; ++q) { int sp = G(m.rx, q * Nr + j); if (sp < 0) break; pf *= cs[sp]; if (sp < Nb) ++gf; }
        for (int q = 4; q < 8; ++q) { int sp = G(m.rx, q * Nr + j); if (sp < 0) break; pr *= cs[sp]; if (sp < Nb) ++gr; }
        double df = pf * ((dom == 0 ? 1.0 : 0.0) - scale * (double)gf / V);
        double dr = pr * ((dom == 0 ? 1.0 : 0.0) - scale * (double)gr / V);
        qb[j] = df - dr;
    }
    __syncwarp();
    coupled_rows<false>(qb, J + Ns * LD, Ns, Nb, V, A, lane);
    __syncwarp();
    for (int c = lane; c <= Ns; c += 32) {                                   // V row from the gas block
        double a = 0.0;
        for (int i = 0; i < Nb; ++i) a += J[i + c * LD];
        J[Ns + c * LD] = a * RMSB_R * T / P;
    }
    __syncwarp();
    for (int k = 0; k < d.nconst; ++k) {
        int s = d.constspc[k] - 1;
        for (int c = lane; c < Ny; c += 32) J[s + c * LD] = 0.0;
    }
    __syncwarp();
}

// per-reactor constants the RHS reads from the workspace (call once per reactor, before the first RHS)
// ParametrizedTPDomain (src/Domain.jl:241-320 ctor, 1437-1512 calcthermo, 2002-2004 derivatives): y = [ns; V], temperature and
// pressure follow time profiles (d.ptab = [T table][P table]).  A variable-k domain: thermo through the vector path and every
// rate coefficient on each call, exactly as rhs_v; closure dV/dt = sum(dn)*R*T/P + dT/dt*V/T - dP/dt*V/P.
// p = [dG_i ; f_j] as for the other variable-k domains (h/RT += dG/(RT), k *= f).
static __device__ __noinline__ void rhs_tp_param(DomainArgs d, Ws w, int lane, bool store_k) {
    const DevMech& m = mech();
    const int Ns = m.Ns, Nr = m.Nr;
    const double* p = d.p;
    const double* yb = w.ybuf();
    double* cs = w.cs();
    double* fb = w.fbuf();
    double* qb = w.q();
    double* gb = w.g();
    double dTdt = 0.0, dPdt = 0.0;
    const double T = profile_eval(d.ptab, d.t, &dTdt);
    const double P = profile_eval(d.ptab + 1 + 3 * (int)d.ptab[0], d.t, &dPdt);
    const double V = yb[Ns], invV = 1.0 / V;
    TPow<double> tp(T);
    const double RT = RMSB_R * T, invRT = tp.T4 * RMSB_INV_R;
    const double Nt = P * V * invRT, C = P * invRT;
    for (int s = lane; s < Ns; s += 32) {
        double cp, h, sr;
        nasa_hscp(s, tp, T, cp, h, sr);
        if (p) h = h + p[s] * invRT;
        gb[s] = h - sr;
        cs[s] = yb[s] * invV;
    }
    __syncwarp();
    KinEnv<double> e;
    e.T = T; e.logT = tp.T6; e.RT = RT; e.invRT = invRT; e.P = P;
    const double logC0 = RMSB_LOG_1E5_OVER_R - tp.T6;
    for (int j = lane; j < Nr; j += 32) {
        double kf;
        if (G(m.ktype, j) == RMSB_KIN_ARRHENIUS_VEC) kf = arr_vec(m.kpar + G(m.koff, j), e);
        else kf = eval_kin<double>(j, &e, ceff_of(j, C, yb, V), false);
        double ex = 0.0;
        for (int k = G(m.kc_ptr, j); k < G(m.kc_ptr, j + 1); ++k) ex = ex + G(m.kc_coef, k) * gb[G(m.kc_spc, k)];
        ex = ex + G(m.nrp, j) * logC0;
        double kr = kf * exp(-ex);
        kf = kf * G(m.fwd, j); kr = kr * G(m.rev, j);
        if (p) { kf = kf * p[Ns + j]; kr = kr * p[Ns + j]; }
        if (store_k) { w.kf()[j] = kf; w.kr()[j] = kr; }
        qb[j] = net_rate<double>(j, kf, kr, cs);
    }
    __syncwarp();
    double nsum = 0.0;
    for (int s = lane; s < Ns; s += 32) {
        double acc = 0.0;
        for (int k = G(m.sr_ptr, s); k < G(m.sr_ptr, s + 1); ++k) {
            int jr = G(m.sr_rxn, k);
            double qv = qb[jr & 0x7fffffff];
            acc = (jr < 0) ? acc - qv : acc + qv;
        }
        const double dn = acc * V;
        fb[s] = dn;
        nsum += dn;
    }
    const double dV = warp_sum(nsum) * RT / P + dTdt * V * tp.T4 - dPdt * V / P;
    if (lane == 0) {
        fb[Ns] = dV;
        double* ax = w.aux();
        ax[0] = T; ax[1] = P; ax[2] = V; ax[3] = Nt; ax[4] = C; ax[5] = dTdt; ax[6] = dPdt;
    }
    __syncwarp();
    for (int k = lane; k < d.nconst; k += 32) fb[d.constspc[k] - 1] = 0.0;
    __syncwarp();
    if (d.flow) {                                        // src/Domain.jl:2007-2044: the ConstantTP interface terms at the current T, P
        const double* fl = d.flow;
        const double Fin = fl[IF_FIN], Fout = fl[IF_FOUT], Vin = fl[IF_VIN], Vout = fl[IF_VOUT];
        const double iNt = 1.0 / Nt, rtp = RT / P;
        for (int s = lane; s < Ns; s += 32) {
            double v = fb[s];
            v += fl[IF_Y + s] * Fin; v -= Fout * yb[s] * iNt; v += Vin * fl[IF_Y + Ns + s]; v -= Vout * yb[s] * invV;
            fb[s] = v;
        }
        __syncwarp();
        double dVi = fb[Ns] + Fin * rtp - Fout * rtp + Vin - Vout;
        __syncwarp();
        if (fl[IF_VM] != 0.0) {
            const double flowm = P * dVi / RT;
            for (int s = lane; s < Ns; s += 32) fb[s] -= flowm * yb[s] * iNt;
            dVi -= dVi;
        }
        if (lane == 0) fb[Ns] = dVi;
        __syncwarp();
    }
}

__device__ __forceinline__ void rhs_prepare(const DomainArgs& d, const Ws& w, int lane) {
    if (lane == 0) w.aux()[7] = (d.kind == RMSB_DOMAIN_CONSTANT_V) ? 1.0 / d.cond[0] : 0.0;
    __syncwarp();
}

__device__ __forceinline__ void rhs_warp(const DomainArgs& d, const Ws& w, int lane, bool store_k) {
    if (d.kind == RMSB_DOMAIN_CONSTANT_V) rhs_v<double>(d, w, lane, store_k);
    else if (d.kind == RMSB_DOMAIN_CONSTANT_P) rhs_v<double, true>(d, w, lane, store_k);
    else if (d.kind == RMSB_DOMAIN_CONSTANT_TP && d.ptab) rhs_tp_param(d, w, lane, store_k);
    else if (d.kind == RMSB_DOMAIN_CONSTANT_TV && d.ptab) { setup_constk(d, w, lane); rhs_ck(d, w, lane); }   // ParametrizedTConstantV: k's at T(t)
    else if (d.kind == RMSB_DOMAIN_COUPLED_TP_TAPHI) rhs_coupled(d, w, lane);
    else rhs_ck(d, w, lane);
}

// once per reactor for every domain whose rate coefficients are constants of the reactor
__device__ __forceinline__ void setup_domain(const DomainArgs& d, const Ws& w, int lane) {
    if (d.kind == RMSB_DOMAIN_COUPLED_TP_TAPHI) setup_coupled(d, w, lane);
    else if (d.kind != RMSB_DOMAIN_CONSTANT_V && d.kind != RMSB_DOMAIN_CONSTANT_P && !d.ptab) setup_constk(d, w, lane);   // ParametrizedTP: k's every call
}

// ------------------------------------------------------------------------------------------------
// jacobiany! for one reactor: in w.ybuf, out dense column-major Ny x LD matrix in w.J.
// ------------------------------------------------------------------------------------------------
static __device__ __noinline__ void jac_warp(DomainArgs d, Ws w, int lane) {
    const DevMech& m = mech();
    const int Ns = m.Ns, Nr = m.Nr, Ny = w.Ny, LD = w.LD;
    if (d.kind == RMSB_DOMAIN_COUPLED_TP_TAPHI) { jac_coupled(d, w, lane); return; }
    rhs_warp(d, w, lane, true);
    double* J = w.J();
    const double* cs = w.cs();
    const double* kfb = w.kf();
    const double* krb = w.kr();
    double* dterm = w.dterm();
    const double* ax = w.aux();
    const double aT = ax[0], aP = ax[1], aV = ax[2], aN = ax[3], aC = ax[4], aCv = ax[5], adotU = ax[6];
    for (int i = lane; i < Ny * LD; i += 32) J[i] = 0.0;
    // (1) derivative terms  m*k*prod(other cs)   (src/Reactor.jl:899-1366)
    for (int t = lane; t < m.n_dt; t += 32) {
        int jr = G(m.dt_rxn, t);
        int j = jr & 0x3fffffff;
        double v = (double)G(m.dt_mult, t) * ((jr >> 30) ? krb[j] : kfb[j]);
        for (int k = 0; k < 3; ++k) { int s = G(m.dt_oth, 3 * t + k); if (s < 0) break; v *= cs[s]; }
        dterm[t] = v;
    }
    __syncwarp();
    // (2) gather into entries, in the reference's accumulation order; entries with more than COOP_ROW terms (the diagonal and the
    //     O2 / ROO. rows and columns of the oxidation mechanisms) are summed by the whole warp, as the long species rows of rhs_ck are
    if (!m.heavy_je) {
        for (int e = lane; e < m.n_je; e += 32) {
            double acc = 0.0;
            for (int k = G(m.je_ptr, e); k < G(m.je_ptr, e + 1); ++k) {
                int ti = G(m.jt_idx, k);
                double v = dterm[ti & 0x7fffffff];
                acc = (ti < 0) ? acc - v : acc + v;
            }
            J[G(m.je_row, e) + G(m.je_col, e) * LD] = acc;
        }
    } else
    for (int e0 = 0; e0 < m.n_je; e0 += 32) {
        const int e = e0 + lane;
        int k0 = 0, k2 = 0;
        if (e < m.n_je) { k0 = G(m.je_ptr, e); k2 = G(m.je_ptr, e + 1); }
        const bool heavy = (k2 - k0) > COOP_ROW;
        unsigned hv = __ballot_sync(FULL, heavy);
        while (hv) {
            const int src = __ffs(hv) - 1;
            hv &= hv - 1;
            const int h0 = __shfl_sync(FULL, k0, src), h2 = __shfl_sync(FULL, k2, src);
            // every lane sums a CONTIGUOUS piece of the list in the reference's order: the +d / -d pairs a species on both sides of a
            // reaction contributes sit next to each other and must cancel inside one partial sum (dealt round-robin they left
            // O(d) partial sums whose rounding cost the propane case 38 % more steps after ignition)
            const int chunk = (h2 - h0 + 31) >> 5;
            const int kb = h0 + lane * chunk, ke = min(h2, kb + chunk);
            double a = 0.0;
            for (int k = kb; k < ke; ++k) { int ti = G(m.jt_idx, k); double v = dterm[ti & 0x7fffffff]; a = (ti < 0) ? a - v : a + v; }
            a = warp_sum(a);
            if (lane == 0) J[G(m.je_row, e0 + src) + G(m.je_col, e0 + src) * LD] = a;
        }
        if (e >= m.n_je || heavy) continue;
        double acc = 0.0;
        for (int k = k0; k < k2; ++k) {
            int ti = G(m.jt_idx, k);
            double v = dterm[ti & 0x7fffffff];
            acc = (ti < 0) ? acc - v : acc + v;
        }
        J[G(m.je_row, e) + G(m.je_col, e) * LD] = acc;
    }
    __syncwarp();
    const bool hasV = (d.kind == RMSB_DOMAIN_CONSTANT_TP || d.kind == RMSB_DOMAIN_CONSTANT_P);
    const int Vcol = (d.kind == RMSB_DOMAIN_CONSTANT_P) ? Ns + 1 : Ns;      // column of the volume in y
    // (3) V column of the ns block (src/Reactor.jl:1368-1413) -- lanes over species rows
    if (hasV) {
        // per-reaction pieces once (one lane per reaction), kept on the derivative-term array, which (1)-(2) are done with and which
        // holds at least 2 Nr entries (one term per direction and distinct species); the species sums then only gather
        double* dfv = dterm;
        double* drv = dterm + Nr;
        for (int j = lane; j < Nr; j += 32) {
            int nre = 0, npr = 0;
            double pf = kfb[j], pr = krb[j];
            for (int q = 0; q < 4; ++q) { int sp = G(m.rx, q * Nr + j); if (sp < 0) break; pf *= cs[sp]; ++nre; }
            for (int q = 4; q < 8; ++q) { int sp = G(m.rx, q * Nr + j); if (sp < 0) break; pr *= cs[sp]; ++npr; }
            dfv[j] = (nre > 1) ? -(double)(nre - 1) * pf : 0.0;
            drv[j] = (npr > 1) ? -(double)(npr - 1) * pr : 0.0;
        }
        __syncwarp();
        for (int s = lane; s < Ns; s += 32) {
            double acc = 0.0;
            for (int k = G(m.sr_ptr, s); k < G(m.sr_ptr, s + 1); ++k) {
                const int jr = G(m.sr_rxn, k);
                const int j = jr & 0x7fffffff;
                const double df = dfv[j], dr = drv[j];
                if (jr < 0) { acc -= df; acc += dr; } else { acc += df; acc -= dr; }
            }
            J[s + Vcol * LD] = acc;
        }
        __syncwarp();
    }
    // (4) efficiency reactions: dk/dn_i = dk/dCeff*(eff_i-1)/V (+ dk/dV for V-state domains)
    //     (src/Domain.jl:2151-2167, src/Reactor.jl:1418-1467).  One reaction at a time (uniform loop).
    if (m.n_effrxn > 0 && d.kind != RMSB_DOMAIN_CONSTANT_TAPHI) {
        KinEnv<double> e;
        e.T = aT; e.logT = log(aT); e.RT = RMSB_R * aT; e.invRT = 1.0 / (RMSB_R * aT); e.P = aP;
        for (int q = 0; q < m.n_effrxn; ++q) {
            int j = G(m.effrxn, q);
            double Kc = kfb[j] / krb[j];
            double acc = 0.0;
            for (int k = G(m.eff_ptr, j); k < G(m.eff_ptr, j + 1); ++k) acc += cs[G(m.eff_spc, k)] * G(m.eff_val, k);
            double Ceff = aC + acc;
            double dkdC = dk_dCeff(j, &e, Ceff);
            double a = 1.0, b = 1.0;
            for (int k = 0; k < 4; ++k) { int sp = G(m.rx, k * Nr + j); if (sp < 0) break; a *= cs[sp]; }
            for (int k = 4; k < 8; ++k) { int sp = G(m.rx, k * Nr + j); if (sp < 0) break; b *= cs[sp]; }
            double pf = a * aV, pr = b * aV;
            int ne = G(m.eff_ptr, j + 1) - G(m.eff_ptr, j);
            int ncol = ne + (hasV ? 1 : 0);
            for (int c = lane; c < ncol; c += 32) {
                int col; double dkf;
                if (c < ne) { col = G(m.eff_spc, G(m.eff_ptr, j) + c); dkf = dkdC * (G(m.eff_val, G(m.eff_ptr, j) + c) / aV); }
                else { col = Vcol; dkf = dkdC * (aC - Ceff) / aV; }
                double dkr = dkf / Kc;
                double dvf = dkf * pf, dvr = dkr * pr;
                for (int k = 0; k < 4; ++k) { int sp = G(m.rx, k * Nr + j); if (sp < 0) break; J[sp + col * LD] -= dvf; }
                for (int k = 4; k < 8; ++k) { int sp = G(m.rx, k * Nr + j); if (sp < 0) break; J[sp + col * LD] += dvf; }
                for (int k = 4; k < 8; ++k) { int sp = G(m.rx, k * Nr + j); if (sp < 0) break; J[sp + col * LD] -= dvr; }
                for (int k = 0; k < 4; ++k) { int sp = G(m.rx, k * Nr + j); if (sp < 0) break; J[sp + col * LD] += dvr; }
            }
            __syncwarp();
        }
    }
    // (5) thermodynamic-variable rows
    if (d.kind == RMSB_DOMAIN_CONSTANT_TP) {
        for (int c = lane; c <= Ns; c += 32) {           // Domain.jl:2211-2214
            double a = 0.0;
            for (int i = 0; i < Ns; ++i) a += J[i + c * LD];
            double v = a * RMSB_R * aT / aP;
            if (d.ptab && c == Ns) v += ax[5] / aT - ax[6] / aP;     // ParametrizedTP: + dT/dt/T - dP/dt/P (Domain.jl:2646)
            J[Ns + c * LD] = v;
        }
    } else if (d.kind == RMSB_DOMAIN_CONSTANT_V) {
        const double* us = w.us();
        const double* cpb = w.cp();
        for (int c = lane; c < Ns; c += 32) {            // Domain.jl:2312-2316
            double dCv = cpb[c] * RMSB_R / aN;
            double a = 0.0, sm = 0.0;
            for (int i = 0; i < Ns; ++i) { double v = J[i + c * LD]; a += us[i] * v; sm += v; }
            double jt = -a / (aN * aCv) + adotU / (aN * aCv * aCv) * dCv;
            J[Ns + c * LD] = jt;
            J[Ns + 1 + c * LD] = sm * RMSB_R * aT / aV + aP / aT * jt;
        }
    } else if (d.kind == RMSB_DOMAIN_CONSTANT_P) {       // Domain.jl:2480-2490 (us holds the enthalpies, adotU = dot(Hs,dn))
        const double* hs = w.us();
        const double* cpb = w.cp();
        const double aCp = aCv + RMSB_R;
        const double dTdt = -adotU / (aN * aCp);
        for (int c = lane; c < Ns + 2; c += 32) {
            if (c == Ns) continue;                       // the T column is the derivative of the whole RHS, step (7)
            double a = 0.0, sm = 0.0;
            for (int i = 0; i < Ns; ++i) { double v = J[i + c * LD]; a += hs[i] * v; sm += v; }
            const double dPdt = d.ptab ? ax[7] : 0.0;
            double jt = -a / (aN * aCp);
            if (c < Ns) jt -= dTdt * ((cpb[c] * RMSB_R / aN) / aCp);
            else jt += dPdt / (aN * aCp);                       // ParametrizedP: d/dV of V*dP/dt (Domain.jl:2895)
            J[Ns + c * LD] = jt;
            double jv = sm * RMSB_R * aT / aP + aV / aT * jt;
            if (c == Ns + 1) jv += dTdt / aT - dPdt / aP;        // :2896
            J[Ns + 1 + c * LD] = jv;
        }
    }
    __syncwarp();
    // (6) constant species rows (Domain.jl:2216-2218,2318-2320,3053-3055)
    for (int k = 0; k < d.nconst; ++k) {
        int s = d.constspc[k] - 1;
        for (int c = lane; c < Ny; c += 32) J[s + c * LD] = 0.0;
    }
    __syncwarp();
    // (6b) boundary interfaces: the analytic terms of the species (and volume) columns exactly as the reference writes them
    //      (src/Domain.jl:2220-2294 ConstantTP, 2322-2448 ConstantV, 2493-2628 ConstantP, 3057-3088 ConstantTV / ConstantTAPhi);
    //      the VolumeMaintainingOutlet has none there.  Lanes own columns / diagonal entries: no two lanes touch one entry.
    if (d.flow) {
        const double* fl = d.flow;
        const double* yb = w.ybuf();
        const double Fin = fl[IF_FIN], Fout = fl[IF_FOUT], Vin = fl[IF_VIN], Vout = fl[IF_VOUT];
        if (d.kind == RMSB_DOMAIN_CONSTANT_TP) {
            for (int i = lane; i < Ns; i += 32) {
                J[i + i * LD] -= Fout / aN;
                J[i + Ns * LD] -= -Fout * yb[i] / (aN * aV);
                J[i + i * LD] -= Vout / aV;
                J[i + Ns * LD] -= -Vout * yb[i] / (aV * aV);
            }
        } else if (d.kind == RMSB_DOMAIN_CONSTANT_V || d.kind == RMSB_DOMAIN_CONSTANT_P) {
            const bool isp = d.kind == RMSB_DOMAIN_CONSTANT_P;
            const double* us = w.us();
            const double* cpb = w.cp();
            double un = 0.0;
            for (int s = lane; s < Ns; s += 32) un += us[s] * yb[s];
            const double dotUn = warp_sum(un);
            const double aC2 = isp ? aCv + RMSB_R : aCv;                      // Cv (ConstantV) or Cp (ConstantP)
            const double NC = aN * aC2;
            if (!isp) {
                const double dT1 = Fin * (fl[IF_HIN] - dotUn / aN) / NC;
                const double dT2 = (aP * aV / aN * Fout) / NC;
                const double dT3 = (Vin * (fl[IF_HPV] - dotUn / aV)) / NC;
                const double dT4 = (aP * Vout) / NC;
                for (int i = lane; i < Ns; i += 32) {
                    const double dC = cpb[i] * RMSB_R / aN;
                    double dd = Fin * (-us[i] / aN) / NC - dT1 * (dC / aC2);     // Inlet
                    J[Ns + i * LD] += dd; J[Ns + 1 + i * LD] += aP / aT * dd;
                    J[i + i * LD] -= Fout / aN;                                   // Outlet
                    dd = -dT2 * (dC / aC2);
                    J[Ns + i * LD] -= dd; J[Ns + 1 + i * LD] -= aP / aT * dd;
                    dd = Vin * (-us[i] / aN) / NC - dT3 * (dC / aC2);            // VolumetricFlowRateInlet
                    J[Ns + i * LD] += dd; J[Ns + 1 + i * LD] += aP / aT * dd;
                    J[i + i * LD] -= Vout / aV;                                   // VolumetricFlowRateOutlet
                    dd = -dT4 * (dC / aC2);
                    J[Ns + i * LD] -= dd; J[Ns + 1 + i * LD] -= aP / aT * dd;
                }
            } else {
                const double H = dotUn / aN;
                const double dT1 = Fin * (fl[IF_HIN] - H) / NC;
                const double dT3 = Vin * (fl[IF_HPV] - dotUn / aV) / NC;
                for (int i = lane; i < Ns; i += 32) {
                    const double dC = cpb[i] * RMSB_R / aN;
                    double dd = Fin * (-us[i] / aN) / NC - dT1 * (dC / aC2);     // Inlet
                    J[Ns + i * LD] += dd; J[Ns + 1 + i * LD] += aV / aT * dd;
                    J[i + i * LD] -= Fout / aN;                                   // Outlet
                    J[i + (Ns + 1) * LD] -= -Fout * yb[i] / aN / aV;
                    J[Ns + i * LD] += Vin * (-us[i] / aN) / NC - dT3 * (dC / aC2);   // VolumetricFlowRateInlet
                    J[i + i * LD] -= Vout / aV;                                   // VolumetricFlowRateOutlet
                    J[i + (Ns + 1) * LD] -= -Vout / (aV * aV) * yb[i];
                }
                if (lane == 0) {
                    const double ddV = Fin * H / aV / NC;
                    J[Ns + (Ns + 1) * LD] += ddV;
                    J[Ns + 1 + (Ns + 1) * LD] += aV / aT * ddV;
                }
            }
        } else {                                                              // ConstantTV / ConstantTAPhi: N = sum(ns)
            for (int i = lane; i < Ns; i += 32) {
                J[i + i * LD] -= Fout / aN;
                const double r = Fout * yb[i] / (aN * aN);
                for (int c = 0; c < Ny; ++c) J[i + c * LD] += r;
                J[i + i * LD] -= Vout / aV;
            }
        }
        __syncwarp();
    }
    // (7) ConstantV: T and P columns = derivative of the complete RHS (src/Reactor.jl:1469-1477)
    if (d.kind == RMSB_DOMAIN_CONSTANT_V) {
        rhs_v<Dual2>(d, w, lane, false);
        const double* fb = w.fbuf();
        for (int i = lane; i < Ny; i += 32) { J[i + Ns * LD] = fb[3 * i + 1]; J[i + (Ns + 1) * LD] = fb[3 * i + 2]; }
        __syncwarp();
    } else if (d.kind == RMSB_DOMAIN_CONSTANT_P) {      // T column only; the V column above is analytic (Domain.jl:2630)
        rhs_v<Dual2, true>(d, w, lane, false);
        const double* fb = w.fbuf();
        for (int i = lane; i < Ny; i += 32) J[i + Ns * LD] = fb[3 * i + 1];
        __syncwarp();
    }
}

// getkfs (src/PhaseState.jl:165-171): forward rate coefficients of every reaction BEFORE diffusion limitation, vector
// forms for the vectorised block -- what the constant-k constructors store in p[Ns+1:end] (kfsnondiff for the liquid
// phase, src/Domain.jl:549-551).  State in w.ybuf, result in w.kf.
static __device__ __noinline__ void kfs_intrinsic(DomainArgs d, Ws w, int lane) {
    const DevMech& m = mech();
    const int Ns = m.Ns, Nr = m.Nr;
    const double* ns = w.ybuf();
    double T, P, V;
    if (d.kind == RMSB_DOMAIN_CONSTANT_V) { T = ns[Ns]; P = ns[Ns + 1]; V = d.cond[0]; }
    else if (d.kind == RMSB_DOMAIN_CONSTANT_P) { T = ns[Ns]; P = d.cond[0]; V = ns[Ns + 1]; }
    else if (d.kind == RMSB_DOMAIN_CONSTANT_TP) { T = d.cond[0]; P = d.cond[1]; V = ns[Ns]; }
    else { T = d.cond[0]; P = (d.kind == RMSB_DOMAIN_CONSTANT_TV) ? 1.0e8 : 0.0; V = d.cond[1]; }
    double N = 0.0;
    for (int s = 0; s < Ns; ++s) N += ns[s];
    const double C = (d.kind == RMSB_DOMAIN_CONSTANT_TP) ? P / (RMSB_R * T) : (d.kind == RMSB_DOMAIN_CONSTANT_TAPHI ? 0.0 : N / V);
    KinEnv<double> e;
    e.T = T; e.logT = log(T); e.RT = RMSB_R * T; e.invRT = 1.0 / (RMSB_R * T); e.P = P;
    double* kfb = w.kf();
    for (int j = lane; j < Nr; j += 32) kfb[j] = eval_kin<double>(j, &e, ceff_of(j, C, ns, V), false);
    __syncwarp();
}

// ------------------------------------------------------------------------------------------------
// jacobianp! for one reactor (src/Domain.jl:3095-3485): d(dy/dt)/dp, Ny x (Ns+Nr).  Columns 0..Ns-1: dG_i (through
// krev = kf/Kc); column Ns+j: k_j (constant-k domains: the rate coefficient itself -- the multiplier for efficiency
// reactions; variable-k domains: the multiplier f_j).  One lane per column: a call produces the 32 columns
// c0 .. c0+31 into w.jp() laid out [row][lane] (JP_CHUNK doubles per row; lanes past Ns+Nr write zeros).
// Needs cs/kf/kr/us/aux of the state in w.ybuf: `fresh` evaluates them first (rhs_warp with store_k).
// ------------------------------------------------------------------------------------------------
constexpr int JP_CHUNK = 32;
static __device__ __noinline__ void jacp_warp(DomainArgs d, Ws w, int lane, int c0, bool fresh) {
    const DevMech& m = mech();
    const int Ns = m.Ns, Nr = m.Nr, Ny = w.Ny, Np = Ns + Nr;
    if (fresh) rhs_warp(d, w, lane, true);
    double* col = w.jp() + lane;
    const double* cs = w.cs();
    const double* kfb = w.kf();
    const double* krb = w.kr();
    const double* ax = w.aux();
    const double aT = ax[0], aP = ax[1], aV = ax[2], aN = ax[3], aCv = ax[5];
    const double RTinv = 1.0 / (RMSB_R * aT);
    // the Parametrized* domains are variable-k domains whatever their layout (ParametrizedTP on the ConstantTP layout,
    // ParametrizedTConstantV on the ConstantTV one): jacobianpnsderiv! of src/Domain.jl:3204-3233, p = [dG; f]
    const bool constk = (d.kind != RMSB_DOMAIN_CONSTANT_V && d.kind != RMSB_DOMAIN_CONSTANT_P && !d.ptab);
    const int c = c0 + lane;
    for (int i = 0; i < Ny; ++i) col[i * JP_CHUNK] = 0.0;
    if (c < Np) {
        double raw = 0.0;       // dn_c/dt / V before constant-species zeroing (G columns of ConstantV)
        if (c >= Ns) {
            // reaction column: flux = fderiv - rderiv scattered with the stoichiometry (Domain.jl:3140-3142)
            const int j = c - Ns;
            double pr = 1.0, pp = 1.0;
            for (int k = 0; k < 4; ++k) { int sp = G(m.rx, k * Nr + j); if (sp < 0) break; pr *= cs[sp]; }
            for (int k = 4; k < 8; ++k) { int sp = G(m.rx, k * Nr + j); if (sp < 0) break; pp *= cs[sp]; }
            const bool eff = G(m.eff_ptr, j + 1) > G(m.eff_ptr, j);
            double fderiv, rderiv;
            if (d.kind == RMSB_DOMAIN_CONSTANT_TV && !d.ptab) { // Domain.jl:3290-3317: chain rule through the diffusion limit
                double pk = d.p[Ns + j];
                fderiv = kfb[j] * kfb[j] / (pk * pk) * pr;
                rderiv = kfb[j] * krb[j] / (pk * pk) * pp;
            } else if (constk && !eff) { fderiv = pr; rderiv = krb[j] / kfb[j] * pp; }
            else { fderiv = kfb[j] * pr; rderiv = krb[j] * pp; }
            const double flux = (fderiv - rderiv) * aV;
            for (int k = 0; k < 4; ++k) { int sp = G(m.rx, k * Nr + j); if (sp < 0) break; col[sp * JP_CHUNK] -= flux; }
            for (int k = 4; k < 8; ++k) { int sp = G(m.rx, k * Nr + j); if (sp < 0) break; col[sp * JP_CHUNK] += flux; }
        } else {
            // dG column of species c: every slot occurrence of the species contributes -/+ gderiv_j (Domain.jl:3144-3199)
            for (int q = G(m.sr_ptr, c); q < G(m.sr_ptr, c + 1); ++q) {
                const int jr = G(m.sr_rxn, q);
                const int j = jr & 0x7fffffff;
                int sl[8];                                   // the reaction's slots, read once for the products and the scatter below
#pragma unroll
                for (int k = 0; k < 8; ++k) sl[k] = G(m.rx, k * Nr + j);
                double pr = 1.0, pp = 1.0;
#pragma unroll
                for (int k = 0; k < 4; ++k) { if (sl[k] < 0) break; pr *= cs[sl[k]]; }
#pragma unroll
                for (int k = 4; k < 8; ++k) { if (sl[k] < 0) break; pp *= cs[sl[k]]; }
                const double net = kfb[j] * pr - krb[j] * pp;
                raw = (jr < 0) ? raw - net : raw + net;
                // gderiv = d(reverse flux)/dG = krev*prod(c_products)/RT (the reference reaches it through rderiv*kfs*RTinv in
                // the constant-k domains, whose rderiv carries 1/kfs -- except for efficiency reactions, Domain.jl:3144)
                const bool eff = G(m.eff_ptr, j + 1) > G(m.eff_ptr, j);
                double gderiv = krb[j] * pp * RTinv;
                if (constk && eff && d.kind != RMSB_DOMAIN_CONSTANT_TV) gderiv *= kfb[j];
                gderiv *= aV;
                const double sg = (jr < 0) ? gderiv : -gderiv;      // reactant slot: reactant rows -= g, product rows += g
#pragma unroll
                for (int k = 0; k < 4; ++k) { if (sl[k] < 0) break; col[sl[k] * JP_CHUNK] -= sg; }
#pragma unroll
                for (int k = 4; k < 8; ++k) { if (sl[k] < 0) break; col[sl[k] * JP_CHUNK] += sg; }
            }
        }
        // thermodynamic-variable rows
        if (d.kind == RMSB_DOMAIN_CONSTANT_TP) {                 // Domain.jl:3386-3388
            double a = 0.0;
            for (int i = 0; i < Ns; ++i) a += col[i * JP_CHUNK];
            col[Ns * JP_CHUNK] = a * RMSB_R * aT / aP;
        } else if (d.kind == RMSB_DOMAIN_CONSTANT_V) {           // Domain.jl:3406-3414
            const double* us = w.us();
            double a = 0.0, sm = 0.0;
            for (int i = 0; i < Ns; ++i) { a += us[i] * col[i * JP_CHUNK]; sm += col[i * JP_CHUNK]; }
            double jt = (c < Ns) ? -(a + raw * aV) / (aN * aCv) : -a / (aN * aCv);
            col[Ns * JP_CHUNK] = jt;
            col[(Ns + 1) * JP_CHUNK] = sm * RMSB_R * aT / aV + aP / aT * jt;
        }
        if (d.kind == RMSB_DOMAIN_CONSTANT_P) {                  // Domain.jl:3447-3456 (us holds the enthalpies)
            const double* hs = w.us();
            const double aCp = aCv + RMSB_R;
            double a = 0.0, sm = 0.0;
            for (int i = 0; i < Ns; ++i) { a += hs[i] * col[i * JP_CHUNK]; sm += col[i * JP_CHUNK]; }
            double jt = (c < Ns) ? -(a + raw * aV) / (aN * aCp) : -a / (aN * aCp);
            col[Ns * JP_CHUNK] = jt;
            col[(Ns + 1) * JP_CHUNK] = sm * RMSB_R * aT / aP + jt * aV / aT;
        }
        for (int k = 0; k < d.nconst; ++k) col[(d.constspc[k] - 1) * JP_CHUNK] = 0.0;
        // boundary interfaces: the only parameter-Jacobian term the reference carries is the Inlet's on the T and P (V) rows of the
        // dG columns of the adiabatic domains (src/Domain.jl:3421-3430,3461-3470, as written there); ConstantTP / TV / TAPhi: none
        if (d.flow && c < Ns && (d.kind == RMSB_DOMAIN_CONSTANT_V || d.kind == RMSB_DOMAIN_CONSTANT_P)) {
            const bool isp = d.kind == RMSB_DOMAIN_CONSTANT_P;
            const double ddG = d.flow[IF_FIN] * (-w.ybuf()[c] / aN) / (aN * (isp ? aCv + RMSB_R : aCv));
            col[Ns * JP_CHUNK] += ddG;
            col[(Ns + 1) * JP_CHUNK] += isp ? ddG * aV / aT : aP / aT * ddG;
        }
    }
    __syncwarp();
}

}  // namespace rmsb
