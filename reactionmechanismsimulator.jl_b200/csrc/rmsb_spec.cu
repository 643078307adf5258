// rmsb_spec.cu -- mechanism-specialised evaluation kernels, generated and compiled at run time (NVRTC).
//
// The generic thread-per-reactor kernels (dydt_tpr_kernel / jac_tpr_kernel, rmsb_api.cu) spend four instructions in five on
// walking the mechanism tables (index loads, loop control, local-memory arrays indexed by table entries): ncu counts 7 690
// instructions per state of which 1 370 are FP64.  For a small all-Arrhenius gas mechanism this unit writes the same arithmetic
// out as straight-line CUDA C -- every NASA coefficient, Arrhenius triplet, stoichiometric index and Jacobian gather list a
// literal, every per-species / per-reaction value a scalar the compiler keeps in registers -- compiles it for sm_100a with NVRTC
// the first time a large batch asks for it, and launches that kernel instead.  The operations and their order are those of
// rhs_thread / jac_tpr_kernel (which follow src/Domain.jl:1064-1113,1141-1217, src/Reactor.jl:465-505,899-1413,1469-1477 and
// src/Domain.jl:2296-2316,2466-2490), so results agree with the generic kernels to rounding of FMA contraction.
//
// The Jacobian kernel builds J two columns at a time in registers, stages them in a shared-memory tile and lets each warp write
// its 32 reactors' column pairs as contiguous 2*Ny-double runs (the generic kernel keeps a dense Ny x Ny block per thread in local
// memory: 6 KB frame, 6.6x DRAM traffic).
//
// Nothing here is a fallback for a missing GPU: without NVRTC (or RMSB_NO_SPEC=1) the generic CUDA kernels run.
#include <dlfcn.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <memory>
#include <mutex>
#include <sstream>
#include <string>
#include <vector>

#include "rmsb_internal.h"
#include "rmsb_solver.cuh"

using namespace rmsb;

namespace {

constexpr int SPEC_THREADS = 128;
int spec_max(const char* name, int dflt) { const char* v = std::getenv(name); int b = v ? std::atoi(v) : dflt; return b > 0 ? b : dflt; }

struct SpecKernels {
    int state = 0;                 // 0 not tried, 1 ready, -1 failed
    std::string why;
    cudaLibrary_t lib = nullptr;
    cudaKernel_t dydt = nullptr, jac = nullptr;
    int ldt = 0, ldj = 0;          // shared-memory row strides (doubles) of the two kernels
    int occ_dydt = 1, occ_jac = 1; // resident blocks per SM
};

}  // namespace

struct rmsb_spec_state {
    int Ns = 0, Nr = 0, maxr = 0, n_dt = 0, n_je = 0;
    std::vector<int> nasa_nr, nasa_nc;
    std::vector<double> tmax, coef;                     // [range][Ns], [range*9 + c][Ns]
    std::vector<int> rx, koff, kcptr, kcspc, srptr, srrxn;
    std::vector<double> kpar, fwd, rev, kccoef, nrp;
    std::vector<int> dtrxn, dtmult, dtoth, jeptr, jerow, jecol, jtidx;
    std::mutex mu;
    SpecKernels k[2];                                    // [0] ConstantV, [1] ConstantP
};

namespace {

// Doubles are written as 17-digit literals.  (Collecting them in a __constant__ array instead -- K[i] operands, LDCU loads into
// uniform registers -- was measured and lost: dydt -3 %, Jacobian 2.2x slower through uniform-register pressure and spills.)
std::string lit(double v) {
    if (v != v) return "(0.0/0.0)";
    if (std::isinf(v)) return v > 0 ? "(1.0/0.0)" : "(-1.0/0.0)";
    char b[48];
    snprintf(b, sizeof b, "%.17g", v);
    std::string s(b);
    if (s.find_first_of(".eE") == std::string::npos) s += ".0";
    if (v < 0 || (v == 0 && std::signbit(v))) s = "(" + s + ")";
    return s;
}
std::string idx(const char* a, int i) { return std::string(a) + "[" + std::to_string(i) + "]"; }

// Dual numbers for the T and P columns (the in-kernel analogue of the reference's ForwardDiff pass, src/Reactor.jl:1469-1477).
// D1 carries d/dT only: everything thermo and kinetics compute depends on T alone, so the d/dP parts the generic kernels carry
// through those expressions are sums of exact zeros.  D2 (d/dT, d/dP) is used for the few quantities that see P: the total moles
// P V / (R T), dT/dt and dP/dt (dV/dt).  A D1 converts to a D2 with a zero P part.
const char* PRELUDE = R"SRC(
struct D1 {
    double v, a;
    __device__ __forceinline__ D1() {}
    __device__ __forceinline__ D1(double v_) : v(v_), a(0.0) {}
    __device__ __forceinline__ D1(double v_, double a_) : v(v_), a(a_) {}
};
__device__ __forceinline__ D1 operator+(D1 x, D1 y) { return D1(x.v + y.v, x.a + y.a); }
__device__ __forceinline__ D1 operator+(D1 x, double y) { return D1(x.v + y, x.a); }
__device__ __forceinline__ D1 operator+(double x, D1 y) { return D1(x + y.v, y.a); }
__device__ __forceinline__ D1 operator-(D1 x, D1 y) { return D1(x.v - y.v, x.a - y.a); }
__device__ __forceinline__ D1 operator-(D1 x, double y) { return D1(x.v - y, x.a); }
__device__ __forceinline__ D1 operator-(double x, D1 y) { return D1(x - y.v, -y.a); }
__device__ __forceinline__ D1 operator-(D1 x) { return D1(-x.v, -x.a); }
__device__ __forceinline__ D1 operator*(D1 x, D1 y) { return D1(x.v * y.v, x.a * y.v + x.v * y.a); }
__device__ __forceinline__ D1 operator*(D1 x, double y) { return D1(x.v * y, x.a * y); }
__device__ __forceinline__ D1 operator*(double x, D1 y) { return D1(x * y.v, x * y.a); }
__device__ __forceinline__ D1 operator/(D1 x, D1 y) { double q = x.v / y.v, iy = 1.0 / y.v; return D1(q, (x.a - q * y.a) * iy); }
__device__ __forceinline__ D1 operator/(D1 x, double y) { return D1(x.v / y, x.a / y); }
__device__ __forceinline__ D1 operator/(double x, D1 y) { double q = x / y.v, s = -q / y.v; return D1(q, s * y.a); }
struct D2 {
    double v, a, b;
    __device__ __forceinline__ D2() {}
    __device__ __forceinline__ D2(double v_) : v(v_), a(0.0), b(0.0) {}
    __device__ __forceinline__ D2(D1 x) : v(x.v), a(x.a), b(0.0) {}
    __device__ __forceinline__ D2(double v_, double a_, double b_) : v(v_), a(a_), b(b_) {}
};
__device__ __forceinline__ D2 operator+(D2 x, D2 y) { return D2(x.v + y.v, x.a + y.a, x.b + y.b); }
__device__ __forceinline__ D2 operator+(D2 x, double y) { return D2(x.v + y, x.a, x.b); }
__device__ __forceinline__ D2 operator+(double x, D2 y) { return D2(x + y.v, y.a, y.b); }
__device__ __forceinline__ D2 operator-(D2 x, D2 y) { return D2(x.v - y.v, x.a - y.a, x.b - y.b); }
__device__ __forceinline__ D2 operator-(D2 x, double y) { return D2(x.v - y, x.a, x.b); }
__device__ __forceinline__ D2 operator-(double x, D2 y) { return D2(x - y.v, -y.a, -y.b); }
__device__ __forceinline__ D2 operator-(D2 x) { return D2(-x.v, -x.a, -x.b); }
__device__ __forceinline__ D2 operator*(D2 x, D2 y) { return D2(x.v * y.v, x.a * y.v + x.v * y.a, x.b * y.v + x.v * y.b); }
__device__ __forceinline__ D2 operator*(D2 x, double y) { return D2(x.v * y, x.a * y, x.b * y); }
__device__ __forceinline__ D2 operator*(double x, D2 y) { return D2(x * y.v, x * y.a, x * y.b); }
__device__ __forceinline__ D2 operator/(D2 x, D2 y) {
    double q = x.v / y.v, iy = 1.0 / y.v;
    return D2(q, (x.a - q * y.a) * iy, (x.b - q * y.b) * iy);
}
__device__ __forceinline__ D2 operator/(D2 x, double y) { return D2(x.v / y, x.a / y, x.b / y); }
__device__ __forceinline__ D2 operator/(double x, D2 y) { double q = x / y.v, s = -q / y.v; return D2(q, s * y.a, s * y.b); }
// mixed D1 / D2 expressions promote the D1 operand
__device__ __forceinline__ D2 operator+(D2 x, D1 y) { return x + D2(y); }
__device__ __forceinline__ D2 operator+(D1 x, D2 y) { return D2(x) + y; }
__device__ __forceinline__ D2 operator-(D2 x, D1 y) { return x - D2(y); }
__device__ __forceinline__ D2 operator-(D1 x, D2 y) { return D2(x) - y; }
__device__ __forceinline__ D2 operator*(D2 x, D1 y) { return x * D2(y); }
__device__ __forceinline__ D2 operator*(D1 x, D2 y) { return D2(x) * y; }
__device__ __forceinline__ D2 operator/(D2 x, D1 y) { return x / D2(y); }
__device__ __forceinline__ D2 operator/(D1 x, D2 y) { return D2(x) / y; }
__device__ __forceinline__ D1 dexp(D1 x) { double e = exp(x.v); return D1(e, e * x.a); }
__device__ __forceinline__ double dexp(double x) { return exp(x); }
__device__ __forceinline__ D1 dlog(D1 x) { double i = 1.0 / x.v; return D1(log(x.v), x.a * i); }
__device__ __forceinline__ double dlog(double x) { return log(x); }
__device__ __forceinline__ double val(double x) { return x; }
__device__ __forceinline__ double val(D1 x) { return x.v; }
__device__ __forceinline__ double val(D2 x) { return x.v; }
template <class N> __device__ __forceinline__ N seedT(double v);
template <> __device__ __forceinline__ double seedT<double>(double v) { return v; }
template <> __device__ __forceinline__ D1 seedT<D1>(double v) { return D1(v, 1.0); }
template <class M> __device__ __forceinline__ M seedP(double v);
template <> __device__ __forceinline__ double seedP<double>(double v) { return v; }
template <> __device__ __forceinline__ D2 seedP<D2>(double v) { return D2(v, 0.0, 1.0); }
)SRC";

// ---- NASA polynomial of species s, range r, in the operation order of nasa_hscp (rmsb_device.cuh) ---------------------------
void emit_nasa(std::ostringstream& o, const rmsb_spec_state& m, int s, int r, const char* ind) {
    const int Ns = m.Ns;
    auto c = [&](int k) { return lit(m.coef[((size_t)r * 9 + k) * Ns + s]); };
    const std::string S = std::to_string(s);
    o << ind << "{ ";
    if (m.nasa_nc[(size_t)r * Ns + s] != 7) {
        o << "N ct0 = " << c(0) << " * T5, ct1 = " << c(1) << " * T4; const double ct2 = " << c(2) << "; N ct3 = " << c(3) << " * T0, ct4 = "
          << c(4) << " * T1, ct5 = " << c(5) << " * T2, ct6 = " << c(6) << " * T3;\n" << ind
          << "  cp[" << S << "] = ct0 + ct1 + ct2 + ct3 + ct4 + ct5 + ct6;\n" << ind
          << "  hh[" << S << "] = -ct0 + T6 * ct1 + ct2 + 0.5 * ct3 + 0.33333333333 * ct4 + 0.25 * ct5 + 0.2 * ct6 + " << c(7) << " * T4;\n" << ind
          << "  sr[" << S << "] = -0.5 * ct0 - ct1 + T6 * ct2 + ct3 + 0.5 * ct4 + 0.33333333333 * ct5 + 0.25 * ct6 + " << c(8) << "; }\n";
    } else {
        o << "const double ct2 = " << c(0) << "; N ct3 = " << c(1) << " * T0, ct4 = " << c(2) << " * T1, ct5 = " << c(3) << " * T2, ct6 = "
          << c(4) << " * T3;\n" << ind
          << "  cp[" << S << "] = ct2 + ct3 + ct4 + ct5 + ct6;\n" << ind
          << "  hh[" << S << "] = ct2 + 0.5 * ct3 + 0.33333333333 * ct4 + 0.25 * ct5 + 0.2 * ct6 + " << c(5) << " * T4;\n" << ind
          << "  sr[" << S << "] = T6 * ct2 + ct3 + 0.5 * ct4 + 0.33333333333 * ct5 + 0.25 * ct6 + " << c(6) << "; }\n";
    }
}

// product of concentrations of one side of reaction j, in slot order, starting from `k`
std::string side_product(const rmsb_spec_state& m, int j, bool products, const std::string& k, const char* cs) {
    std::string e = k;
    for (int q = 0; q < 4; ++q) {
        int s = m.rx[(size_t)(products ? q + 4 : q) * m.Nr + j];
        if (s < 0) break;
        e += " * " + idx(cs, s);
    }
    return e;
}
int side_count(const rmsb_spec_state& m, int j, bool products) {
    int n = 0;
    for (int q = 0; q < 4; ++q) if (m.rx[(size_t)(products ? q + 4 : q) * m.Nr + j] >= 0) ++n;
    return n;
}

// ---- dy/dt of one reactor: rhs_thread (rmsb_api.cu) written out for this mechanism -------------------------------------------
void emit_rhs(std::ostringstream& o, const rmsb_spec_state& m, bool isp) {
    const int Ns = m.Ns, Nr = m.Nr;
    o << "struct Keep { double kf[" << Nr << "], kr[" << Nr << "], us[" << Ns << "], cp[" << Ns << "], cs[" << Ns << "]; double T, P, V, Nt, C, Cv, dotU; };\n";
    // N: type of everything that depends on T only; M: type of what also sees P (value pass: double, double)
    o << "template <class N, class M, bool KEEP>\n__device__ __forceinline__ void spec_rhs(const double* y, const double cond, const double* __restrict__ p, N* f, M& fT, M& fP, Keep& kp) {\n";
    o << "    const double V = " << (isp ? "y[" + std::to_string(Ns + 1) + "]" : std::string("cond")) << ";\n";
    o << "    const double invV = 1.0 / V;\n";
    o << "    N T = seedT<N>(y[" << Ns << "]);\n    M P = " << (isp ? std::string("M(cond)") : "seedP<M>(y[" + std::to_string(Ns + 1) + "])") << ";\n";
    o << "    const double Tv = val(T);\n";
    o << "    N T0 = T, T1 = T * T, T2 = T1 * T, T3 = T2 * T, T4 = 1.0 / T, T5 = T4 * T4, T6 = dlog(T);\n";
    o << "    N RT = 8.314 * T;\n    N invRT = T4 * (1.0 / 8.314);\n    M Nt = P * V * invRT, C = P * invRT;\n";
    o << "    N cp[" << Ns << "], hh[" << Ns << "], sr[" << Ns << "];\n";
    // species grouped by their range boundaries: one branch chain per group (selectPoly: first range with T <= Tmax, else the last)
    std::map<std::vector<double>, std::vector<int>> groups;
    for (int s = 0; s < Ns; ++s) {
        std::vector<double> key;
        for (int r = 0; r + 1 < m.nasa_nr[s]; ++r) key.push_back(m.tmax[(size_t)r * Ns + s]);
        groups[key].push_back(s);
    }
    for (auto& kv : groups) {
        const auto& key = kv.first;
        if (key.empty()) { for (int s : kv.second) emit_nasa(o, m, s, 0, "    "); continue; }
        for (size_t r = 0; r <= key.size(); ++r) {
            if (r == 0) o << "    if (Tv <= " << lit(key[0]) << ") {\n";
            else if (r < key.size()) o << "    } else if (Tv <= " << lit(key[r]) << ") {\n";
            else o << "    } else {\n";
            for (int s : kv.second) emit_nasa(o, m, s, (int)r, "        ");
        }
        o << "    }\n";
    }
    o << "    N g[" << Ns << "], u[" << Ns << "]; double cs[" << Ns << "];\n    N cpsum = 0.0;\n";
    o << "    if (p) {";
    for (int s = 0; s < Ns; ++s) o << " hh[" << s << "] = hh[" << s << "] + p[" << s << "] * invRT;";
    o << " }\n";
    for (int s = 0; s < Ns; ++s) {
        o << "    u[" << s << "] = " << (isp ? "hh[" + std::to_string(s) + "] * RT" : "(hh[" + std::to_string(s) + "] - 1.0) * RT") << "; g[" << s << "] = hh[" << s
          << "] - sr[" << s << "]; cs[" << s << "] = y[" << s << "] * invV; cpsum = cpsum + cp[" << s << "] * y[" << s << "];\n";
    }
    o << "    if (KEEP) {";
    for (int s = 0; s < Ns; ++s) o << " kp.us[" << s << "] = val(u[" << s << "]); kp.cp[" << s << "] = val(cp[" << s << "]); kp.cs[" << s << "] = cs[" << s << "];";
    o << " }\n";
    o << "    N logC0 = 9.394984624139783 - T6;\n";
    o << "    N q[" << Nr << "];\n";
    for (int j = 0; j < Nr; ++j) {
        const double* a = m.kpar.data() + m.koff[j];
        const bool hasf = m.fwd[j] != 0.0, hasr = m.rev[j] != 0.0;
        o << "    {   // reaction " << j << "\n";
        std::string arg;
        if (a[1] != 0.0) arg = lit(a[1]) + " * T6";
        if (a[2] != 0.0) arg += (arg.empty() ? "-(" + lit(a[2]) + " * invRT)" : " - " + lit(a[2]) + " * invRT");
        if (arg.empty()) o << "        N kf = " << lit(a[0]) << ";\n";
        else o << "        N kf = " << lit(a[0]) << " * dexp(" << arg << ");\n";
        if (hasr) {
            std::string ex;
            for (int k = m.kcptr[j]; k < m.kcptr[j + 1]; ++k) {
                const double cf = m.kccoef[k];
                const std::string gs = idx("g", m.kcspc[k]);
                std::string term = (std::fabs(cf) == 1.0) ? gs : lit(std::fabs(cf)) + " * " + gs;
                if (ex.empty()) ex = (cf < 0 ? "-" : "") + term;     // a -(c*g) first term equals (-c)*g exactly
                else ex += (cf < 0 ? " - " : " + ") + term;
            }
            if (ex.empty()) ex = "N(0.0)";
            if (m.nrp[j] != 0.0) ex += " + " + lit(m.nrp[j]) + " * logC0";
            o << "        N ex = " << ex << ";\n        N kr = kf * dexp(-ex);\n";
        } else o << "        N kr = 0.0;\n";
        if (!hasf) o << "        kf = N(0.0);\n";
        o << "        if (p) { kf = kf * p[" << Ns + j << "]; kr = kr * p[" << Ns + j << "]; }\n";
        o << "        if (KEEP) { kp.kf[" << j << "] = val(kf); kp.kr[" << j << "] = val(kr); }\n";
        o << "        q[" << j << "] = " << side_product(m, j, false, "kf", "cs") << " - " << side_product(m, j, true, "kr", "cs") << ";\n    }\n";
    }
    o << "    N usum = 0.0, nsum = 0.0;\n";
    for (int s = 0; s < Ns; ++s) {
        std::string acc;
        for (int k = m.srptr[s]; k < m.srptr[s + 1]; ++k) {
            const int jr = m.srrxn[k];
            const std::string qv = idx("q", jr & 0x7fffffff);
            if (acc.empty()) acc = (jr < 0 ? "-" : "") + qv;
            else acc += (jr < 0 ? " - " : " + ") + qv;
        }
        if (acc.empty()) acc = "N(0.0)";
        o << "    { N dn = (" << acc << ") * V; f[" << s << "] = dn; usum = usum + u[" << s << "] * dn; nsum = nsum + dn; }\n";
    }
    if (isp) {
        o << "    M dT = -usum / (8.314 * cpsum);\n    fT = dT;\n    fP = nsum * RT / P + dT * V * T4;\n";
    } else {
        o << "    M dT = -usum / (8.314 * (cpsum - Nt));\n    fT = dT;\n    fP = nsum * RT * invV + P * T4 * dT;\n";
    }
    o << "    if (KEEP) { kp.T = val(T); kp.P = val(P); kp.V = V; kp.Nt = val(Nt); kp.C = val(C); kp.Cv = val(cpsum) * (8.314 / val(Nt)) - 8.314; kp.dotU = val(usum); }\n";
    o << "}\n\n";
}

// resident blocks per SM the kernels are compiled for (register budget = 65536 / (128 * blocks)); RMSB_SPEC_DYDT_BLOCKS /
// RMSB_SPEC_JAC_BLOCKS override the defaults for tuning runs
int env_blocks(const char* name, int dflt) {
    const char* v = std::getenv(name);
    int b = v ? std::atoi(v) : dflt;
    return (b >= 0 && b <= 32) ? b : dflt;
}

void emit_dydt_kernel(std::ostringstream& o, const rmsb_spec_state& m, int ldt) {
    const int Ny = m.Ns + 2;
    // every warp stages its own 32 states (coalesced global accesses, no block-wide barrier: the warps of a block drift apart and
    // overlap each other's memory phases)
    o << "extern \"C\" __global__ void __launch_bounds__(" << SPEC_THREADS << ", " << env_blocks("RMSB_SPEC_DYDT_BLOCKS", 3) << ") rmsb_spec_dydt(const double* __restrict__ Y, const double* __restrict__ COND, int ncond,\n"
         "        const double* __restrict__ PP, long long p_stride, const int* __restrict__ constspc, int nconst, long long n, double* __restrict__ OUT) {\n"
         "    extern __shared__ double tile[];\n    const int lane = threadIdx.x & 31;\n    double* wt = tile + (threadIdx.x >> 5) * (32 * " << ldt << ");\n"
         "    const long long nwarps = (long long)gridDim.x * (blockDim.x >> 5);\n"
         "    for (long long base = ((long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * 32; base < n; base += nwarps * 32) {\n"
         "        const int cnt = (int)((n - base < 32) ? n - base : 32);\n"
         "        for (int i = lane; i < cnt * " << Ny << "; i += 32) { int r = i / " << Ny << "; wt[r * " << ldt << " + (i - r * " << Ny << ")] = Y[base * " << Ny << " + i]; }\n"
         "        __syncwarp();\n"
         "        if (lane < cnt) {\n            const long long ir = base + lane;\n            double* row = wt + lane * " << ldt << ";\n"
         "            double y[" << Ny << "], f[" << Ny << "];\n";
    for (int i = 0; i < Ny; ++i) o << "            y[" << i << "] = row[" << i << "];\n";
    o << "            Keep kp;\n            spec_rhs<double, double, false>(y, COND[ir * ncond], PP ? PP + ir * p_stride : nullptr, f, f[" << Ny - 2 << "], f[" << Ny - 1 << "], kp);\n";
    for (int i = 0; i < Ny; ++i) o << "            row[" << i << "] = f[" << i << "];\n";
    o << "            for (int k = 0; k < nconst; ++k) row[constspc[k] - 1] = 0.0;\n        }\n        __syncwarp();\n"
         "        for (int i = lane; i < cnt * " << Ny << "; i += 32) { int r = i / " << Ny << "; OUT[base * " << Ny << " + i] = wt[r * " << ldt << " + (i - r * " << Ny << ")]; }\n"
         "        __syncwarp();\n    }\n}\n\n";
}

// ---- jacobiany!: jac_tpr_kernel (rmsb_api.cu) written out; columns are produced in pairs and leave through a shared-memory tile --
// Order of work (chosen for short live ranges, the results do not depend on it): the dual-number pass over the complete RHS first
// (T and P columns, src/Reactor.jl:1469-1477) -- for ConstantV those two columns leave at once --, then the value pass, then the
// species columns with their derivative terms formed column by column (a term belongs to the column of the species it differentiates).
void emit_jac_kernel(std::ostringstream& o, const rmsb_spec_state& m, bool isp, int ldj, int chunk) {
    const int Ns = m.Ns, Nr = m.Nr, Ny = Ns + 2;
    o << "extern \"C\" __global__ void __launch_bounds__(" << SPEC_THREADS << ", " << env_blocks("RMSB_SPEC_JAC_BLOCKS", 2) << ") rmsb_spec_jac(const double* __restrict__ Y, const double* __restrict__ COND, int ncond,\n"
         "        const double* __restrict__ PP, long long p_stride, const int* __restrict__ constspc, int nconst, long long n, double* __restrict__ OUT) {\n"
         "    extern __shared__ double tile[];\n    const int tid = threadIdx.x, lane = tid & 31, wbase = tid & ~31;\n"
         "    for (long long base = (long long)blockIdx.x * " << SPEC_THREADS << "; base < n; base += (long long)gridDim.x * " << SPEC_THREADS << ") {\n"
         "        const int cnt = (int)((n - base < " << SPEC_THREADS << ") ? n - base : " << SPEC_THREADS << ");\n"
         "        for (int i = tid; i < cnt * " << Ny << "; i += " << SPEC_THREADS << ") { int r = i / " << Ny << "; tile[r * " << ldj << " + (i - r * " << Ny << ")] = Y[base * " << Ny << " + i]; }\n"
         "        __syncthreads();\n"
         "        // threads past the end of the batch recompute the last reactor (their rows are never written out)\n"
         "        const int me = tid < cnt ? tid : cnt - 1;\n        const long long ir = base + me;\n        double* row = tile + tid * " << ldj << ";\n"
         "        double y[" << Ny << "], f[" << Ny << "];\n";
    for (int i = 0; i < Ny; ++i) o << "        y[" << i << "] = tile[me * " << ldj << " + " << i << "];\n";
    o << "        __syncthreads();\n";
    o << "        const double* p = PP ? PP + ir * p_stride : nullptr;\n        const double cond = COND[ir * ncond];\n        Keep kp;\n";
    // (6) constant species rows, then the warp writes its 32 reactors' runs of `ncols` columns starting at column c0
    auto emit_write = [&](int c0, int ncols) {
        const int w = ncols * Ny;
        o << "            for (int k = 0; k < nconst; ++k) { const int s = constspc[k] - 1; for (int cc = 0; cc < " << ncols << "; ++cc) row[cc * " << Ny << " + s] = 0.0; }\n"
             "            __syncwarp();\n";
        if (env_blocks("RMSB_SPEC_WRITER", 4) == 0) {
            o << "            for (int r = 0; r < 32 && wbase + r < cnt; ++r) {\n"
                 "                double* dst = OUT + (base + wbase + r) * " << Ny * Ny << " + " << c0 * Ny << ";\n"
                 "                const double* src = tile + (wbase + r) * " << ldj << ";\n"
                 "                for (int i = lane; i < " << w << "; i += 32) dst[i] = src[i];\n"
                 "            }\n";
        } else {
            o << "            {\n"
                 "                double* dst = OUT + (base + wbase) * " << Ny * Ny << " + " << c0 * Ny << " + lane;\n"
                 "                const double* src = tile + wbase * " << ldj << " + lane;\n"
                 "                const int nr = (cnt - wbase < 32) ? cnt - wbase : 32;\n"
                 "#pragma unroll " << env_blocks("RMSB_SPEC_WRITER", 4) << "\n"
                 "                for (int r = 0; r < nr; ++r)\n"
                 "                    for (int i = 0; i < " << w << "; i += 32) if (lane + i < " << w << ") dst[r * " << Ny * Ny << " + i] = src[r * " << ldj << " + i];\n"
                 "            }\n";
        }
        o << "            __syncwarp();\n";
    };
    // (7) T (and P) columns: derivative of the complete RHS
    if (!isp) {
        o << "        {\n            D1 fd[" << Ns << "]; D2 fT, fP;\n            spec_rhs<D1, D2, false>(y, cond, p, fd, fT, fP, kp);\n";
        for (int i = 0; i < Ns; ++i) o << "            row[" << i << "] = fd[" << i << "].a; row[" << Ny + i << "] = 0.0;\n";
        o << "            row[" << Ns << "] = fT.a; row[" << Ns + 1 << "] = fP.a; row[" << Ny + Ns << "] = fT.b; row[" << Ny + Ns + 1 << "] = fP.b;\n";
        if (chunk >= 2) emit_write(Ns, 2);
        else {
            emit_write(Ns, 1);
            for (int i = 0; i < Ns; ++i) o << "            row[" << i << "] = 0.0;\n";
            o << "            row[" << Ns << "] = fT.b; row[" << Ns + 1 << "] = fP.b;\n";
            emit_write(Ns + 1, 1);
        }
        o << "        }\n";
    } else {
        o << "        double tcol[" << Ny << "];      // the T column waits for the V column next to it\n"
             "        {\n            D1 fd[" << Ns << "], fT, fP;\n            spec_rhs<D1, D1, false>(y, cond, p, fd, fT, fP, kp);\n"
             "            for (int i = 0; i < " << Ns << "; ++i) tcol[i] = fd[i].a;\n            tcol[" << Ns << "] = fT.a; tcol[" << Ns + 1 << "] = fP.a;\n        }\n";
    }
    o << "        spec_rhs<double, double, true>(y, cond, p, f, f[" << Ns << "], f[" << Ns + 1 << "], kp);\n";
    if (isp) {
        // per-reaction pieces of the V column (src/Reactor.jl:1368-1413)
        o << "        double dfv[" << Nr << "], drv[" << Nr << "];\n";
        for (int j = 0; j < Nr; ++j) {
            const int nre = side_count(m, j, false), npr = side_count(m, j, true);
            if (nre > 1) o << "        dfv[" << j << "] = -" << lit((double)(nre - 1)) << " * (" << side_product(m, j, false, idx("kp.kf", j), "kp.cs") << ");\n";
            if (npr > 1) o << "        drv[" << j << "] = -" << lit((double)(npr - 1)) << " * (" << side_product(m, j, true, idx("kp.kr", j), "kp.cs") << ");\n";
        }
        o << "        const double aCp = kp.Cv + 8.314, dTdt = -kp.dotU / (kp.Nt * aCp);\n";
    }
    // entries by column (the entry list is sorted by column, then row); derivative terms by the column they belong to
    std::vector<std::vector<int>> col_entries(Ns);
    for (int e = 0; e < m.n_je; ++e) col_entries[m.jecol[e]].push_back(e);
    // chunks: the species columns in groups of `chunk`; for ConstantP one more chunk {T column, V column}
    std::vector<std::pair<int, int>> chunks;
    for (int c0 = 0; c0 < Ns; c0 += chunk) chunks.push_back({c0, std::min(Ns, c0 + chunk)});
    if (isp) { if (chunk >= 2) chunks.push_back({Ns, Ns + 2}); else { chunks.push_back({Ns, Ns + 1}); chunks.push_back({Ns + 1, Ns + 2}); } }
    for (auto& ch : chunks) {
        const int c0 = ch.first, c1 = ch.second;
        o << "        {   // columns " << c0 << " .. " << c1 - 1 << "\n";
        for (int c = c0; c < c1; ++c) {
            const int off = (c - c0) * Ny;
            if (c == Ns) { for (int i = 0; i < Ny; ++i) o << "            row[" << off + i << "] = tcol[" << i << "];\n"; continue; }
            std::vector<std::string> nz(Ns);                 // expression name of the species rows of this column
            o << "            {\n";
            if (c < Ns) {
                // (1) the derivative terms this column's entries gather (src/Reactor.jl:899-1366)
                std::map<int, std::string> tname;
                for (int e : col_entries[c])
                    for (int k = m.jeptr[e]; k < m.jeptr[e + 1]; ++k) {
                        const int t = m.jtidx[k] & 0x7fffffff;
                        if (tname.count(t)) continue;
                        const int jr = m.dtrxn[t], j = jr & 0x3fffffff;
                        std::string ex = (jr >> 30) ? idx("kp.kr", j) : idx("kp.kf", j);
                        if (m.dtmult[t] != 1) ex = lit((double)m.dtmult[t]) + " * " + ex;
                        for (int q = 0; q < 3; ++q) { int sp = m.dtoth[3 * t + q]; if (sp < 0) break; ex += " * " + idx("kp.cs", sp); }
                        tname[t] = "d" + std::to_string(t);
                        o << "                const double d" << t << " = " << ex << ";\n";
                    }
                // (2) entries in the reference's accumulation order
                for (int e : col_entries[c]) {
                    std::string acc;
                    for (int k = m.jeptr[e]; k < m.jeptr[e + 1]; ++k) {
                        const int ti = m.jtidx[k];
                        const std::string v = tname[ti & 0x7fffffff];
                        if (acc.empty()) acc = (ti < 0 ? "-" : "") + v;
                        else acc += (ti < 0 ? " - " : " + ") + v;
                    }
                    const std::string nm = "e" + std::to_string(m.jerow[e]);
                    o << "                const double " << nm << " = " << (acc.empty() ? "0.0" : acc) << ";\n";
                    nz[m.jerow[e]] = nm;
                }
            } else {                                          // (3) V column of the species block (ConstantP)
                for (int s = 0; s < Ns; ++s) {
                    std::string acc;
                    for (int k = m.srptr[s]; k < m.srptr[s + 1]; ++k) {
                        const int jr = m.srrxn[k], j = jr & 0x7fffffff;
                        const bool reactant = jr < 0;
                        if (side_count(m, j, false) > 1) { acc += (acc.empty() ? (reactant ? "-" : "") : (reactant ? " - " : " + ")) + idx("dfv", j); }
                        if (side_count(m, j, true) > 1) { acc += (acc.empty() ? (reactant ? "" : "-") : (reactant ? " + " : " - ")) + idx("drv", j); }
                    }
                    if (acc.empty()) continue;
                    const std::string nm = "e" + std::to_string(s);
                    o << "                const double " << nm << " = " << acc << ";\n";
                    nz[s] = nm;
                }
            }
            // (5) thermodynamic-variable rows: weighted column sums in ascending row order (structural zeros add nothing);
            //     a + b + c parses as (a + b) + c, the accumulation order of the generic kernel
            std::string sa, sm;
            for (int i = 0; i < Ns; ++i) {
                if (nz[i].empty()) continue;
                sa += (sa.empty() ? "" : " + ") + idx("kp.us", i) + " * " + nz[i];
                sm += (sm.empty() ? "" : " + ") + nz[i];
            }
            if (sa.empty()) { sa = "0.0"; sm = "0.0"; }
            o << "                const double sa = " << sa << ";\n                const double sm = " << sm << ";\n";
            if (!isp) {
                o << "                const double dCv = kp.cp[" << c << "] * 8.314 / kp.Nt;\n"
                     "                const double jt = -sa / (kp.Nt * kp.Cv) + kp.dotU / (kp.Nt * kp.Cv * kp.Cv) * dCv;\n"
                     "                const double jp = sm * 8.314 * kp.T / kp.V + kp.P / kp.T * jt;\n";
            } else {
                o << "                double jt = -sa / (kp.Nt * aCp);\n";
                if (c < Ns) o << "                jt -= dTdt * ((kp.cp[" << c << "] * 8.314 / kp.Nt) / aCp);\n";
                o << "                double jp = sm * 8.314 * kp.T / kp.P + kp.V / kp.T * jt;\n";
                if (c == Ns + 1) o << "                jp += dTdt / kp.T;\n";
            }
            for (int i = 0; i < Ns; ++i) o << "                row[" << off + i << "] = " << (nz[i].empty() ? std::string("0.0") : nz[i]) << ";\n";
            o << "                row[" << off + Ns << "] = jt; row[" << off + Ns + 1 << "] = jp;\n            }\n";
        }
        emit_write(c0, c1 - c0);
        o << "        }\n";
    }
    o << "        __syncthreads();\n    }\n}\n";
}

std::string make_source(const rmsb_spec_state& m, bool isp, int* ldt, int* ldj) {
    const int Ny = m.Ns + 2;
    const int chunk = (2 * Ny <= 64) ? 2 : 1;
    *ldt = Ny | 1;
    *ldj = (chunk * Ny) | 1;
    std::ostringstream o;
    o << "// generated by librmsb200 (rmsb_spec.cu): " << m.Ns << " species, " << m.Nr << " reactions, " << (isp ? "ConstantP" : "ConstantV") << "\n";
    o << PRELUDE << "\n";
    emit_rhs(o, m, isp);
    emit_dydt_kernel(o, m, *ldt);
    emit_jac_kernel(o, m, isp, *ldj, chunk);
    return o.str();
}

// ---- NVRTC through dlopen: the library has no link-time dependency on it -----------------------------------------------------
struct Nvrtc {
    void* lib = nullptr;
    int (*create)(void**, const char*, const char*, int, const char* const*, const char* const*) = nullptr;
    int (*compile)(void*, int, const char* const*) = nullptr;
    int (*log_size)(void*, size_t*) = nullptr;
    int (*get_log)(void*, char*) = nullptr;
    int (*cubin_size)(void*, size_t*) = nullptr;
    int (*get_cubin)(void*, char*) = nullptr;
    int (*destroy)(void**) = nullptr;
    std::string why;
};
const Nvrtc& nvrtc() {
    static Nvrtc n = [] {
        Nvrtc r;
        std::vector<std::string> names;
        if (const char* e = std::getenv("RMSB_NVRTC")) names.push_back(e);
        for (const char* s : {"libnvrtc.so.12", "/usr/local/cuda/lib64/libnvrtc.so.12", "libnvrtc.so", "/usr/local/cuda/lib64/libnvrtc.so"}) names.push_back(s);
        for (auto& nm : names) { r.lib = dlopen(nm.c_str(), RTLD_NOW | RTLD_LOCAL); if (r.lib) break; }
        if (!r.lib) { r.why = "libnvrtc not found"; return r; }
#define SYM(field, name) *reinterpret_cast<void**>(&r.field) = dlsym(r.lib, name); if (!r.field) { r.why = std::string("libnvrtc lacks ") + name; r.lib = nullptr; return r; }
        SYM(create, "nvrtcCreateProgram"); SYM(compile, "nvrtcCompileProgram"); SYM(log_size, "nvrtcGetProgramLogSize"); SYM(get_log, "nvrtcGetProgramLog");
        SYM(cubin_size, "nvrtcGetCUBINSize"); SYM(get_cubin, "nvrtcGetCUBIN"); SYM(destroy, "nvrtcDestroyProgram");
#undef SYM
        return r;
    }();
    return n;
}

bool compile_cubin(const std::string& src, std::vector<char>* cubin, std::string* why) {
    const Nvrtc& n = nvrtc();
    if (!n.lib) { *why = n.why; return false; }
    void* prog = nullptr;
    if (n.create(&prog, src.c_str(), "rmsb_spec.cu", 0, nullptr, nullptr) != 0) { *why = "nvrtcCreateProgram failed"; return false; }
    const char* opts[] = {"--gpu-architecture=sm_100a", "--std=c++17", "-lineinfo"};
    const int rc = n.compile(prog, 3, opts);
    if (rc != 0) {
        size_t ls = 0;
        n.log_size(prog, &ls);
        std::string log(ls, '\0');
        if (ls) n.get_log(prog, &log[0]);
        *why = "nvrtcCompileProgram: " + log.substr(0, 2000);
        n.destroy(&prog);
        return false;
    }
    size_t cs = 0;
    n.cubin_size(prog, &cs);
    cubin->resize(cs);
    n.get_cubin(prog, cubin->data());
    n.destroy(&prog);
    return cs > 0;
}

// RMSB_NO_SPEC=1 in the environment keeps large batches on the generic kernels (A/B measurements, tests of the generic path);
// read on every call so that a test can flip it
bool spec_disabled() { const char* v = std::getenv("RMSB_NO_SPEC"); return v && v[0] == '1'; }

}  // namespace

// ---- interface to rmsb_api.cu -------------------------------------------------------------------------------------------------
int rmsb_spec_capture(rmsb_mech* h, const std::vector<int>& rx, const std::vector<int>& srptr, const std::vector<int>& srrxn, const std::vector<int>& kcptr,
                      const std::vector<int>& kcspc, const std::vector<double>& kccoef, const std::vector<double>& nrp, const std::vector<int>& ktype,
                      const std::vector<int>& koff, const std::vector<double>& kpar, const std::vector<double>& fwd, const std::vector<double>& rev,
                      const std::vector<int>& nasa_nr, const std::vector<int>& nasa_nc, const std::vector<double>& nasa_tmax, const std::vector<double>& nasa_coef,
                      int maxr, const std::vector<int>& dtrxn, const std::vector<int>& dtmult, const std::vector<int>& dtoth, const std::vector<int>& jeptr,
                      const std::vector<int>& jerow, const std::vector<int>& jecol, const std::vector<int>& jtidx) {
    const DevMech& dm = h->dm;
    h->spec.reset();
    if (dm.Ns > spec_max("RMSB_SPEC_MAXNS", 32) || dm.Nr > spec_max("RMSB_SPEC_MAXNR", 128) || dm.Nr < 1 || dm.coupled || dm.difflim || dm.phase_kind != RMSB_PHASE_GAS || dm.n_effrxn > 0) return RMSB_OK;
    for (int j = 0; j < dm.Nr; ++j) if (ktype[j] != RMSB_KIN_ARRHENIUS_VEC) return RMSB_OK;
    auto s = std::make_shared<rmsb_spec_state>();
    s->Ns = dm.Ns; s->Nr = dm.Nr; s->maxr = maxr; s->n_dt = (int)dtrxn.size(); s->n_je = (int)jerow.size();
    s->nasa_nr = nasa_nr; s->nasa_nc = nasa_nc; s->tmax = nasa_tmax; s->coef = nasa_coef;
    s->rx = rx; s->koff = koff; s->kcptr = kcptr; s->kcspc = kcspc; s->srptr = srptr; s->srrxn = srrxn;
    s->kpar = kpar; s->fwd = fwd; s->rev = rev; s->kccoef = kccoef; s->nrp = nrp;
    s->dtrxn = dtrxn; s->dtmult = dtmult; s->dtoth = dtoth; s->jeptr = jeptr; s->jerow = jerow; s->jecol = jecol; s->jtidx = jtidx;
    h->spec = s;
    return RMSB_OK;
}

int rmsb_spec_source(const rmsb_mech* h, int kind, std::string* out) {
    if (!h->spec) return fail(RMSB_ERR_UNSUPPORTED, "mechanism does not qualify for specialised kernels (small all-Arrhenius gas mechanisms only)");
    if (kind != RMSB_DOMAIN_CONSTANT_V && kind != RMSB_DOMAIN_CONSTANT_P) return fail(RMSB_ERR_UNSUPPORTED, "specialised kernels: ConstantV / ConstantP only");
    int a, b;
    *out = make_source(*h->spec, kind == RMSB_DOMAIN_CONSTANT_P, &a, &b);
    return RMSB_OK;
}

// 1 when the specialised kernels of (handle, domain) are loaded on the handle's device (compiling them on first use), else 0
int rmsb_spec_ready(const rmsb_mech* h, int kind, std::string* why) {
    if (!h->spec || spec_disabled()) { if (why) *why = h->spec ? "RMSB_NO_SPEC=1" : "mechanism does not qualify"; return 0; }
    if (kind != RMSB_DOMAIN_CONSTANT_V && kind != RMSB_DOMAIN_CONSTANT_P) { if (why) *why = "domain does not qualify"; return 0; }
    rmsb_spec_state& s = *h->spec;
    SpecKernels& k = s.k[kind == RMSB_DOMAIN_CONSTANT_P ? 1 : 0];
    std::lock_guard<std::mutex> lock(s.mu);
    if (k.state == 0) {
        k.state = -1;
        const std::string src = make_source(s, kind == RMSB_DOMAIN_CONSTANT_P, &k.ldt, &k.ldj);
        std::vector<char> cubin;
        if (compile_cubin(src, &cubin, &k.why)) {
            DeviceGuard g(h->device);
            cudaError_t e = cudaLibraryLoadData(&k.lib, cubin.data(), nullptr, nullptr, 0, nullptr, nullptr, 0);
            if (e == cudaSuccess) e = cudaLibraryGetKernel(&k.dydt, k.lib, "rmsb_spec_dydt");
            if (e == cudaSuccess) e = cudaLibraryGetKernel(&k.jac, k.lib, "rmsb_spec_jac");
            if (e == cudaSuccess && sizeof(double) * SPEC_THREADS * k.ldj > 48 * 1024)
                e = cudaFuncSetAttribute((const void*)k.jac, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(sizeof(double) * SPEC_THREADS * k.ldj));
            if (e == cudaSuccess) e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&k.occ_dydt, (const void*)k.dydt, SPEC_THREADS, sizeof(double) * SPEC_THREADS * k.ldt);
            if (e == cudaSuccess) e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&k.occ_jac, (const void*)k.jac, SPEC_THREADS, sizeof(double) * SPEC_THREADS * k.ldj);
            if (e == cudaSuccess && (k.occ_dydt < 1 || k.occ_jac < 1)) e = cudaErrorLaunchOutOfResources;
            if (e == cudaSuccess) k.state = 1;
            else { k.why = std::string("loading the specialised kernels: ") + cudaGetErrorString(e); cudaGetLastError(); }
        }
    }
    if (why) *why = k.why;
    return k.state == 1 ? 1 : 0;
}

// mode 0: dydt, 1: dense column-major Jacobian.  Call only after rmsb_spec_ready returned 1.
int rmsb_spec_launch_eval(const rmsb_mech* h, int mode, const BatchArgs& a, cudaStream_t st) {
    const SpecKernels& k = h->spec->k[a.domain_kind == RMSB_DOMAIN_CONSTANT_P ? 1 : 0];
    cudaDeviceProp prop;
    CU(rmsb_cached_props(h->device, &prop));
    const int ld = mode == 0 ? k.ldt : k.ldj;
    const size_t smem = sizeof(double) * (size_t)SPEC_THREADS * ld;
    // one resident wave of blocks striding over the batch: every block then does (nearly) the same number of tiles
    const int64_t blocks = (a.n + SPEC_THREADS - 1) / SPEC_THREADS, cap = (int64_t)prop.multiProcessorCount * (mode == 0 ? k.occ_dydt : k.occ_jac) * env_blocks("RMSB_SPEC_WAVES", 1);
    const int grid = (int)std::max<int64_t>(1, std::min(blocks, cap));
    const double* Y = a.y; const double* COND = a.cond; int ncond = a.ncond; const double* PP = a.p; long long p_stride = a.p_stride;
    const int* cs = a.constspc; int nconst = a.nconst; long long n = a.n; double* OUT = a.out;
    void* args[] = {&Y, &COND, &ncond, &PP, &p_stride, &cs, &nconst, &n, &OUT};
    CU(cudaLaunchKernel((const void*)(mode == 0 ? k.dydt : k.jac), dim3(grid), dim3(SPEC_THREADS), args, smem, st));
    ++g_launches;
    return RMSB_OK;
}

void rmsb_spec_release(rmsb_mech* h) {
    if (!h->spec) return;
    for (auto& k : h->spec->k) if (k.lib) { cudaLibraryUnload(k.lib); k.lib = nullptr; k.state = 0; }
}
