/* rmsb200.h -- C ABI of librmsb200.so, the B200-native batched stiff-kinetics engine that
 * replaces ReactionMechanismSimulator.jl's data-parallel hot path.
 *
 * The reference (pure Julia) has no FFI seam; the seam this library sits behind is the SciML
 * problem/solution contract the reference builds in src/Reactor.jl:31-35,65,76
 *   dydt(dy,y,p,t) = dydtreactor!(dy,y,t,domain,interfaces,p=p)     src/Reactor.jl:653-660
 *   jacy!(J,y,p,t) = jacobiany!(J,y,p,t,domain,interfaces,nothing)   src/Domain.jl:2204-3091
 *   jacp!(J,y,p,t) = jacobianp!(J,y,p,t,domain,interfaces,nothing)   src/Domain.jl:3376-3485
 *   solve(react.ode, CVODE_BDF(); abstol, reltol) -> ODESolution     (call sites src/TestReactors.jl:22,197,305)
 * evaluated for MANY independent reactors that share one flattened mechanism
 * (IdealGas/IdealDiluteSolution/IdealSurface, src/Phase.jl:40-136).
 *
 * Conventions
 *  - every entry point returns int: 0 = ok, <0 = error; text via rmsb_last_error() (thread local).
 *  - all pointers in the *_batch entry points are CALLER-OWNED HOST buffers, row = reactor,
 *    contiguous, Float64; the library owns all device memory and copies in/out itself.
 *  - the *_dev entry points take DEVICE pointers (same layouts) and a cudaStream_t (as void*);
 *    they only enqueue work -- no allocation of caller-visible memory, no synchronisation.
 *  - a mechanism handle is immutable after creation and may be shared by host threads.
 *  - species / reaction indices in the descriptor are 1-based with 0 = none, exactly the
 *    reference's rxnarray convention (src/Phase.jl:320-348), so a Julia caller passes its arrays as is.
 *  - there is NO CPU fallback: without a CUDA device every compute entry point fails with
 *    RMSB_ERR_CUDA.
 */
#ifndef RMSB200_H
#define RMSB200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RMSB_VERSION 200

/* ---- error codes -------------------------------------------------------------------------- */
#define RMSB_OK             0
#define RMSB_ERR_ARG       -1   /* bad argument / inconsistent descriptor            */
#define RMSB_ERR_CUDA      -2   /* CUDA runtime error or no device                   */
#define RMSB_ERR_UNSUPPORTED -3 /* feature outside the hot path      */
#define RMSB_ERR_NOMEM     -4

/* ---- phase kinds: which getC0 / calcthermo family applies (src/Phase.jl:301-307) ---------- */
#define RMSB_PHASE_GAS      0   /* IdealGas:            C0 = 1e5/(R*T)                */
#define RMSB_PHASE_LIQUID   1   /* IdealDiluteSolution: C0 = 1e5/(R*T), diffusion limits */
#define RMSB_PHASE_SURFACE  2   /* IdealSurface:        C0 = sitedensity             */

/* ---- domain kinds and their state / condition layouts (src/Domain.jl) ---------------------
 *  kind            y (Ny)            cond (Ncond)   reference
 *  CONSTANT_V      [ns; T; P] Ns+2   [V]            src/Domain.jl:120-170,1064-1113,1901-1906
 *  CONSTANT_TP     [ns; V]    Ns+1   [T, P]         src/Domain.jl:42-102,916-966,1863-1867
 *  CONSTANT_TV     ns         Ns     [T, V]         src/Domain.jl:500-560,1514-1561
 *  CONSTANT_TAPHI  ns         Ns     [T, A]         src/Domain.jl:660-711,1597-1633  (phi = 0)
 *  COUPLED_TP_TAPHI [ns_bulk; ns_surf; V] Ns+1  [T, P, A]   ConstantTP gas + ConstantTAPhi surface + ReactiveInternalInterfaceConstantTPhi
 *                                                 in one system: src/Reactor.jl:99-258,662-725, src/Interface.jl:71-133 (needs a
 *                                                 two-phase descriptor: species_phase / rxn_domain below)
 */
#define RMSB_DOMAIN_CONSTANT_V      0
#define RMSB_DOMAIN_CONSTANT_TP     1
#define RMSB_DOMAIN_CONSTANT_TV     2
#define RMSB_DOMAIN_CONSTANT_TAPHI  3
#define RMSB_DOMAIN_COUPLED_TP_TAPHI 4
/*  CONSTANT_P      [ns; T; V] Ns+2   [P]            src/Domain.jl:173-239,1141-1192,1950-1953 (adiabatic, constant pressure) */
#define RMSB_DOMAIN_CONSTANT_P      5

/* ---- kinetics type tags (src/Calculators/Rate.jl, Ratevec.jl) -----------------------------
 * kin_params[kin_offset[j] .. kin_offset[j+1]) holds, as doubles:
 *  ARRHENIUS_VEC   A n Ea                         k = A*exp(n*log(T) - Ea*(1/(R*T)))   Ratevec.jl:24
 *  ARRHENIUS       A n Ea                         k = A*T^n*exp(-Ea/(R*T))            Rate.jl:16
 *  MULTIARRHENIUS  m, m x (A n Ea)                sum of scalar Arrhenius             Rate.jl:68-74
 *  PLOG_VEC        nP, Ps[nP], nP x (A n Ea)      PdepArrheniusvec                    Ratevec.jl:204-215
 *  PLOG            nP, Ps[nP], nP x {m, m x (A n Ea)}   PdepArrhenius (entries Arrhenius/MultiArrhenius) Rate.jl:48-60
 *  MULTIPLOG       nK, nK x {PLOG block}          MultiPdepArrhenius                  Rate.jl:82-88
 *  THIRDBODY       A n Ea                         k = Ceff*Arrhenius                  Rate.jl:98
 *  LINDEMANN       high(A n Ea) low(A n Ea)                                           Rate.jl:109-114
 *  TROE            high(A n Ea) low(A n Ea) a T3 T1 T2                                Rate.jl:129-147
 *  CHEBYSHEV_VEC   Tmin Tmax Pmin Pmax nT nP coefs[nT][nP] (row major)                Ratevec.jl:89-100
 */
#define RMSB_KIN_ARRHENIUS_VEC   0
#define RMSB_KIN_ARRHENIUS       1
#define RMSB_KIN_MULTIARRHENIUS  2
#define RMSB_KIN_PLOG_VEC        3
#define RMSB_KIN_PLOG            4
#define RMSB_KIN_MULTIPLOG       5
#define RMSB_KIN_THIRDBODY       6
#define RMSB_KIN_LINDEMANN       7
#define RMSB_KIN_TROE            8
#define RMSB_KIN_CHEBYSHEV_VEC   9

#define RMSB_VISC_NONE     0
#define RMSB_VISC_CONSTANT 1   /* params[0] = mu                          src/Calculators/Viscosity.jl:9-13 */
#define RMSB_VISC_RIEDEL   2   /* params = A B C D E                      src/Calculators/Viscosity.jl:15-24 */

/* Flattened mechanism, reactions in PHASE ORDER (vectorised kinetics first, src/Phase.jl:42). */
typedef struct rmsb_mech_desc {
    int32_t struct_size;          /* = sizeof(rmsb_mech_desc); ABI guard                                  */
    int32_t phase_kind;           /* RMSB_PHASE_*                                                         */
    int32_t n_species;
    int32_t n_reactions;
    double  sitedensity;          /* IdealSurface only (src/Phase.jl:133)                                 */
    int32_t diffusion_limited;    /* phase.diffusionlimited (src/Phase.jl:56,95)                          */
    int32_t nasa_max_ranges;      /* max #polynomials per species                                         */
    const int32_t* nasa_nranges;  /* [Ns]                                                                 */
    const int32_t* nasa_ncoef;    /* [Ns][max_ranges]  7 or 9 (src/Calculators/Thermo.jl:16-63)           */
    const double*  nasa_tmax;     /* [Ns][max_ranges]  selectPoly: first range with T <= Tmax (:101-111)  */
    const double*  nasa_coefs;    /* [Ns][max_ranges][9]                                                  */
    const int32_t* rxnarray;      /* [Nr][8] 1-based, 0 = none; cols 0-3 reactants, 4-7 products          */
    const int32_t* kin_type;      /* [Nr] RMSB_KIN_*                                                      */
    const int32_t* kin_offset;    /* [Nr+1]                                                               */
    const double*  kin_params;    /* [kin_offset[Nr]]                                                     */
    const int32_t* eff_ptr;       /* [Nr+1] CSR over reactions of third-body efficiencies                 */
    const int32_t* eff_species;   /* 1-based species index                                                */
    const double*  eff_value;     /* efficiency - 1 (src/Parse.jl:411-413)                                */
    const uint8_t* reversible;    /* [Nr]                                                                 */
    const uint8_t* forwardable;   /* [Nr]                                                                 */
    const double*  radius;        /* [Ns] species.radius (liquid, Smoluchowski), may be NULL              */
    const double*  diff_radius;   /* [Ns] StokesDiffusivity.r; <= 0 means EmptyDiffusivity (Inf); may be NULL */
    int32_t viscosity_kind;       /* RMSB_VISC_*  (phase.solvent.mu)                                      */
    double  viscosity_params[5];
    /* two-phase (gas + surface) systems only, NULL otherwise: species of the bulk phase first, then the surface phase */
    const int32_t* species_phase; /* [Ns] 0 = bulk phase (concentration = n/V), 1 = surface phase (n/A)              */
    const int32_t* rxn_domain;    /* [Nr] 0 = bulk reaction (rate*V), 1 = surface reaction (rate*A),                 */
                                  /*      2 = interface reaction (rate*A, constant k, two-phase Kc; src/Interface.jl:89-108) */
} rmsb_mech_desc;

typedef struct rmsb_mech rmsb_mech;    /* opaque, immutable */

/* Per-reactor integration statistics (all counters are per reactor). */
typedef struct rmsb_stats {
    int32_t n_steps;        /* accepted steps                 */
    int32_t n_rhs;          /* dydt evaluations               */
    int32_t n_jac;          /* Jacobian evaluations           */
    int32_t n_lu;           /* LU factorisations of I - c*J   */
    int32_t n_newton_fail;  /* Newton convergence failures    */
    int32_t n_err_fail;     /* error-test failures            */
    int32_t last_order;
    int32_t reserved;
    double  last_h;
} rmsb_stats;

/* solve status per reactor */
#define RMSB_SOLVE_OK            0
#define RMSB_SOLVE_MAX_STEPS     1
#define RMSB_SOLVE_STEP_TOO_SMALL 2
#define RMSB_SOLVE_NAN           3

typedef struct rmsb_solve_opts {
    int32_t struct_size;     /* = sizeof(rmsb_solve_opts)                                          */
    int32_t max_steps;       /* per reactor, default 500000                                        */
    int32_t max_order;       /* 1..5, default 5                                                    */
    int32_t idt_index;       /* 0-based y index whose max d/dt defines idt; -1 = domain default (T) */
    double  first_step;      /* 0 = automatic                                                      */
    int32_t warps_per_block; /* 0 = automatic (most resident warps per SM); 1..4 fixed; results
                                are bit-identical for every value                                   */
    int32_t reserved;        /* integrator flags: 0 = defaults; bit 0 = scipy-style Newton test; bit 1 = no step hysteresis;
                              * bit 2 / bit 3 (diagnostics): do not impose the unit rows of inert species / recompute the step as
                              * fl(t+h)-t like SciPy; bit 4 (large-mechanism path): refactorise on every change of h instead of
                              * keeping the factors while c = h/alpha stays within 30 % (CVODE's dgmax); bit 6 (lane-group path): hand the reactors out in
                              * the caller's order instead of hottest-first (for callers that sorted by a measured cost); bits 8-15 = v: Newton tolerance coefficient 0.1/2^(v-1) (0 = default 0.1/8) */
    /* ---- version 2 (struct_size >= 40; version-1 callers with struct_size 32 are still accepted) ---- */
    double* dysave;          /* rmsb_solve_batch only (host buffer, may be NULL): [n][nsave][Ny], dy/dt = dydtreactor!(u) at every
                              * saved state -- the `du` of sol.interp (src/Simulation.jl:27-34), so the caller's Hermite interpolation
                              * needs no host RHS.  NaN where ysave is NaN. */
} rmsb_solve_opts;

/* ---- library ----------------------------------------------------------------------------- */
int         rmsb_version(void);
const char* rmsb_last_error(void);
int         rmsb_device_count(void);

/* ---- mechanism --------------------------------------------------------------------------- */
/* Flatten once into SoA device tables on `device` (replaces the per-call object walks of
 * src/PhaseState.jl:61-76,165-171 and the Dict iteration of :53-58). */
int  rmsb_mech_create(const rmsb_mech_desc* desc, int device, rmsb_mech** out);
void rmsb_mech_destroy(rmsb_mech* mech);
int  rmsb_mech_ny(const rmsb_mech* mech, int domain_kind);     /* state length Ny    */
int  rmsb_mech_ncond(const rmsb_mech* mech, int domain_kind);  /* #per-reactor consts */
int  rmsb_mech_device(const rmsb_mech* mech);

/* ---- batched hot-path evaluations (host buffers) ------------------------------------------
 * y     [n][Ny]      cond [n][Ncond]
 * p     NULL, or [Ns+Nr] shared by all reactors (p_stride = 0), or [n][Ns+Nr] (p_stride = Ns+Nr):
 *       CONSTANT_V  : p = [dG_i ; f_j]    h/RT += dG/(RT), kf,krev *= f          src/Domain.jl:1100,1112
 *       constant-k  : p = [G_i ; k_j]     absolute values (efficiency reactions: multiplier) :89-93,551,702
 * const_species [n_const] 1-based species whose dydt / Jacobian rows are zeroed (:1904-1906,2318-2320)
 */
int rmsb_dydt_batch(const rmsb_mech* mech, int domain_kind, int64_t n,
                    const double* y, const double* cond, const double* p, int64_t p_stride,
                    const int32_t* const_species, int32_t n_const,
                    double* dydt /* [n][Ny] */);

/* J[r] is dense COLUMN-MAJOR Ny x Ny, J[eqn + Ny*var] = d(dy_eqn/dt)/dy_var (jacobiany!). */
int rmsb_jac_batch(const rmsb_mech* mech, int domain_kind, int64_t n,
                   const double* y, const double* cond, const double* p, int64_t p_stride,
                   const int32_t* const_species, int32_t n_const,
                   double* J /* [n][Ny*Ny] */);

/* Jp[r] is dense COLUMN-MAJOR Ny x (Ns+Nr), Jp[eqn + Ny*par] = d(dy_eqn/dt)/dp_par (jacobianp!,
 * src/Domain.jl:3095-3485): par < Ns -> G_par, par = Ns+j -> k_j (or its multiplier, see p above).
 * CONSTANT_TV needs an explicit p (chain rule through the diffusion limit, :3290-3317). */
int rmsb_jacp_batch(const rmsb_mech* mech, int domain_kind, int64_t n,
                    const double* y, const double* cond, const double* p, int64_t p_stride,
                    const int32_t* const_species, int32_t n_const,
                    double* Jp /* [n][Ny*(Ns+Nr)] */);

/* kf,krev as calcthermo returns them (for rops / rates accessors, src/Simulation.jl:239-254,763-778). */
int rmsb_rates_batch(const rmsb_mech* mech, int domain_kind, int64_t n,
                     const double* y, const double* cond, const double* p, int64_t p_stride,
                     double* kf /* [n][Nr] */, double* krev /* [n][Nr] */);

/* getkfs (src/PhaseState.jl:165-171): forward rate coefficients BEFORE diffusion limitation -- what the constant-k
 * domain constructors store in p[Ns..] (kfsnondiff for liquid phases, src/Domain.jl:549-551). */
int rmsb_kfs_batch(const rmsb_mech* mech, int domain_kind, int64_t n,
                   const double* y, const double* cond, double* kf /* [n][Nr] */);

/* Species Gibbs energies G[r][s] at temperatures T[r]: scalar_path != 0 -> getGibbs through
 * getEnthalpy - T*getEntropy (src/Calculators/Thermo.jl:6-8,75-93, what the constant-T domain
 * constructors store in p[1:Ns]); 0 -> (h/RT - s/R)*R*T of calcHSCpdless (Thermovec.jl:76-119). */
int rmsb_gibbs_batch(const rmsb_mech* mech, int64_t n, const double* T, int scalar_path, double* G /* [n][Ns] */);

/* Integrate every reactor from t = 0 to t_end[r] with a variable-order (1-5) variable-step
 * BDF/NDF method, modified Newton, dense LU with partial pivoting, WRMS error control
 * |e_i| / (atol + rtol*|y_i|) -- the role CVODE_BDF plays in the reference (src/Reactor.jl:76-81).
 *  tsave  [nsave] shared output times (ascending); ysave [n][nsave][Ny] dense output (NaN past t_end)
 *  yend   [n][Ny] state at t_end (may be NULL)
 *  idt    [n] time of max d y[idt_index]/dt (continuous version of src/TestReactors.jl:307-308); NaN if none
 */
int rmsb_solve_batch(const rmsb_mech* mech, int domain_kind, int64_t n,
                     const double* y0, const double* cond, const double* p, int64_t p_stride,
                     const int32_t* const_species, int32_t n_const,
                     const double* t_end, double rtol, double atol,
                     int32_t nsave, const double* tsave, double* ysave,
                     double* yend, double* idt, int32_t* status, rmsb_stats* stats,
                     const rmsb_solve_opts* opts);

/* Forward sensitivities S = dy/dp of every reactor w.r.t. all Ns+Nr parameters, integrated together with the
 * state: the role of `Reactor(...; forwardsensitivities=true)` + ODEForwardSensitivityProblem
 * (src/Reactor.jl:63-75) and of threadedsensitivities (src/ThreadedSensitivities.jl:15-79).  Staggered direct
 * method: after each accepted BDF step of the state, the linear sensitivity corrector
 * (I - c*J(y_new)) dS = c*(J(y_new)*S_pred + Jp(y_new)) - psi_S is solved for all parameters with one fresh LU
 * (one lane per parameter column); sensitivities follow the state's step/order sequence and are not in its
 * error test (CVODES errconS = false).  S(0) = 0.
 *  ssave [n][nsave][Ns+Nr][Ny] (NaN past t_end), send [n][Ns+Nr][Ny]; either may be NULL. */
int rmsb_sens_batch(const rmsb_mech* mech, int domain_kind, int64_t n,
                    const double* y0, const double* cond, const double* p, int64_t p_stride,
                    const int32_t* const_species, int32_t n_const,
                    const double* t_end, double rtol, double atol,
                    int32_t nsave, const double* tsave, double* ysave, double* ssave,
                    double* yend, double* send, int32_t* status, rmsb_stats* stats,
                    const rmsb_solve_opts* opts);

/* ---- device-resident variants (inputs/outputs already in HBM; enqueue only) ---------------- */
int rmsb_dydt_batch_dev(const rmsb_mech* mech, int domain_kind, int64_t n,
                        const double* d_y, const double* d_cond, const double* d_p, int64_t p_stride,
                        const int32_t* d_const_species, int32_t n_const,
                        double* d_dydt, void* stream);
int rmsb_jac_batch_dev(const rmsb_mech* mech, int domain_kind, int64_t n,
                       const double* d_y, const double* d_cond, const double* d_p, int64_t p_stride,
                       const int32_t* d_const_species, int32_t n_const,
                       double* d_J, void* stream);
int rmsb_jacp_batch_dev(const rmsb_mech* mech, int domain_kind, int64_t n,
                        const double* d_y, const double* d_cond, const double* d_p, int64_t p_stride,
                        const int32_t* d_const_species, int32_t n_const,
                        double* d_Jp, void* stream);
/* d_work: rmsb_sens_work_bytes() bytes of device scratch (work-queue counter + the backward-difference
 * history of the sensitivities of every resident warp). */
int64_t rmsb_sens_work_bytes(const rmsb_mech* mech, int domain_kind, int64_t n);
int rmsb_sens_batch_dev(const rmsb_mech* mech, int domain_kind, int64_t n,
                        const double* d_y0, const double* d_cond, const double* d_p, int64_t p_stride,
                        const int32_t* d_const_species, int32_t n_const,
                        const double* d_t_end, double rtol, double atol,
                        int32_t nsave, const double* d_tsave, double* d_ysave, double* d_ssave,
                        double* d_yend, double* d_send, int32_t* d_status, rmsb_stats* d_stats,
                        const rmsb_solve_opts* opts, void* d_work, void* stream);
/* d_work: rmsb_solve_work_bytes() bytes of device scratch (work-queue counter etc.). */
int64_t rmsb_solve_work_bytes(const rmsb_mech* mech, int domain_kind, int64_t n);
int rmsb_solve_batch_dev(const rmsb_mech* mech, int domain_kind, int64_t n,
                         const double* d_y0, const double* d_cond, const double* d_p, int64_t p_stride,
                         const int32_t* d_const_species, int32_t n_const,
                         const double* d_t_end, double rtol, double atol,
                         int32_t nsave, const double* d_tsave, double* d_ysave,
                         double* d_yend, double* d_idt, int32_t* d_status, rmsb_stats* d_stats,
                         const rmsb_solve_opts* opts, void* d_work, void* stream);

/* ---- boundary interfaces (CSTR / PFR ensembles) ----------------------------------------------------
 * Inlet, Outlet, VolumetricFlowRateInlet, VolumetricFlowRateOutlet and VolumeMaintainingOutlet of the reference
 * (src/Interface.jl:234-365; consumed by calcdomainderivatives!, src/Domain.jl:1839-1998, and by jacobiany!, :2220-2294,
 * 2322-2448, 2493-2628, 3057-3088) with CONSTANT flows, for the ConstantTP, ConstantV, ConstantP, ConstantTV and ConstantTAPhi
 * domains with at most 128 state variables.  `interfaces` holds one row of 7 + 2*Ns doubles per reactor (if_stride = 7 + 2*Ns)
 * or one row shared by all (if_stride = 0):
 *   [0] Inlet molar flow F (mol/s)   [1] Outlet molar flow   [2] VolumetricFlowRateInlet Vin (m^3/s)   [3] VolumetricFlowRateOutlet Vout
 *   [4] VolumeMaintainingOutlet present (0 / 1; ConstantTP and ConstantP)   [5] Inlet enthalpy H (J/mol)
 *   [6] VolumetricFlowRateInlet enthalpy per volume (J/m^3)   [7 .. 7+Ns) Inlet mole fractions   [7+Ns .. 7+2Ns) inlet concentrations
 * Several interfaces of one type are merged by the caller (the terms are linear in flow * composition and flow * enthalpy).
 * Time-dependent flows F(t), the kLAkH reservoir interface, parameter Jacobians and sensitivities with interfaces: not built. */
int rmsb_dydt_batch_if(const rmsb_mech* mech, int domain_kind, int64_t n,
                       const double* y, const double* cond, const double* p, int64_t p_stride,
                       const int32_t* const_species, int32_t n_const,
                       const double* interfaces, int64_t if_stride, double* dydt);
int rmsb_jac_batch_if(const rmsb_mech* mech, int domain_kind, int64_t n,
                      const double* y, const double* cond, const double* p, int64_t p_stride,
                      const int32_t* const_species, int32_t n_const,
                      const double* interfaces, int64_t if_stride, double* J);
int rmsb_solve_batch_if(const rmsb_mech* mech, int domain_kind, int64_t n,
                        const double* y0, const double* cond, const double* p, int64_t p_stride,
                        const int32_t* const_species, int32_t n_const,
                        const double* interfaces, int64_t if_stride,
                        const double* t_end, double rtol, double atol,
                        int32_t nsave, const double* tsave, double* ysave,
                        double* yend, double* idt, int32_t* status, rmsb_stats* stats,
                        const rmsb_solve_opts* opts);

/* Parameter Jacobian and forward sensitivities with boundary interfaces (the reference's own tests run both on ConstantTP with an Inlet
 * and an Outlet, src/TestReactors.jl:271-296).  jacobianp! gains terms only for an Inlet on the adiabatic domains
 * (src/Domain.jl:3421-3430,3461-3470, reproduced as written); the sensitivity equations see the interfaces through the RHS and
 * jacobiany!. */
int rmsb_jacp_batch_if(const rmsb_mech* mech, int domain_kind, int64_t n, const double* y, const double* cond,
                       const double* p, int64_t p_stride, const int32_t* const_species, int32_t n_const,
                       const double* interfaces, int64_t if_stride, double* Jp);
int rmsb_sens_batch_if(const rmsb_mech* mech, int domain_kind, int64_t n, const double* y0, const double* cond,
                       const double* p, int64_t p_stride, const int32_t* const_species, int32_t n_const,
                       const double* interfaces, int64_t if_stride,
                       const double* t_end, double rtol, double atol, int32_t nsave, const double* tsave,
                       double* ysave, double* ssave, double* yend, double* send, int32_t* status, rmsb_stats* stats,
                       const rmsb_solve_opts* opts);


/* ---- time profiles (ParametrizedVDomain / ParametrizedPDomain) and the general extended calls ---------
 * The reference's rapid-compression-machine / shock-tube domains (src/Domain.jl:331-498): the ConstantV domain whose volume
 * follows V(t), the ConstantP domain whose pressure follows P(t) -- dT/dt and dP/dt (dV/dt) gain the dV/dt (dP/dt) terms of
 * src/Domain.jl:2003-2090, the Jacobian those of :2726-2748,2874-2896.  The profile is a natural cubic spline (the reference's
 * getspline, src/Tools.jl:46-85) handed over as [nt, ts[nt], values[nt], gamma[nt]] (gamma = second derivatives at the knots,
 * gamma[0] = gamma[nt-1] = 0; linear extrapolation outside the knots), one row per reactor (profile_stride = 1 + 3*nt) or one row
 * shared by all (0).  `cond` keeps its layout (its V / P entry is ignored).  domain_kind stays RMSB_DOMAIN_CONSTANT_V / _P.
 * ParametrizedTPDomain (src/Domain.jl:241-320,1437-1512,2002-2004,2635-2725): domain_kind RMSB_DOMAIN_CONSTANT_TP with `profile` =
 * the T table followed by the P table; y = [ns; V]; rate coefficients are evaluated on every call (vector thermo path) and
 * p = [dG_i ; f_j] as for ConstantV (not the [G ; k] of ConstantTP).
 * ParametrizedTConstantVDomain (src/Domain.jl:563-633,1372-1434; Jacobian of the constant-T family :3049-3055): domain_kind
 * RMSB_DOMAIN_CONSTANT_TV with `profile` = the T table; y = ns; Gibbs energies through the vector thermo path, viscosity,
 * diffusivities and the diffusion-limited kf / krev are evaluated at T(t) on every call; p = [dG_i ; f_j] (not the [G ; kf] of
 * ConstantTV); cond = [T (ignored), V].
 * t: evaluation time of every state for dydt / jac (NULL = 0); solve integrates from t = 0. */
typedef struct rmsb_extras {
    int32_t struct_size;          /* = sizeof(rmsb_extras) */
    const double* interfaces;     /* boundary interfaces, see rmsb_dydt_batch_if; NULL = none */
    int64_t if_stride;
    const double* profile;        /* time profile(s); NULL = none */
    int64_t profile_stride;
    const double* t;              /* [n] evaluation times (dydt / jac only); NULL = 0 */
} rmsb_extras;
int rmsb_dydt_batch_ex(const rmsb_mech* mech, int domain_kind, int64_t n,
                       const double* y, const double* cond, const double* p, int64_t p_stride,
                       const int32_t* const_species, int32_t n_const, const rmsb_extras* extras, double* dydt);
int rmsb_jac_batch_ex(const rmsb_mech* mech, int domain_kind, int64_t n,
                      const double* y, const double* cond, const double* p, int64_t p_stride,
                      const int32_t* const_species, int32_t n_const, const rmsb_extras* extras, double* J);
int rmsb_solve_batch_ex(const rmsb_mech* mech, int domain_kind, int64_t n,
                        const double* y0, const double* cond, const double* p, int64_t p_stride,
                        const int32_t* const_species, int32_t n_const, const rmsb_extras* extras,
                        const double* t_end, double rtol, double atol,
                        int32_t nsave, const double* tsave, double* ysave,
                        double* yend, double* idt, int32_t* status, rmsb_stats* stats,
                        const rmsb_solve_opts* opts);

/* Parameter Jacobian and forward sensitivities with interfaces and / or the time profile of a Parametrized* domain (the reference tests jacobianp of all
 * four, src/TestReactors.jl:163-165,423-425,444-446,465-467); jacobianp! of a Parametrized domain is the variable-k one
 * (p = [dG; f]) whatever its layout (src/Domain.jl:3204-3233). */
int rmsb_jacp_batch_ex(const rmsb_mech* mech, int domain_kind, int64_t n, const double* y, const double* cond,
                       const double* p, int64_t p_stride, const int32_t* const_species, int32_t n_const,
                       const rmsb_extras* extras, double* Jp);
int rmsb_sens_batch_ex(const rmsb_mech* mech, int domain_kind, int64_t n, const double* y0, const double* cond,
                       const double* p, int64_t p_stride, const int32_t* const_species, int32_t n_const,
                       const rmsb_extras* extras,
                       const double* t_end, double rtol, double atol, int32_t nsave, const double* tsave,
                       double* ysave, double* ssave, double* yend, double* send, int32_t* status, rmsb_stats* stats,
                       const rmsb_solve_opts* opts);


/* ---- one host process, several GPUs -------------------------------------------------------------
 * rmsb_solve_batch over n_mech handles of the SAME mechanism created on different devices: the reactors are split into
 * contiguous blocks by index (block g = reactors [g*n/n_mech, (g+1)*n/n_mech)), each block runs rmsb_solve_batch on its
 * handle's device from its own host thread, the outputs land in the caller's buffers at the reactor's index.  No exchange
 * between devices (reactors are independent); a one-process Julia caller reaches every GPU of the node through
 * this call.  Returns the first non-zero status of any block. */
int rmsb_solve_batch_multi(const rmsb_mech* const* mechs, int32_t n_mech, int domain_kind, int64_t n,
                           const double* y0, const double* cond, const double* p, int64_t p_stride,
                           const int32_t* const_species, int32_t n_const,
                           const double* t_end, double rtol, double atol,
                           int32_t nsave, const double* tsave, double* ysave,
                           double* yend, double* idt, int32_t* status, rmsb_stats* stats,
                           const rmsb_solve_opts* opts);

/* ---- large mechanisms (more than 128 state variables; config C3) ---------------------------
 * rmsb_mech_create analyses the Jacobian pattern of mechanisms with more than 128 species once (minimum-degree ordering,
 * fill pattern, dependency levels); dydt / jac / rates / solve then run one thread block per reactor with a sparse LU of
 * I - c*J on that fixed pattern -- the role of the GMRES + ILU path of the reference (src/Reactor.jl:342-382).
 * ConstantTV, ConstantTAPhi and ConstantTP domains (ConstantTP: the species block bordered by the dense volume row and
 * column; info describes the y = ns pattern).  info: [0] n, [1] nnz(L+U), [2] forward levels, [3] backward levels,
 * [4] longest row, [5] multiply-adds per factorisation, [6] first row of the dense core, [7] scratch doubles per resident reactor. */
int rmsb_large_info(const rmsb_mech* mech, int64_t info[8]);
/* cycle counters of the block-per-reactor solver summed over blocks: [0] rhs, [1] Jacobian, [2] factorisation,
 * [3] triangular solves, [4] whole solves; reset != 0 clears them after reading. */
int rmsb_large_profile(const rmsb_mech* mech, int reset, double cycles[8]);

/* ---- kernels generated for the mechanism ---------------------------------------------------------
 * For small all-Arrhenius gas mechanisms (n_species <= 32, n_reactions <= 128, no third bodies) on ConstantV / ConstantP, batches of
 * >= 4096 states in rmsb_dydt_batch / rmsb_jac_batch run on CUDA kernels the library writes out for the mechanism (every table
 * entry a literal, csrc/rmsb_spec.cu) and compiles with NVRTC (libnvrtc.so.12 of the CUDA toolkit, found through the dynamic
 * loader or RMSB_NVRTC=path) on first use; the arithmetic and its order are those of the generic kernels.  Without NVRTC, or with
 * RMSB_NO_SPEC=1 in the environment, the generic kernels run.
 * rmsb_mech_specialized: 1 if such kernels are loaded for (mech, domain) -- compiling them now if need be -- else 0 with the reason
 * in rmsb_last_error.  rmsb_spec_source_desc: the generated source (needs no device): writes at most cap bytes, NUL-terminated,
 * to buf (may be NULL) and the full length to *need. */
int rmsb_mech_specialized(const rmsb_mech* mech, int domain_kind);
int rmsb_spec_source_desc(const rmsb_mech_desc* desc, int domain_kind, char* buf, size_t cap, size_t* need);

/* ---- measurement helpers ------------------------------------------------------------------ */
/* Launch counter: number of kernels this library has launched in this process (bench gpu_launches). */
int64_t rmsb_kernel_launch_count(void);
/* Device time (CUDA events on the call's stream, milliseconds) of the kernels of the last rmsb_solve_batch / rmsb_sens_batch
 * call made by this host thread, copies excluded; < 0 if none. */
double  rmsb_last_kernel_ms(void);
/* FP64 FMA microbenchmark on `device`: returns sustained DFMA TFLOP/s (2 flop per FMA), <0 on error.
 * Used as the FP64 roofline denominator (MEASURED_PEAKS.json has no FP64 entry). */
double  rmsb_fp64_peak_tflops(int device, int iters);

#ifdef __cplusplus
}
#endif
#endif /* RMSB200_H */
