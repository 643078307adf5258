# RMSB200.jl -- Julia host shim for librmsb200.so (include/rmsb200.h).
#
# STATUS: written against the C ABI (rmsb200.h version 200), NOT EXECUTED in the build image (no `julia` binary there; SURVEY.md
# "Container facts") -- expect to fix small things on first load.  The executable host mirror used by the tests is the Python package `rmsb200/`; this file is
# the binding a ReactionMechanismSimulator.jl maintainer would load.  It keeps the reference API surface:
# `readinput`-parsed `Species`/`ElementaryReaction`, `IdealGas`, `ConstantVDomain`/`ConstantTPDomain`/
# `ConstantTVDomain`, `Reactor`, `solve`, `Simulation` all stay RMS's own; only the hot path moves:
#
#     sols = solve(reactors::Vector{<:Reactor}, RMSB200(); abstol=1e-16, reltol=1e-6)   # one kernel launch
#     sols = solve(reactors, RMSB200(devices=[0,1,2,3,4,5,6,7]); ...)                  # one process, eight GPUs (rmsb_solve_batch_multi)
#     sol  = solve(react.ode, RMSB200(react); abstol, reltol)                          # drop-in for CVODE_BDF()
#     sols = solve(systems::Vector{<:Reactor}, RMSB200(); ...)                         # gas + surface + ReactiveInternalInterfaceConstantTPhi (C4)
#
# Rules the device path imposes (rmsb_last_error names them when violated): more than 128 state variables -> ConstantTV /
# ConstantTAPhi / ConstantTP domains only (large-mechanism path); ConstantTV parameter Jacobians / sensitivities need the explicit `p`.
#
# Every `ccall` below names the header entry point it binds; layouts are row = reactor, Float64, 1-based
# species indices exactly as `phase.rxnarray` stores them (src/Phase.jl:320-348).
module RMSB200Shim

using ReactionMechanismSimulator
using SciMLBase
const RMS = ReactionMechanismSimulator

const LIB = get(ENV, "RMSB200_LIB", joinpath(@__DIR__, "..", "csrc", "librmsb200.so"))

# ---- mirrors of the C structs (include/rmsb200.h) -------------------------------------------------
struct MechDesc
    struct_size::Int32; phase_kind::Int32; n_species::Int32; n_reactions::Int32
    sitedensity::Float64; diffusion_limited::Int32; nasa_max_ranges::Int32
    nasa_nranges::Ptr{Int32}; nasa_ncoef::Ptr{Int32}; nasa_tmax::Ptr{Float64}; nasa_coefs::Ptr{Float64}
    rxnarray::Ptr{Int32}; kin_type::Ptr{Int32}; kin_offset::Ptr{Int32}; kin_params::Ptr{Float64}
    eff_ptr::Ptr{Int32}; eff_species::Ptr{Int32}; eff_value::Ptr{Float64}
    reversible::Ptr{UInt8}; forwardable::Ptr{UInt8}; radius::Ptr{Float64}; diff_radius::Ptr{Float64}
    viscosity_kind::Int32; viscosity_params::NTuple{5,Float64}
    species_phase::Ptr{Int32}; rxn_domain::Ptr{Int32}      # two-phase (gas + surface + interface) systems only, else C_NULL
end

struct Stats
    n_steps::Int32; n_rhs::Int32; n_jac::Int32; n_lu::Int32; n_newton_fail::Int32; n_err_fail::Int32
    last_order::Int32; reserved::Int32; last_h::Float64
end

struct SolveOpts                                          # rmsb_solve_opts, version 2 (40 bytes)
    struct_size::Int32; max_steps::Int32; max_order::Int32; idt_index::Int32
    first_step::Float64; warps_per_block::Int32; reserved::Int32
    dysave::Ptr{Float64}                                  # [Ny, nsave, n] du at the saved states, or C_NULL
end

const PHASE_GAS, PHASE_LIQUID, PHASE_SURFACE = Int32(0), Int32(1), Int32(2)
const DOMAIN_V, DOMAIN_TP, DOMAIN_TV, DOMAIN_TAPHI, DOMAIN_COUPLED_TP_TAPHI, DOMAIN_P = Int32(0), Int32(1), Int32(2), Int32(3), Int32(4), Int32(5)
const KIN = (ARRHENIUS_VEC=0, ARRHENIUS=1, MULTIARRHENIUS=2, PLOG_VEC=3, PLOG=4, MULTIPLOG=5, THIRDBODY=6, LINDEMANN=7, TROE=8, CHEBYSHEV_VEC=9)

lasterror() = unsafe_string(ccall((:rmsb_last_error, LIB), Cstring, ()))
check(rc) = rc == 0 || error("librmsb200 error $rc: $(lasterror())")

# ---- flattening: phase (already flattened by RMS, src/Phase.jl:40-57) -> descriptor -------------------
arr(a::Arrhenius) = Float64[a.A, a.n, a.Ea]
function plogblock(k::PdepArrhenius)
    out = Float64[length(k.Ps); k.Ps]
    for a in k.arrs
        a isa Arrhenius ? append!(out, [1.0; arr(a)]) : append!(out, [Float64(length(a.arrs)); vcat(arr.(a.arrs)...)])
    end
    out
end
"(tag, params) in the layout documented in include/rmsb200.h; `isvec` = reaction lies in a vectorised group (index <= veckineticsinds[end])"
function packkinetics(k, isvec::Bool)
    k isa Arrhenius && return (isvec ? KIN.ARRHENIUS_VEC : KIN.ARRHENIUS, arr(k))
    k isa MultiArrhenius && return (KIN.MULTIARRHENIUS, [Float64(length(k.arrs)); vcat(arr.(k.arrs)...)])
    k isa PdepArrhenius && return isvec ? (KIN.PLOG_VEC, [Float64(length(k.Ps)); k.Ps; vcat(arr.(k.arrs)...)]) : (KIN.PLOG, plogblock(k))
    k isa MultiPdepArrhenius && return (KIN.MULTIPLOG, [Float64(length(k.parrs)); vcat(plogblock.(k.parrs)...)])
    k isa ThirdBody && return (KIN.THIRDBODY, arr(k.arr))
    k isa Lindemann && return (KIN.LINDEMANN, [arr(k.arrhigh); arr(k.arrlow)])
    k isa Troe && return (KIN.TROE, [arr(k.arrhigh); arr(k.arrlow); k.a; k.T3; k.T1; k.T2])
    k isa Chebyshev && return (KIN.CHEBYSHEV_VEC, [k.Tmin; k.Tmax; k.Pmin; k.Pmax; Float64.(size(k.coefs))...; Base.vec(permutedims(k.coefs))])
    error("kinetics type $(typeof(k)) is outside the accelerated hot path")
end

mutable struct Mech
    handle::Ptr{Cvoid}
    keep::Vector{Any}          # arrays the descriptor points into
    Ns::Int; Nr::Int
end

"viscosity of the solvent -> (RMSB_VISC_*, params); anything the device does not evaluate is an error, not a silent VISC_NONE"
function packviscosity(phase)
    (hasproperty(phase, :solvent) && phase.solvent !== nothing) || return (Int32(0), (0.0, 0.0, 0.0, 0.0, 0.0))
    mu = phase.solvent.mu
    mu isa RiedelViscosity && return (Int32(2), (mu.A, mu.B, mu.C, mu.D, mu.E))
    mu isa ConstantViscosity && return (Int32(1), (mu.mu, 0.0, 0.0, 0.0, 0.0))
    error("viscosity type $(typeof(mu)) is outside the accelerated hot path")
end

"""
Flatten species / reactions (already ordered by RMS, src/Phase.jl:40-57) into the descriptor.  `species_phase` / `rxn_domain`
are C_NULL for a single phase; the two-phase constructor below fills them (include/rmsb200.h, rmsb_mech_desc).
"""
function Mech(species, reactions, isvec::Vector{Bool}, kind::Int32; sitedensity=0.0, diffusionlimited=false, visc=(Int32(0), (0.0, 0.0, 0.0, 0.0, 0.0)),
              spcindex=Dict(s.name => i for (i, s) in enumerate(species)), species_phase=Int32[], rxn_domain=Int32[], device::Integer=0)
    Ns, Nr = length(species), length(reactions)
    maxr = maximum(length(s.thermo.polys) for s in species)
    nr = Int32[length(s.thermo.polys) for s in species]
    nc = fill(Int32(7), maxr, Ns); tmax = zeros(maxr, Ns); coefs = zeros(9, maxr, Ns)        # column-major == C [Ns][maxr][9]
    for (i, s) in enumerate(species), (r, p) in enumerate(s.thermo.polys)
        nc[r, i] = length(p.coefs); tmax[r, i] = p.Tmax; coefs[1:length(p.coefs), r, i] = p.coefs
    end
    rx = zeros(Int32, 8, Nr)                                                                  # 8 x Nr column-major == C [Nr][8]
    for (j, rxn) in enumerate(reactions)
        for (k, sp) in enumerate(rxn.reactants); rx[k, j] = spcindex[sp.name]; end
        for (k, sp) in enumerate(rxn.products); rx[4 + k, j] = spcindex[sp.name]; end
    end
    ktype = Int32[]; koff = Int32[0]; kpar = Float64[]; effptr = Int32[0]; effspc = Int32[]; effval = Float64[]
    for (j, rxn) in enumerate(reactions)
        tag, pars = packkinetics(rxn.kinetics, isvec[j])
        push!(ktype, tag); append!(kpar, pars); push!(koff, length(kpar))
        if hasproperty(rxn.kinetics, :efficiencies)
            for i in sort(collect(keys(rxn.kinetics.efficiencies)))
                push!(effspc, i); push!(effval, rxn.kinetics.efficiencies[i])                 # already eff-1 (src/Parse.jl:411-413)
            end
        end
        push!(effptr, length(effspc))
    end
    rev = UInt8[r.reversible for r in reactions]; fwd = UInt8[r.forwardable for r in reactions]
    radius = Float64[s.radius for s in species]
    diffr = Float64[s.diffusion isa StokesDiffusivity ? s.diffusion.r : -1.0 for s in species]
    isempty(kpar) && push!(kpar, 0.0); isempty(effspc) && push!(effspc, 0); isempty(effval) && push!(effval, 0.0)
    keep = Any[nr, nc, tmax, coefs, rx, ktype, koff, kpar, effptr, effspc, effval, rev, fwd, radius, diffr, species_phase, rxn_domain]
    desc = Ref(MechDesc(sizeof(MechDesc), kind, Ns, Nr, sitedensity, Int32(diffusionlimited), maxr, pointer(nr), pointer(nc), pointer(tmax),
        pointer(coefs), pointer(rx), pointer(ktype), pointer(koff), pointer(kpar), pointer(effptr), pointer(effspc), pointer(effval),
        pointer(rev), pointer(fwd), pointer(radius), pointer(diffr), visc[1], visc[2],
        isempty(species_phase) ? Ptr{Int32}(C_NULL) : pointer(species_phase), isempty(rxn_domain) ? Ptr{Int32}(C_NULL) : pointer(rxn_domain)))
    h = Ref{Ptr{Cvoid}}(C_NULL)
    GC.@preserve keep check(ccall((:rmsb_mech_create, LIB), Cint, (Ref{MechDesc}, Cint, Ref{Ptr{Cvoid}}), desc, device, h))   # rmsb_mech_create
    m = Mech(h[], keep, Ns, Nr)
    finalizer(x -> ccall((:rmsb_mech_destroy, LIB), Cvoid, (Ptr{Cvoid},), x.handle), m)                                       # rmsb_mech_destroy
    m
end

"single phase: IdealGas / IdealDiluteSolution / IdealSurface"
function Mech(phase; device::Integer=0)
    nvec = isempty(phase.veckineticsinds) ? 0 : phase.veckineticsinds[end]
    kind = phase isa IdealSurface ? PHASE_SURFACE : (phase isa IdealDiluteSolution ? PHASE_LIQUID : PHASE_GAS)
    Mech(phase.species, phase.reactions, Bool[j <= nvec for j in 1:length(phase.reactions)], kind;
         sitedensity=hasproperty(phase, :sitedensity) ? phase.sitedensity : 0.0, diffusionlimited=phase.diffusionlimited,
         visc=packviscosity(phase), device=device)
end

"""
Two-phase mechanism for config C4: gas (ConstantTP) + surface (ConstantTAPhi) + the reactions of a
ReactiveInternalInterfaceConstantTPhi (src/Interface.jl:71-133).  Species: gas then surface; reactions: gas | surface | interface
(rxn_domain 0 / 1 / 2).  `inter.reactions` have already been upgraded by RMS (upgradekinetics turns sticking coefficients into
Arrhenius, src/Interface.jl:196-232), and interface reactions use the scalar rate forms.
"""
function Mech(gas::IdealGas, surf::IdealSurface, inter; device::Integer=0)
    species = vcat(gas.species, surf.species)
    reactions = vcat(gas.reactions, surf.reactions, inter.reactions)
    nvg = isempty(gas.veckineticsinds) ? 0 : gas.veckineticsinds[end]
    nvs = isempty(surf.veckineticsinds) ? 0 : surf.veckineticsinds[end]
    isvec = vcat(Bool[j <= nvg for j in 1:length(gas.reactions)], Bool[j <= nvs for j in 1:length(surf.reactions)], falses(length(inter.reactions)))
    sphase = vcat(zeros(Int32, length(gas.species)), ones(Int32, length(surf.species)))
    rdom = vcat(zeros(Int32, length(gas.reactions)), ones(Int32, length(surf.reactions)), fill(Int32(2), length(inter.reactions)))
    Mech(species, reactions, collect(isvec), PHASE_GAS; sitedensity=surf.sitedensity, species_phase=sphase, rxn_domain=rdom, device=device)
end

# ---- domain -> (kind, cond, constant species) -----------------------------------------------------------
domainkind(d::ConstantVDomain) = (DOMAIN_V, Float64[d.V])
domainkind(d::ConstantPDomain) = (DOMAIN_P, Float64[d.P])
domainkind(d::ConstantTPDomain) = (DOMAIN_TP, Float64[d.T, d.P])
domainkind(d::ConstantTVDomain) = (DOMAIN_TV, Float64[d.T, d.V])
domainkind(d::ConstantTAPhiDomain) = (DOMAIN_TAPHI, Float64[d.T, d.A])

"p is passed to the device only when it differs from the domain's own default (the nothermochg && nokfchg branch, src/Domain.jl:941-946)"
devicep(d, p) = (p isa SciMLBase.NullParameters || p == d.p) ? (C_NULL, 0) : (pointer(p), 0)

"""
Parameters of an ensemble: (buffer to keep alive, pointer, p_stride).  All reactors on their domain defaults -> (nothing, C_NULL, 0);
one shared non-default vector -> p_stride 0; reactors with different vectors -> an [Ns+Nr, n] matrix, p_stride = Ns+Nr.
"""
function ensemblep(reactors)
    ps = [r.p for r in reactors]
    all(i -> ps[i] isa SciMLBase.NullParameters || ps[i] == reactors[i].domain.p, eachindex(ps)) &&
        all(r -> r.domain.p == reactors[1].domain.p, reactors) && return (nothing, Ptr{Float64}(C_NULL), 0)
    pv = [p isa SciMLBase.NullParameters ? r.domain.p : p for (p, r) in zip(ps, reactors)]
    all(==(pv[1]), pv) && return (pv[1], pointer(pv[1]), 0)
    P = hcat(pv...)
    (P, pointer(P), size(P, 1))
end

# ---- batched hot-path evaluations ---------------------------------------------------------------------------
"dydt for n states of reactors sharing `domain`'s phase: Y is Ny x n (column = reactor)"
function dydt_batch(mech::Mech, domain, Y::Matrix{Float64}, cond::Matrix{Float64}; p=SciMLBase.NullParameters())
    kind, _ = domainkind(domain)
    out = similar(Y); pp, ps = devicep(domain, p)
    cs = Int32.(domain.constantspeciesinds)
    GC.@preserve p check(ccall((:rmsb_dydt_batch, LIB), Cint,                                                                 # rmsb_dydt_batch
        (Ptr{Cvoid}, Cint, Int64, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Int64, Ptr{Int32}, Int32, Ptr{Float64}),
        mech.handle, kind, size(Y, 2), Y, cond, pp, ps, isempty(cs) ? C_NULL : cs, length(cs), out))
    out
end

"dense Jacobians, Ny x Ny x n, J[eqn, var, r] (column-major per reactor == jacobiany!'s layout)"
function jac_batch(mech::Mech, domain, Y::Matrix{Float64}, cond::Matrix{Float64}; p=SciMLBase.NullParameters())
    kind, _ = domainkind(domain)
    Ny, n = size(Y); out = zeros(Ny, Ny, n); pp, ps = devicep(domain, p)
    cs = Int32.(domain.constantspeciesinds)
    GC.@preserve p check(ccall((:rmsb_jac_batch, LIB), Cint,                                                                  # rmsb_jac_batch
        (Ptr{Cvoid}, Cint, Int64, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Int64, Ptr{Int32}, Int32, Ptr{Float64}),
        mech.handle, kind, n, Y, cond, pp, ps, isempty(cs) ? C_NULL : cs, length(cs), out))
    out
end

"parameter Jacobians, Ny x (Ns+Nr) x n, Jp[eqn, par, r] (== jacobianp!'s layout, src/Domain.jl:3095-3485)"
function jacp_batch(mech::Mech, domain, Y::Matrix{Float64}, cond::Matrix{Float64}; p=domain.p)
    kind, _ = domainkind(domain)
    Ny, n = size(Y); Np = length(domain.phase.species) + length(domain.phase.reactions); out = zeros(Ny, Np, n)
    cs = Int32.(domain.constantspeciesinds)
    pp, ps = domain isa ConstantTVDomain ? (pointer(p), 0) : devicep(domain, p)      # ConstantTV needs p itself (diffusion-limit chain rule)
    GC.@preserve p check(ccall((:rmsb_jacp_batch, LIB), Cint,                                                                 # rmsb_jacp_batch
        (Ptr{Cvoid}, Cint, Int64, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Int64, Ptr{Int32}, Int32, Ptr{Float64}),
        mech.handle, kind, n, Y, cond, pp, ps, isempty(cs) ? C_NULL : cs, length(cs), out))
    out
end

"""
Forward sensitivities of many reactors in ONE call (rmsb_sens_batch): what `Reactor(...; forwardsensitivities=true)` + `solve`
+ `extract_local_sensitivities(sol, t)` (src/Reactor.jl:63-75, src/TestReactors.jl:239) or `threadedsensitivities(react)`
(src/ThreadedSensitivities.jl:15-79) compute one reactor / one parameter at a time.  Returns (yend[Ny,n], S[Ny,Np,n]) at the
end time and, with `saveat`, (ysave[Ny,k,n], ssave[Ny,Np,k,n]).
"""
function sens_batch(reactors::Vector{<:RMS.Reactor}; device=0, abstol=1e-6, reltol=1e-3, saveat=Float64[], max_steps=500000)
    r1 = reactors[1]; dom = r1.domain; mech = mechfor(dom.phase, device)
    kind, _ = domainkind(dom)
    n, Ny = length(reactors), length(r1.y0); Np = length(dom.phase.species) + length(dom.phase.reactions)
    Y0 = hcat((r.y0 for r in reactors)...)
    cond = hcat((domainkind(r.domain)[2] for r in reactors)...)
    tend = Float64[r.tspan[2] for r in reactors]
    ts = collect(Float64, saveat); nsave = length(ts)
    ysave = zeros(Ny, nsave, n); ssave = zeros(Ny, Np, nsave, n); yend = zeros(Ny, n); send = zeros(Ny, Np, n)
    status = zeros(Int32, n); stats = Vector{Stats}(undef, n)
    opts = Ref(SolveOpts(sizeof(SolveOpts), max_steps, 5, -1, 0.0, 0, 0, Ptr{Float64}(C_NULL)))
    cs = Int32.(dom.constantspeciesinds)
    pp, ps = dom isa ConstantTVDomain ? (pointer(r1.p), 0) : devicep(dom, r1.p)
    GC.@preserve r1 check(ccall((:rmsb_sens_batch, LIB), Cint,                                                                  # rmsb_sens_batch
        (Ptr{Cvoid}, Cint, Int64, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Int64, Ptr{Int32}, Int32, Ptr{Float64}, Float64, Float64,
         Int32, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Int32}, Ptr{Stats}, Ref{SolveOpts}),
        mech.handle, kind, n, Y0, cond, pp, ps, isempty(cs) ? C_NULL : cs, length(cs), tend, reltol, abstol,
        nsave, nsave > 0 ? ts : C_NULL, nsave > 0 ? ysave : C_NULL, nsave > 0 ? ssave : C_NULL, yend, send, status, stats, opts))
    all(status .== 0) || @warn "rmsb_sens_batch: $(count(!=(0), status)) reactors did not reach t_end"
    (yend, send, ysave, ssave, status)
end

# ---- solve ------------------------------------------------------------------------------------------------------
struct RMSB200 <: SciMLBase.AbstractODEAlgorithm
    max_steps::Int32; max_order::Int32; devices::Vector{Int}
    reactor::Any                      # RMSB200(react): the reactor whose `react.ode` is being solved (single-reactor drop-in)
end
RMSB200(; max_steps=500000, max_order=5, device=0, devices=[device]) = RMSB200(max_steps, max_order, collect(Int, devices), nothing)
RMSB200(react::RMS.Reactor; kw...) = (a = RMSB200(; kw...); RMSB200(a.max_steps, a.max_order, a.devices, react))

const MECHCACHE = Dict{Tuple{UInt,Int},Mech}()
mechfor(phase, device) = get!(() -> Mech(phase; device=device), MECHCACHE, (objectid(phase), device))

"(mechanism per device, domain kind, Y0 [Ny,n], cond [ncond,n]) of an ensemble: single-domain reactors, or gas + surface systems (C4)"
function ensembleinputs(reactors, devices)
    r1 = reactors[1]
    if r1.domain isa Tuple                                        # multi-domain Reactor (src/Reactor.jl:99-258)
        (length(r1.domain) == 2 && length(r1.interfaces) == 1 && r1.interfaces[1] isa ReactiveInternalInterfaceConstantTPhi &&
         r1.domain[1] isa ConstantTPDomain && r1.domain[2] isa ConstantTAPhiDomain) ||
            error("multi-domain systems other than gas(ConstantTP) + surface(ConstantTAPhi) + ReactiveInternalInterfaceConstantTPhi are outside the accelerated hot path")
        dg, dc, inter = r1.domain[1], r1.domain[2], r1.interfaces[1]
        mechs = [get!(() -> Mech(dg.phase, dc.phase, inter; device=d), MECHCACHE, (objectid(inter), d)) for d in devices]
        # y = [ns_gas; ns_surf; V] is the reference's own layout (src/Reactor.jl:106-132); cond = [T, P, A]
        Y0 = hcat((r.y0 for r in reactors)...)
        cond = hcat((Float64[r.domain[1].T, r.domain[1].P, r.interfaces[1].A] for r in reactors)...)
        return mechs, DOMAIN_COUPLED_TP_TAPHI, Y0, cond, Int32[]
    end
    mechs = [mechfor(r1.domain.phase, d) for d in devices]
    kind, _ = domainkind(r1.domain)
    Y0 = hcat((r.y0 for r in reactors)...)
    cond = hcat((domainkind(r.domain)[2] for r in reactors)...)
    mechs, kind, Y0, cond, Int32.(r1.domain.constantspeciesinds)
end

"""
Integrate many reactors that share one phase and domain type in ONE call (rmsb_solve_batch; rmsb_solve_batch_multi when
`alg.devices` names several GPUs: contiguous blocks of reactors per device, one host thread each).  Returns one `ODESolution`
per reactor built exactly as the reference itself fabricates solutions
(`SciMLBase.build_solution(prob, alg, t, u; interp, retcode)`, src/Simulation.jl:72, src/ThreadedSensitivities.jl:165):
cubic Hermite interpolation over the saved (t, u, du) -- du comes back from the device (rmsb_solve_opts.dysave), so
`Simulation(sol, domain)` works unchanged and no host RHS is evaluated.
"""
function SciMLBase.solve(reactors::Vector{<:RMS.Reactor}, alg::RMSB200; abstol=1e-6, reltol=1e-3, saveat=Float64[])
    mechs, kind, Y0, cond, cs = ensembleinputs(reactors, alg.devices)
    Ny, n = size(Y0)
    tend = Float64[r.tspan[2] for r in reactors]
    ts = isempty(saveat) ? collect(range(0.0, maximum(tend), length=201)) : collect(Float64, saveat)
    nsave = length(ts)
    ysave = zeros(Ny, nsave, n); dysave = zeros(Ny, nsave, n); yend = zeros(Ny, n); idt = zeros(n)
    status = zeros(Int32, n); stats = Vector{Stats}(undef, n)
    opts = Ref(SolveOpts(sizeof(SolveOpts), alg.max_steps, alg.max_order, -1, 0.0, 0, 0, pointer(dysave)))
    pkeep, pp, ps = ensemblep(reactors)
    handles = Ptr{Cvoid}[m.handle for m in mechs]
    GC.@preserve pkeep dysave handles begin
        if length(handles) == 1
            check(ccall((:rmsb_solve_batch, LIB), Cint,                                                                       # rmsb_solve_batch
                (Ptr{Cvoid}, Cint, Int64, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Int64, Ptr{Int32}, Int32, Ptr{Float64}, Float64, Float64,
                 Int32, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Int32}, Ptr{Stats}, Ref{SolveOpts}),
                handles[1], kind, n, Y0, cond, pp, ps, isempty(cs) ? C_NULL : cs, length(cs), tend, reltol, abstol,
                nsave, ts, ysave, yend, idt, status, stats, opts))
        else
            check(ccall((:rmsb_solve_batch_multi, LIB), Cint,                                                                 # rmsb_solve_batch_multi
                (Ptr{Ptr{Cvoid}}, Int32, Cint, Int64, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Int64, Ptr{Int32}, Int32, Ptr{Float64}, Float64, Float64,
                 Int32, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Int32}, Ptr{Stats}, Ref{SolveOpts}),
                handles, length(handles), kind, n, Y0, cond, pp, ps, isempty(cs) ? C_NULL : cs, length(cs), tend, reltol, abstol,
                nsave, ts, ysave, yend, idt, status, stats, opts))
        end
    end
    map(1:n) do i
        keepidx = findall(t -> t <= tend[i], ts)
        u = [ysave[:, k, i] for k in keepidx]
        du = [dysave[:, k, i] for k in keepidx]
        interp = SciMLBase.HermiteInterpolation(ts[keepidx], u, du)
        SciMLBase.build_solution(reactors[i].ode, alg, ts[keepidx], u; interp=interp,
                                 retcode=status[i] == 0 ? SciMLBase.ReturnCode.Success : SciMLBase.ReturnCode.Failure)
    end
end

"`solve(react.ode, RMSB200(react); abstol, reltol)`: the single-reactor drop-in for `solve(react.ode, CVODE_BDF(); ...)` (src/Reactor.jl:76-81)"
function SciMLBase.solve(prob::SciMLBase.ODEProblem, alg::RMSB200; kw...)
    alg.reactor === nothing && error("construct the algorithm with the reactor: RMSB200(react)")
    prob === alg.reactor.ode || error("RMSB200(react) must be used with react.ode")
    SciMLBase.solve([alg.reactor], alg; kw...)[1]
end

end # module
