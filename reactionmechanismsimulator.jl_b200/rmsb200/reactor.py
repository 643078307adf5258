"""Reactor + solve: the reference's `Reactor(domain, y0, tspan; p)` / `solve(react.ode, CVODE_BDF();
abstol, reltol)` surface (src/Reactor.jl:30-98; call sites src/TestReactors.jl:22,197,305) for one
reactor or an ensemble.  `solve` runs the whole ensemble in ONE call to `rmsb_solve_batch`.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Sequence, Tuple

import numpy as np

from . import _lib as L
from .domain import Domain


class RMSB200:
    """Algorithm tag standing where CVODE_BDF() stands in the reference (`react.recommendedsolver`)."""

    def __init__(self, max_steps=500000, max_order=5, first_step=0.0, warps_per_block=0, flags=0):
        """flags: 0 = production defaults (CVODE-style Newton convergence test, step-size hysteresis);
        bit 0 = scipy-style Newton test; bit 1 = no hysteresis (3 reproduces scipy.integrate.BDF's step logic)."""
        self.max_steps, self.max_order, self.first_step, self.warps_per_block = max_steps, max_order, first_step, warps_per_block
        self.flags = flags


class Reactor:
    """Reactor(domain, y0, tspan; p) (src/Reactor.jl:13-28,30-98).  `domain` may hold an ensemble."""

    def __new__(cls, domain, y0=None, tspan=(0.0, 1.0), interfaces=(), p=None, **kw):
        """Reactor((domaingas, domaincat), (y0gas, y0cat), tspan, (inter,), (pgas, pcat, pinter)) -> (react, y0, p):
        the multi-domain constructor of the reference (src/Reactor.jl:99-258) for one ConstantTP gas domain, one
        ConstantTAPhi surface domain and one ReactiveInternalInterfaceConstantTPhi."""
        if isinstance(domain, (tuple, list)):
            from .coupled import ReactiveInternalInterfaceConstantTPhi, coupled_domain
            if interfaces is None or len(interfaces) == 0:
                # domains without interfaces (src/TestReactors.jl:470-531): a block-diagonal system, each block solved by its own call
                from .multidomain import MultiDomainReactor
                mr = MultiDomainReactor(domain, y0, tspan, p, forwardsensitivities=bool(kw.get("forwardsensitivities", False)))
                return mr, mr.y0.copy(), mr.p
            if len(domain) != 2 or len(interfaces) != 1 or not isinstance(interfaces[0], ReactiveInternalInterfaceConstantTPhi):
                raise NotImplementedError("multi-domain systems other than gas(TP) + surface(TAPhi) + ReactiveInternalInterfaceConstantTPhi "
                                          "are outside the accelerated hot path (DESIGN.md)")
            dom = coupled_domain(interfaces[0])
            pall = None if p is None else np.concatenate([np.ravel(np.asarray(x, dtype=float)) for x in p])   # src/Reactor.jl:135-200
            dom.p, dom._p_default = pall, (None if pall is None else pall.copy())
            react = object.__new__(cls)
            Reactor.__init__(react, dom, None, tspan, (), pall)
            react.interfaces = list(interfaces)
            return react, dom.y0[0].copy(), pall
        return object.__new__(cls)

    def __init__(self, domain: Domain, y0=None, tspan: Tuple[float, float] = (0.0, 1.0), interfaces: Sequence = (), p=None,
                 forwardsensitivities: bool = False):
        if forwardsensitivities and domain.kind == L.DOMAIN_COUPLED_TP_TAPHI:
            raise NotImplementedError("forward sensitivities of coupled gas-surface systems (DESIGN.md, 'next')")
        self.domain = domain
        self.y0 = domain.y0 if y0 is None else np.ascontiguousarray(np.atleast_2d(y0), dtype=float)
        assert tspan[0] == 0.0, "integration starts at t = 0"
        self.tspan = (float(tspan[0]), tspan[1])
        self.p = domain.p if p is None else p
        self.interfaces = interfaces if isinstance(interfaces, np.ndarray) else list(interfaces)     # interface objects, or packed rows (pack_interfaces)
        self.recommendedsolver = RMSB200()
        self.forwardsensitivities = bool(forwardsensitivities)

    @property
    def ode(self):          # `solve(react.ode, ...)` reads naturally
        return self


class Solution:
    """What comes back from `solve`: the role of the ODESolution the reference's Simulation consumes
    (src/Simulation.jl:25-39): saved times `t`, states `u[r, k, :]`, final state, status and per-reactor
    integrator statistics; `sol(t)` interpolates (cubic Hermite on the saved points, like the
    Sundials.jl solution object, when derivatives were requested; linear otherwise)."""

    def __init__(self, reactor, t, u, yend, idt, status, stats, t_end, sens=None, sens_end=None):
        self.reactor, self.t, self.u, self.yend = reactor, t, u, yend
        self.idt, self.status, self.stats, self.t_end = idt, status, stats, t_end
        self.sens, self.sens_end = sens, sens_end          # [n][nsave][Np][Ny], [n][Np][Ny] (forward sensitivities only)
        self._du = None

    def extract_local_sensitivities(self, t=None, r: int = 0):
        """(x, dp) as SciMLSensitivity's extract_local_sensitivities(sol, t) returns them for reactor r
        (call site src/TestReactors.jl:239): x = y(t), dp[j] = dy/dp_j (t).  t = None -> t_end; otherwise t must be one
        of the saved times."""
        if self.sens_end is None:
            raise ValueError("solution was computed without forwardsensitivities=True")
        if t is None:
            return self.yend[r].copy(), [v.copy() for v in self.sens_end[r]]
        k = int(np.argmin(np.abs(self.t - t)))
        if abs(self.t[k] - t) > 1e-12 * max(1.0, abs(t)):
            raise ValueError("sensitivities are available at the saved times (saveat=...) and at t_end")
        return self.u[r, k].copy(), [v.copy() for v in self.sens[r, k]]

    @property
    def retcode(self):
        return ["Success" if s == L.SOLVE_OK else {1: "MaxIters", 2: "DtLessThanMin", 3: "Unstable"}[int(s)] for s in self.status]

    def derivatives(self):
        """du/dt at the saved points: returned by the solve call itself when it was asked to (save_derivatives=True, the
        `dysave` buffer of rmsb_solve_opts), otherwise one batched dydt call per saved time."""
        if self._du is None:
            dom = self.reactor.domain
            n, k, Ny = self.u.shape
            du = np.empty_like(self.u)
            for j in range(k):
                yj = np.ascontiguousarray(self.u[:, j, :])
                ok = np.isfinite(yj).all(axis=1)
                d = np.full_like(yj, np.nan)
                if ok.all():
                    d = dom.dydt(yj, self.reactor.p)
                du[:, j, :] = d
            self._du = du
        return self._du

    def __call__(self, tq, r: int = 0, hermite: bool = True):
        """State of reactor r at time tq."""
        t = self.t
        if tq <= t[0]:
            return self.u[r, 0].copy()
        if tq >= t[-1]:
            return self.u[r, -1].copy()
        k = int(np.searchsorted(t, tq, side="right")) - 1
        h = t[k + 1] - t[k]
        s = (tq - t[k]) / h
        y0, y1 = self.u[r, k], self.u[r, k + 1]
        if not hermite:
            return (1 - s) * y0 + s * y1
        du = self.derivatives()
        d0, d1 = du[r, k], du[r, k + 1]
        h00, h10 = 2 * s ** 3 - 3 * s ** 2 + 1, s ** 3 - 2 * s ** 2 + s
        h01, h11 = -2 * s ** 3 + 3 * s ** 2, s ** 3 - s ** 2
        return h00 * y0 + h10 * h * d0 + h01 * y1 + h11 * h * d1


def threadedsensitivities(react: Reactor, paramindices=None, odesolver: Optional[RMSB200] = None, senssolver=None,
                          odekwargs: Optional[dict] = None, senskwargs: Optional[dict] = None) -> "Solution":
    """threadedsensitivities(react; paramindices, odesolver, senssolver, odekwargs, senskwargs)
    (src/ThreadedSensitivities.jl:15-79).  The reference solves the state once and then one linear sensitivity ODE per
    parameter on a thread each; here the state and ALL parameter columns advance together in one rmsb_sens_batch call, so
    the result is the same object `solve` returns for a reactor built with forwardsensitivities=True.  `paramindices`
    (1-based, as in the reference) selects which columns are kept; tolerances are taken from `odekwargs`."""
    kw = dict(odekwargs or {})
    kw.update({k: v for k, v in (senskwargs or {}).items() if k not in kw})
    r2 = Reactor(react.domain, react.y0, react.tspan, p=react.p, forwardsensitivities=True)
    sol = solve(r2, odesolver, abstol=kw.get("abstol", 1e-6), reltol=kw.get("reltol", 1e-3), saveat=kw.get("saveat"))
    if paramindices is not None:
        idx = np.asarray(paramindices, dtype=int) - 1
        sol.sens_end = sol.sens_end[:, idx]
        if sol.sens is not None:
            sol.sens = sol.sens[:, :, idx]
    return sol


def solve(react: Reactor, alg: Optional[RMSB200] = None, *, abstol: float = 1e-6, reltol: float = 1e-3,
          saveat: Optional[Sequence[float]] = None, idt_index: int = -1, t_end=None, save_derivatives: bool = False,
          devices: Optional[Sequence[int]] = None) -> Solution:
    """solve(react.ode, alg; abstol, reltol[, saveat]) -- default tolerances are SciML's (1e-6, 1e-3).
    save_derivatives: also return du at the saved states (what sol.interp.du holds in the reference) from the same call.
    devices: GPU ordinals to shard the ensemble over from this one process (rmsb_solve_batch_multi: contiguous blocks by
    reactor index, one host thread per device); None = the device of the phase's handle."""
    if type(react).__name__ == "MultiDomainReactor":
        from .multidomain import solve_multidomain
        kw = dict(abstol=abstol, reltol=reltol, saveat=saveat, save_derivatives=save_derivatives)
        return solve_multidomain(react, alg, **kw)
    alg = alg or react.recommendedsolver
    dom = react.domain
    y0 = react.y0
    n, Ny = y0.shape
    te = react.tspan[1] if t_end is None else t_end
    te = np.ascontiguousarray(np.broadcast_to(np.asarray(te, dtype=float), (n,)))
    tsave = None if saveat is None else np.ascontiguousarray(np.asarray(saveat, dtype=float))
    nsave = 0 if tsave is None else tsave.size
    ysave = np.empty((n, nsave, Ny)) if nsave else None
    yend = np.empty((n, Ny))
    idt = np.empty(n)
    status = np.empty(n, dtype=np.int32)
    stats = np.zeros(n, dtype=L.STATS_DTYPE)
    pp, ps = dom._p_for_sens(react.p) if react.forwardsensitivities else dom._p_for_device(react.p)
    cs = dom.constantspeciesinds
    opts = L.SolveOpts(C.sizeof(L.SolveOpts), alg.max_steps, alg.max_order, idt_index, alg.first_step, alg.warps_per_block, alg.flags)
    if react.forwardsensitivities:
        # ODEForwardSensitivityProblem path (src/Reactor.jl:63-75): state + all Ns+Nr sensitivities in one call
        Np = dom.phase.Ns + dom.phase.Nr
        ssave = np.empty((n, nsave, Np, Ny)) if nsave else None
        send = np.empty((n, Np, Ny))
        xtr, keep = dom._extras(react.interfaces, n)               # boundary interfaces (src/TestReactors.jl:284-296) / time profile
        if xtr is not None:
            if getattr(dom, "profile", None) is not None:
                pp, ps = dom._p_for_device(react.p)
            L.check(L.lib().rmsb_sens_batch_ex(dom.phase.handle, dom.kind, n, L.ptr(y0), L.ptr(dom.cond), L.ptr(pp), ps,
                                               L.ptr(cs) if cs.size else None, cs.size, C.byref(xtr), L.ptr(te), reltol, abstol, nsave,
                                               L.ptr(tsave), L.ptr(ysave), L.ptr(ssave), L.ptr(yend), L.ptr(send), L.ptr(status), L.ptr(stats),
                                               C.byref(opts)))
        else:
            L.check(L.lib().rmsb_sens_batch(dom.phase.handle, dom.kind, n, L.ptr(y0), L.ptr(dom.cond), L.ptr(pp), ps,
                                            L.ptr(cs) if cs.size else None, cs.size, L.ptr(te), reltol, abstol, nsave, L.ptr(tsave),
                                            L.ptr(ysave), L.ptr(ssave), L.ptr(yend), L.ptr(send), L.ptr(status), L.ptr(stats), C.byref(opts)))
        idt[:] = np.nan
        sol = Solution(react, tsave, ysave, yend, idt, status, stats, te, ssave, send)
        sol.kernel_ms = float(L.lib().rmsb_last_kernel_ms())
        return sol
    dysave = np.empty((n, nsave, Ny)) if (save_derivatives and nsave) else None
    if dysave is not None:
        opts.dysave = dysave.ctypes.data
    xtr, keep = (None, ()) if dom.kind == L.DOMAIN_COUPLED_TP_TAPHI else dom._extras(react.interfaces, n)   # internal interfaces live in the merged mechanism
    if xtr is not None:
        # boundary interfaces and / or a time profile (Parametrized* domains)
        if devices is not None and len(devices) > 1:
            raise NotImplementedError("boundary interfaces / time profiles with the multi-device call")
        L.check(L.lib().rmsb_solve_batch_ex(dom.phase.handle, dom.kind, n, L.ptr(y0), L.ptr(dom.cond), L.ptr(pp), ps,
                                            L.ptr(cs) if cs.size else None, cs.size, C.byref(xtr), L.ptr(te), reltol, abstol, nsave, L.ptr(tsave),
                                            L.ptr(ysave), L.ptr(yend), L.ptr(idt), L.ptr(status), L.ptr(stats), C.byref(opts)))
    elif devices is not None and len(devices) > 0:
        handles = (C.c_void_p * len(devices))(*[dom.phase.handle_on(int(d)) for d in devices])
        L.check(L.lib().rmsb_solve_batch_multi(handles, len(devices), dom.kind, n, L.ptr(y0), L.ptr(dom.cond), L.ptr(pp), ps,
                                               L.ptr(cs) if cs.size else None, cs.size, L.ptr(te), reltol, abstol, nsave, L.ptr(tsave),
                                               L.ptr(ysave), L.ptr(yend), L.ptr(idt), L.ptr(status), L.ptr(stats), C.byref(opts)))
    else:
        L.check(L.lib().rmsb_solve_batch(dom.phase.handle, dom.kind, n, L.ptr(y0), L.ptr(dom.cond), L.ptr(pp), ps,
                                         L.ptr(cs) if cs.size else None, cs.size, L.ptr(te), reltol, abstol, nsave, L.ptr(tsave),
                                         L.ptr(ysave), L.ptr(yend), L.ptr(idt), L.ptr(status), L.ptr(stats), C.byref(opts)))
    sol = Solution(react, tsave, ysave, yend, idt, status, stats, te)
    sol._du = dysave
    sol.kernel_ms = float(L.lib().rmsb_last_kernel_ms())
    return sol
