"""Plain elementwise relative error of the CUDA path against the oracle -- dydt, jacobiany, jacobianp, kf, krev on every parity
case -- on all entries above the cancellation floor derived in oracle/floors.py, at the north-star bar of 1e-12; every entry
below the floor must sit within a few floors.  The table is printed (pytest -s) and written to gpurun_out/elementwise_parity.txt
(copied to profiles/ and DESIGN.md section 6)."""
import os

import numpy as np
import pytest

import cases
import rmsb200 as rb
from oracle import floors as fl
from oracle import rms_oracle as ro
from rmsb200 import _lib as L

pytestmark = pytest.mark.gpu
BAR = 1e-12
ROWS = []


def _ens(gdom, ys):
    ens = rb.Domain(gdom.kind, gdom.phase, ys, np.repeat(gdom.cond, len(ys), axis=0))
    ens.constantspeciesinds = gdom.constantspeciesinds
    ens.p, ens._p_default = gdom.p, gdom._p_default
    return ens


def _row(case, what, rep):
    ROWS.append(f"{case:18s} {what:5s} entries {rep['n']:6d}  above floor {rep['n_above']:6d}  max plain rel err above floor {rep['max_rel_above']:.2e}  "
                f"worst |err| / (1e-12 |ref| + k floors) {rep['worst_excess']:.2e}   (k = 4 for dydt, 8 otherwise)")
    print(ROWS[-1])


@pytest.mark.parametrize("case", list(cases.CASES))
def test_elementwise_parity(case):
    assert L.lib().rmsb_device_count() > 0, "GPU tests need a CUDA device (no CPU fallback exists)"
    odom = cases.oracle_domain(case)
    gdom, y0, p = cases.product_domain(case)
    Ns, Nr = odom.phase.Ns, odom.phase.Nr
    ys = cases.states(case, odom.y0, Ns, odom.kind)[:2]
    ens = _ens(gdom, ys)
    variable_k = odom.kind in ("ConstantV", "ConstantP")
    pset = [odom.p if odom.kind in ("ConstantV", "ConstantTP") else None]
    if variable_k:
        pset.append(cases.parameter_vector(case, Ns, Nr))
    agg = {}

    def merge(what, rep):
        a = agg.setdefault(what, {"n": 0, "n_above": 0, "max_rel_above": 0.0, "worst_excess": 0.0})
        a["n"] += rep["n"]; a["n_above"] += rep["n_above"]
        a["max_rel_above"] = max(a["max_rel_above"], rep["max_rel_above"])
        if rep["worst_excess"] >= a["worst_excess"]:
            a["worst_excess"] = rep["worst_excess"]
            a["worst"] = {k: rep[k] for k in ("worst_at", "worst_x", "worst_ref", "worst_floor")}

    for po in pset:
        pg = p if po is odom.p or po is None else po
        f, J = ens.dydt(ys, pg), ens.jacobiany(ys, pg)
        kf, kr = ens.rates(ys, pg)
        Jp = ens.jacobianp(ys, pg if (pg is not None or odom.kind != "ConstantTV") else p)
        for k, y in enumerate(ys):
            fo, M = fl.dydt_terms(odom, y, 0.0, po)
            merge("dydt", fl.report(f[k], fo, fl.EPS * M, BAR, slack=4.0))
            Jo, Jf = fl.jac_floor(lambda yy, pp: ro.jacobiany(odom, yy, 0.0, pp), odom, y, po)
            merge("J", fl.report(J[k], Jo, Jf, BAR))
            pj = po if po is not None else odom.p
            Po, Pf = fl.jacp_floor(odom, y, pj)
            merge("Jp", fl.report(Jp[k], Po, Pf, BAR))
            th = odom.calcthermo(y, 0.0, po)
            # rate coefficients are products and exponentials, no cancellation: the floor is the conditioning of exp() only
            _, kff = fl.numeric_floor(lambda yy, pp: odom.calcthermo(yy, 0.0, pp).kfs, [y, po])
            _, krf = fl.numeric_floor(lambda yy, pp: odom.calcthermo(yy, 0.0, pp).krevs, [y, po])
            merge("kf", fl.report(kf[k], th.kfs, kff, BAR))
            merge("krev", fl.report(kr[k], th.krevs, krf, BAR))
    for what, rep in agg.items():
        _row(case, what, rep)
        if rep["worst_excess"] > 0.5 or rep["max_rel_above"] > 0.5 * BAR:
            print("    worst entry:", rep.get("worst"), "Ns", Ns, "Nr", Nr)
    for what, rep in agg.items():
        assert rep["max_rel_above"] <= BAR, (case, what, rep.get("worst"))
        assert rep["worst_excess"] <= 1.0, (case, what, rep.get("worst"))


def test_elementwise_parity_thread_per_reactor_and_group_kernels():
    """The large-batch (thread-per-reactor) and small-batch (lane-group) evaluation kernels of the headline mechanism against the
    oracle with the same floors: they replace the looser kernel-vs-kernel bars of tests/test_gpu_tpr.py."""
    case = "superminimal_V"
    odom = cases.oracle_domain(case)
    gdom, y0, p = cases.product_domain(case)
    Ns = odom.phase.Ns
    rng = np.random.default_rng(3)
    n = 4608
    ys = np.repeat(odom.y0[None], n, axis=0)
    tot = odom.y0[:Ns].sum()
    ys[:, :Ns] = np.abs(ys[:, :Ns] + tot * 10.0 ** rng.uniform(-8, -2, (n, Ns)) * rng.uniform(0.1, 1, (n, Ns)))
    ys[:, Ns] *= rng.uniform(0.9, 1.6, n)
    ys[:, Ns + 1] *= rng.uniform(0.8, 1.5, n)
    big = _ens(gdom, ys)
    f_big, J_big = big.dydt(ys, p), big.jacobiany(ys, p)
    sub = _ens(gdom, ys[:64])
    f_small, J_small = sub.dydt(ys[:64], p), sub.jacobiany(ys[:64], p)
    for name, f, J in (("thread/reactor", f_big, J_big), ("lane group", f_small, J_small)):
        agg_f = {"n": 0, "n_above": 0, "max_rel_above": 0.0, "worst_excess": 0.0}
        agg_J = dict(agg_f)
        for k in range(48):
            fo, M = fl.dydt_terms(odom, ys[k], 0.0, odom.p)
            Jo, Jf = fl.jac_floor(lambda yy, pp: ro.jacobiany(odom, yy, 0.0, pp), odom, ys[k], odom.p, nsample=4)
            for agg, rep in ((agg_f, fl.report(f[k], fo, fl.EPS * M, BAR, slack=4.0)), (agg_J, fl.report(J[k], Jo, Jf, BAR))):
                agg["n"] += rep["n"]; agg["n_above"] += rep["n_above"]
                agg["max_rel_above"] = max(agg["max_rel_above"], rep["max_rel_above"]); agg["worst_excess"] = max(agg["worst_excess"], rep["worst_excess"])
        _row(f"C2 {name}", "dydt", agg_f)
        _row(f"C2 {name}", "J", agg_J)
        assert agg_f["max_rel_above"] <= BAR and agg_f["worst_excess"] <= 1.0
        assert agg_J["max_rel_above"] <= BAR and agg_J["worst_excess"] <= 1.0


def test_write_table():
    out = cases.ROOT / "gpurun_out"
    if ROWS and out.is_dir() and os.access(out, os.W_OK):
        (out / "elementwise_parity.txt").write_text("\n".join(ROWS) + "\n")
