"""Host-side logic of the product (parser, flattening, descriptor packing, C-ABI surface) -- CPU only, no
compute calls.  The product's flattening is compared with the oracle's independent one."""
import ctypes as C
import re

import numpy as np
import pytest

import cases
import rmsb200 as rb
from oracle import cpu_oracle as co
from oracle import rms_oracle as ro
from rmsb200 import _lib as L
from rmsb200 import phase as ph


def test_library_exports_every_header_symbol():
    header = (cases.ROOT / "include" / "rmsb200.h").read_text()
    declared = set(re.findall(r"\b(rmsb_[a-z0-9_]+)\s*\(", header))
    assert declared, "no declarations found"
    lib = L.lib()
    for name in declared:
        assert hasattr(lib, name), f"librmsb200.so does not export {name}"
    assert set(L.EXPORTS) == declared
    assert lib.rmsb_version() == 200


def test_struct_layouts_match_header():
    assert C.sizeof(L.Stats) == 40 and C.sizeof(L.SolveOpts) == 40        # v2: + dysave (a version-1 struct of 32 bytes is still accepted)
    assert C.sizeof(L.MechDesc) == C.sizeof(co.Desc)


def test_no_gpu_fails_loudly():
    lib = L.lib()
    if lib.rmsb_device_count() > 0:
        pytest.skip("a CUDA device is present")
    d = rb.readinput(cases.fixture("superminimal"))
    ig = rb.IdealGas(d["phase"]["Species"], d["phase"]["Reactions"])
    with pytest.raises(rb.RMSBError, match="no CUDA device"):
        _ = ig.handle
    assert lib.rmsb_fp64_peak_tflops(0, 16) < 0


@pytest.mark.parametrize("fx,key", [("superminimal", "phase"), ("ethane", "phase"), ("propane", "phase"), ("liquid_phase", "phase"),
                                   ("minimal_edge", "phase"), ("ch4o2cat", "gas"), ("ch4o2cat", "surface")])
def test_flattening_agrees_with_oracle(fx, key):
    do, dp = ro.readinput(cases.fixture(fx)), rb.readinput(cases.fixture(fx))
    if fx == "liquid_phase":
        po = ro.IdealDiluteSolution(do[key]["Species"], do[key]["Reactions"], do["Solvents"][0])
        pp = rb.IdealDiluteSolution(dp[key]["Species"], dp[key]["Reactions"], dp["Solvents"][0])
    elif key == "surface":
        po = ro.IdealSurface(do[key]["Species"], do[key]["Reactions"], 2.486e-5)
        pp = rb.IdealSurface(dp[key]["Species"], dp[key]["Reactions"], 2.486e-5)
    else:
        po, pp = ro.IdealGas(do[key]["Species"], do[key]["Reactions"]), rb.IdealGas(dp[key]["Species"], dp[key]["Reactions"])
    assert (po.Ns, po.Nr, po.nvec) == (pp.Ns, pp.Nr, pp.nvec)
    np.testing.assert_array_equal(po.rxnarray.T + 1, pp.rxnarray)                    # reaction reordering (src/Phase.jl:42)
    assert [i + 1 for i in po.efficiencyinds] == pp.efficiencyinds
    cm = co.CpuMech(po)                                                            # independent packing of the same descriptor
    for name in ("kt", "ko", "ep", "es"):
        np.testing.assert_array_equal(cm._keep[name], getattr(pp, {"kt": "_ktype", "ko": "_koff", "ep": "_effptr", "es": "_effspc"}[name]))
    np.testing.assert_array_equal(cm._keep["kp"], pp._kpar)
    np.testing.assert_array_equal(cm._keep["coef"], pp._coefs)
    np.testing.assert_allclose(cm._keep["rad"], pp._radius, rtol=1e-15)
    np.testing.assert_allclose(cm._keep["dr"], pp._diffr, rtol=1e-15)


def test_reaction_order_vectorised_first():
    d = rb.readinput(cases.fixture("ethane"))
    ig = rb.IdealGas(d["phase"]["Species"], d["phase"]["Reactions"])
    types = [r.kinetics["type"] for r in ig.reactions]
    assert all(t in ("Arrhenius", "PdepArrhenius", "Chebyshev") for t in types[:ig.nvec])
    assert all(t in ("MultiArrhenius", "MultiPdepArrhenius", "ThirdBody", "Lindemann", "Troe", "PdepArrhenius") for t in types[ig.nvec:])
    assert ig.veckineticsinds[-1] == ig.nvec
    assert [r.index for r in ig.reactions] == list(range(1, ig.Nr + 1))
    eff = [r for r in ig.reactions if r.kinetics.get("efficiencies")]
    assert len(eff) == 7 and all(v >= -1.0 for r in eff for v in r.kinetics["efficiencies"].values())   # stored as eff-1


def test_kinetics_packing_layout():
    a = {"type": "Arrhenius", "A": 1.0, "n": 2.0, "Ea": 3.0}
    assert ph.pack_kinetics(a, True) == (L.KIN_ARRHENIUS_VEC, [1.0, 2.0, 3.0])
    assert ph.pack_kinetics(a, False) == (L.KIN_ARRHENIUS, [1.0, 2.0, 3.0])
    pl = {"type": "PdepArrhenius", "Ps": [1e4, 1e6], "arrs": [a, {"type": "MultiArrhenius", "arrs": [a, a]}]}
    assert ph.pack_kinetics(pl, False) == (L.KIN_PLOG, [2.0, 1e4, 1e6, 1.0, 1.0, 2.0, 3.0, 2.0, 1.0, 2.0, 3.0, 1.0, 2.0, 3.0])
    tr = {"type": "Troe", "arrhigh": a, "arrlow": a, "a": 0.5, "T3": 10.0, "T1": 20.0, "T2": 0.0, "efficiencies": {}}
    assert ph.pack_kinetics(tr, False) == (L.KIN_TROE, [1.0, 2.0, 3.0, 1.0, 2.0, 3.0, 0.5, 10.0, 20.0, 0.0])
    with pytest.raises(NotImplementedError):
        ph.pack_kinetics({"type": "Arrheniusq"}, False)


def test_molecule_counts():
    from rmsb200.mechanism import getspeciesradius, molecule_counts
    assert molecule_counts("CCCCCCCC")[:2] == ({"C": 8, "H": 18}, 25)
    assert molecule_counts("[O]O")[:2] == ({"O": 2, "H": 1}, 2)
    assert molecule_counts("CCCCC(CCC)O[O]")[:2] == ({"C": 8, "H": 17, "O": 2}, 26)
    assert molecule_counts("[H][H]")[:2] == ({"H": 2}, 1)
    assert molecule_counts("O=C=O")[2] == pytest.approx(0.044009, rel=1e-4)
    with pytest.raises(KeyError):
        molecule_counts("[Pt]")
    assert getspeciesradius({"C": 8, "H": 18}, 25) == pytest.approx(3.659e-10, rel=1e-3)


def test_domain_constructors_host_side():
    d = rb.readinput(cases.fixture("superminimal"))
    ig = rb.IdealGas(d["phase"]["Species"], d["phase"]["Reactions"], name="phase")
    dom, y0, p = rb.ConstantVDomain(phase=ig, initialconds={"T": 1000.0, "P": 10e5, "H2": 0.67, "O2": 0.33})
    assert y0.shape == (14,) and p.shape == (29,)
    assert dom.V == pytest.approx(8.314 * 1000.0 / 10e5)                  # V = N*R*T/P  (src/Domain.jl:143-145)
    assert dom.thermovariabledict == {"T": 13, "P": 14} and dom.indexes == [1, 12, 13, 14]
    with pytest.raises(ValueError):
        rb.ConstantVDomain(phase=ig, initialconds={"T": 1000.0, "P": 1e5, "V": 1.0, "H2": 1.0})
    with pytest.raises(AssertionError):
        rb.ConstantVDomain(phase=ig, initialconds={"T": 1000.0, "P": 1e5, "XYZ": 1.0})
    ens = rb.ConstantVEnsemble(ig, [1000.0, 1200.0], [1e5, 2e5], np.array([[0, 0, 0, 0, 0.67, 0.33] + [0] * 6] * 2))
    assert ens.n == 2 and ens.cond.shape == (2, 1) and ens.y0.shape == (2, 14)


def test_product_never_touches_the_oracle():
    """oracle/ is test infrastructure: nothing under the product package may import, load or execute it."""
    pkg = cases.ROOT / "reactionmechanismsimulator.jl_b200"
    for path in list(pkg.rglob("*.py")) + list(pkg.rglob("*.cu")) + list(pkg.rglob("*.cuh")) + list(pkg.rglob("*.jl")):
        text = path.read_text()
        assert "oracle" not in text, f"{path} mentions the oracle"


def test_coupled_descriptor_host_side():
    """C4: the merged two-phase descriptor (rmsb200/coupled.py) -- species order, reaction blocks, sticking -> Arrhenius."""
    from oracle import coupled as oc
    from rmsb200.coupled import CoupledPhase, stickingcoefficient2arrhenius
    d = rb.readinput(cases.fixture("ch4o2cat"))
    ig = rb.IdealGas(d["gas"]["Species"], d["gas"]["Reactions"], name="gas")
    cat = rb.IdealSurface(d["surface"]["Species"], d["surface"]["Reactions"], 2.486e-5, name="surface")
    cp = CoupledPhase(ig, cat, d[frozenset(("gas", "surface"))]["Reactions"])
    assert (cp.Ns, cp.Nr, cp.n_gas_rxn, cp.n_surf_rxn, cp.n_iface_rxn) == (41, 222, 3, 198, 21)
    assert list(cp._species_phase) == [0] * 12 + [1] * 29 and list(cp._rxn_domain) == [0] * 3 + [1] * 198 + [2] * 21
    osys = oc.ch4o2cat_system(cases.fixture("ch4o2cat"))
    assert cp.names == osys.names
    np.testing.assert_array_equal(cp.rxnarray[-21:] - 1, osys.irxnarray.T)
    for r_p, r_o in zip(cp.reactions[-21:], osys.irxns):                      # closed-form sticking conversion agrees
        assert r_p.kinetics["type"] == "Arrhenius"
        for key in ("A", "n", "Ea"):
            assert r_p.kinetics[key] == pytest.approx(r_o.kinetics[key], rel=1e-14)
    with pytest.raises(NotImplementedError):
        stickingcoefficient2arrhenius({"A": 5.0, "n": 0.0, "Ea": 0.0}, 2.486e-5, 1, 0.016)
    # molecular weights come from the SMILES (RDKit MolWt in the reference, src/Parse.jl:143)
    mw = {s.name: s.molecularweight for s in ig.species}
    assert mw["CH4"] == pytest.approx(0.016043, rel=1e-4) and mw["CO2"] == pytest.approx(0.044009, rel=1e-4)


def test_transitory_time_scale_selection_matches_the_oracle_restatement():
    """rmsb200.transitory.getdividingtimescale (host post-processing of the batched Jacobians) against the oracle's independent
    restatement of src/TransitorySensitivities.jl:5-113 on random bimodal time-scale spectra."""
    from oracle import transitory as tr
    from rmsb200 import transitory as pt
    rng = np.random.default_rng(0)
    seen = 0
    for _ in range(100):
        n = int(rng.integers(20, 80))
        d = -10.0 ** np.concatenate([rng.normal(8, 1, n // 2), rng.normal(0, 1.5, n - n // 2)])
        J = np.diag(d)
        try:
            a = tr.getdividingtimescale(J)
        except ValueError:
            a = None
        try:
            b = pt.getdividingtimescale(J)
        except ValueError:
            b = None
        assert a == b
        seen += a is not None
    assert seen > 50


def test_pack_interfaces_merges_by_type_and_matches_oracle_enthalpy():
    """Host packing of boundary interfaces (rmsb200/interfaces.py): two inlets merge into one row (flow-weighted composition and
    enthalpy), the inlet enthalpy equals the oracle's getEnthalpy sum, per-reactor flows override the row."""
    from types import SimpleNamespace
    d = rb.readinput(cases.fixture("superminimal"))
    gdom = SimpleNamespace(phase=rb.IdealGas(d["phase"]["Species"], d["phase"]["Reactions"], name="phase"), kind=L.DOMAIN_CONSTANT_TP)   # no device needed
    odom = cases.oracle_domain("superminimal_TP")
    a = rb.Inlet(gdom, {"H2": 0.67, "O2": 0.33, "T": 1000.0, "P": 1e5}, lambda t: 0.001)
    b = rb.Inlet(gdom, {"H2": 1.0, "T": 800.0, "P": 1e5}, 0.003)
    oa = ro.Inlet(odom, {"H2": 0.67, "O2": 0.33, "T": 1000.0, "P": 1e5}, 0.001)
    assert a.H == pytest.approx(oa.H, rel=1e-14) and np.allclose(a.y, oa.y, rtol=1e-15)
    row = rb.pack_interfaces(gdom, [a, b, rb.Outlet(gdom, 0.004), rb.VolumeMaintainingOutlet(gdom)])
    Ns = gdom.phase.Ns
    assert row.shape == (7 + 2 * Ns,) and row[0] == pytest.approx(0.004) and row[1] == 0.004 and row[4] == 1.0
    np.testing.assert_allclose(row[7:7 + Ns] * row[0], a.y * a.F + b.y * b.F, rtol=1e-14)
    assert row[5] * row[0] == pytest.approx(a.H * a.F + b.H * b.F, rel=1e-14)
    rows = rb.pack_interfaces(gdom, [a], np.array([[1e-3, 2e-3, 0, 0], [5e-3, 0, 0, 0]]))
    assert rows.shape == (2, 7 + 2 * Ns) and rows[1, 0] == 5e-3 and rows[0, 1] == 2e-3
    with pytest.raises(NotImplementedError):
        rb.Outlet(gdom, lambda t: 1e-3 * (1.0 + t))


@pytest.mark.parametrize("fx", ["superminimal", "ethane", "liquid_phase", "ch4o2cat"])
def test_rms_writer_round_trip(fx, tmp_path):
    """writeyml (src/yml.jl:14-16) round trip: readinput -> getmechdict -> writeyml -> readinput gives the same species, thermo,
    reactions and kinetics (compared through the flattened descriptor tables the device receives)."""
    a = rb.readinput(cases.fixture(fx))
    path = str(tmp_path / f"{fx}_copy.rms")
    rb.writeyml(rb.getmechdict(a), path)
    b = rb.readinput(path)
    assert set(map(str, a.keys())) == set(map(str, b.keys()))
    for key in a:
        if key == "Solvents":
            assert [(s.name, s.mu) for s in a[key]] == [(s.name, s.mu) for s in b[key]]
            continue
        ra, rbb = a[key]["Reactions"], b[key]["Reactions"]
        assert [str(r) for r in ra] == [str(r) for r in rbb]
        for x, y in zip(ra, rbb):
            tx, px = ph.pack_kinetics(x.kinetics, False) if x.kinetics["type"] != "StickingCoefficient" else (0, [x.kinetics["A"], x.kinetics["n"], x.kinetics["Ea"]])
            ty, py = ph.pack_kinetics(y.kinetics, False) if y.kinetics["type"] != "StickingCoefficient" else (0, [y.kinetics["A"], y.kinetics["n"], y.kinetics["Ea"]])
            assert tx == ty and px == py and x.kinetics.get("efficiencies") == y.kinetics.get("efficiencies")
            assert (x.reversible, x.forwardable) == (y.reversible, y.forwardable)
        if "Species" in a[key]:
            for sx, sy in zip(a[key]["Species"], b[key]["Species"]):
                assert sx.name == sy.name and sx.smiles == sy.smiles and sx.radius == sy.radius and sx.diffusion_r == sy.diffusion_r
                assert [(q.coefs, q.Tmin, q.Tmax) for q in sx.thermo.polys] == [(q.coefs, q.Tmin, q.Tmax) for q in sy.thermo.polys]


def test_generated_kernel_source_compiles_for_sm100a(tmp_path):
    """csrc/rmsb_spec.cu: the CUDA source the library writes out for a small all-Arrhenius gas mechanism comes from the descriptor
    alone (no device), holds both kernels, compiles for sm_100a, and mechanisms outside that class are refused by name."""
    import shutil
    import subprocess
    d = rb.readinput(cases.fixture("superminimal"))
    ig = rb.IdealGas(d["phase"]["Species"], d["phase"]["Reactions"], name="phase")
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    for kind in (L.DOMAIN_CONSTANT_V, L.DOMAIN_CONSTANT_P):
        src = ig.spec_source(kind)
        assert "rmsb_spec_dydt" in src and "rmsb_spec_jac" in src
        assert src.count("// reaction ") == ig.Nr                # value pass and dual-number pass share one template
        f = tmp_path / f"spec_{kind}.cu"
        f.write_text(src)
        r = subprocess.run([nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-std=c++17", "-cubin", "-o", str(f.with_suffix(".cubin")), str(f)],
                           capture_output=True, text=True)
        assert r.returncode == 0, r.stderr[-2000:]
    with pytest.raises(L.RMSBError, match="ConstantV / ConstantP"):
        ig.spec_source(L.DOMAIN_CONSTANT_TP)
    d = rb.readinput(cases.fixture("ethane"))                        # falloff kinetics: generic kernels only
    eg = rb.IdealGas(d["phase"]["Species"], d["phase"]["Reactions"], name="phase")
    with pytest.raises(L.RMSBError, match="does not qualify"):
        eg.spec_source(L.DOMAIN_CONSTANT_V)
