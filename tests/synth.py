"""Synthetic large mechanisms for config C3 (~1000+ species, sparse Jacobian).  The reference ships no mechanism of that
size (the octane fixture has 35 species), so the parity and throughput cases use a seeded generator
with the structure of a hydrocarbon oxidation mechanism: a small pool of 'hub' species (radicals / small molecules that
take part in most reactions) and a long tail of large molecules that only react with the hubs and with their neighbours.
The file written is an ordinary `.rms` YAML mechanism, read by both the product (`rmsb200.readinput`) and the oracle."""
from __future__ import annotations

import numpy as np

N_HUB = 16


def write_synthetic_mechanism(path: str, ns: int = 1000, rxn_per_species: float = 4.0, seed: int = 7) -> dict:
    rng = np.random.default_rng(seed)
    names = [f"S{i:04d}" for i in range(ns)]
    lines = ["Phases:", "- Species:"]
    # generations (carbon number) of geometrically growing size; each generation adds a CH2-like increment to cp, H and S
    n_gen, fam = 10, 8
    w = 1.45 ** np.arange(n_gen)
    bounds = N_HUB + np.concatenate([[0], np.cumsum(np.round((ns - N_HUB) * w / w.sum()).astype(int))])
    bounds[-1] = ns
    gen_of = np.zeros(ns, dtype=int)
    for g in range(n_gen):
        gen_of[bounds[g]:bounds[g + 1]] = g
    for i, nm in enumerate(names):
        if i < N_HUB:    # small radicals / molecules: high enthalpy, low entropy (additions to them are exothermic)
            a0, a1, a2 = rng.uniform(3.5, 4.5), 5e-4, -5e-8
            a5, a6 = rng.uniform(1.0e4, 2.0e4), rng.uniform(3.0, 8.0)
        else:
            g = gen_of[i]
            a0, a1, a2 = 4.0 + 1.5 * g + rng.uniform(-0.2, 0.2), 1e-3 * (1 + g), -1e-7 * (1 + g)
            a5, a6 = -2.5e3 * g + rng.uniform(-1.0e3, 1.0e3), 10.0 + 1.0 * g + rng.uniform(-1.0, 1.0)
        coefs = "[" + ", ".join(repr(float(v)) for v in (a0, a1, a2, 0.0, 0.0, a5, a6)) + "]"
        lines += [f"  - name: {nm}", "    radicalelectrons: 0", "    thermo:", "      polys:",
                  f"      - {{Tmax: 1000.0, Tmin: 100.0, coefs: {coefs}, type: NASApolynomial}}",
                  f"      - {{Tmax: 5000.0, Tmin: 1000.0, coefs: {coefs}, type: NASApolynomial}}",
                  "      type: NASA", "    type: Species"]
    lines += ["  name: phase", "Reactions:"]
    nr = int(rxn_per_species * ns)
    seen = set()
    count = 0
    # generations (carbon number) of geometrically growing size, split into families of isomers; a family reacts within
    # itself, with the hubs, and with two parent families of the previous generation (decomposition / recombination)
    families = []                      # (generation, [species])
    for g in range(n_gen):
        idx = np.arange(bounds[g], bounds[g + 1])
        for k in range(0, len(idx), fam):
            families.append((g, idx[k:k + fam]))
    by_gen = [[k for k, (g, _) in enumerate(families) if g == gg] for gg in range(n_gen)]
    parents = []
    for k, (g, _) in enumerate(families):
        parents.append([] if g == 0 or not by_gen[g - 1] else list(rng.choice(by_gen[g - 1], size=min(2, len(by_gen[g - 1])), replace=False)))
    fam_of = np.zeros(ns, dtype=int)
    for k, (_, idx) in enumerate(families):
        fam_of[idx] = k

    def same(i):
        return int(rng.choice(families[fam_of[i]][1]))

    def parent(i):
        ps = parents[fam_of[i]]
        if not ps:
            return int(rng.integers(0, N_HUB))
        return int(rng.choice(families[int(rng.choice(ps))][1]))

    def hub():
        return int(rng.integers(0, N_HUB))

    # every reaction conserves 'mass' (hubs weigh 1, generation g weighs g + 2), so the total amount stays bounded
    while count < nr:
        kind = rng.integers(0, 10)
        i = int(rng.integers(N_HUB, ns))
        if kind < 4:        # abstraction within the family: X_i + hub_a <=> X_j + hub_b
            re, pr = [i, hub()], [same(i), hub()]
            bimol = True
        elif kind < 7:      # decomposition to the parent generation: X_i <=> X_parent + hub   (generation 0: two hubs)
            re, pr = [i], [parent(i), hub()]
            bimol = False
        elif kind < 9:      # hub-assisted decomposition: X_i + hub <=> X_parent + hub + hub
            re, pr = [i, hub()], [parent(i), hub(), hub()]
            bimol = True
        else:               # isomerisation within the family
            re, pr = [i], [same(i)]
            bimol = False
        if sorted(re) == sorted(pr) or i in pr:
            continue
        key = (tuple(sorted(re)), tuple(sorted(pr)))
        if key in seen:
            continue
        seen.add(key)
        A = 10.0 ** (rng.uniform(0.0, 4.0) if bimol else rng.uniform(5.0, 10.0))
        n = float(rng.uniform(0.0, 0.5))
        Ea = float(rng.uniform(2.0e4, 1.2e5))
        lines += [f"- kinetics: {{A: {A!r}, Ea: {Ea!r}, n: {n!r}, type: Arrhenius}}",
                  f"  products: [{', '.join(names[k] for k in pr)}]", "  radicalchange: 0",
                  f"  reactants: [{', '.join(names[k] for k in re)}]", "  type: ElementaryReaction"]
        count += 1
    lines += ["Solvents:", "- mu: {mu: 0.0005, type: ConstantViscosity}", "  name: solvent", "  type: Solvent", "Units: {}"]
    with open(path, "w") as fh:
        fh.write("\n".join(lines) + "\n")
    return {"names": names, "n_hub": N_HUB, "nr": nr}


def initial_moles(ns: int, seed: int = 3) -> dict:
    """A 'fuel' (a heavy species), an 'oxidiser' hub and a trace of everything else left at zero."""
    rng = np.random.default_rng(seed)
    ic = {f"S{ns - 1:04d}": 6000.0, f"S{ns // 2:04d}": 500.0, "S0000": 5.0, "S0001": 50.0}
    for k in rng.integers(N_HUB, ns, 20):
        ic[f"S{int(k):04d}"] = float(rng.uniform(0.1, 10.0))
    return ic


def species_weights(ns: int) -> np.ndarray:
    """The 'mass' every generated reaction conserves: hubs weigh 1, generation g weighs g + 2."""
    n_gen = 10
    w = 1.45 ** np.arange(n_gen)
    bounds = N_HUB + np.concatenate([[0], np.cumsum(np.round((ns - N_HUB) * w / w.sum()).astype(int))])
    bounds[-1] = ns
    out = np.ones(ns)
    for g in range(n_gen):
        out[bounds[g]:bounds[g + 1]] = g + 2.0
    return out


# ------------------------------------------------------------------------------------------------------------------
# SURVEY.md section 8(d) recipe for config C3 at its stated size: the reference's own 35-species liquid octane oxidation
# mechanism (tests/fixtures/liquid_phase.rms, src/TestReactors.jl:7-27) replicated over K homologous "fuel families" that
# share the small-species pool (inert gases, O2, HO2, OH).  Family species keep the template's SMILES (so radii, diffusivities
# and thermo are the template's) under a family-prefixed name; every family reaction is the template's with a seeded rate
# perturbation A * 10^U(-0.3, 0.3).  K = 36, keep = 0.85  ->  Ns = 7 + 36*28 = 1015 species, Nr ~ 11.6 k reactions.
# ------------------------------------------------------------------------------------------------------------------
SHARED = ("Ar", "He", "Ne", "N2", "oxygen", "[O]O", "OH(D)")


def _flow(v):
    """YAML flow scalar for a name (SMILES need quoting)."""
    return "'" + v.replace("'", "''") + "'"


def write_octane_families(path: str, families: int = 36, keep: float = 0.85, seed: int = 20261019, template: str | None = None,
                          falloff: bool = False) -> dict:
    """falloff=True (for the gas-phase variant of the same mechanism, read as an IdealGas): the association reactions R + O2 -> ROO of
    every family (six per family in the template) become, in turn, Troe, Lindemann and ThirdBody reactions with third-body
    efficiencies on pool species and on the family's fuel, so the large-mechanism path sees falloff kinetics and efficiencies."""
    import yaml
    from pathlib import Path
    tpl = template or str(Path(__file__).resolve().parent / "fixtures" / "liquid_phase.rms")
    with open(tpl) as fh:
        doc = yaml.load(fh, Loader=getattr(yaml, "CSafeLoader", yaml.SafeLoader))
    rng = np.random.default_rng(seed)
    species = doc["Phases"][0]["Species"]
    shared = [s for s in species if s["name"] in SHARED]
    own = [s for s in species if s["name"] not in SHARED]
    fam_name = lambda k, nm: f"F{k:02d}_{nm}"
    lines = ["Phases:", "- Species:"]

    def emit_species(sp, name):
        lines.append(f"  - name: {_flow(name)}")
        lines.append(f"    radicalelectrons: {sp.get('radicalelectrons', 0)}")
        lines.append(f"    smiles: {_flow(sp['smiles'])}")
        lines.append("    thermo:")
        lines.append("      polys:")
        for pl in sp["thermo"]["polys"]:
            coefs = "[" + ", ".join(repr(float(c)) for c in pl["coefs"]) + "]"
            lines.append(f"      - {{Tmax: {float(pl['Tmax'])!r}, Tmin: {float(pl['Tmin'])!r}, coefs: {coefs}, type: NASApolynomial}}")
        lines.append("      type: NASA")
        lines.append("    type: Species")

    names = []
    for sp in shared:
        emit_species(sp, sp["name"]); names.append(sp["name"])
    for k in range(families):
        for sp in own:
            emit_species(sp, fam_name(k, sp["name"])); names.append(fam_name(k, sp["name"]))
    lines += ["  name: phase", "Reactions:"]
    nr = 0

    n_assoc = [0]

    def emit_reaction(r, k, scale):
        nonlocal nr
        ren = lambda nm: nm if (nm in SHARED or k is None) else fam_name(k, nm)
        kin = r["kinetics"]
        assert kin["type"] == "Arrhenius"
        A, Ea, n = float(kin["A"]) * scale, float(kin["Ea"]), float(kin["n"])
        if falloff and k is not None and len(r["reactants"]) == 2 and len(r["products"]) == 1:
            form = n_assoc[0] % 3
            n_assoc[0] += 1
            eff = "{" + ", ".join(f"{_flow(a)}: {b!r}" for a, b in (("N2", 1.4), ("Ar", 0.7), ("oxygen", 1.2), (fam_name(k, "octane"), 3.0))) + "}"
            hi = f"{{A: {A!r}, Ea: {Ea!r}, n: {n!r}, type: Arrhenius}}"
            lo = f"{{A: {A * 2.0e1!r}, Ea: {Ea * 0.5!r}, n: {n - 1.0!r}, type: Arrhenius}}"
            if form == 0:
                lines.extend(["- kinetics:", "    T1: 1500.0", "    T2: 5000.0", "    T3: 100.0", "    a: 0.6", f"    arrhigh: {hi}", f"    arrlow: {lo}",
                              f"    efficiencies: {eff}", "    type: Troe"])
            elif form == 1:
                lines.extend(["- kinetics:", f"    arrhigh: {hi}", f"    arrlow: {lo}", f"    efficiencies: {eff}", "    type: Lindemann"])
            else:
                lines.extend(["- kinetics:", f"    arr: {lo}", f"    efficiencies: {eff}", "    type: ThirdBody"])
        else:
            lines.append(f"- kinetics: {{A: {A!r}, Ea: {Ea!r}, n: {n!r}, type: Arrhenius}}")
        lines.append("  products: [" + ", ".join(_flow(ren(x)) for x in r["products"]) + "]")
        lines.append(f"  radicalchange: {r.get('radicalchange', 0)}")
        lines.append("  reactants: [" + ", ".join(_flow(ren(x)) for x in r["reactants"]) + "]")
        lines.append("  type: ElementaryReaction")
        nr += 1

    for r in doc["Reactions"]:
        if all(x in SHARED for x in list(r["reactants"]) + list(r["products"])):
            emit_reaction(r, None, 1.0)                    # pool-only chemistry: once
    for k in range(families):
        for r in doc["Reactions"]:
            if all(x in SHARED for x in list(r["reactants"]) + list(r["products"])):
                continue
            if rng.uniform() > keep:
                continue
            emit_reaction(r, k, 10.0 ** rng.uniform(-0.3, 0.3))
    mu = doc["Solvents"][0]["mu"]
    lines += ["Solvents:", "- mu: {" + ", ".join(f"{a}: {mu[a]!r}" for a in mu) + "}", f"  name: {doc['Solvents'][0]['name']}", "  type: Solvent", "Units: {}"]
    with open(path, "w") as fh:
        fh.write("\n".join(lines) + "\n")
    return {"names": names, "families": families, "nr": nr, "ns": len(names)}


def octane_family_moles(families: int = 36, seed: int = 5) -> dict:
    """Initial amounts: the reference's C3 charge (octane 6154 mol, oxygen 4.953 mol held constant, src/TestReactors.jl:18-22) with
    the fuel split unevenly over the families."""
    rng = np.random.default_rng(seed)
    w = rng.uniform(0.5, 1.5, families)
    w /= w.sum()
    ic = {f"F{k:02d}_octane": 6154.0 * float(w[k]) for k in range(families)}
    ic["oxygen"] = 4.953
    return ic
