"""Host logic of the large-mechanism path (csrc/rmsb_sparse.h): minimum-degree ordering, filled pattern, dependency levels and
dense core, checked on the CPU against a dense LU -- no GPU involved."""
import subprocess

import numpy as np
import pytest

import cases
import synth
from oracle import rms_oracle as ro

DRIVER = cases.ROOT / "tests" / "cpp" / "sparse_driver.cpp"


@pytest.fixture(scope="module")
def analysed(tmp_path_factory):
    tmp = tmp_path_factory.mktemp("sparse")
    exe = tmp / "sparse_driver"
    subprocess.run(["g++", "-O2", "-std=c++17", "-o", str(exe), str(DRIVER)], check=True)
    ns = 200
    path = str(tmp / "synth.rms")
    synth.write_synthetic_mechanism(path, ns=ns, seed=11)
    d = ro.readinput(path)
    ph = ro.IdealDiluteSolution(d["phase"]["Species"], d["phase"]["Reactions"], d["Solvents"][0], diffusionlimited=False)
    ra = ph.rxnarray
    ent = set()
    for j in range(ra.shape[1]):          # Jacobian entries of mass action: every species of a reaction couples to every other
        sl = [int(ra[k, j]) for k in range(8) if ra[k, j] >= 0]
        ent.update((a, b) for a in sl for b in sl)
    ent = sorted(ent)
    inp = f"{ns} {len(ent)}\n" + "\n".join(f"{a} {b}" for a, b in ent) + "\n"
    out = subprocess.run([str(exe)], input=inp, capture_output=True, text=True, check=True).stdout.splitlines()
    head = out[0].split()
    res = {"n": int(head[1]), "core": int(head[3]), "max_row": int(head[5]), "flops": int(head[7])}
    for line in out[1:]:
        p = line.split()
        res[p[0]] = np.array(p[2:], dtype=np.int64)
        assert len(res[p[0]]) == int(p[1])
    return ns, ent, res


def test_permutation_and_pattern_contain_the_jacobian(analysed):
    ns, ent, r = analysed
    assert sorted(r["perm"]) == list(range(ns)) and np.array_equal(r["iperm"][r["perm"]], np.arange(ns))
    rows = [set(r["col"][r["row_ptr"][i]:r["row_ptr"][i + 1]]) for i in range(ns)]
    for i in range(ns):
        cols = r["col"][r["row_ptr"][i]:r["row_ptr"][i + 1]]
        assert np.all(np.diff(cols) > 0) and r["col"][r["diag"][i]] == i
    for a, b in ent:
        assert int(r["iperm"][b]) in rows[int(r["iperm"][a])]
    assert r["max_row"] == max(len(s) for s in rows)


def test_fill_pattern_is_closed_under_elimination(analysed):
    """LU without pivoting of a diagonally dominant matrix on the Jacobian pattern: every nonzero of the dense factors lies
    inside the analysed pattern, and the flop count matches."""
    ns, ent, r = analysed
    rng = np.random.default_rng(0)
    A = np.zeros((ns, ns))
    for a, b in ent:
        A[r["iperm"][a], r["iperm"][b]] = rng.uniform(-1.0, 1.0)
    A += np.diag(np.abs(A).sum(axis=1) + 1.0)
    LU = A.copy()
    for k in range(ns):
        LU[k + 1:, k] /= LU[k, k]
        LU[k + 1:, k + 1:] -= np.outer(LU[k + 1:, k], LU[k, k + 1:])
    inside = np.zeros((ns, ns), dtype=bool)
    for i in range(ns):
        inside[i, r["col"][r["row_ptr"][i]:r["row_ptr"][i + 1]]] = True
    assert np.all(np.abs(LU[~inside]) < 1e-13)
    assert inside.sum() == len(r["col"]) and np.array_equal(inside, inside.T)
    nk = np.array([r["row_ptr"][k + 1] - r["diag"][k] - 1 for k in range(ns)])
    assert r["flops"] == int((nk * nk).sum())


def test_levels_and_dense_core(analysed):
    ns, ent, r = analysed
    lv = np.empty(ns, dtype=int)
    for l in range(len(r["lvl_l_ptr"]) - 1):
        lv[r["lvl_l_rows"][r["lvl_l_ptr"][l]:r["lvl_l_ptr"][l + 1]]] = l
    lu = np.empty(ns, dtype=int)
    for l in range(len(r["lvl_u_ptr"]) - 1):
        lu[r["lvl_u_rows"][r["lvl_u_ptr"][l]:r["lvl_u_ptr"][l + 1]]] = l
    for i in range(ns):
        cols = r["col"][r["row_ptr"][i]:r["row_ptr"][i + 1]]
        assert np.all(lv[cols[cols < i]] < lv[i]) and np.all(lu[cols[cols > i]] < lu[i])
    core = r["core"]
    assert 0 < core < ns
    for i in range(core, ns):      # every core row ends with all core columns: the dense kernels address them by arithmetic
        cols = r["col"][r["row_ptr"][i]:r["row_ptr"][i + 1]]
        assert np.array_equal(cols[-(ns - core):], np.arange(core, ns))
