"""CPU oracle for the RMS hot path -- TEST INFRASTRUCTURE ONLY.

Nothing under ``oracle/`` is part of the shipped product.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference``
legs may import or execute it, and only as the checker / CPU comparator.

Parity status: the reference (Julia, ReactionMechanismSimulator.jl v1.0.1) cannot be run
in this environment (no julia, no network).  The oracle is therefore pinned against the
reference's OWN golden values (src/TestParse.jl, src/Calculators/TestRate.jl,
src/Calculators/TestThermo.jl, src/TestReactors.jl) -- see tests/test_oracle_golden.py.
Integrator internals (CVODE step sequence) are parity-unpinned; only converged solution
values are pinned, exactly as in the reference's own tests.

Files: rms_oracle.py (NumPy restatement of every first-party routine on the path, all five domains), integrate.py (SciPy
trajectories), cpu/rms_cpu.cpp + cpu_oracle.py (C++ port: fast checker and CPU baseline), coupled.py (gas + surface + interface
system, config C4), sensitivity.py (augmented forward-sensitivity system, config C5; absolute values parity-unpinned, the
reference pins none), transitory.py (transitory sensitivities: the reference's 160-value known-answer test for jacobiany /
jacobianp, tests/test_oracle_constp.py).
"""
