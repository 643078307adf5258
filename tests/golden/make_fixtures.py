"""Regenerates tests/fixtures/*.rms from the reference's mechanism DATA files.

The `.rms` files under /root/reference/src/testing are YAML mechanism data (species thermo, reaction
kinetics) -- the inputs BASELINE.json's configs name.  They are re-serialised here (sorted keys,
normalised flow style) so the parity tests and the bench can run on the GPU box, where /root/reference
does not exist.  Run in the build container:  python tests/golden/make_fixtures.py
"""
import sys
from pathlib import Path

import yaml

SRC = Path("/root/reference/src/testing")
DST = Path(__file__).resolve().parent.parent / "fixtures"
NAMES = ["superminimal", "superminimal_broken", "minimal", "ethane", "liquid_phase", "ch4o2cat", "propane", "minimal_edge"]

if __name__ == "__main__":
    DST.mkdir(exist_ok=True)
    for n in NAMES:
        doc = yaml.safe_load((SRC / f"{n}.rms").read_text())
        (DST / f"{n}.rms").write_text(yaml.safe_dump(doc, sort_keys=True, default_flow_style=None, width=160))
        print(n, (DST / f"{n}.rms").stat().st_size)
