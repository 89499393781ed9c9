#!/usr/bin/env python
"""Copies the potential PARAMETER files the reference ships (data, not source) into tests/golden/potentials/ as .gz,
byte for byte, so that the GPU box (which has no /root/reference) can run the parity tests and bench.py on the very
files BASELINE.md / SURVEY.md 8(d) name.  Run in the build container:  python tests/golden/make_potential_fixtures.py

  Cu.eam.alloy               scripts/python/pytab-eam-alloy/Cu.eam.alloy           (configs[1]/[3] potential, rc 7.29)
  AlCu.eam.alloy             data/potentials/AlCu.eam.alloy                        (configs[4]; two species, rc 6.6825)
  Ta1_Ravelo_2013.eam.alloy  data/potentials/Ta1_Ravelo_2013.eam.alloy
  WBe_Wood_PRB2019.snap*     data/regression_new/potentials/snap/                  (2J = 8, two elements; configs[2] uses the W block)
  Ta06A.snap*                data/regression_new/potentials/snap/                  (2J = 6)
"""
import gzip
import hashlib
import os
import shutil

REF = os.environ.get("XS_REFERENCE", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))
FILES = {
    "Cu.eam.alloy": "scripts/python/pytab-eam-alloy/Cu.eam.alloy",
    "AlCu.eam.alloy": "data/potentials/AlCu.eam.alloy",
    "Ta1_Ravelo_2013.eam.alloy": "data/potentials/Ta1_Ravelo_2013.eam.alloy",
    "WBe_Wood_PRB2019.snapcoeff": "data/regression_new/potentials/snap/WBe_Wood_PRB2019.snapcoeff",
    "WBe_Wood_PRB2019.snapparam": "data/regression_new/potentials/snap/WBe_Wood_PRB2019.snapparam",
    "Ta06A.snapcoeff": "data/regression_new/potentials/snap/Ta06A.snapcoeff",
    "Ta06A.snapparam": "data/regression_new/potentials/snap/Ta06A.snapparam",
}

if __name__ == "__main__":
    out = os.path.join(HERE, "potentials")
    os.makedirs(out, exist_ok=True)
    sums = []
    for name, rel in FILES.items():
        src = os.path.join(REF, rel)
        raw = open(src, "rb").read()
        with open(os.path.join(out, name + ".gz"), "wb") as f:
            with gzip.GzipFile(filename="", mode="wb", fileobj=f, mtime=0, compresslevel=9) as g:      # reproducible bytes
                g.write(raw)
        sums.append("%s  %s  %d bytes  <- %s" % (hashlib.sha256(raw).hexdigest(), name, len(raw), rel))
    open(os.path.join(out, "SHA256SUMS"), "w").write("\n".join(sums) + "\n")
    print("\n".join(sums))
