#!/usr/bin/env python
"""Extracts the MRT moment-transform literals of the reference into tests/golden/mrt_tables.npz.

Run in the build container (needs /root/reference):  python tests/golden/make_mrt_golden.py
Source: src/library/natrium/collision_advanced/AuxiliaryMRTFunctions.cpp -- the `moment_trafo` /
`inverse_trafo` initialisers of MRTDellarD2Q9 (:15-45), MRTLallemandD2Q9 (:51-80) and MRTDHumieresD3Q19
(:86-205).  Every entry there is a rational literal `a. / b.`; they are evaluated in double precision exactly as
the C++ compiler would.  The fixture pins oracle/ and natrium_b200/mrt.py (which rebuild the bases from their
definitions) to the reference's own numbers; nothing at test time reads /root/reference.
"""
import os
import re

import numpy as np

SRC = "/root/reference/src/library/natrium/collision_advanced/AuxiliaryMRTFunctions.cpp"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "mrt_tables.npz")


def main():
    txt = open(SRC).read()
    txt = re.sub(r"//[^\n]*", "", txt)
    out = {}
    for m in re.finditer(r"const\s+array<array<double,\s*(\d+)>,\s*\1>\s+(\w+)::(\w+)\s*=\s*(.*?);", txt, flags=re.S):
        q, cls, name, body = int(m.group(1)), m.group(2), m.group(3), m.group(4)
        vals = [float(a) / float(b) for a, b in re.findall(r"(-?\d+\.?\d*)\s*/\s*(\d+\.?\d*)", body)]
        assert len(vals) == q * q, (cls, name, len(vals))
        out[f"{cls}_{name}"] = np.array(vals).reshape(q, q)
    assert sorted(out) == sorted(f"{c}_{n}" for c in ("MRTDellarD2Q9", "MRTLallemandD2Q9", "MRTDHumieresD3Q19")
                                 for n in ("moment_trafo", "inverse_trafo")), sorted(out)
    for c in ("MRTDellarD2Q9", "MRTLallemandD2Q9", "MRTDHumieresD3Q19"):
        M, T = out[f"{c}_moment_trafo"], out[f"{c}_inverse_trafo"]
        err = np.max(np.abs(M @ T - np.eye(M.shape[0])))
        print(c, M.shape, "max |M T - I| =", err)
        assert err < 1e-13
    np.savez(OUT, **out)
    print("wrote", OUT)


if __name__ == "__main__":
    main()
