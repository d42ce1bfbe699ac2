#!/usr/bin/env python
"""Extracts the pseudo-entropic stabilizer matrices of the reference into tests/golden/stabilizer_tables.npz.

Run in the build container (needs /root/reference):  python tests/golden/make_stabilizer_golden.py
Source: src/library/natrium/dataprocessors/PseudoEntropicStabilizer.cpp -- `n` / `d` (D2Q9: entry = n/d, :27-40),
`nd_d2q9_with_e` (:42-57) and `nd_d3q19` (:59-150), rational literals evaluated in double precision.
"""
import os
import re

import numpy as np

SRC = "/root/reference/src/library/natrium/dataprocessors/PseudoEntropicStabilizer.cpp"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "stabilizer_tables.npz")


def block(txt, name):
    m = re.search(r"double\s+" + name + r"\s*\[\]\[(\d+)\]\s*=\s*(.*?);", txt, flags=re.S)
    return int(m.group(1)), m.group(2)


def main():
    txt = re.sub(r"//[^\n]*", "", open(SRC).read())
    out = {}
    q, body = block(txt, "n")
    nn = np.array([float(v) for v in re.findall(r"-?\d+\.?\d*", body)]).reshape(q, q)
    q, body = block(txt, "d")
    dd = np.array([float(v) for v in re.findall(r"-?\d+\.?\d*", body)]).reshape(q, q)
    out["d2q9"] = nn / dd
    for name, key in (("nd_d2q9_with_e", "d2q9_with_e"), ("nd_d3q19", "d3q19")):
        q, body = block(txt, name)
        vals = [float(a) / float(b) for a, b in re.findall(r"(-?\d+\.?\d*)\s*/\s*(\d+\.?\d*)", body)]
        assert len(vals) == q * q, (name, len(vals))
        out[key] = np.array(vals).reshape(q, q)
    for k, A in out.items():
        print(k, A.shape, "max |A A - A| =", np.max(np.abs(A @ A - A)), "column sums", np.round(A.sum(0), 12)[:4])
    np.savez(OUT, **out)
    print("wrote", OUT)


if __name__ == "__main__":
    main()
