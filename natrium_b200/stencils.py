"""Velocity stencils on the hot path -- host mirror of NATriuM's ``Stencil`` interface.

Mirrors ``natrium::Stencil`` (L/stencils/Stencil.h:53-171): ``getD/getQ/getDirections/
getDirection/getWeight(s)/getSpeedOfSoundSquare/getScaling/getMaxParticleVelocityMagnitude/
getIndexOfOppositeDirection``.  Direction order and weights are those of
L/stencils/{D2Q9,D3Q19,D3Q15,D2Q25H,D3Q45}.cpp, because the reference's collision code
hard-codes index sums against that order (AuxiliaryCollisionFunctions.h:258-287).
Only the stencils BASELINE.json's configs name are provided.
"""
import math

import numpy as np

Stencil_D2Q9, Stencil_D3Q19, Stencil_D3Q15, Stencil_D2Q25H, Stencil_D3Q45 = (
    "Stencil_D2Q9", "Stencil_D3Q19", "Stencil_D3Q15", "Stencil_D2Q25H", "Stencil_D3Q45")


def _d2q9_units():
    return [(0, 0), (1, 0), (0, 1), (-1, 0), (0, -1), (1, 1), (-1, 1), (-1, -1), (1, -1)]


def _d3q19_units():
    return [(0, 0, 0), (1, 0, 0), (0, 0, 1), (-1, 0, 0), (0, 0, -1), (0, -1, 0), (0, 1, 0),
            (1, 0, 1), (-1, 0, 1), (-1, 0, -1), (1, 0, -1), (1, -1, 0), (1, 1, 0), (-1, 1, 0), (-1, -1, 0),
            (0, -1, 1), (0, 1, 1), (0, 1, -1), (0, -1, -1)]


def _d3q15_units():
    return [(0, 0, 0), (1, 0, 0), (-1, 0, 0), (0, 1, 0), (0, -1, 0), (0, 0, 1), (0, 0, -1),
            (1, 1, 1), (-1, -1, -1), (1, 1, -1), (-1, -1, 1), (1, -1, 1), (-1, 1, -1), (1, -1, -1), (-1, 1, 1)]


def _d2q25h_dirs(s):
    # operation order of D2Q25H::makeDirections (bit-identical tables)
    m = s * math.sqrt(5. - math.sqrt(10.)) / math.sqrt(3.)
    n = math.sqrt(5. + math.sqrt(10.)) * s / math.sqrt(3.)
    ring = _d2q9_units()[1:]
    out = [(0.0, 0.0)] + [(m * x, m * y) for x, y in ring] + [(n * x, n * y) for x, y in ring]
    out += [(m, n), (m, -n), (-m, -n), (-m, n), (n, m), (n, -m), (-n, -m), (-n, m)]
    return out


def _d2q25h_weights():
    r = (math.sqrt(5.) - math.sqrt(2.)) / math.sqrt(3.)
    # same evaluation order as D2Q25H::makeWeights so the tables are bit-identical to the reference's
    w0 = (-3 - 3 * r * r * r * r + 54 * r * r) / (75 * r * r)
    wm = (9 * r * r * r * r - 6 - 27 * r * r) / (300 * r * r * (r * r - 1))
    wn = (9 - 6 * r * r * r * r - 27 * r * r) / (300 * (1 - r * r))
    return [w0 * w0] + [w0 * wm] * 4 + [wm * wm] * 4 + [w0 * wn] * 4 + [wn * wn] * 4 + [wm * wn] * 8


# D3Q45 (off-lattice, Hermite-quadrature based): rows before the 1/sqrt(3) normalisation
_Q45 = dict(a=0.06386083877343968, b=1.2239121278243665, c=1.5766994272507744, d=0.5069610024977665,
            e=2.9239876105912574, f=0.4744978678080795, g=1.7320508075688787, h=2.403092127540177,
            i=0.8892242114059369, j=1.5602655313772367, k=2.7367507163016924, l=0.14279717659756475,
            m=3.5256070994177073, n=1.1335992635264445)
_Q45_ROWS = """
0 0 0 | a -b -b | -b a -b | -b -b a | c -d -d | -d c -d | -d -d c | d d -c | d -c d | -c d d |
b b -a | b -a b | -a b b | e f f | f e f | f f e | g g g | h i -j | h -j i | j -i -h | j -h -i |
i h -j | i -j h | -i j -h | -i -h j | -j h i | -j i h | -h j -i | -h -i j | -g -g -g | -f -f -e |
-f -e -f | -e -f -f | k k -l | k -l k | -l k k | -m n n | n -m n | n n -m | -n -n m | -n m -n |
m -n -n | l -k -k | -k l -k | -k -k l
"""
_Q45_W = [0.20740740740740618] + [0.05787037037037047] * 12 + [0.00462962962962958] * 20 + [0.0004629629629629939] * 12


def _d3q45_dirs(s):
    rows = []
    for row in _Q45_ROWS.replace("\n", " ").split("|"):
        vals = []
        for tok in row.split():
            if tok == "0":
                vals.append(0.0)
            else:
                sgn = -1.0 if tok[0] == "-" else 1.0
                vals.append(sgn * _Q45[tok[-1]])
        rows.append(tuple(s * v / math.sqrt(3) for v in vals))
    assert len(rows) == 45
    return rows


def _scaled(units):
    return lambda s: [tuple(s * c if c else 0.0 for c in row) for row in units()]


# name -> (scaled directions(s), weights(), |e|_max / scaling as the reference's headers state it)
_TABLES = {
    Stencil_D2Q9: (_scaled(_d2q9_units), lambda: [4 / 9.] + [1 / 9.] * 4 + [1 / 36.] * 4, math.sqrt(2)),
    Stencil_D3Q19: (_scaled(_d3q19_units), lambda: [1 / 3.] + [1 / 18.] * 6 + [1 / 36.] * 12, math.sqrt(2)),
    Stencil_D3Q15: (_scaled(_d3q15_units), lambda: [2 / 9.] + [1 / 9.] * 6 + [1 / 72.] * 8, math.sqrt(3)),
    Stencil_D2Q25H: (_d2q25h_dirs, _d2q25h_weights, math.sqrt(2)),
    Stencil_D3Q45: (_d3q45_dirs, lambda: list(_Q45_W), math.sqrt(2)),
}


class Stencil:
    """Host mirror of natrium::Stencil for the stencils on the path."""

    def __init__(self, stencil_type, scaling=1.0):
        name = stencil_type if stencil_type.startswith("Stencil_") else "Stencil_" + stencil_type
        dirs, weights, vmax = _TABLES[name]
        self._type = name
        self._scaling = float(scaling)
        self._e = np.ascontiguousarray(np.array(dirs(self._scaling), dtype=np.float64))
        self._w = np.ascontiguousarray(np.array(weights(), dtype=np.float64))
        self._vmax = self._scaling * vmax if name == Stencil_D3Q15 else vmax * self._scaling
        self._opposite = None

    def getStencilType(self):
        return self._type

    def getD(self):
        return self._e.shape[1]

    def getQ(self):
        return self._e.shape[0]

    def getDirections(self):
        return self._e

    def getDirection(self, i):
        return self._e[i]

    def getWeights(self):
        return self._w

    def getWeight(self, i):
        return float(self._w[i])

    def getScaling(self):
        return self._scaling

    def getSpeedOfSoundSquare(self):
        return self._scaling * self._scaling / 3.

    def getSpeedOfSound(self):
        return self._scaling / math.sqrt(3.)

    def getMaxParticleVelocityMagnitude(self):
        return self._vmax

    def getIndexOfOppositeDirection(self, i):
        if self._opposite is None:
            opp = []
            for a in range(self.getQ()):
                d = np.abs(self._e + self._e[a]).sum(axis=1)
                opp.append(int(np.argmin(d)))
            self._opposite = opp
        return self._opposite[i]
