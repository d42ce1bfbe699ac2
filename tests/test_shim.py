"""shim/B200Backend.h -- the C++ shim for a NATriuM build -- compiled here against stand-ins for the Epetra /
deal.II classes it touches (shim/mock/Epetra_mock.h) and linked with the real libnatrium_b200.so.

CPU: buildOwnedFirstNumbering / haloPlanFromImporter on block partitions with up to 8 neighbours.
GPU: the real library driven through the shim (matrix blocks via ExtractMyRowView, populations via ExtractView, reference-
ordered and fused steps, lazy host mirror, exception translation) against the oracle."""
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def shim_exe(tmp_path_factory):
    from natrium_b200 import _capi
    assert os.path.exists(_capi.LIB_PATH)
    exe = str(tmp_path_factory.mktemp("shim") / "shim_check")
    libdir = os.path.dirname(_capi.LIB_PATH)
    subprocess.run(["g++", "-O1", "-std=c++17", "-Wall", "-Werror", "-I" + os.path.join(ROOT, "include"), "-I" + os.path.join(ROOT, "shim"),
                    "-o", exe, os.path.join(ROOT, "tests", "cpp", "shim_check.cpp"), "-L" + libdir, "-lnatrium_b200",
                    "-Wl,-rpath," + libdir], check=True)
    return exe


@pytest.mark.parametrize("args", ["12 12 2 2 1", "16 12 4 2 2", "9 9 3 3 3", "8 8 1 1 4", "10 10 1 4 5", "24 24 2 4 6"])
def test_shim_numbering_and_halo_plan(shim_exe, args):
    """Owned-first numbering and the ghost plan from mock Epetra importers: block partitions (px x py) with edge and corner
    ghosts, column maps and remote lists in scrambled order; what A sends to B is what B's ghost slots for A hold, in order."""
    out = subprocess.run([shim_exe, "halo", *args.split()], capture_output=True, text=True).stdout
    assert out.startswith("OK"), out
    px, py = int(args.split()[2]), int(args.split()[3])
    if px >= 3 and py >= 3:
        assert "max_neighbours=8" in out


@pytest.mark.gpu
@pytest.mark.parametrize("case", ["tgv2d_small", "tgv3d_d3q19_p2"])
def test_shim_drives_the_library(shim_exe, case, oracle_lib, tmp_path):
    from tests import common
    o = common.oracle_problem(case)
    c, st, pb, dt = common.product_problem(case)
    steps = 4
    stepper = oracle_lib.ReferenceOrderStepper(o["st"], o["blocks"], o["dofs"].N, c["nu"], dt)
    f = o["f"].copy()
    for _ in range(steps):
        assert stepper.step(f, None) == 0
    Q, D, n = st.getQ(), st.getD(), o["dofs"].N
    path = str(tmp_path / "shim.bin")
    with open(path, "wb") as fh:
        fh.write(np.array([D, Q, n, Q - 1, steps], dtype=np.int64).tobytes())
        fh.write(np.array([st.getScaling(), st.getSpeedOfSoundSquare(), c["nu"], dt], dtype=np.float64).tobytes())
        fh.write(np.ascontiguousarray(st.getDirections(), dtype=np.float64).tobytes())
        fh.write(np.ascontiguousarray(st.getWeights(), dtype=np.float64).tobytes())
        for a in range(Q - 1):
            m = o["blocks"][(a, a)].tocsr()
            fh.write(np.array([m.nnz], dtype=np.int64).tobytes())
            fh.write(np.ascontiguousarray(m.indptr, dtype=np.int64).tobytes())
            fh.write(np.ascontiguousarray(m.indices, dtype=np.int32).tobytes())
            fh.write(np.ascontiguousarray(m.data, dtype=np.float64).tobytes())
        fh.write(np.ascontiguousarray(o["f"], dtype=np.float64).tobytes())
        fh.write(np.ascontiguousarray(f, dtype=np.float64).tobytes())
    r = subprocess.run([shim_exe, "run", path], capture_output=True, text=True)
    print(r.stdout, r.stderr)
    assert r.returncode == 0 and r.stdout.startswith("OK"), r.stdout + r.stderr
