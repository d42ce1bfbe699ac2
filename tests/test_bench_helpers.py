"""Host-side pieces of bench.py that need no GPU: the algorithmic-bytes model of SURVEY 8(d), the BASELINE configurations as
data, the reference arm's behaviour on ranks other than 0, and that the product path does not import the oracle."""
import math
import os
import subprocess
import sys
import types

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def test_algorithmic_bytes_match_the_survey():
    """D2Q9 1 608, D3Q19 4 296, D2Q25H f+g 6 120 (p = 4) / 2 856 (p = 2), D3Q45 f+g 67 488 B per DoF."""
    assert bench.algorithmic_bytes_per_dof(120, 2, 9) == 1608
    assert bench.algorithmic_bytes_per_dof(330, 3, 19) == 4296
    assert bench.algorithmic_bytes_per_dof(440, 2, 25, with_g=True) == 6120
    assert bench.algorithmic_bytes_per_dof(8 * 3 + 16 * 9, 2, 25, with_g=True) == 2856
    assert bench.algorithmic_bytes_per_dof(5500, 3, 45, with_g=True) == 67488


def test_case_specs_are_the_baseline_configurations():
    args = types.SimpleNamespace(cells=32, order=4, stencil="D3Q19", stretch=0.0, jitter=0.0, row_noise=0.0)
    c2 = bench.case_spec("c2", args)
    assert c2["dim"] == 3 and c2["cells"] == [32, 32, 32] and c2["p"] == 4 and not any(c2["walls"]) and not c2["with_g"]
    assert abs(c2["scaling"] - math.sqrt(3) / 0.05) < 1e-12 and abs(c2["nu"] - 2 * math.pi) < 1e-12 and c2["cfl"] == 0.4
    c1, c3, c4, c5 = (bench.case_spec(k, args) for k in ("c1", "c3", "c4", "c5"))
    assert c1["stencil"] == "D2Q9" and c1["cells"] == [8, 8] and c1["p"] == 4
    assert c3["stencil"] == "D3Q45" and c3["with_g"] and c3["prandtl"] == 0.71 and c3["sutherland"] and c3["cells"] == [16, 16, 16]
    assert c4["stencil"] == "D2Q25H" and c4["p"] == 2 and all(c4["walls"]) and c4["cells"] == [512, 512] and c4["wall_kind"] == "velocity"
    assert c5["stencil"] == "D3Q45" and c5["walls"] == [False, True, False] and c5["wall_kind"] == "thermal" and c5["force_type"] == "EXACT_DIFFERENCE"
    assert c5["stretch"] == 0.8 and abs(c5["wall_T"] - 0.85) < 1e-15
    # stretched vertices: the reference's channel mapping y -> y - 0.8 sin(2 pi y) / (2 pi), monotone, end points kept
    v = bench.case_vertices(c5, c5["cells"], c5["length"])[1]
    assert np.all(np.diff(v) > 0) and abs(v[0]) < 1e-15 and abs(v[-1] - c5["length"][1]) < 1e-12
    assert np.diff(v)[0] < np.diff(v)[len(v) // 2]            # cells are thin at the walls


def test_block_layouts_cover_the_scaling_run():
    for world, blocks in bench.BLOCKS_OF_WORLD.items():
        assert int(np.prod(blocks)) == world and len(blocks) == 3
    assert set(bench.BLOCKS_OF_WORLD) == {2, 4, 8}


def test_reference_arm_is_silent_on_other_ranks():
    """Under torchrun only rank 0 runs and prints the reference arm; the other ranks exit 0 without work."""
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "0"],
                       env=env, capture_output=True, text=True, timeout=120)
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_product_package_does_not_touch_the_oracle():
    for root, _, files in os.walk(os.path.join(ROOT, "natrium_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(root, f), errors="ignore").read()
                assert "import oracle" not in src and "from oracle" not in src and "liboracle" not in src, os.path.join(root, f)
    src = open(os.path.join(ROOT, "shim", "B200Backend.h")).read()
    assert "oracle" not in src
