import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def _library_is_stale():
    lib = os.path.join(ROOT, "natrium_b200", "libnatrium_b200.so")
    if not os.path.exists(lib):
        return True
    t = os.path.getmtime(lib)
    src_dirs = [os.path.join(ROOT, "natrium_b200", "csrc"), os.path.join(ROOT, "include")]
    for d in src_dirs:
        for fn in os.listdir(d):
            if fn.endswith((".cu", ".cuh", ".h")) and os.path.getmtime(os.path.join(d, fn)) > t:
                return True
    return False


def pytest_sessionstart(session):
    """The built library is git-ignored: a fresh checkout has none.  Build it (nvcc cross-compiles without a GPU) so
    that the export / header tests and every GPU test exercise the current sources, never a missing or stale .so."""
    if _library_is_stale():
        import shutil
        import subprocess
        if shutil.which("nvcc") is None and not os.path.exists("/usr/local/cuda/bin/nvcc"):
            return          # nothing we can do here; _capi.load() will fail loudly where the library is needed
        env = dict(os.environ)
        env["PATH"] = env.get("PATH", "") + os.pathsep + "/usr/local/cuda/bin"
        subprocess.run(["make", "-C", os.path.join(ROOT, "natrium_b200", "csrc"), "-j", str(min(8, os.cpu_count() or 1))],
                       check=True, env=env, stdout=subprocess.DEVNULL)


@pytest.fixture(scope="session")
def oracle_lib():
    from oracle import cpu
    cpu.build()
    return cpu
