import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "cfd-with-cuda_b200"))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def meshdir(tmp_path_factory):
    return str(tmp_path_factory.mktemp("meshes"))


@pytest.fixture(scope="session")
def mesh(meshdir):
    """mesh(kind, **kw) -> path of a generated .inp (cached per session)."""
    import meshgen
    cache = {}

    def make(name, kind="cube", **kw):
        if name not in cache:
            path = os.path.join(meshdir, name + ".inp")
            getattr(meshgen, kind)(path, **kw)
            cache[name] = path
        return cache[name]
    return make


@pytest.fixture(scope="session", autouse=True)
def _built():
    import subprocess
    import fracstep
    import orc
    orc.build()
    fracstep.build()
    # the rest of the checkers: stand-in library, and -- where the reference tree is mounted -- the reference programs themselves
    # (oracle/_ref: git-ignored binaries, so a fresh checkout has none; a no-op on the GPU box, which gets them prebuilt)
    subprocess.call(["make", "-C", os.path.join(ROOT, "oracle")], stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
