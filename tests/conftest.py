"""Test configuration.  `-m "not gpu"` = oracle vs golden vectors, host flow, ABI surface, CPU emulation of the
device math, sharding logic (gloo, world 2).  `-m gpu` = the parity tests proper, through the C ABI on a B200."""
import gzip
import json
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session", autouse=True)
def built():
    """Builds anything missing in-tree (libgelcu.so, libgelhost.so, oracle, emu)."""
    need = [os.path.join(ROOT, "gel_b200", n) for n in ("libgelcu.so", "libgelhost.so")] + \
           [os.path.join(ROOT, "oracle", "libgeloracle.so"), os.path.join(ROOT, "tests", "emu", "libgelemu.so")]
    if not all(os.path.exists(p) for p in need):
        import __graft_entry__
        __graft_entry__.build()
    return True


@pytest.fixture(scope="session")
def golden():
    return json.load(open(os.path.join(GOLDEN, "golden.json")))


@pytest.fixture(scope="session")
def cfg1_paths(tmp_path_factory):
    d = tmp_path_factory.mktemp("cfg1")
    obj, bmp = str(d / "sphere50.obj"), str(d / "tex256.bmp")
    open(obj, "wb").write(gzip.open(os.path.join(GOLDEN, "sphere50.obj.gz")).read())
    open(bmp, "wb").write(gzip.open(os.path.join(GOLDEN, "tex256.bmp.gz")).read())
    return obj, bmp


@pytest.fixture(scope="session")
def cfg1(cfg1_paths):
    """(tv, tn, tt, tex) of the cfg-1 inputs through the PRODUCT's host flow (gel_host.c)."""
    import gel_b200
    tv, tn, tt = gel_b200.load_obj(cfg1_paths[0])
    return tv, tn, tt, gel_b200.load_bmp(cfg1_paths[1])


def bits(a):
    return np.ascontiguousarray(a).view(np.uint32)


def random_soup(rng, n, *, size=(0.02, 0.25), zspread=0.3, xr=(-0.6, 0.6), yr=(-0.3, 1.2)):
    """n random triangles that stay on screen for modest view angles; unit normals; uv in [0,1]."""
    c = np.stack([rng.uniform(*xr, n), rng.uniform(*yr, n), rng.uniform(-zspread, zspread, n)], 1)
    tv = np.empty((n, 3, 3), np.float32)
    for k in range(3):
        ang = rng.uniform(0, 2 * np.pi, n)
        rad = rng.uniform(*size, n)
        tv[:, k] = c + np.stack([rad * np.cos(ang), rad * np.sin(ang), rng.uniform(-0.05, 0.05, n)], 1)
    tn = rng.normal(size=(n, 3, 3)) * 0.5 + np.array([0, 0, 1.0])
    tn /= np.linalg.norm(tn, axis=2, keepdims=True)
    tt = np.zeros((n, 3, 3), np.float32)
    tt[:, :, :2] = rng.uniform(0, 1, (n, 3, 2))
    return tv.reshape(n, 9).astype(np.float32), tn.reshape(n, 9).astype(np.float32), tt.reshape(n, 9)
