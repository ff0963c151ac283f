import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

MODELS = os.path.join(ROOT, "hslabs_b200", "models")
PRESETS = os.path.join(MODELS, "pgs_presets.txt")
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def orc():
    """The CPU oracle (test infrastructure)."""
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "-s"])
    from oracle import orc as o
    return o


@pytest.fixture(scope="session")
def refb():
    """oracle/_ref: the reference's own sources compiled against shim headers (test infrastructure).  Built here when
    /root/reference is present; on the GPU box the prebuilt library that travelled with the snapshot is used."""
    from oracle import ref as r
    r.build()
    if not r.available():
        pytest.skip("oracle/_ref/libhslref.so not built and /root/reference not present")
    return r


def ref_xml(name):
    """Model file for the reference build: the reference's own XML when /root/reference is mounted, else the generated
    one (bit-identical load-time constants: test_generated_xml_equals_reference_xml)."""
    name = name if name.endswith(".xml") else name + ".xml"
    p = os.path.join("/root/reference", name)
    return p if os.path.exists(p) else os.path.join(MODELS, name)


def ref_presets():
    p = "/root/reference/pgs_config.txt"
    return p if os.path.exists(p) else PRESETS


@pytest.fixture(scope="session")
def hsl():
    """The product package, with its CUDA library built (nvcc cross-compiles without a GPU)."""
    import hslabs_b200
    from hslabs_b200 import build
    build.build()
    return hslabs_b200


def model_xml(name):
    return os.path.join(MODELS, name if name.endswith(".xml") else name + ".xml")


def rel_err(a, b):
    import numpy as np
    a = np.asarray(a, float)
    b = np.asarray(b, float)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-300))
