import glob
import importlib
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")
FIXTURES = os.path.join(ROOT, "oracle", "_ref", "fixtures")        # full liam set; present only where staged


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on the B200 box)")


@pytest.fixture(scope="session")
def built():
    """Builds the oracle, the synthetic generators and (if missing) the CUDA library."""
    subprocess.run(["make", "-s", "-C", os.path.join(ROOT, "oracle")], check=True)
    subprocess.run(["make", "-s", "-C", os.path.join(ROOT, "tools", "synth")], check=True)
    lib = os.path.join(ROOT, "universal-volumetric_b200", "libuvol_b200.so")
    if not os.path.exists(lib):
        subprocess.run(["make", "-s", "-C", os.path.join(ROOT, "universal-volumetric_b200", "csrc")], check=True)
    return True


@pytest.fixture(scope="session")
def uv(built):
    return importlib.import_module("universal-volumetric_b200")


@pytest.fixture(scope="session")
def ctx(uv):
    c = uv.Context(0)
    yield c
    c.close()


def golden_drc():
    return sorted(glob.glob(os.path.join(GOLDEN, "liam", "*.drc")))


def golden_ktx2():
    return sorted(glob.glob(os.path.join(GOLDEN, "liam", "*.ktx2")))


def fixture_drc(limit=None):
    f = sorted(glob.glob(os.path.join(FIXTURES, "geometry_draco", "*.drc")))
    return f[:limit] if limit else f


def fixture_ktx2(limit=None):
    f = sorted(glob.glob(os.path.join(FIXTURES, "texture_ktx2", "*.ktx2")))
    return f[:limit] if limit else f


def read(path):
    with open(path, "rb") as fh:
        return fh.read()


def with_draco_metadata(blob):
    """The same .drc with the metadata flag set and a metadata section spliced in behind the header (one attribute metadata, a
    geometry metadata with two entries and a nested sub-metadata), as draco::MetadataEncoder writes it."""
    def varint(v):
        o = b""
        while True:
            c = v & 0x7F; v >>= 7
            if v:
                o += bytes([c | 0x80])
            else:
                return o + bytes([c])

    def md(entries, subs=()):
        o = varint(len(entries))
        for k, v in entries:
            o += bytes([len(k)]) + k + varint(len(v)) + v
        o += varint(len(subs))
        for k, m in subs:
            o += bytes([len(k)]) + k + m
        return o
    meta = varint(1) + varint(0) + md([(b"name", b"position")]) + md([(b"clip", b"liam"), (b"fps", b"\x1e\0\0\0")], [(b"sub", md([(b"k", b"v" * 300)]))])
    hdr = bytearray(blob[:11]); hdr[10] |= 0x80
    return bytes(hdr) + meta + blob[11:]
