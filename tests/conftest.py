import os
import shutil
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu on the GPU box")


def pytest_collection_modifyitems(config, items):
    """A plain `pytest tests` on a box without a CUDA device skips the GPU tests instead of failing them."""
    if have_gpu():
        return
    skip = pytest.mark.skip(reason="no CUDA device (GPU tests run with -m gpu on a B200)")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def built():
    """Native artefacts: built once per session (no-ops when the in-tree files are current)."""
    from frog_b200 import build
    from oracle import oracle
    if shutil.which("nvcc"):
        build.build_all()
    else:  # GPU box without nvcc would still have the shipped binaries
        assert os.path.exists(build.LIB), "libfrogmatch.so missing and nvcc unavailable"
    oracle.build(ref=os.path.exists("/root/reference/match/match.cpp"))
    return True


@pytest.fixture(scope="session")
def golden_dir(tmp_path_factory):
    """A scratch copy of tests/golden with the {DIR} list templates materialised."""
    src = os.path.join(ROOT, "tests", "golden")
    dst = str(tmp_path_factory.mktemp("golden"))
    for f in os.listdir(src):
        if f.startswith("points") or f == "quirks.csv" or f.endswith(".pairs.bin") or f == "manifest.json":
            shutil.copy(os.path.join(src, f), os.path.join(dst, f))
        elif f.startswith("list_"):
            text = open(os.path.join(src, f)).read().replace("{DIR}", dst)
            open(os.path.join(dst, f), "w").write(text)
    return dst


def have_gpu() -> bool:
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False
