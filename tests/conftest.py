import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")
REFERENCE = "/root/reference"  # exists only in the build container; never on the GPU box


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu under gpurun)")
    config.addinivalue_line("markers", "reference: needs /root/reference (build container only)")
    # The C-ABI library is a build artefact (git-ignored): a fresh checkout builds it once, like __graft_entry__.build()
    lib = os.path.join(ROOT, "pytorchcv_b200", "libpcv_b200.so")
    if not os.path.exists(lib):
        import shutil
        import subprocess
        if shutil.which("nvcc") or os.path.exists("/usr/local/cuda/bin/nvcc"):
            subprocess.run(["make", "-C", os.path.join(ROOT, "pytorchcv_b200", "csrc"), "-j", "8"], check=False,
                           capture_output=True)


def pytest_collection_modifyitems(config, items):
    import torch
    have_gpu = torch.cuda.is_available()
    have_ref = os.path.isdir(os.path.join(REFERENCE, "pytorchcv"))
    for item in items:
        if "gpu" in item.keywords and not have_gpu:
            item.add_marker(pytest.mark.skip(reason="no CUDA device"))
        if "reference" in item.keywords and not have_ref:
            item.add_marker(pytest.mark.skip(reason="/root/reference not present"))


@pytest.fixture(scope="session")
def reference_pkg():
    """The real osmr/pytorchcv package, importable only where /root/reference is mounted."""
    if REFERENCE not in sys.path:
        sys.path.insert(0, REFERENCE)
    import pytorchcv  # noqa: F401
    return pytorchcv
