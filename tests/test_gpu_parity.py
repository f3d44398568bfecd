"""`-m gpu`: the parity tests proper.  Every case calls the sm_100a kernels through the C ABI and compares with the
oracle / golden fixtures (see tests/gpu_checks.py for the individual checks and their tolerances)."""
import pytest

import gpu_checks

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(600)]


@pytest.mark.parametrize("name", list(gpu_checks.CHECKS))
def test_gpu_check(name):
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    result = gpu_checks.CHECKS[name]()
    torch.cuda.synchronize()
    print(name, result)


def test_native_library_is_loaded():
    """The GPU path must run on libvap_b200.so, not on a torch fallback: the library is mapped into this process."""
    import importlib
    vap = importlib.import_module("video-as-prompt_b200")
    vap._lib.load()
    maps = open("/proc/self/maps").read()
    assert "libvap_b200.so" in maps
