"""`-m gpu`: parity against the reference's OWN classes on the same GPU (tests/ref_gpu_checks.py) — stock forward, then
`vap_b200.install()` on that same instance, at Wan-14B / CogVideoX-5B widths.  Needs the reference staged under the git-ignored
baseline/_ref (`python baseline/ref_loader.py`, done by `__graft_entry__.build()` in the authoring container; the staged tree
travels to the GPU box with the snapshot)."""
import pytest
import torch

import ref_gpu_checks

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(900)]


@pytest.mark.parametrize("name", list(ref_gpu_checks.CHECKS))
def test_against_reference_on_gpu(name):
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    if not ref_gpu_checks.available():
        pytest.skip("the reference is not staged under baseline/_ref (run `python baseline/ref_loader.py` where /root/reference exists)")
    result = ref_gpu_checks.CHECKS[name]()
    torch.cuda.synchronize()
    print(name, result)
