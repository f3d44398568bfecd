import importlib
import os
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200, sm_100a); run with `-m gpu` on the GPU box")


def pytest_collection_modifyitems(config, items):
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def vap():
    return importlib.import_module("video-as-prompt_b200")


@pytest.fixture(scope="session")
def wan_golden():
    return torch.load(os.path.join(GOLDEN, "wan_tiny.pt"), map_location="cpu", weights_only=False)


@pytest.fixture(scope="session")
def cog_golden():
    return torch.load(os.path.join(GOLDEN, "cog_tiny.pt"), map_location="cpu", weights_only=False)


def rel_err(a: torch.Tensor, b: torch.Tensor) -> float:
    """max|a-b| / max|b|  — the 'max-abs relative' metric of BASELINE.json's north_star (tolerance 2e-2)."""
    a, b = a.detach().float().cpu(), b.detach().float().cpu()
    return ((a - b).abs().max() / b.abs().max().clamp_min(1e-12)).item()


def cosine(a: torch.Tensor, b: torch.Tensor) -> float:
    a, b = a.detach().double().cpu().flatten(), b.detach().double().cpu().flatten()
    return (torch.dot(a, b) / (a.norm() * b.norm())).item()
