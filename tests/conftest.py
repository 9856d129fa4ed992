import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


# Unit-level parity first, whole-step machinery last: with `-x` a failure in the CUDA-graph tests must not hide the kernel,
# mel and inference tests behind it.
_ORDER = ["test_boundary", "test_oracle_pinned", "test_mel", "test_kernels", "test_gpu_parity", "test_synthesize",
          "test_evaluate", "test_fgd", "test_pipeline", "test_gan_baseline", "test_checkpoint", "test_data", "test_dp_gloo", "test_dp_nccl", "test_b128_parity",
          "test_graph_step"]


def pytest_collection_modifyitems(session, config, items):
    def key(item):
        mod = os.path.splitext(os.path.basename(str(item.fspath)))[0]
        return _ORDER.index(mod) if mod in _ORDER else len(_ORDER) - 1
    items.sort(key=key)   # stable: keeps the order inside each module


@pytest.fixture(scope="session")
def golden_modules():
    import torch
    return torch.load(os.path.join(GOLDEN, "modules.pt"), weights_only=False)


@pytest.fixture(scope="session")
def golden_steps():
    import torch
    return {v: torch.load(os.path.join(GOLDEN, f"step_{v}.pt"), weights_only=False) for v in ("gesture", "expressive")}
