import os
import sys

import pytest
import torch

os.environ.setdefault("S2AG_ALLOW_EMU", "1")   # the CPU logic-emulator build may be injected by this harness only
ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


_EMU = {"lib": None}


@pytest.fixture(params=[pytest.param("emu"), pytest.param("cuda", marks=pytest.mark.gpu)])
def dev(request):
    """Device under test.  'cuda' = the product library on the GPU (parity tests proper).
    'emu' = the same kernel sources compiled against the CPU execution-model emulator in
    tests/emu (test infrastructure; checks kernel/host LOGIC in the GPU-less container)."""
    from speech2affective_gestures_b200 import _C
    if request.param == "cuda":
        assert torch.cuda.is_available(), "gpu-marked test needs a CUDA device"
        if _C.is_emulated():  # never let an earlier emulator injection leak into a GPU test
            _C._lib, _C._emulated = None, False
        assert _C.lib().s2ag_is_device_build() == 1
        assert not _C.is_emulated()
        return torch.device("cuda:0")
    cs = getattr(request.node, "callspec", None)
    if cs is not None and cs.params.get("kind") == "full":
        pytest.skip("full-width configuration runs on the GPU only")
    if _EMU["lib"] is None:
        sys.path.insert(0, os.path.join(ROOT, "tests", "emu"))
        import build_emu
        _EMU["lib"] = build_emu.build()
    _C._inject_for_tests(_EMU["lib"], strict=os.environ.get("S2AG_EMU_STRICT", "1") == "1")
    return torch.device("cpu")


@pytest.fixture(autouse=True)
def _reset_injected_noise():
    """the parity harness injects re-parametrisation noise through a module global: never let it leak between tests"""
    yield
    from speech2affective_gestures_b200.net import embedding_net as men
    men.eps_source = None
