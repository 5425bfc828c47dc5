import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def oracle():
    from oracle import oracle as O
    O.build()
    return O


@pytest.fixture(scope="session")
def _engine_module():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from rsoccer_b200 import _lib, engine as E
    _lib.lib()          # fails loudly if the extension is not built
    return E


@pytest.fixture(params=["per_match", "per_match_packed", "per_body", "per_match_overlap", "per_body_overlap"])
def engine(request, monkeypatch, _engine_module):
    """The CUDA engine, once per kernel family: rs_create reads RS_PER_MATCH (1 = one lane
    per match, rs_device.cuh; 0 = one lane per body, rs_lanes.cuh; unset = by world size) and
    RS_PACKED (1 = the packed fp32x2 instruction forms of the large-world VSS-v0 kernel, 0 = the
    scalar forms; unset = by world size), so every GPU test exercises all of them whatever the
    size heuristics would pick.  "per_match_overlap" also turns the step-to-step overlap protocol
    on (RS_STEP_OVERLAP=3, include/rsoccer_b200.h): every test then runs with the tile flags."""
    monkeypatch.setenv("RS_PER_MATCH", "0" if request.param.startswith("per_body") else "1")
    monkeypatch.setenv("RS_PACKED", "0" if request.param in ("per_match", "per_body", "per_body_overlap") else "1")
    if request.param == "per_match_overlap":
        monkeypatch.setenv("RS_STEP_OVERLAP", "3")
    elif request.param == "per_body_overlap":
        monkeypatch.setenv("RS_STEP_OVERLAP", "2")
    else:
        monkeypatch.delenv("RS_STEP_OVERLAP", raising=False)
    return _engine_module
