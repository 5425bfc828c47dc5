"""-m gpu: the unmodified reference package on the CUDA engine (README.md:116-133)."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_unmodified_reference_envs_run_on_the_cuda_engine(_engine_module):
    """gym.make of all five registered ids from the installed reference package (baseline/_ref), `robosim`
    served by librsoccer_b200.so, compared step by step with the same classes on the oracle
    (tests/dropin_check.py).  Runs in its own interpreter: it swaps sys.modules entries."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from dropin_check import reference_path
    if reference_path() is None:
        pytest.skip("reference package not installed (baseline/_ref: __graft_entry__.build() installs it where /root/reference exists)")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "dropin_check.py"), "80"],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and r.stdout.strip().endswith("OK"), r.stdout[-3000:] + r.stderr[-3000:]
