"""CPU check of the HOST logic of the engine's callers -- the eight linear-response parametrisations, WaveFunctionSAUPS,
extended-space embedding / projection, the rdm3 / rdm4 contractions, the two-step optimisation drivers, the engine calls of one optimiser iteration -- against the reference goldens, with the libsqsv calls
replaced by the oracle in a SUBPROCESS (tests/host_standin.py; test infrastructure, never on a product path).  The kernels
behind these callers are checked by the ``-m gpu`` tests against the same goldens."""
import os
import subprocess
import sys

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.mark.parametrize("which", ["lr", "saups", "extended", "rdm34", "opt2", "ucc", "strings", "attributes", "iteration"])
def test_host_logic_of_callers(which):
    res = subprocess.run([sys.executable, os.path.join(HERE, "host_callers_check.py"), which], capture_output=True, text=True, timeout=600)
    sys.stdout.write(res.stdout[-4000:])
    sys.stderr.write(res.stderr[-2000:])
    assert res.returncode == 0 and "HOST_CALLERS_OK" in res.stdout
