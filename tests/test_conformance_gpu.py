"""BASELINE.json's three conformance criteria on a reduced sample of scripts/conformance.py (-m gpu): 8 cases (corpus kind x rate mode x
block size) x 6 streams x 6 s through the CUDA encoder and decoder against the oracle.  The full-size run is
`python scripts/conformance.py > profiles/rNN_conformance.json`."""
import json
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_three_criteria_reduced():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "scripts", "conformance.py"), "--streams", "6", "--seconds", "6"],
                       cwd=ROOT, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-2000:]
    out = json.loads(r.stdout)
    assert len(out["cases"]) == 8
    assert out["flip_rate"] <= out["allowed_flip_rate"], out                       # criterion 2: <= 0.1 % of blocks
    for c in out["cases"]:
        assert c["window_ctrl_mismatches"] == 0, c
        assert c["decode_status_bad"] == 0, c
        assert c["decode_pcm16_max_lsb"] <= 1, c                                   # criterion 1: +-1 LSB ...
        assert c["decode_float_err_rel_peak"] < 1e-5, c                            # ... and 1e-5 of the peak
        assert c["snr_delta_db_max_abs"] < 0.1, c                                  # criterion 3: within 0.1 dB
