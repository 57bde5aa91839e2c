"""Opt-in checks of the experimental library build (videoseal_b200/libvsb200_exp.so: -DVSB_EXP, see __graft_entry__.py and
DESIGN.md section 8).  Skipped unless VSB_TEST_EXP=1: the default library is the product; this file is the first thing to run
when the experimental kernels get GPU time:

    VSB_TEST_EXP=1 python -m pytest tests/test_experimental_gpu.py -m gpu -q

Each case runs the layer-by-layer network walk of tests/gpu_diag.py in a subprocess (the library is chosen at import time through
VSB200_LIB) and checks every reported tensor against the oracle with the tolerances of tests/test_e2e_gpu.py."""
import json
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXP_LIB = os.path.join(ROOT, "videoseal_b200", "libvsb200_exp.so")


def _walk(env_extra):
    env = dict(os.environ, VSB200_LIB=EXP_LIB, **env_extra)
    p = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "gpu_diag.py"), "--net", "videoseal_1.0"], capture_output=True,
                       text=True, timeout=600, env=env)
    rows = [json.loads(l[7:]) for l in p.stdout.splitlines() if l.startswith("RESULT ")]
    assert rows, p.stderr[-2000:]
    return {r["tensor"]: r for r in rows}


@pytest.mark.skipif(os.environ.get("VSB_TEST_EXP") != "1" or not os.path.exists(EXP_LIB), reason="experimental build: opt in with VSB_TEST_EXP=1")
@pytest.mark.parametrize("env_extra", [{"VSB_NO_PDL": "1"}, {}, {"VSB_NO_PDL": "1", "VSB_DW2": "1"}, {"VSB_DW2": "1"}],
                         ids=["epilogue-trims", "pdl", "rolling-dwconv", "pdl+rolling-dwconv"])
def test_experimental_build_walks_the_network_like_the_oracle(env_extra):
    rows = _walk(env_extra)
    for name, r in rows.items():
        assert "error" not in r and "shape_got" not in r, r
        assert r.get("nan", 0) == 0, r
    assert rows["imgs_w"]["maxerr"] <= 1e-3
    assert rows["logits"]["bit_mismatch"] == 0 and rows["logits"]["maxerr"] <= 1.5e-3 * rows["logits"]["ref_absmax"]
    for name in ("s0b0_a", "s1b0_a"):            # outputs of the depthwise 7x7 + LayerNorm of the first two stages (fp16, O(1) values)
        assert rows[name]["maxerr"] <= 2e-2 * max(1.0, rows[name]["ref_absmax"]), rows[name]
