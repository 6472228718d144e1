"""CPU: oracle/shapes.py (parameter names / shapes restated from the reference's constructors) against the product model's
state_dict and against the shape tables stored in the golden fixtures (recorded from the UNMODIFIED reference model);
and the isolation of the reference arm of bench.py from the product package."""
import json
import os
import subprocess
import sys

import pytest

from util import ROOT, model_case


def _bench_workloads():
    sys.path.insert(0, ROOT)
    import importlib
    return importlib.import_module("bench").WORKLOADS


@pytest.mark.parametrize("name", ["zigzag8_b1", "sweep2_b1", "faceshq1024", "ucf101_sst"])
def test_shapes_match_product_model(name):
    from oracle.shapes import zigma_state_shapes
    from zigma_b200 import ZigMa
    cfg = dict(_bench_workloads()[name]["cfg"], depth=3)
    m = ZigMa(device="cpu", **cfg)
    assert zigma_state_shapes(cfg) == {k: tuple(v.shape) for k, v in m.state_dict().items()}


@pytest.mark.parametrize("name", ["tiny_zigzag8", "tiny_sweep2", "tiny_hilbert2", "tiny_patch2_cls", "tiny_video_sst", "tiny_text", "tiny_video_text"])
def test_shapes_match_reference_recorded_tables(name):
    from oracle.shapes import zigma_state_shapes
    _, cfg, shapes = model_case(name)
    assert zigma_state_shapes(cfg) == shapes


def test_reference_arm_never_imports_the_product_package():
    """`bench.py --impl reference` times the CPU restatement only: no zigma_b200 import (so no libzigma_b200.so in the
    process), same metric / config keys as the `ours` arm, unscaled ms_per_step."""
    code = ("import sys, runpy\n"
            "sys.argv = ['bench.py', '--impl', 'reference', '--steps', '1', '--bs', '1']\n"
            "runpy.run_path('bench.py', run_name='__main__')\n"
            "print('PRODUCT_IMPORTED', any(m == 'zigma_b200' or m.startswith('zigma_b200.') for m in sys.modules))\n")
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, cwd=ROOT, timeout=600,
                       env=dict(os.environ, ZIGMA_REF_BUDGET_S="5"))
    assert r.returncode == 0, r.stderr[-2000:]
    assert "PRODUCT_IMPORTED False" in r.stdout
    line = json.loads([l for l in r.stdout.splitlines() if l.startswith("{")][-1])
    assert line["impl"] == "reference" and line["unit"] == "tokens/s" and line["gpu_launches"] == 0
    assert line["config"]["global_batch"] == 1 and line["steps"] >= 1
    assert abs(line["value"] - 1024 / (line["ms_per_step"] * 1e-3)) < 1e-6 * line["value"]      # unscaled: tokens of ONE evaluation / its time
