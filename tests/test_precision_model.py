"""CPU check of the precision DESIGN.md section 2 commits to: a torch model of the network that rounds activations to fp16
exactly where the engine stores fp16 (tools/precision_twin.py) meets the parity contract against the reference goldens.
This pins the *design* (fp32 Conv_0, fp16 operands with fp32 accumulation, fp16 + int8-extension dense-block residual, fp32
RRDB skip stream) independently of the GPU
run; the gpu tests then check that the kernels implement it."""
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))


@pytest.mark.parametrize("name,natural", [("noise_40x36_jpeg_t200", False), ("readme0_crop64_df2k_t200", True)])
def test_engine_precision_model_meets_contract(name, natural, models_dir):
    import precision_twin as pt
    s = pt.run(name, "engine")
    if natural:
        assert s["max"] <= 1 and s["frac_diff"] <= 0.03, s  # >= 97 % of bytes exact on natural images
    else:
        assert s["max"] <= 2 and s["n_gt1"] <= max(2, s["n"] // 10000), s
    exact = pt.run(name, "all32")  # without any rounding the model IS the reference up to summation order
    assert exact["max"] <= 1 and exact["frac_diff"] < 0.002, exact
