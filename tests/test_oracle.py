"""CPU tests: the oracle restatement (oracle/realsr_oracle.c) against the golden vectors produced by the REAL
reference (oracle/_ref = reference process_cpu + ncnn b4ba207, see oracle/gen_golden.py), plus the reference binary
itself when it is present.  This is what pins the oracle (the reference ships no tests of its own)."""
import os
import subprocess
import tempfile

import numpy as np
import pytest

from common import ROOT, Oracle, diff_stats, golden, manifest, model_paths

SMALL = [c for c in manifest()["cases"] if c["h"] * c["w"] <= 2500]


@pytest.mark.parametrize("case", SMALL, ids=[c["name"] for c in SMALL])
def test_port_matches_reference_golden(case, models_dir):
    """direct-conv fp32 port vs ncnn (Winograd fp32): <= 1 LSB, < 0.5 % of bytes differ (measured 0.05 %)"""
    img, gold = golden(case["name"])
    out = Oracle(case["model"]).process(img, case["tilesize"], tta=bool(case["tta"]))
    s = diff_stats(out, gold)
    assert s["max"] <= 1, s
    assert s["frac_diff"] < 0.005, s


def test_port_tile_geometry_matters(models_dir):
    """tile size is part of the numerical contract: T=32 vs T=200 outputs differ"""
    img, _ = golden("noise_50x33_df2k_t32")
    o = Oracle("models-DF2K")
    a = o.process(img, 32)
    b = o.process(img, 200)
    assert diff_stats(a, b)["max"] > 1


@pytest.mark.skipif(not os.path.exists(os.path.join(ROOT, "oracle", "_ref", "oracle_ref_avx2")), reason="oracle/_ref not built")
def test_reference_binary_reproduces_golden(models_dir):
    """the real reference binary is deterministic and reproduces its committed golden bit-exactly (any thread count)"""
    case = [c for c in manifest()["cases"] if c["name"] == "noise_40x36_jpeg_t200"][0]
    img, gold = golden(case["name"])
    p, b = model_paths(case["model"])
    for threads in (1, 3):
        with tempfile.TemporaryDirectory() as td:
            fi, fo = os.path.join(td, "i.raw"), os.path.join(td, "o.raw")
            img.tofile(fi)
            subprocess.check_call([os.path.join(ROOT, "oracle", "_ref", "oracle_ref_avx2"), p, b, fi, str(img.shape[1]), str(img.shape[0]), "3", fo,
                                   str(case["tilesize"]), "0", str(threads)], stdout=subprocess.DEVNULL)
            out = np.fromfile(fo, dtype=np.uint8).reshape(gold.shape)
        assert np.array_equal(out, gold)
