"""The CUDA path against the COMMITTED golden fixtures (tests/golden/golden_points.json, generated in the build container by
tests/golden/make_golden.py from the two agreeing CPU restatements): nothing but the fixture file and the C ABI is involved."""
import json
import os

import numpy as np
import pytest

import _data
import _gpu
from picard_ica_b200 import DensityType, Picard, PicardConfig
from picard_ica_b200.utils import amari_distance

pytestmark = pytest.mark.gpu
GOLD = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "golden_points.json")))
DENS = {0: DensityType.tanh_with_alpha, 1: DensityType.exp_with_alpha, 2: lambda a: DensityType.cube()}


@pytest.mark.parametrize("case", GOLD["points"], ids=lambda c: f"n{c['n']}-k{c['kind']}-o{int(c['ortho'])}-e{int(c['extended'])}")
def test_point_matches_golden(case):
    x = _data.whitened(case["n"], case["t"], seed=case["seed"])
    w = _data.orthogonal(case["n"], case["seed"] + 1)
    got = _gpu.eval_point(x, w, case["kind"], case["alpha"], case["ortho"], case["extended"], 0.01, c=w @ w.T)
    assert _data.rel_err(got["g"], np.array(case["g"])) <= 1e-10
    assert _data.rel_err(got["h"], np.array(case["h"])) <= 1e-10
    assert abs(got["loss"] - case["loss"]) <= 1e-10 * max(1.0, abs(case["loss"]))
    assert abs(got["gradient_norm"] - case["gradient_norm"]) <= 1e-10
    np.testing.assert_array_equal(got["signs"], np.array(case["signs"]))


@pytest.mark.parametrize("case", GOLD["fits"], ids=lambda c: f"{c['data']}-n{c['n']}")
def test_fit_matches_golden(case):
    x = _data.lcg_bench_data(case["n"], case["t"], 42) if case["data"] == "lcg" else _data.mixture(case["n"], case["t"], seed=case["seed"], kind=case["data"])[0]
    res = Picard.fit_with_config(x, PicardConfig(density=DENS[case["kind"]](case["alpha"]), ortho=case["ortho"], extended=case["extended"],
                                                 w_init=_data.orthogonal(case["n"], 43)))
    assert abs(res.n_iterations - case["n_iterations"]) <= 1 and res.converged == case["converged"]
    assert amari_distance(res.full_unmixing(), np.linalg.inv(np.array(case["full_unmixing"]))) <= 1e-6
