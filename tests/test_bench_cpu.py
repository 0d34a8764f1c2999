"""bench.py's CPU legs (the reference's compiled kernels, oracle/_ref) on tiny problems: every
kind returns a positive rate and describes its sample.  No GPU."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


@pytest.mark.parametrize("kind", ["linear", "probit", "block", "nprior"])
def test_reference_throughput_kinds(kind):
    from oracle import ref_loader
    if not ref_loader.have_compiled():
        pytest.skip("oracle/_ref not built")
    import bench
    rng = np.random.default_rng(3)
    n, p = 60, 45
    X = rng.standard_normal((n, p))
    y = 2.0 * X[:, 0] + rng.standard_normal(n)
    target = (y < 0).astype(float) if kind == "probit" else y
    res = bench.reference_throughput(X, target, 30, workers=2, kind=kind)
    assert res["value"] > 0 and res["cores"] == 2 and res["kind"] == "reference"
    assert f"p={p}" in res["sample"]


def test_workload_config_names_the_baseline_configuration():
    import argparse
    import bench
    args = argparse.Namespace(sweeps=5000, burn=100, ref_sweeps=200, gpus=1, mode="gram")
    cfg = bench.workload_config(args)
    assert "configs[1]" in cfg["workload"] and cfg["chains"] == 512 and cfg["sweeps_per_step"] == 5100
    weak = bench.workload_config(args, chains=1024)
    assert "1024 chains in total" in weak["workload"] and weak["chains"] == 1024
