"""The oracle and the CUDA kernels against outputs of the REFERENCE'S OWN SOURCE.

``tests/golden/ref_steps.npz`` was produced by executing the reference's hot-path files unchanged under a NumPy-backed
``jax.numpy`` stand-in (``oracle/refexec.py``, generator ``tests/golden/make_reference_golden.py``): its
``update_E`` / ``update_H`` / ``update_detector_states`` / reverse updates / Bloch pad correction / symmetry mirror /
detector ``update`` methods / TFSF + dipole injection, on 24 seeded scenes.  This pins the oracle's transcription of
those functions against the reference code itself (what it cannot pin is XLA's float32 code generation).

* CPU: the oracle reproduces every fixture array (most bit-for-bit; bound 2e-6 of the array's max).
* CPU, only where ``/root/reference`` exists: re-executing the reference source reproduces the committed fixtures.
* GPU: the CUDA kernels against the same fixtures - fields <= 1e-5, detectors <= 1e-4 (the north star's bounds)."""

import os
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))
import make_reference_golden as G  # noqa: E402
from scenes import rel_l2  # noqa: E402

GOLD = np.load(G.PATH)


def _close(name, got, ref, tol):
    got, ref = np.asarray(got), np.asarray(ref)
    assert got.shape == ref.shape and (np.iscomplexobj(got) == np.iscomplexobj(ref) or not np.iscomplexobj(ref)), (name, got.shape, ref.shape)
    scale = max(float(np.abs(ref).max()), 1e-30)
    err = float(np.abs(got - ref).max()) / scale
    assert err <= tol, (name, err)
    return err


@pytest.mark.parametrize("name", sorted(G.SCENES))
def test_oracle_matches_reference_source(name):
    out = G.run_oracle(name)
    keys = [k for k in GOLD.files if k.startswith(name + "/")]
    assert sorted(keys) == sorted(out), (sorted(set(keys) ^ set(out)))
    worst = max(_close(k, out[k], GOLD[k], 2e-6) for k in keys)
    print(f"[{name}] oracle vs reference source: worst max-abs error / max|ref| = {worst:.2e} over {len(keys)} arrays")


def test_oracle_pure_functions_match_reference_source():
    out = G.extra_oracle()
    keys = [k for k in GOLD.files if k.startswith("extra/")]
    assert sorted(keys) == sorted(out)
    for k in keys:
        _close(k, out[k], GOLD[k], 1e-6)


def test_host_cpml_tables_match_reference_place_on_grid():
    """The CPML a / b / 1/kappa tables are built once by fdtdx_b200/boundaries.py and consumed by BOTH the oracle and
    the kernels, so a CUDA-vs-oracle test cannot see a wrong table.  Here they are compared with the tables the
    reference's own ``PerfectlyMatchedLayer.place_on_grid`` body produced (uniform / stretched grid, kappa 1 / graded
    to 4, all six faces): float32-identical up to 1 ulp."""
    host = G.pml_tables_host()
    keys = [k for k in GOLD.files if k.startswith("pml_tables/")]
    assert sorted(keys) == sorted(host) and len(keys) == 4 * 6 * 6
    for k in keys:
        a, b = host[k].reshape(-1), GOLD[k].reshape(-1)
        assert a.shape == b.shape and a.dtype == np.float32
        np.testing.assert_allclose(a, b, rtol=2e-7, atol=0, err_msg=k)


@pytest.mark.skipif(not os.path.isdir("/root/reference/src/fdtdx"), reason="reference sources not present (GPU box)")
def test_fixtures_are_what_the_reference_source_produces():
    """Guards the committed fixtures: three scenes re-executed from the reference's files."""
    from oracle import refexec

    ref = refexec.Reference()
    for name in ("pml_kappa_nonuniform", "detectors_all_nonuniform", "bloch_xy"):
        for k, v in G.run_reference(ref, name).items():
            assert np.array_equal(np.asarray(v), GOLD[k]), k


def _cuda_run(name):
    import torch

    import fdtdx_b200 as fx

    _, _, fwd, rev = G.SCENES[name]
    objects, arrays, cfg = G.scene(name)
    dev = arrays.to_torch("cuda")
    dets = len(objects.detectors) > 0
    t, out = fx.custom_fdtd_forward(dev, objects, cfg, reset_container=False, record_detectors=dets, start_time=0, end_time=fwd)
    torch.cuda.synchronize()
    return out.to_numpy()


FWD_ONLY = sorted(n for n, (_, _, _, rev) in G.SCENES.items() if rev == 0)


@pytest.mark.gpu
@pytest.mark.parametrize("name", FWD_ONLY)
def test_cuda_matches_reference_source(name):
    out = G.flatten(_cuda_run(name), f"{name}/fwd")
    worst_f = worst_d = 0.0
    for k in (k for k in GOLD.files if k.startswith(f"{name}/fwd/")):
        e = rel_l2(out[k], GOLD[k]) if np.abs(GOLD[k]).max() > 0 else float(np.abs(out[k]).max())
        if "/det/" in k:
            worst_d = max(worst_d, e)
            assert e <= 1e-4, (k, e)
        else:
            worst_f = max(worst_f, e)
            assert e <= 1e-5, (k, e)
    print(f"[{name}] CUDA vs reference source: fields / psi / P rel-L2 <= {worst_f:.2e}, detectors <= {worst_d:.2e}")
