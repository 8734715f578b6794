"""The torch restatement of the forward step (oracle/yee_torch.py: the differentiable gradient
reference and the multi-threaded CPU arm of bench.py) is pinned to the NumPy oracle, which is itself
pinned by the reference's known-answer and physics tests (test_oracle_known_answers.py)."""

import numpy as np
import pytest
import torch

from oracle import yee, yee_torch
from scenes import build_scene, rel_l2, seed_fields

CASES = {
    "pml_plane_source_detectors": dict(shape=(12, 10, 16), source="plane_z", detectors=("energy", "phasor", "poynting", "field"), time=6e-15),
    "nonuniform_diag_sigma": dict(shape=(10, 9, 12), nonuniform=True, eps_tier=3, sigma_E=True, sigma_H=True, mu_tier=3),
    "pec_pmc_dipole": dict(shape=(9, 9, 10), source="dipole", time=5e-15,
                           boundaries={"min_x": "pec", "max_x": "pmc", "min_y": "pmc", "max_y": "pec", "min_z": "pec", "max_z": "pml"}),
    "periodic_kappa": dict(shape=(8, 8, 12), thickness=3, kappa=True,
                           boundaries={"min_x": "periodic", "max_x": "periodic", "min_y": "pml", "max_y": "pml", "min_z": "pml", "max_z": "pml"}),
}


@pytest.mark.parametrize("name", list(CASES))
def test_torch_oracle_matches_numpy_oracle(name):
    objects, arrays, cfg = build_scene(**CASES[name])
    seeded = "source" not in CASES[name]
    if seeded:
        seed_fields(arrays, seed=2)
    steps = min(cfg.time_steps_total, 25)
    st = (0, arrays)
    for _ in range(steps):
        st = yee.forward(st, cfg, objects, None, True, False, True)
    with torch.no_grad():
        E, H, det = yee_torch.run_forward(arrays, objects, cfg, steps, dtype=torch.float32)
    assert np.abs(st[1].fields.E).max() > 0
    assert rel_l2(E.numpy(), st[1].fields.E) <= 2e-6
    assert rel_l2(H.numpy(), st[1].fields.H) <= 2e-6
    for dname, state in st[1].detector_states.items():
        for key, ref in state.items():
            assert rel_l2(det[dname][key].numpy(), ref) <= 1e-5, (dname, key)


def test_torch_oracle_float64_is_the_float32_limit():
    objects, arrays, cfg = build_scene(shape=(10, 9, 12), source="plane_z", time=5e-15)
    steps = min(cfg.time_steps_total, 20)
    with torch.no_grad():
        E32, H32, _ = yee_torch.run_forward(arrays, objects, cfg, steps, dtype=torch.float32)
        E64, H64, _ = yee_torch.run_forward(arrays, objects, cfg, steps, dtype=torch.float64)
    assert rel_l2(E32.numpy(), E64.numpy()) <= 1e-5
    assert rel_l2(H32.numpy(), H64.numpy()) <= 1e-5
