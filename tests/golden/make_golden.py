"""Generates the committed golden vectors: oracle outputs on seeded scenes.

    PYTHONPATH=. python tests/golden/make_golden.py

The reference cannot be imported in this image (no jax), so these vectors are produced by the CPU
oracle (``oracle/yee.py``), which is itself pinned by the reference's known-answer and physics
tests (tests/test_oracle_*.py).  They freeze the oracle (regression pin, CPU test) and give the
CUDA parity tests a fixture that does not depend on running the oracle on the GPU box.
"""
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, os.path.dirname(HERE))

import numpy as np  # noqa: E402

from oracle import yee  # noqa: E402
from scenes import build_scene, seed_fields  # noqa: E402

CASES = {
    "pml_source_detectors_nonuniform": (dict(source="plane_z", detectors=("energy_slices", "phasor", "poynting", "field_reduce"), nonuniform=True, eps_tier=3, time=6e-15), 40, False),
    "ade_c4_sigma_seeded": (dict(poles=2, c4=True, sigma_E=True, eps_tier=3, coeff_tier=3), 8, True),
    "periodic_sigma_mu_seeded": (dict(boundaries="periodic", sigma_E=True, sigma_H=True, mu_tier=3, eps_tier=3), 10, True),
    # ragged rows (Nz % 4 != 0, z-padded shadows / interleaved kernels), x-normal plane source, odd CPML thickness
    "ragged_nz21_plane_x": (dict(shape=(14, 11, 21), thickness=3, source="plane_x", nonuniform=True, eps_tier=3,
                                 detectors=("energy_slices", "phasor", "poynting"), time=6e-15), 40, False),
    # full-tensor tier with the energy detectors (3x3 inverse per cell)
    "tensor_eps9_mu9_energy": (dict(shape=(12, 10, 16), eps_tier=9, mu_tier=9, source="plane_z", detectors=("energy", "energy_reduce", "poynting"), time=5e-15), 30, False),
}


def run_case(name):
    kw, steps, seeded = CASES[name]
    objects, arrays, cfg = build_scene(**kw)
    if seeded:
        seed_fields(arrays, seed=11)
    st = (0, arrays)
    for _ in range(steps):
        st = yee.forward(st, cfg, objects, None, True, False, True)
    out = {"E": st[1].fields.E, "H": st[1].fields.H}
    for k, (a, b) in st[1].fields.psi_E.items():
        out[f"psiE1_{k}"], out[f"psiE2_{k}"] = a, b
    if st[1].fields.dispersive_P_curr is not None:
        out["P_curr"] = st[1].fields.dispersive_P_curr
    for d, s in st[1].detector_states.items():
        for k, v in s.items():
            out[f"det_{d}_{k}"] = v
    return out


if __name__ == "__main__":
    for name in CASES:
        np.savez_compressed(os.path.join(HERE, f"{name}.npz"), **run_case(name))
        print("wrote", name)
