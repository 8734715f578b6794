"""Precomputed per-cell 3x3 update matrices for the full-tensor tier (setup-time, on device).

The reference recomputes ``A = M1^-1 M2``, ``B = c M1^-1 inv`` every step with two batched solves
(``fdtd/misc.py:69-129``); they depend only on the materials, so they are
solved once per material assignment with the same float32 solve and read back as 9-component
arrays by the tensor kernels.
"""

from __future__ import annotations

from fdtdx_b200.constants import eta0


def _expand(x, spatial, device):
    import torch

    out = torch.zeros((3, 3, *spatial), dtype=torch.float32, device=device)
    if x is None:
        return None
    if not hasattr(x, "shape") or len(x.shape) == 0:
        for i in range(3):
            out[i, i] = float(x)
        return out
    n = x.shape[0]
    if n == 9:
        return x.reshape(3, 3, *spatial).to(torch.float32)
    for i in range(3):
        out[i, i] = x[0] if n == 1 else x[i]
    return out


def update_matrices(inv_prop, sigma, courant_number: float, kind: str):
    """Returns (A, B, A_rev, B_rev), each (9, Nx, Ny, Nz) float32 on the device of the inputs."""
    import torch

    ref = sigma if (not hasattr(inv_prop, "shape") or len(inv_prop.shape) == 0) else inv_prop
    spatial, device = tuple(ref.shape[1:]), ref.device
    inv = _expand(inv_prop, spatial, device)
    sig = _expand(sigma, spatial, device)
    eta = eta0 if kind == "E" else 1.0 / eta0
    eye = torch.eye(3, dtype=torch.float32, device=device)[:, :, None, None, None].expand(3, 3, *spatial)
    M1, M2 = eye.clone(), eye.clone()
    if sig is not None:
        factor = torch.tensor(courant_number * eta / 2, dtype=torch.float32, device=device) * torch.einsum("ijxyz,jkxyz->ikxyz", inv, sig)
        M1 = M1 + factor
        M2 = M2 - factor
    perm, inv_perm = (2, 3, 4, 0, 1), (3, 4, 0, 1, 2)
    c = torch.tensor(courant_number, dtype=torch.float32, device=device)

    def solve(a, b):
        return torch.linalg.solve(a.permute(perm).contiguous(), b.permute(perm).contiguous()).permute(inv_perm)

    A = solve(M1, M2)
    B = c * solve(M1, inv)
    Ar = solve(M2, M1)
    Br = c * solve(M2, inv)
    flat = lambda m: m.reshape(9, *spatial).contiguous()
    return flat(A), flat(B), flat(Ar), flat(Br)
