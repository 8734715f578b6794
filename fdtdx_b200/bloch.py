"""Complex-valued runs: Bloch boundaries with a non-zero wave vector.

The reference promotes ``E``, ``H``, the CPML ``psi`` arrays and the interface recordings to
complex64 when a ``BlochBoundary`` has ``k != 0`` along its axis (``fdtd/initialization.py:581-596``)
and multiplies the wrapped ghost planes by ``exp(-+ i k L)`` (``objects/boundaries/bloch.py:61-96``).
Every other operation of the step is linear with real coefficients, so the real and the imaginary
part evolve independently except at those ghost planes.  The kernels therefore run the complex
simulation as TWO real systems (Re, Im) in the reference's planar float32 layouts, one ``Plan`` each;
the half-step kernels and the detector stencil mix the partner system's value into every wrapped
ghost read (``StepParams::bH`` / ``GridDev::Ep``, ``fdtdx_b200_set_bloch``):

    low-side ghost   F[N-1] * conj(phase):  Re' = Re*c + Im*s,   Im' = Im*c - Re*s
    high-side ghost  F[0]   * phase:        Re' = Re*c - Im*s,   Im' = Im*c + Re*s

Sources inject real amplitudes (``tfsf.py:298-307``), i.e. only into the Re system.  Detector states
recombine after each call: linear detectors (phasor) as ``Re + i Im``, field recordings keep the real
part (a complex value stored into the real state array), quadratic ones (energy ``|E|^2``, Poynting
``Re(E x conj H)``) as the sum of the two systems' values.
"""

from __future__ import annotations

from fdtdx_b200.detectors import EnergyDetector, FieldDetector, PhasorDetector, PoyntingFluxDetector
from fdtdx_b200.plan import Plan


def is_complex_run(arrays) -> bool:
    E = arrays.fields.E
    return bool(getattr(E, "is_complex", lambda: False)())


class ComplexPlan:
    """Two lock-stepped real plans behind the ``Plan`` interface the drivers use."""

    def __init__(self, objects, config, arrays):
        self.objects, self.config = objects, config
        self._ident = None
        self._split(arrays)
        self.re = Plan(objects, config, self.sub[0], bloch_role="re")
        self.im = Plan(objects, config, self.sub[1], bloch_role="im")
        self.plans = (self.re, self.im)
        for det in objects.detectors:
            if not isinstance(det, (EnergyDetector, FieldDetector, PhasorDetector, PoyntingFluxDetector)):
                raise NotImplementedError(f"detector type {type(det).__name__} with complex (Bloch) fields")
        if arrays.dispersive_c1 is not None:
            raise NotImplementedError("dispersive media with complex (Bloch) fields")

    # ------------------------------------------------------------------ Re / Im containers
    @staticmethod
    def _parts(t):
        return t.real.contiguous(), t.imag.contiguous()

    def _split(self, arrays):
        """(Re, Im) containers over planar float32 copies of the complex leaves.  Materials are shared;
        detector sub-states and recording sub-buffers persist while the caller keeps passing the same
        state tensors (the functional drivers hand back the tensors they were given)."""
        import torch

        f = arrays.fields
        E, H = self._parts(f.E), self._parts(f.H)
        psiE = {k: (self._parts(a), self._parts(b)) for k, (a, b) in f.psi_E.items()}
        psiH = {k: (self._parts(a), self._parts(b)) for k, (a, b) in f.psi_H.items()}
        ident = (
            tuple(id(v) for st in arrays.detector_states.values() for v in st.values()),
            None if arrays.recording_state is None else tuple(id(v) for v in arrays.recording_state.data.values()),
        )
        fresh = ident != self._ident
        self._ident = ident
        if fresh:
            self.det_sub = [
                {d: {k: torch.zeros_like(v) for k, v in st.items()} for d, st in arrays.detector_states.items()} for _ in range(2)
            ]
            self.rec_sub = None
            if arrays.recording_state is not None:
                self.rec_sub = [{k: self._parts(v)[w] for k, v in arrays.recording_state.data.items()} for w in range(2)]
        sub = []
        for w in range(2):
            a = arrays.aset("fields->E", E[w]).aset("fields->H", H[w])
            a = a.aset("fields->psi_E", {k: (p[0][w], p[1][w]) for k, p in psiE.items()})
            a = a.aset("fields->psi_H", {k: (p[0][w], p[1][w]) for k, p in psiH.items()})
            a = a.aset("detector_states", self.det_sub[w])
            if arrays.recording_state is not None:
                from fdtdx_b200.container import RecordingState

                a = a.aset("recording_state", RecordingState(data=self.rec_sub[w], state={}))
            sub.append(a)
        self.sub = sub

    def bind(self, arrays):
        self._split(arrays)
        self.arrays = arrays
        for w, p in enumerate(self.plans):
            p.bind(self.sub[w])
            other = self.sub[1 - w].fields
            p.bind_bloch_partner(other.E, other.H)

    # ------------------------------------------------------------------ execution
    def run_forward(self, t0: int, n: int, record_detectors: bool, record_boundaries: bool, simulate_boundaries: bool = True):
        """``forward`` (forward.py:83-156) on both systems, half-step by half-step: each half-step reads
        the partner's *other* field only, which neither launch of that half-step writes."""
        for t in range(int(t0), int(t0) + int(n)):
            if record_detectors:
                for p in self.plans:
                    p.run_forward_phase(t, 3, True, False, simulate_boundaries)      # H_prev for the co-location stencil
            for ph in (0, 1):
                for p in self.plans:
                    p.run_forward_phase(t, ph, False, False, simulate_boundaries)    # update_E, then update_H
            for p in self.plans:
                p.run_forward_phase(t, 2, record_detectors, record_boundaries, simulate_boundaries)

    def run_reverse(self, t_from: int, n: int, record_detectors: bool, reset_fields: bool):
        """``backward`` (backward.py:62-135), phase-interleaved like the forward step."""
        for t in range(int(t_from) - 1, int(t_from) - 1 - int(n), -1):
            for ph in range(6):
                for p in self.plans:
                    p.run_reverse_phase(t, ph, record_detectors, reset_fields)

    def finish(self, arrays):
        """Write the two systems back into the caller's complex leaves and recombine detector states."""
        import torch

        f = arrays.fields
        A, B = self.sub[0].fields, self.sub[1].fields
        f.E.copy_(torch.complex(A.E, B.E))
        f.H.copy_(torch.complex(A.H, B.H))
        for name in f.psi_E:
            for w in range(2):
                f.psi_E[name][w].copy_(torch.complex(A.psi_E[name][w], B.psi_E[name][w]))
                f.psi_H[name][w].copy_(torch.complex(A.psi_H[name][w], B.psi_H[name][w]))
        for det in self.objects.detectors:
            for key, out in arrays.detector_states[det.name].items():
                a, b = self.det_sub[0][det.name][key], self.det_sub[1][det.name][key]
                if isinstance(det, PhasorDetector):
                    out.copy_(a + 1j * b)                       # linear: Re + i Im
                elif isinstance(det, FieldDetector):
                    out.copy_(a)                                # the real part is what a real state keeps
                else:
                    out.copy_(a + b)                            # |E|^2, Re(E x conj H): sum of the systems
        if arrays.recording_state is not None:
            for k, out in arrays.recording_state.data.items():
                out.copy_(torch.complex(self.rec_sub[0][k], self.rec_sub[1][k]))
        return arrays

    def launch_count(self) -> int:
        return self.re.launch_count() + self.im.launch_count()
