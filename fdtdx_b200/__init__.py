"""fdtdx_b200: B200-native (sm_100a) Yee time-stepping backend behind the fdtdx call surface.

Public names mirror ``fdtdx/__init__.py:9-156`` for the hot path only.
"""

from fdtdx_b200 import constants
from fdtdx_b200.boundaries import (
    BlochBoundary,
    PerfectElectricConductor,
    PerfectlyMatchedLayer,
    PerfectMagneticConductor,
    PeriodicBoundary,
    SimulationVolume,
    boundary_objects_from_config,
)
from fdtdx_b200.config import GradientConfig, RectilinearGrid, SimulationConfig, UniformGrid
from fdtdx_b200.container import ArrayContainer, FieldState, ObjectContainer, RecordingState
from fdtdx_b200.detectors import (
    ClosedSurfacePhasorPoyntingFluxDetector,
    ClosedSurfacePoyntingFluxDetector,
    EnergyDetector,
    FieldDetector,
    ModeOverlapDetector,
    PhasorDetector,
    PhasorPoyntingFluxDetector,
    PoyntingFluxDetector,
)
from fdtdx_b200.device import ClosestIndex, Device, GaussianSmoothing2D, TanhProjection
from fdtdx_b200.initialization import Material, UniformMaterialObject, init_arrays, place_objects
from fdtdx_b200.profile import CustomTimeSignalProfile, GaussianPulseProfile, SingleFrequencyProfile
from fdtdx_b200.recorder import DtypeConversion, LinearReconstructEveryK, Recorder
from fdtdx_b200.sources import (
    PointDipoleSource,
    TFSFPlaneSource,
    gaussian_amplitude_profile,
    make_plane_source,
)
from fdtdx_b200.stop_conditions import (
    DetectorConvergenceCondition,
    EnergyThresholdCondition,
    StoppingCondition,
    TimeStepCondition,
)
from fdtdx_b200.switch import OnOffSwitch, WaveCharacter


def __getattr__(name):
    # loop drivers need the CUDA extension; import lazily so that host-only code (plan compiler,
    # containers) stays importable on a CPU box, while any call fails loudly without the .so.
    if name in (
        "run_fdtd",
        "reversible_fdtd",
        "checkpointed_fdtd",
        "custom_fdtd_forward",
        "full_backward",
        "forward",
        "backward",
        "apply_params",
    ):
        from fdtdx_b200 import fdtd as _fdtd

        return getattr(_fdtd, name)
    raise AttributeError(name)
from fdtdx_b200.symmetry import unfold_array, unfold_detector_states, unfold_fields  # noqa: E402,F401
