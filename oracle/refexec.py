"""TEST INFRASTRUCTURE ONLY - executes the reference's OWN source files for the hot path on the CPU.

The reference (``/root/reference/src/fdtdx``) cannot be imported here: it needs jax, equinox, pytreeclass,
tidy3d ... none of which is installed.  Its hot-path arithmetic, however, is plain ``jax.numpy`` array code.
This module

* provides a NumPy-backed stand-in for ``jax`` / ``jax.numpy`` (functional ``.at[...]`` updates, float32 /
  complex64 results like JAX with x64 disabled, ``lax.cond`` as a Python branch), and
* ``exec``s the reference's source files *unchanged and where they lie* (nothing is copied into this repo)
  with every other import replaced by a permissive stub (decorators return their argument, base classes
  are empty), and
* wraps this repo's host-mirror objects in proxies that borrow the reference classes' methods
  (``PerfectlyMatchedLayer.step_cpml``, ``BlochBoundary.apply_pad_correction``, ``*Detector.update`` ...),

so that ``ref.update_E(...)``, ``ref.curl_H(...)``, ``ref.interpolate_fields(...)`` ... are the reference's own
functions.  ``tests/golden/make_reference_golden.py`` uses it to generate the fixtures
``tests/golden/ref_*.npz`` (committed; the GPU box has no ``/root/reference``) that pin the oracle
(``oracle/yee.py``) - its transcription of those functions - against the reference source itself.  What this
does NOT pin: XLA's own float32 code generation (fusion / FMA contraction), which only a real JAX run could show.
"""

from __future__ import annotations

import abc
import ast
import os
import sys
import types

import numpy as np

REF_ROOT = os.environ.get("FDTDX_REFERENCE_SRC", "/root/reference/src/fdtdx")


def available() -> bool:
    return os.path.isdir(REF_ROOT)


# ------------------------------------------------------------------------------------------ jax.numpy stand-in
def _narrow(x):
    """JAX without x64: float64 -> float32, complex128 -> complex64."""
    if isinstance(x, np.ndarray):
        if x.dtype == np.float64:
            x = x.astype(np.float32)
        elif x.dtype == np.complex128:
            x = x.astype(np.complex64)
        return x.view(JArr)
    if isinstance(x, np.floating) and x.dtype == np.float64:
        return np.float32(x)
    if isinstance(x, np.complexfloating) and x.dtype == np.complex128:
        return np.complex64(x)
    if isinstance(x, (tuple, list)):
        return type(x)(_narrow(v) for v in x)
    return x


def _weak(x):
    """NumPy float64 / complex128 scalars (from this repo's host objects) act like weakly typed Python scalars."""
    if isinstance(x, np.ndarray) and not isinstance(x, JArr):
        x = x.view(JArr)
    if isinstance(x, (np.ndarray,)) and x.ndim == 0 and x.dtype in (np.float64, np.complex128):
        return x.item()
    if isinstance(x, np.ndarray) and x.dtype == np.float64:
        return x.astype(np.float32)
    if isinstance(x, np.ndarray) and x.dtype == np.complex128:
        return x.astype(np.complex64)
    if isinstance(x, (np.float64, np.complex128)):
        return x.item()
    return x


class _AtIdx:
    def __init__(self, arr, idx):
        self.arr, self.idx = arr, idx

    def _apply(self, fn):
        out = np.array(self.arr, copy=True)
        fn(out)
        return _narrow(out)

    def set(self, v):
        v = _weak(v)
        if np.iscomplexobj(v) and not np.iscomplexobj(self.arr):
            v = np.real(v)  # JAX casts (with a warning); the imaginary part is dropped
        return self._apply(lambda o: o.__setitem__(self.idx, v))

    def add(self, v):
        return self._apply(lambda o: o.__setitem__(self.idx, o[self.idx] + _weak(v)))

    def multiply(self, v):
        return self._apply(lambda o: o.__setitem__(self.idx, o[self.idx] * _weak(v)))

    def get(self):
        return _narrow(np.asarray(self.arr)[self.idx])


class _At:
    def __init__(self, arr):
        self.arr = arr

    def __getitem__(self, idx):
        return _AtIdx(self.arr, idx)


class JArr(np.ndarray):
    """ndarray with JAX's functional ``.at`` and float32 / complex64 closure under ufuncs."""

    @property
    def at(self):
        return _At(self)

    # JAX arrays are immutable: `a += b` rebinds to a new (broadcast) array
    def __iadd__(self, o):
        return self + o

    def __isub__(self, o):
        return self - o

    def __imul__(self, o):
        return self * o

    def __itruediv__(self, o):
        return self / o

    def __iter__(self):  # (the sequence protocol would never see an IndexError from the clamped __getitem__)
        if self.ndim == 0:
            raise TypeError("iteration over a 0-d array")
        return (np.ndarray.__getitem__(self, k) for k in range(self.shape[0]))

    def __getitem__(self, idx):
        # JAX clamps out-of-range integer indices instead of raising (the reference relies on it:
        # inv_permittivity_slice[b_axis] on a 1-component array, tfsf.py:298-303)
        tup = idx if isinstance(idx, tuple) else (idx,)
        if any(isinstance(i, (int, np.integer)) and not isinstance(i, bool) for i in tup):
            fixed, ax = [], 0
            for i in tup:
                if i is Ellipsis:
                    ax = self.ndim - (len(tup) - len(fixed) - 1)
                    fixed.append(i)
                    continue
                if i is None:
                    fixed.append(i)
                    continue
                if isinstance(i, (int, np.integer)) and not isinstance(i, bool) and ax < self.ndim:
                    n = self.shape[ax]
                    i = int(i)
                    if i >= n:
                        i = n - 1
                    elif i < -n:
                        i = 0
                fixed.append(i)
                ax += 1
            idx = tuple(fixed) if isinstance(idx, tuple) else fixed[0]
        return super().__getitem__(idx)

    def __array_ufunc__(self, ufunc, method, *inputs, out=None, **kw):
        ins = tuple(np.asarray(_weak(i)) if isinstance(i, np.ndarray) else _weak(i) for i in inputs)
        if out is not None:
            kw["out"] = tuple(np.asarray(o) for o in out)
        res = getattr(ufunc, method)(*ins, **kw)
        return _narrow(res)

    def astype(self, dtype, *a, **k):
        return _narrow(np.asarray(self).astype(dtype, *a, **k)) if np.dtype(dtype) not in (np.float64, np.complex128) else np.asarray(self).astype(dtype, *a, **k).view(JArr)


class _NpProxy(types.ModuleType):
    """``jax.numpy`` / ``jnp.linalg`` / ``jnp.fft``: NumPy functions whose results are narrowed like JAX's."""

    def __init__(self, name, target):
        super().__init__(name)
        self._target = target

    def __getattr__(self, name):
        if name in ("linalg", "fft"):
            return _NpProxy(f"{self.__name__}.{name}", getattr(self._target, name))
        obj = getattr(self._target, name)
        if isinstance(obj, type) or not callable(obj):
            return obj

        def fn(*a, **k):
            a = tuple(_weak(v) if not isinstance(v, (list, tuple)) else type(v)(_weak(q) for q in v) for v in a)
            k = {q: _weak(v) for q, v in k.items()}
            return _narrow(obj(*a, **k))

        fn.__name__ = name
        return fn


def _make_jax():
    jax = types.ModuleType("jax")
    jnp = _NpProxy("jax.numpy", np)
    jax.numpy = jnp
    jax.Array = JArr  # isinstance(x, jax.Array) must hold for the arrays of this stand-in
    lax = types.ModuleType("jax.lax")
    lax.cond = lambda pred, t, f, *ops: t(*ops) if bool(pred) else f(*ops)
    lax.stop_gradient = lambda x: x
    jax.lax = lax
    jax.ShapeDtypeStruct = _Any
    dbg = types.ModuleType("jax.debug")
    dbg.callback = lambda *a, **k: None
    jax.debug = dbg
    return jax, jnp, lax


# ------------------------------------------------------------------------------------------ permissive stubs
class _AnyMeta(abc.ABCMeta):
    def __getattr__(cls, name):
        if name.startswith("__") and name.endswith("__"):
            raise AttributeError(name)
        return _Any

    def __or__(cls, other):
        return _Any

    def __ror__(cls, other):
        return _Any

    def __getitem__(cls, item):
        return _Any

    def __call__(cls, *a, **k):
        if cls is _Any:
            if len(a) == 1 and not k and (isinstance(a[0], type) or callable(a[0])):
                return a[0]  # decorator: @autoinit, @override ...
            return _Any      # field(default=...) and friends
        return super().__call__(*a, **k)


class _Any(metaclass=_AnyMeta):
    """Stands for every name the hot-path files import but never compute with (types, decorators, bases)."""


class _StubModule(types.ModuleType):
    def __getattr__(self, name):
        if name.startswith("__") and name.endswith("__"):
            raise AttributeError(name)
        sub = sys.modules.get(f"{self.__name__}.{name}")  # `from fdtdx import constants`: a module executed for real
        return sub if sub is not None else _Any


class Reference:
    """The reference's hot-path modules, executed from their own source under the stand-ins above."""

    FILES = {
        "fdtdx.constants": "constants.py",
        "fdtdx.core.physics.symmetry": "core/physics/symmetry.py",
        "fdtdx.core.misc": "core/misc.py",
        "fdtdx.core.physics.curl": "core/physics/curl.py",
        "fdtdx.core.physics.metrics": "core/physics/metrics.py",
        "fdtdx.objects.boundaries.boundary": "objects/boundaries/boundary.py",
        "fdtdx.objects.boundaries.perfectly_matched_layer": "objects/boundaries/perfectly_matched_layer.py",
        "fdtdx.objects.boundaries.bloch": "objects/boundaries/bloch.py",
        "fdtdx.objects.boundaries.pec": "objects/boundaries/pec.py",
        "fdtdx.objects.boundaries.pmc": "objects/boundaries/pmc.py",
        "fdtdx.fdtd.misc": "fdtd/misc.py",
        "fdtdx.core.null": "core/null.py",
        "fdtdx.core.linalg": "core/linalg.py",
        "fdtdx.core.axis": "core/axis.py",
        "fdtdx.core.window": "core/window.py",
        "fdtdx.core.wavelength": "core/wavelength.py",
        "fdtdx.core.switch": "core/switch.py",
        "fdtdx.objects.sources.profile": "objects/sources/profile.py",
        "fdtdx.objects.sources.source": "objects/sources/source.py",
        "fdtdx.objects.sources.tfsf": "objects/sources/tfsf.py",
        "fdtdx.objects.sources.dipole": "objects/sources/dipole.py",
        "fdtdx.objects.detectors.detector": "objects/detectors/detector.py",
        "fdtdx.objects.detectors.energy": "objects/detectors/energy.py",
        "fdtdx.objects.detectors.field": "objects/detectors/field.py",
        "fdtdx.objects.detectors.poynting_flux": "objects/detectors/poynting_flux.py",
        "fdtdx.objects.detectors.phasor": "objects/detectors/phasor.py",
        "fdtdx.fdtd.update": "fdtd/update.py",
    }

    def __init__(self):
        if not available():
            raise RuntimeError(f"reference sources not found under {REF_ROOT}")
        self.jax, self.jnp, _ = _make_jax()
        self.modules: dict[str, types.ModuleType] = {}
        saved = dict(sys.modules)
        try:
            for name in list(sys.modules):
                if name == "jax" or name.startswith("jax.") or name == "fdtdx" or name.startswith("fdtdx."):
                    del sys.modules[name]
            sys.modules["jax"] = self.jax
            sys.modules["jax.numpy"] = self.jnp
            sys.modules["jax.lax"] = self.jax.lax
            finder = _StubFinder(self)
            sys.meta_path.insert(0, finder)
            try:
                for name, rel in self.FILES.items():
                    self._exec(name, rel)
            finally:
                sys.meta_path.remove(finder)
        finally:
            for name in list(sys.modules):
                if name not in saved:
                    del sys.modules[name]
            sys.modules.update(saved)
        m = self.modules
        for mod in m.values():  # convenience: ref.update_E, ref.curl_H, ref.interpolate_fields, ...
            for k, v in vars(mod).items():
                if callable(v) and getattr(v, "__module__", None) == mod.__name__ and not k.startswith("__"):
                    setattr(self, k, v)

    def _exec(self, name: str, rel: str):
        path = os.path.join(REF_ROOT, rel)
        with open(path) as f:
            src = f.read()
        ast.parse(src)  # the file is executed exactly as it is on disk
        mod = types.ModuleType(name)
        mod.__file__ = path
        sys.modules[name] = mod
        exec(compile(src, path, "exec"), mod.__dict__)
        self.modules[name] = mod

    # -------------------------------------------------------------------------------------- proxies
    def wrap(self, obj, config=None):
        """This repo's host-mirror object with the reference class's methods on top."""
        table = {
            "PerfectlyMatchedLayer": ("fdtdx.objects.boundaries.perfectly_matched_layer", "PerfectlyMatchedLayer"),
            "BlochBoundary": ("fdtdx.objects.boundaries.bloch", "BlochBoundary"),
            "PerfectElectricConductor": ("fdtdx.objects.boundaries.pec", "PerfectElectricConductor"),
            "PerfectMagneticConductor": ("fdtdx.objects.boundaries.pmc", "PerfectMagneticConductor"),
            "EnergyDetector": ("fdtdx.objects.detectors.energy", "EnergyDetector"),
            "FieldDetector": ("fdtdx.objects.detectors.field", "FieldDetector"),
            "PoyntingFluxDetector": ("fdtdx.objects.detectors.poynting_flux", "PoyntingFluxDetector"),
            "PhasorDetector": ("fdtdx.objects.detectors.phasor", "PhasorDetector"),
            # objects/detectors/mode.py: ModeOverlapDetector(PhasorDetector) keeps PhasorDetector.update (its overlap integral is post-run)
            "ModeOverlapDetector": ("fdtdx.objects.detectors.phasor", "PhasorDetector"),
            "TFSFPlaneSource": ("fdtdx.objects.sources.tfsf", "TFSFPlaneSource"),
            "PointDipoleSource": ("fdtdx.objects.sources.dipole", "PointDipoleSource"),
            "SingleFrequencyProfile": ("fdtdx.objects.sources.profile", "SingleFrequencyProfile"),
            "GaussianPulseProfile": ("fdtdx.objects.sources.profile", "GaussianPulseProfile"),
            "CustomTimeSignalProfile": ("fdtdx.objects.sources.profile", "CustomTimeSignalProfile"),
            "WaveCharacter": ("fdtdx.core.wavelength", "WaveCharacter"),
            "OnOffSwitch": ("fdtdx.core.switch", "OnOffSwitch"),
        }
        key = type(obj).__name__
        if key not in table:
            return obj
        mod, cls = table[key]
        extras = {}
        if key == "PointDipoleSource":
            # the reference caches the oriented material at apply() time; NULL makes update_E / update_H contract it
            # from the arrays they are given (dipole.py:216-222) - the host mirror only has axis-aligned dipoles
            null = self.modules["fdtdx.core.null"].NULL
            extras = {"_inv_eps_oriented": null, "_inv_mu_oriented": null, "azimuth_angle": 0.0, "elevation_angle": 0.0}
        return RefProxy(obj, getattr(self.modules[mod], cls), config, self, extras)

    def wrap_objects(self, objects, config):
        """An ObjectContainer look-alike whose boundaries / detectors carry the reference methods."""
        return _Objects(self, objects, config)


class _StubFinder:
    """Every ``fdtdx.*`` / third-party import the executed files make that is not one of Reference.FILES."""

    def __init__(self, ref):
        self.ref = ref

    def find_spec(self, name, path=None, target=None):
        import importlib.machinery

        root = name.split(".")[0]
        if root in ("fdtdx", "jax", "equinox", "pytreeclass", "tidy3d", "matplotlib", "seaborn", "optax", "moviepy", "trimesh", "gdstk",
                    "loguru", "rich", "tqdm", "PIL", "imageio", "plotly", "jaxtyping", "chex", "orbax"):
            if name in sys.modules:
                return None
            return importlib.machinery.ModuleSpec(name, self, is_package=True)
        return None

    def create_module(self, spec):
        m = _StubModule(spec.name)
        m.__path__ = []
        return m

    def exec_module(self, module):
        pass


class RefProxy:
    """Attribute access goes to the reference class first (methods bound to this proxy, properties evaluated
    on it), then to the wrapped host-mirror object - whose field names are the reference's."""

    def __init__(self, obj, ref_cls, config=None, ref=None, extras=None):
        object.__setattr__(self, "_extras", extras or {})
        object.__setattr__(self, "_obj", obj)
        object.__setattr__(self, "_ref_cls", ref_cls)
        object.__setattr__(self, "_config", config)
        object.__setattr__(self, "_ref", ref)

    def __getattr__(self, name):
        obj, ref_cls = object.__getattribute__(self, "_obj"), object.__getattribute__(self, "_ref_cls")
        if name == "_config":
            return object.__getattribute__(self, "_config")
        extras = object.__getattribute__(self, "_extras")
        if name in extras:
            return extras[name]
        for klass in ref_cls.__mro__:
            if klass in (object, _Any):
                continue
            if name in vars(klass):
                v = vars(klass)[name]
                if v is _Any or (isinstance(v, type) and issubclass(v, _Any)):
                    break  # a stubbed field default: the real value lives on the host object
                if isinstance(v, property):
                    return v.fget(self)
                if hasattr(v, "func") and hasattr(v, "attrname"):  # functools.cached_property
                    return v.func(self)
                if isinstance(v, types.FunctionType):
                    return types.MethodType(v, self)
                if isinstance(v, staticmethod):
                    return v.__func__
                break  # a data default of the reference class: the real value lives on the host object
        if not hasattr(obj, name) and name.startswith("_") and hasattr(obj, name[1:]):
            name = name[1:]  # the reference keeps `_grid_slice_tuple` private; the host mirror's field is public
        val = getattr(obj, name)
        if isinstance(val, np.ndarray):
            return _narrow(val)
        ref = object.__getattribute__(self, "_ref")
        if ref is not None and hasattr(val, "__dict__") and not isinstance(val, (type, types.FunctionType, types.MethodType)):
            return ref.wrap(val, object.__getattribute__(self, "_config"))  # nested host objects (profile, switch, wave character)
        return val

    def __setattr__(self, name, value):
        setattr(object.__getattribute__(self, "_obj"), name, value)


class _Objects:
    def __init__(self, ref, objects, config):
        self._o = objects
        self._w = {id(o): ref.wrap(o, config) for o in objects.object_list}

    def _wl(self, lst):
        return [self._w[id(o)] for o in lst]

    @property
    def volume(self):
        return self._o.volume

    @property
    def boundary_objects(self):
        return self._wl(self._o.boundary_objects)

    @property
    def pml_objects(self):
        return self._wl(self._o.pml_objects)

    @property
    def pec_objects(self):
        return self._wl(self._o.pec_objects)

    @property
    def pmc_objects(self):
        return self._wl(self._o.pmc_objects)

    @property
    def sources(self):
        return self._wl(self._o.sources)

    @property
    def detectors(self):
        return self._wl(self._o.detectors)

    @property
    def forward_detectors(self):
        return self._wl(self._o.forward_detectors)

    @property
    def backward_detectors(self):
        return self._wl(self._o.backward_detectors)

    def __getattr__(self, name):
        return getattr(self._o, name)


def to_jarr(arrays):
    """ArrayContainer of NumPy leaves -> the same container with JArr leaves (functional ``.at``)."""
    return arrays.map_arrays(lambda a: _narrow(np.array(a)) if isinstance(a, np.ndarray) else a)


def reference_pml_tables(ref: Reference, host_pml, config) -> dict:
    """CPML coefficient tables from the reference's own ``PerfectlyMatchedLayer.place_on_grid`` body
    (perfectly_matched_layer.py:97-136, profiles :231-303), run on an instance of the executed class whose user
    fields (axis, direction, slice, sigma/kappa/alpha start/end/order) come from ``host_pml``."""
    RefPML = ref.modules["fdtdx.objects.boundaries.perfectly_matched_layer"].PerfectlyMatchedLayer
    _Any.place_on_grid = lambda self, *a, **k: self  # SimulationObject.place_on_grid: sets slice + config, which Inst serves

    class Inst(RefPML):
        def __getattribute__(self, name):
            if name == "_config":
                return config
            if name == "_grid_slice_tuple":
                return host_pml.grid_slice_tuple
            try:
                v = object.__getattribute__(self, name)
            except AttributeError:
                return getattr(host_pml, name)
            return getattr(host_pml, name) if v is _Any else v

        def aset(self, name, value, **kw):
            object.__setattr__(self, name, value)
            return self

    inst = RefPML.place_on_grid(object.__new__(Inst), host_pml.grid_slice_tuple, config, None)
    return {k: np.asarray(getattr(inst, k)) for k in ("pml_a_E", "pml_b_E", "inv_kappa_E", "pml_a_H", "pml_b_H", "inv_kappa_H")}
