"""Stand-ins for the reference's missing third-party imports (oracle; test infrastructure).

``install()`` registers fake ``e3nn``, ``ml_collections``, ``torch_runstats``, ``ase`` and
``h5py`` modules in ``sys.modules`` so that the UNMODIFIED reference package
(``/root/reference/e3_layers``) can be imported and executed in the build container, with
``import e3nn`` resolving to the restatement in ``oracle/e3nn_ops.py``.  This is how
``tests/golden/make_golden.py`` produces fixtures that are pinned by the reference's own code.
Never used on the GPU box (``/root/reference`` does not exist there) and never by the product.
"""
import sys
import types

import torch

from . import e3nn_ops, irreps, wigner


class ConfigDict(object):
    """Minimal ml_collections.ConfigDict: attribute + item access, nested dict conversion."""

    def __init__(self, initial=None, **kw):
        object.__setattr__(self, "_fields", {})
        if initial is not None:
            for k, v in dict(initial.items() if hasattr(initial, "items") else initial).items():
                self[k] = v
        for k, v in kw.items():
            self[k] = v

    def __setitem__(self, k, v):
        if isinstance(v, dict):
            v = ConfigDict(v)
        self._fields[k] = v

    def __getitem__(self, k):
        return self._fields[k]

    def __setattr__(self, k, v):
        self[k] = v

    def __getattr__(self, k):
        try:
            return object.__getattribute__(self, "_fields")[k]
        except KeyError:
            raise AttributeError(k)

    def __contains__(self, k):
        return k in self._fields

    def __iter__(self):
        return iter(self._fields)

    def __len__(self):
        return len(self._fields)

    def keys(self):
        return self._fields.keys()

    def values(self):
        return self._fields.values()

    def items(self):
        return self._fields.items()

    def get(self, k, default=None):
        return self._fields.get(k, default)

    def pop(self, k, *d):
        return self._fields.pop(k, *d)

    def update(self, *other, **kw):
        for o in other:
            for k, v in (o.items() if hasattr(o, "items") else o):
                if k in self._fields and isinstance(self._fields[k], ConfigDict) and isinstance(v, (dict, ConfigDict)):
                    self._fields[k].update(v)
                else:
                    self[k] = v
        for k, v in kw.items():
            self[k] = v

    def update_from_flattened_dict(self, flat, strip_prefix=""):
        for k, v in flat.items():
            node = self
            parts = k[len(strip_prefix):].split(".")
            for p in parts[:-1]:
                node = node[p]
            node[parts[-1]] = v

    def to_dict(self):
        return {k: (v.to_dict() if isinstance(v, ConfigDict) else v) for k, v in self._fields.items()}

    def __repr__(self):
        return f"ConfigDict({self._fields!r})"


_ELEMENTS = (
    "X H He Li Be B C N O F Ne Na Mg Al Si P S Cl Ar K Ca Sc Ti V Cr Mn Fe Co Ni Cu Zn Ga Ge "
    "As Se Br Kr Rb Sr Y Zr Nb Mo Tc Ru Rh Pd Ag Cd In Sn Sb Te I Xe"
).split()


def _compile_mode(mode):
    def deco(cls):
        return cls
    return deco


def install():
    if "e3nn" in sys.modules and getattr(sys.modules["e3nn"], "_oracle_shim", False):
        return

    def mod(name):
        m = types.ModuleType(name)
        sys.modules[name] = m
        return m

    e3nn = mod("e3nn")
    e3nn._oracle_shim = True
    o3 = mod("e3nn.o3")
    for n in ("Irrep", "Irreps"):
        setattr(o3, n, getattr(irreps, n))
    for n in ("Linear", "TensorProduct", "FullyConnectedTensorProduct", "SphericalHarmonics",
              "ElementwiseTensorProduct"):
        setattr(o3, n, getattr(e3nn_ops, n))
    o3.wigner_3j = wigner.wigner_3j
    o3.rand_matrix = lambda *shape: wigner.rand_rotation()
    e3nn.o3 = o3
    enn = mod("e3nn.nn")
    for n in ("Gate", "NormActivation", "FullyConnectedNet", "Activation"):
        setattr(enn, n, getattr(e3nn_ops, n))
    e3nn.nn = enn
    util = mod("e3nn.util")
    jit = mod("e3nn.util.jit")
    jit.compile_mode = _compile_mode
    jit.script = lambda m: m
    jit.trace = lambda m, *a, **k: m
    util.jit = jit
    e3nn.util = util
    emath = mod("e3nn.math")
    emath.soft_one_hot_linspace = e3nn_ops.soft_one_hot_linspace
    emath.normalize2mom = e3nn_ops.normalize2mom
    e3nn.math = emath

    mlc = mod("ml_collections")
    cd = mod("ml_collections.config_dict")
    cd.ConfigDict = ConfigDict
    mlc.config_dict = cd
    mlc.ConfigDict = ConfigDict

    tr = mod("torch_runstats")
    trs = mod("torch_runstats.scatter")
    trs.scatter = e3nn_ops.scatter
    trs.scatter_std = None
    trs.scatter_mean = None
    tr.scatter = trs

    ase = mod("ase")
    atom = mod("ase.atom")
    atom.atomic_numbers = {s: i for i, s in enumerate(_ELEMENTS)}
    ase.atom = atom
    nl = mod("ase.neighborlist")
    ase.neighborlist = nl

    if "h5py" not in sys.modules:
        try:
            import h5py  # noqa: F401
        except Exception:
            mod("h5py")


def import_reference(path="/root/reference"):
    """Import the genuine reference package on top of the shims; returns the module."""
    install()
    if path not in sys.path:
        sys.path.insert(0, path)
    # the product mirror is also called e3_layers: make sure we get the reference one
    for k in [k for k in sys.modules if k == "e3_layers" or k.startswith("e3_layers.")]:
        del sys.modules[k]
    import e3_layers  # noqa: F401
    import e3_layers.nn, e3_layers.data, e3_layers.utils, e3_layers.configs  # noqa: F401,E401
    mod = sys.modules["e3_layers"]
    assert mod.__file__.startswith(path), mod.__file__
    return mod
