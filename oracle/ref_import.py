"""Import the *unmodified* McQuic reference sources from /root/reference with import-time stubs.

TEST INFRASTRUCTURE ONLY.  This module exists so that `oracle/gen_golden.py` and the
`-m "not gpu"` tests can execute the reference's own hot-path files
(mcquic/modules/compressor.py, mcquic/modules/quantizer.py, mcquic/nn/*.py,
mcquic/data/transforms.py) in this container and pin `oracle/mcquic_oracle.py` against them.
It is never imported by the product (`mcquic_b200/`), by `bench.py`'s GPU arm, or on the GPU box
(where /root/reference does not exist).

Recipe = SURVEY.md Appendix B:
  * third-party packages that are absent here (vlutils, marshmallow, fairscale, distutils) are
    replaced by minimal stand-ins registered in sys.modules *before* anything from mcquic is imported;
  * `mcquic/__init__.py` and `mcquic/data/__init__.py` are bypassed by pre-seeding package objects
    whose __path__ points into the read-only reference tree;
  * `mcquic.rans` (pybind11 extension, not needed by encode/decode) is a dummy module;
  * the three `raise NotImplementedError` gates in mcquic/modules/entropyCoder.py (lines 17, 107,
    140 at reference HEAD 866672cc) are neutralised -- the only deviation from the reference source.
"""
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("MCQUIC_REFERENCE_ROOT", "/root/reference")


def available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "mcquic", "modules"))


def _mod(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


def _pkg(name, path=None, **attrs):
    m = _mod(name, **attrs)
    m.__path__ = [path] if path else []
    return m


class _Registry:
    """Stand-in for vlutils.base.Registry (decorator registry, subscriptable)."""
    _maps = {}

    def __class_getitem__(cls, item):
        return cls

    def __init_subclass__(cls, **kw):
        cls._map = {}

    @classmethod
    def register(cls, key=None):
        if isinstance(key, str):
            def deco(fn):
                cls._map[key] = fn
                return fn
            return deco
        cls._map[key.__name__] = key
        return key

    @classmethod
    def get(cls, key, logger=None):
        return cls._map[key]

    @classmethod
    def summary(cls):
        return str(sorted(cls._map))


def load():
    """Returns the reference `mcquic` package namespace (idempotent)."""
    if "mcquic.modules.compressor" in sys.modules:
        return sys.modules["mcquic"]
    if not available():
        raise RuntimeError(f"reference tree not found at {REFERENCE_ROOT}")
    sys.dont_write_bytecode = True  # the reference tree is read-only

    # ---- third-party stand-ins --------------------------------------------------------------
    class _Restorable:  # vlutils.base.Restorable
        pass

    class _FrequecyHook:  # (sic) vlutils.base.FrequecyHook
        pass

    _pkg("vlutils")
    _pkg("vlutils.base", Registry=_Registry, Restorable=_Restorable, FrequecyHook=_FrequecyHook)
    _mod("vlutils.base.registry", Registry=_Registry)
    _mod("vlutils.logger", readableSize=lambda *a, **k: "", configLogging=lambda *a, **k: None,
         LoggerBase=object)

    class _Field:
        def __init__(self, *a, **k):
            pass

    class _Schema:
        def __init__(self, *a, **k):
            pass

    fields = types.SimpleNamespace(Field=_Field, Int=_Field, Str=_Field, List=_Field, Nested=_Field,
                                   Dict=_Field, Bool=_Field, Float=_Field, Raw=_Field)
    _mod("marshmallow", Schema=_Schema, fields=fields, post_load=lambda f=None, **k: (f if f else (lambda g: g)),
         RAISE="raise", ValidationError=ValueError)

    _pkg("fairscale")
    _pkg("fairscale.nn")
    _pkg("fairscale.nn.checkpoint")
    _mod("fairscale.nn.checkpoint.checkpoint_activations", checkpoint_wrapper=lambda m, *a, **k: m)

    if "distutils" not in sys.modules:
        try:
            import distutils.version  # noqa: F401  (setuptools shim, if present)
        except Exception:
            class StrictVersion:
                def __init__(self, s):
                    self.version = tuple(int(p) for p in s.split("."))

                def __lt__(self, o):
                    return self.version < o.version
            _pkg("distutils")
            _mod("distutils.version", StrictVersion=StrictVersion)

    # ---- the reference package, bypassing its __init__ files ----------------------------------
    ref = os.path.join(REFERENCE_ROOT, "mcquic")
    pkg = _pkg("mcquic", ref, __version__="0.1.40")
    from mcquic.consts import Consts  # reference file, unmodified
    pkg.Consts = Consts
    _pkg("mcquic.data", os.path.join(ref, "data"))
    _mod("mcquic.rans", pmfToQuantizedCDF=None, RansEncoder=lambda: None, RansDecoder=lambda: None)

    # ---- entropyCoder.py with the three gates neutralised --------------------------------------
    path = os.path.join(ref, "modules", "entropyCoder.py")
    with open(path) as fp:
        lines = fp.read().split("\n")
    for ln in (17, 107, 140):
        assert lines[ln - 1].strip() == "raise NotImplementedError", (ln, lines[ln - 1])
        lines[ln - 1] = lines[ln - 1].replace("raise NotImplementedError", "pass")
    import mcquic.modules  # noqa: F401  (namespace from the reference tree)
    ec = _mod("mcquic.modules.entropyCoder", __file__=path)
    exec(compile("\n".join(lines), path, "exec"), ec.__dict__)

    import mcquic.modules.compressor  # noqa: F401
    return sys.modules["mcquic"]


def build_reference_compressor(channel, m, k, seed=0):
    """`Compressor(channel, m, k)` from the reference, random init under `torch.manual_seed(seed)`,
    true fp32."""
    import torch
    load()
    from mcquic.modules.compressor import Compressor
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.manual_seed(seed)
    return Compressor(channel, m, list(k)).eval()
