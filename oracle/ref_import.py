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


def _mini_marshmallow():
    """A FUNCTIONAL minimal stand-in for the parts of marshmallow 3 the reference's schemas use
    (mcquic/utils/specification.py:14-54, mcquic/config.py): declared-field collection in declaration order, `dump` (object
    or mapping -> plain dict, field by field, nested schemas and lists recursed), `load` (unknown keys raise, field
    deserialisation + `_validate`, then the schema's `@post_load` hook).  With it the reference's own `File.serialize` /
    `File.deserialize` run unmodified, so the `.mcq` container bytes can be pinned against them."""

    class ValidationError(ValueError):
        pass

    class Field:
        def __init__(self, *a, **k):
            self.required = k.get("required", False)

        def _serialize(self, value):
            return value

        def _deserialize(self, value):
            return value

        def _validate(self, value):
            pass

        def serialize(self, value):
            return None if value is None else self._serialize(value)

        def deserialize(self, value):
            out = self._deserialize(value)
            self._validate(out)
            return out

    class Int(Field):
        def _serialize(self, value):
            return int(value)

        def _deserialize(self, value):
            if isinstance(value, bool) or not isinstance(value, (int, float, str)):
                raise ValidationError("Not a valid integer.")
            try:
                return int(value)
            except (TypeError, ValueError):
                raise ValidationError("Not a valid integer.")

    class Float(Field):
        def _serialize(self, value):
            return float(value)

        _deserialize = _serialize

    class Bool(Field):
        def _serialize(self, value):
            return bool(value)

        _deserialize = _serialize

    class Str(Field):
        def _serialize(self, value):
            return value.decode() if isinstance(value, bytes) else str(value)

        def _deserialize(self, value):
            if not isinstance(value, (str, bytes)):
                raise ValidationError("Not a valid string.")
            return value.decode() if isinstance(value, bytes) else value

    class List(Field):
        def __init__(self, inner, *a, **k):
            super().__init__(*a, **k)
            self.inner = inner() if isinstance(inner, type) else inner

        def _serialize(self, value):
            return [self.inner.serialize(v) for v in value]

        def _deserialize(self, value):
            if isinstance(value, (str, bytes, dict)) or not hasattr(value, "__iter__"):
                raise ValidationError("Not a valid list.")
            return [self.inner.deserialize(v) for v in value]

    class Dict(Field):
        pass

    class Raw(Field):
        pass

    class Nested(Field):
        def __init__(self, schema, *a, **k):
            super().__init__(*a, **k)
            self.schema = schema() if isinstance(schema, type) else schema

        def _serialize(self, value):
            return self.schema.dump(value)

        def _deserialize(self, value):
            return self.schema.load(value)

    def post_load(fn=None, **kw):
        def mark(f):
            f.__post_load__ = True
            return f
        return mark(fn) if fn is not None else mark

    class SchemaMeta(type):
        def __new__(mcs, name, bases, attrs):
            declared = [(k, v) for k, v in attrs.items() if isinstance(v, Field)]      # class body order
            hooks = [v for v in attrs.values() if getattr(v, "__post_load__", False)]
            for k, _ in declared:
                del attrs[k]
            cls = super().__new__(mcs, name, bases, attrs)
            inherited = []
            for base in bases:
                inherited += getattr(base, "_declared_fields", [])
            cls._declared_fields = inherited + declared
            cls._post_load_hooks = hooks or [h for base in bases for h in getattr(base, "_post_load_hooks", [])]
            return cls

    class Schema(metaclass=SchemaMeta):
        def __init__(self, *a, **k):
            pass

        def dump(self, obj):
            out = {}
            for name, field in self._declared_fields:
                value = obj[name] if isinstance(obj, dict) else getattr(obj, name)
                out[name] = field.serialize(value)
            return out

        def load(self, data):
            if not isinstance(data, dict):
                raise ValidationError("Invalid input type.")
            known = dict(self._declared_fields)
            unknown = [k for k in data if k not in known]
            if unknown:
                raise ValidationError({k: ["Unknown field."] for k in unknown})           # Meta.unknown = RAISE (default)
            out = {name: field.deserialize(data[name]) for name, field in self._declared_fields if name in data}
            for hook in self._post_load_hooks:
                out = hook(self, out)
            return out

    import types as _types
    fields = _types.SimpleNamespace(Field=Field, Int=Int, Integer=Int, Str=Str, String=Str, List=List, Nested=Nested,
                                    Dict=Dict, Bool=Bool, Boolean=Bool, Float=Float, Raw=Raw)
    return dict(Schema=Schema, fields=fields, post_load=post_load, RAISE="raise", ValidationError=ValidationError)


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
    def _readable_size(size, floating=2):      # vlutils.logger.readableSize: binary units, two decimals
        value = float(size)
        for unit in ("B", "KiB", "MiB", "GiB", "TiB"):
            if value < 1024.0 or unit == "TiB":
                return f"{int(value)} {unit}" if unit == "B" else f"{value:.{floating}f} {unit}"
            value /= 1024.0

    sys.modules["vlutils"].logger = _mod("vlutils.logger", readableSize=_readable_size,
                                         configLogging=lambda *a, **k: None, LoggerBase=object)

    _mod("marshmallow", **_mini_marshmallow())

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
