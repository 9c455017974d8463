"""Import the reference's model files UNCHANGED on top of the ME-compatible surface.

Recipe from SURVEY.md §8b (import-harness caveat): put /root/reference on sys.path, pre-seed a stub
`gin` and synthetic parent packages `co3d_3d.src` / `co3d_3d.src.models` (so their heavyweight
__init__.py files never run), then import `co3d_3d.src.models.mink.resnet` / `.res16unet` normally.
Nothing is copied out of /root/reference.
"""
import importlib
import sys
import types
from pathlib import Path

REF = Path("/root/reference")


def available() -> bool:
    return (REF / "co3d_3d/src/models/mink/resnet.py").exists()


def _gin_stub():
    gin = types.ModuleType("gin")

    def configurable(*args, **kwargs):
        if len(args) == 1 and callable(args[0]) and not kwargs:
            return args[0]

        def deco(fn):
            return fn
        return deco

    gin.configurable = configurable
    gin.query_parameter = lambda name: None
    gin.parse_config_files_and_bindings = lambda *a, **k: None
    gin.REQUIRED = object()
    return gin


def install():
    root = str(Path(__file__).resolve().parents[1])
    if root not in sys.path:
        sys.path.insert(0, root)
    if str(REF) not in sys.path:
        sys.path.append(str(REF))
    sys.modules.setdefault("gin", _gin_stub())
    for name, rel in [("co3d_3d", "co3d_3d"), ("co3d_3d.src", "co3d_3d/src"),
                      ("co3d_3d.src.models", "co3d_3d/src/models")]:
        if name not in sys.modules:
            mod = types.ModuleType(name)
            mod.__path__ = [str(REF / rel)]
            sys.modules[name] = mod


def load(module: str):
    install()
    return importlib.import_module(module)
