"""A small gin-config compatible layer for the subset of gin the reference uses.

The reference configures its trainer with `*.gin` files (co3d_3d/configs/*.gin) and reads them back through
  * `@gin.configurable` on functions and classes, with and without parentheses
    (co3d_3d/train.py:50, src/modules/optim.py:12-14, src/models/__init__.py:18);
  * `gin.parse_config_files_and_bindings(files, bindings)` (train.py:252);
  * `gin.query_parameter("scope.param")` (optim.py:106-110,193; segmentation_training.py:54-55; eval.py:164-165).
The config files only hold `name.param = <python literal>` bindings (possibly spanning several lines), comments and
blank lines — no includes, macros, scopes or `@references` — so that is what this module implements.  It has the
same call surface, so `sys.modules["gin"] = ginlite` lets the reference's `optim.py` run unchanged
(tests/golden/make_schedules.py does exactly that to pin `schedules.py`).

gin-config itself is not installed in this image; this is not a port of it.
"""
from __future__ import annotations

import ast
import functools
import inspect
from typing import Any, Callable, Dict, Iterable, Optional, Sequence, Tuple

REQUIRED = object()

_BINDINGS: Dict[Tuple[str, str], Any] = {}
_REGISTRY: Dict[str, Callable] = {}


class GinError(ValueError):
    pass


# ---- parsing -----------------------------------------------------------------------------------------------
def _scan(text: str):
    """Yield (index, char, inside_string) with backslash escapes inside string literals honoured."""
    quote, escaped = None, False
    for i, ch in enumerate(text):
        if quote:
            yield i, ch, True
            if escaped:
                escaped = False
            elif ch == "\\":
                escaped = True
            elif ch == quote:
                quote = None
        else:
            if ch in "'\"":
                quote = ch
                yield i, ch, True
            else:
                yield i, ch, False


def _strip_comment(line: str) -> str:
    """Remove a trailing `# ...` that is not inside a string literal."""
    for i, ch, in_string in _scan(line):
        if ch == "#" and not in_string:
            return line[:i]
    return line


def _depth(text: str) -> int:
    """Bracket nesting depth at the end of `text`, ignoring brackets inside strings."""
    depth = 0
    for _, ch, in_string in _scan(text):
        if not in_string:
            if ch in "([{":
                depth += 1
            elif ch in ")]}":
                depth -= 1
    return depth


def _statements(text: str) -> Iterable[Tuple[int, str]]:
    buf, start = "", 0
    for no, raw in enumerate(text.splitlines(), 1):
        line = _strip_comment(raw).rstrip()
        if not buf:
            if not line.strip():
                continue
            start = no
        buf += (" " if buf else "") + line.strip()
        if _depth(buf) <= 0 and not buf.endswith("\\") and not buf.endswith("="):
            yield start, buf
            buf = ""
    if buf:
        raise GinError(f"line {start}: unterminated statement: {buf[:60]!r}")


def _split_selector(selector: str) -> Tuple[str, str]:
    selector = selector.strip()
    if "/" in selector or selector.startswith(("@", "%")):
        raise GinError(f"scopes, references and macros are not supported: {selector!r}")
    name, dot, param = selector.rpartition(".")
    if not dot or not name or not param.isidentifier():
        raise GinError(f"expected '<configurable>.<parameter>', got {selector!r}")
    return name.split(".")[-1], param          # module-qualified names bind by their last component


def parse_value(text: str) -> Any:
    text = text.strip()
    try:
        return ast.literal_eval(text)
    except (ValueError, SyntaxError) as e:
        raise GinError(f"not a python literal: {text!r}") from e


def parse_config(text: str) -> None:
    """Parse bindings (`a.b = literal`, one per statement) and add them to the global configuration."""
    for no, stmt in _statements(text):
        if stmt.startswith(("import ", "include ")):
            raise GinError(f"line {no}: '{stmt.split()[0]}' statements are not supported")
        selector, eq, value = stmt.partition("=")
        if not eq:
            raise GinError(f"line {no}: expected a binding, got {stmt!r}")
        try:
            name, param = _split_selector(selector)
            bind_parameter_raw(name, param, parse_value(value))
        except GinError as e:
            raise GinError(f"line {no}: {e}") from None


def parse_config_file(path: str) -> None:
    with open(path) as f:
        try:
            parse_config(f.read())
        except GinError as e:
            raise GinError(f"{path}: {e}") from None


def parse_config_files_and_bindings(config_files: Optional[Sequence[str]], bindings: Optional[Sequence[str]] = None,
                                    finalize_config: bool = True, skip_unknown: bool = False) -> None:
    """train.py:252 — files first (in order), then the `--ginb` bindings, later values overriding earlier ones."""
    for path in config_files or []:
        parse_config_file(path)
    if isinstance(bindings, str):
        bindings = [bindings]
    for b in bindings or []:
        parse_config(b)


# ---- bindings ----------------------------------------------------------------------------------------------
def bind_parameter_raw(name: str, param: str, value: Any) -> None:
    _BINDINGS[(name, param)] = value


def bind_parameter(binding_key: str, value: Any) -> None:
    name, param = _split_selector(binding_key)
    bind_parameter_raw(name, param, value)


def query_parameter(binding_key: str) -> Any:
    name, param = _split_selector(binding_key)
    try:
        return _BINDINGS[(name, param)]
    except KeyError:
        raise ValueError(f"Configurable '{name}' has no bound value for parameter '{param}'.") from None


def clear_config() -> None:
    _BINDINGS.clear()


def config_dict() -> Dict[str, Any]:
    return {f"{n}.{p}": v for (n, p), v in sorted(_BINDINGS.items())}


def config_str() -> str:
    return "".join(f"{k} = {v!r}\n" for k, v in config_dict().items())


# ---- @configurable -----------------------------------------------------------------------------------------
def _bound_kwargs(name: str, fn: Callable, args: tuple, kwargs: dict) -> dict:
    """Bound values for the parameters of `fn` the caller did not pass (callers win over the config)."""
    mine = {p: v for (n, p), v in _BINDINGS.items() if n == name}
    if not mine:
        return {}
    sig = inspect.signature(fn)
    params = sig.parameters
    has_var_kw = any(p.kind is p.VAR_KEYWORD for p in params.values())
    positional = [p.name for p in params.values() if p.kind in (p.POSITIONAL_ONLY, p.POSITIONAL_OR_KEYWORD)]
    passed = set(positional[:len(args)]) | set(kwargs)
    extra = {}
    for p, v in mine.items():
        if p in passed:
            continue
        if p not in params and not has_var_kw:
            raise GinError(f"configurable '{name}' has no parameter '{p}'")
        extra[p] = v
    return extra


def _wrap(obj: Callable, name: str) -> Callable:
    if inspect.isclass(obj):
        orig_init = obj.__init__

        @functools.wraps(orig_init)
        def init(self, *args, **kwargs):
            kwargs = {**_bound_kwargs(name, orig_init, (self, *args), kwargs), **kwargs}
            orig_init(self, *args, **kwargs)

        obj.__init__ = init
        _REGISTRY[name] = obj
        return obj

    @functools.wraps(obj)
    def wrapper(*args, **kwargs):
        kwargs = {**_bound_kwargs(name, obj, args, kwargs), **kwargs}
        missing = [k for k, v in kwargs.items() if v is REQUIRED]
        if missing:
            raise GinError(f"required parameters of '{name}' not bound: {missing}")
        return obj(*args, **kwargs)

    _REGISTRY[name] = wrapper
    return wrapper


def configurable(name_or_fn=None, module: Optional[str] = None, allowlist=None, denylist=None, **_ignored):
    """`@configurable`, `@configurable()` and `@configurable("name")`."""
    if callable(name_or_fn):
        return _wrap(name_or_fn, name_or_fn.__name__)

    def deco(fn):
        return _wrap(fn, name_or_fn or fn.__name__)
    return deco


def get_configurable(name: str) -> Callable:
    return _REGISTRY[name]
