"""Loader for the UNMODIFIED reference (read-only, /root/reference) -- test infrastructure only.

Only usable in the build container (the GPU box has no /root/reference).  Used by
tests/golden/make_golden.py to generate the committed fixtures and by the optional
`reference`-marked CPU tests that pin oracle/ against the live reference.

The reference keeps every size as an import-time module constant (libs/config.py:19-65) and
imports matplotlib at module level (libs/utils.py:5), so we (1) stub matplotlib, (2) exec a
regex-patched copy of config.py into sys.modules['libs.config'], (3) exec libs/__init__.py.
"""
import contextlib
import io
import os
import re
import sys
import types

REF_ROOT = os.environ.get("LOCATE_REFERENCE_ROOT", "/root/reference")


def available() -> bool:
    return os.path.isfile(os.path.join(REF_ROOT, "libs", "config.py"))


def load_ref(**overrides):
    """Return the reference `libs` package built with patched config constants."""
    if not available():
        raise RuntimeError(f"reference not mounted at {REF_ROOT}")
    for name in [m for m in sys.modules if m == "libs" or m.startswith("libs.")]:
        del sys.modules[name]
    if "matplotlib" not in sys.modules:
        mpl, plt = types.ModuleType("matplotlib"), types.ModuleType("matplotlib.pyplot")
        mpl.pyplot = plt
        sys.modules["matplotlib"], sys.modules["matplotlib.pyplot"] = mpl, plt
    pkg = types.ModuleType("libs")
    pkg.__path__ = [os.path.join(REF_ROOT, "libs")]
    pkg.__package__ = "libs"
    sys.modules["libs"] = pkg
    with open(os.path.join(REF_ROOT, "libs", "config.py")) as fh:
        src = fh.read()
    for key, val in overrides.items():
        src, hits = re.subn(rf"^{key} = .*$", f"{key} = {val!r}", src, flags=re.M)
        assert hits == 1, f"config constant {key} not found"
    cfg = types.ModuleType("libs.config")
    cfg.__package__ = "libs"
    with contextlib.redirect_stdout(io.StringIO()):
        exec(compile(src, "config.py", "exec"), cfg.__dict__)
        sys.modules["libs.config"] = cfg
        pkg.config = cfg
        with open(os.path.join(REF_ROOT, "libs", "__init__.py")) as fh:
            exec(compile(fh.read(), "__init__.py", "exec"), pkg.__dict__)
    return pkg


@contextlib.contextmanager
def quiet():
    with contextlib.redirect_stdout(io.StringIO()):
        yield
