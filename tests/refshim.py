"""Test-only helper: make the *reference* (`/root/reference`, lhotse) importable in the build
container, where `soundfile`, `intervaltree` and `cytoolz` are absent.
Never used on the GPU box (the reference does not exist there) and never by the product."""
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("LHOTSE_REFERENCE_ROOT", "/root/reference")


class _Stub(types.ModuleType):
    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        return type(name, (), {})


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "lhotse"))


def import_reference():
    """Returns the imported `lhotse` package from the reference tree (or raises ImportError)."""
    if not reference_available():
        raise ImportError("reference tree not present")
    for m in ("soundfile", "intervaltree", "cytoolz"):
        if m not in sys.modules:
            try:
                __import__(m)
            except Exception:
                sys.modules[m] = _Stub(m)
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    sys.dont_write_bytecode = True
    import lhotse  # noqa

    return lhotse
