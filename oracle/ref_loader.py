"""Locate and import the live reference (build container only).  TEST INFRASTRUCTURE."""
import os
import sys

REFERENCE_ROOT = os.environ.get("XB_REFERENCE_ROOT", "/root/reference")
_STUBS = os.path.join(os.path.dirname(os.path.abspath(__file__)), "ref_stubs")


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "xuance"))


def import_reference():
    """Put the stubs (gymnasium, pyglet) and the reference on sys.path and import ``xuance``.

    The reference stays read-only and unmodified; the stubs live in this repo."""
    if not reference_available():
        raise RuntimeError("reference not present at %s" % REFERENCE_ROOT)
    try:
        import gymnasium  # noqa: F401  (a real gymnasium wins if one is installed)
    except Exception:
        if _STUBS not in sys.path:
            sys.path.insert(0, _STUBS)
    if _STUBS not in sys.path:
        sys.path.append(_STUBS)  # pyglet stub
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(1, REFERENCE_ROOT)
    import xuance  # noqa: F401
    return xuance
