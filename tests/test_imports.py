"""CPU: every product module imports (catches syntax / wiring errors without a GPU)."""
import importlib
import os
import pkgutil


def test_all_modules_import():
    import xuance_b200
    root = os.path.dirname(xuance_b200.__file__)
    mods = [m.name for m in pkgutil.walk_packages([root], "xuance_b200.") if not m.name.endswith("libxb200")]
    assert len(mods) > 20
    for name in mods:
        importlib.import_module(name)
