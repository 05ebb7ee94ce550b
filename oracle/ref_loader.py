"""TEST INFRASTRUCTURE ONLY.  Import the *real* reference ``loss.py`` when present.

``/root/reference`` exists only in the build container.  Tests that pin the
oracle against the live reference are skipped when it is absent; the committed
fixtures under ``tests/golden/`` (made by ``oracle/gen_golden.py`` from this
loader) carry the same information to the GPU box.
"""
import importlib.util
import os

REFERENCE_ROOT = os.environ.get("SUPCON_REFERENCE_ROOT", "/root/reference")


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "loss.py"))


def load_reference_module(name: str = "loss"):
    """Load ``/root/reference/<name>.py`` under a private module name so it can
    never shadow (or be shadowed by) the drop-in ``loss`` module of this repo."""
    path = os.path.join(REFERENCE_ROOT, name + ".py")
    if not os.path.isfile(path):
        raise FileNotFoundError(path)
    spec = importlib.util.spec_from_file_location("_reference_" + name, path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod
