"""Make the UNMODIFIED Python reference (/root/reference) importable in the build container.

TEST INFRASTRUCTURE ONLY.  Used by tests/golden/make_golden.py (to generate the committed golden
vectors) and by optional cross-checks that skip when /root/reference is absent (it does not exist
on the GPU box).  Nothing in colosseum_b200/ imports this.

What it does: puts oracle/ref_shim (stand-ins for dm_env, gin, toolz, pydtmc,
sparse, gym -- all absent from the image) and /root/reference on sys.path, and restores two names the
reference needs that numpy-2 / py3.12 dropped (`numpy.core._exceptions`, `collections.Container`).
"""
import collections
import collections.abc
import os
import sys
import types

# /root/reference in the build container; on the GPU box (where that path does not exist) the same UNMODIFIED package
# as installed once by `pip install --no-index --no-deps --target baseline/_ref /root/reference` (git-ignored, travels
# with the gpurun snapshot) -- used only by the live drop-in test tests/test_gpu_dropin_live.py
_REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REFERENCE_ROOT = os.environ.get("COLOSSEUM_REFERENCE_ROOT", "/root/reference")
if not os.path.isdir(os.path.join(REFERENCE_ROOT, "colosseum")) and os.path.isdir(os.path.join(_REPO, "baseline", "_ref", "colosseum")):
    REFERENCE_ROOT = os.path.join(_REPO, "baseline", "_ref")
_SHIM = os.path.join(os.path.dirname(os.path.abspath(__file__)), "ref_shim")


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "colosseum"))


def import_reference():
    """Returns the imported `colosseum` reference package (raises if /root/reference is absent)."""
    if not reference_available():
        raise ImportError(f"reference not present at {REFERENCE_ROOT}")
    for p in (_SHIM, REFERENCE_ROOT):
        if p not in sys.path:
            sys.path.insert(0, p)
    if not hasattr(collections, "Container"):
        collections.Container = collections.abc.Container
    import numpy as np

    if "numpy.core._exceptions" not in sys.modules:
        try:
            import numpy._core._exceptions as _exc
        except Exception:  # pragma: no cover
            _exc = types.ModuleType("numpy.core._exceptions")
            _exc._ArrayMemoryError = MemoryError
        sys.modules["numpy.core._exceptions"] = _exc
    # importing the reference copies its hardness cache (3,007 files) and creates `tmp/` in the CURRENT directory
    # (colosseum/config.py:255-270): do the import from a scratch directory so nothing lands in the repository
    import tempfile

    cwd = os.getcwd()
    os.chdir(tempfile.mkdtemp(prefix="colosseum_ref_"))
    try:
        import colosseum  # noqa: E402
    finally:
        os.chdir(cwd)

    colosseum.config.disable_multiprocessing()
    colosseum.config.VERBOSE_LEVEL = 0
    return colosseum
