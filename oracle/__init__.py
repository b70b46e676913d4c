"""CPU oracles -- TEST INFRASTRUCTURE ONLY.

Nothing under ``topo4d_b200/`` or ``diff_gaussian_rasterization/`` may import this
package.  Allowed importers: ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py``.
"""
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))


def build(quiet: bool = True) -> None:
    """Run the committed recipe (oracle/Makefile).  Safe to call repeatedly."""
    subprocess.run(["make", "-C", _HERE, "all"], check=True,
                   stdout=subprocess.DEVNULL if quiet else None)


def lib_path(name: str) -> str:
    """Path of a built oracle library, building it on first use."""
    for sub in ("_build", "_ref"):
        p = os.path.join(_HERE, sub, name)
        if os.path.exists(p):
            return p
    build()
    for sub in ("_build", "_ref"):
        p = os.path.join(_HERE, sub, name)
        if os.path.exists(p):
            return p
    raise FileNotFoundError(name)
