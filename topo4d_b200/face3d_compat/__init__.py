"""Drop-in twins of the two face3d modules on Topo4D's bake path:

    face3d.mesh.cython.mesh_core_cython   -> topo4d_b200.face3d_compat.mesh_core_cython
    face3d.mesh.render                    -> topo4d_b200.face3d_compat.render

A maintainer switches by replacing `from .cython import mesh_core_cython` in
face3d/mesh/render.py:21; signatures and in-place semantics are identical.
"""
from . import mesh_core_cython, render  # noqa: F401
