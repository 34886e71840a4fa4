"""Mesh generators with the reference's signatures (fedoo/mesh/simple.py:460-812), vectorised."""

from __future__ import annotations

from . import meshgen
from .core import Mesh


def box_mesh(nx=11, ny=11, nz=11, x_min=0, x_max=1, y_min=0, y_max=1, z_min=0, z_max=1, elm_type="hex8", name=""):
    """Structured hex8 box, same node/element numbering and node sets as ``fd.mesh.box_mesh``."""
    if elm_type != "hex8":
        raise NameError("Element not implemented. Only support hex8 elements")
    nodes, elements = meshgen.box_hex8(nx, ny, nz, x_min, x_max, y_min, y_max, z_min, z_max)
    return Mesh(nodes, elements, "hex8", meshgen.box_node_sets(nx, ny, nz), name=name)


def rectangle_mesh(nx=11, ny=11, x_min=0, x_max=1, y_min=0, y_max=1, elm_type="quad4", name=""):
    if elm_type != "quad4":
        raise NameError("Element not implemented. Only support quad4 elements")
    nodes, elements = meshgen.rect_quad4(nx, ny, x_min, x_max, y_min, y_max)
    return Mesh(nodes, elements, "quad4", name=name)
