"""Recipe for ``oracle/_ref``: the UNMODIFIED reference, placed where it can travel to the GPU box.

TEST / BENCH INFRASTRUCTURE ONLY.  fedoo is pure Python (``/root/reference/pyproject.toml:27-30``: numpy + scipy are its
only hard dependencies), so "building" the reference is a file copy: this script copies the package ``fedoo/``, the
three reference tests that can run without optional dependencies and the mesh files they read from
``/root/reference`` into ``oracle/_ref/`` (git-ignored, NOT gpurun-ignored -- it ships to the GPU box like the built
``_fdk.so``).  Nothing is edited; no reference source enters the repository history.

Used by
  * ``bench.py --impl reference`` and ``bench.py``'s ``cpu_baseline`` leg: the real ``Assembly.assemble_global_mat``
    timed on the box's host cores (``kind: "reference"``);
  * ``tests/test_adapter_gpu.py``: the bodies of the reference's own tests run with ``import fedoo`` from here after
    ``fedoo_b200.install(fedoo)`` has put the CUDA path under ``fedoo.Assembly``;
  * the golden generators (``oracle/gen_golden*.py``) import ``/root/reference`` directly instead.

Run by ``__graft_entry__.build()`` whenever ``/root/reference`` exists (this container); on the GPU box the copy made
here is used as is.
"""

from __future__ import annotations

import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF_SRC = os.environ.get("FEDOO_REFERENCE", "/root/reference")
REF_DST = os.path.join(HERE, "_ref")

TESTS = ["test_cantilever_beam_3D_model.py", "test_platewithhol.py", "test_thermal3D.py", "test_octet.py",
         "test_periodic.py", "gyroid.msh", "octet_truss.msh"]  # fmt: skip
MESHES = ["octet_truss_quad.msh"]


def ref_path():
    """Directory to put on sys.path to ``import fedoo`` (the copy), or None when it has not been made."""
    return REF_DST if os.path.isdir(os.path.join(REF_DST, "fedoo")) else None


def make(force=False, verbose=True):
    if not os.path.isdir(os.path.join(REF_SRC, "fedoo")):
        if verbose:
            print(f"make_ref: {REF_SRC} not present; using the existing copy: {ref_path()}")
        return ref_path()
    stamp = os.path.join(REF_DST, ".stamp")
    if not force and os.path.exists(stamp):
        return REF_DST
    if os.path.isdir(REF_DST):
        shutil.rmtree(REF_DST)
    ignore = shutil.ignore_patterns("__pycache__", "*.pyc", ".DS_Store", "_viewer")
    shutil.copytree(os.path.join(REF_SRC, "fedoo"), os.path.join(REF_DST, "fedoo"), ignore=ignore)
    os.makedirs(os.path.join(REF_DST, "tests"), exist_ok=True)
    for f in TESTS:
        shutil.copy2(os.path.join(REF_SRC, "tests", f), os.path.join(REF_DST, "tests", f))
    os.makedirs(os.path.join(REF_DST, "util", "meshes"), exist_ok=True)
    for f in MESHES:
        shutil.copy2(os.path.join(REF_SRC, "util", "meshes", f), os.path.join(REF_DST, "util", "meshes", f))
    with open(stamp, "w") as fh:
        fh.write(f"copied from {REF_SRC}\n")
    if verbose:
        print(f"make_ref: copied the reference to {REF_DST}")
    return REF_DST


def import_fedoo():
    """``import fedoo`` from the copy (bench reference arm, adapter tests).  Raises when the copy is missing."""
    p = ref_path()
    if p is None:
        raise ImportError("oracle/_ref/fedoo is missing: run `python oracle/make_ref.py` where /root/reference exists")
    if p not in sys.path:
        sys.path.insert(0, p)
    import fedoo

    return fedoo


if __name__ == "__main__":
    make(force="--force" in sys.argv)
