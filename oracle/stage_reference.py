"""TEST / BENCH INFRASTRUCTURE ONLY -- never imported by the product package.

Stages the UNMODIFIED reference files of the hot path into the git-ignored `baseline/_ref/` so that they travel to the
GPU box with the gpurun snapshot (`/root/reference` does not exist there) and `bench.py --impl reference` / the
`cpu_baseline` leg can time the reference's own modules on the box's host cores:

    python/niantic/modules/{my_gnn_layer,att,posenet,criterion}.py, python/niantic/utils/pose_utils.py (+ __init__.py)

Nothing is edited and nothing lands in the tracked tree (`.gitignore`: baseline/_ref/).  Run by `__graft_entry__.build()`
whenever /root/reference is present; `reference_model()` then builds PoseNetX_R2 (posenet.py:920-1091) through
oracle/pyg_shim.py (stand-ins for the uninstallable torch_geometric / torch_cluster) with a stub feature extractor,
exactly like oracle/make_golden.py does for the fixtures.
"""
import os
import shutil
import sys
import types

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.environ.get("RPG_REFERENCE_PYTHON", "/root/reference/python")
DST = os.path.join(ROOT, "baseline", "_ref", "python")
FILES = ["niantic/modules/__init__.py", "niantic/modules/my_gnn_layer.py", "niantic/modules/att.py",
         "niantic/modules/posenet.py", "niantic/modules/criterion.py", "niantic/utils/__init__.py",
         "niantic/utils/pose_utils.py"]


def stage():
    """Copies the files (if the reference tree is present).  Returns the staged python/ directory or None."""
    if not os.path.isfile(os.path.join(SRC, FILES[1])):
        return DST if staged() else None
    for f in FILES:
        dst = os.path.join(DST, f)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        src = os.path.join(SRC, f)
        if os.path.isfile(src):
            shutil.copyfile(src, dst)
        elif f.endswith("__init__.py"):
            open(dst, "a").close()
    return DST


def staged():
    return os.path.isfile(os.path.join(DST, FILES[1]))


def reference_python_dir():
    """The reference's python/ directory to import from: the real tree in the build container, the staged copy elsewhere."""
    if os.path.isfile(os.path.join(SRC, FILES[1])):
        return SRC
    return DST if staged() else None


def import_reference_modules():
    """(posenet, criterion) modules of the unmodified reference, imported through the shim; None if unavailable."""
    d = reference_python_dir()
    if d is None:
        return None
    from oracle import pyg_shim
    pyg_shim.REFERENCE_PYTHON = d
    pyg_shim.install()
    sys.modules.setdefault("transforms3d", types.ModuleType("transforms3d"))      # pose_utils.py:7-8 (unused by the path)
    for sub in ("euler", "quaternions"):
        sys.modules.setdefault("transforms3d." + sub, types.ModuleType("transforms3d." + sub))
        setattr(sys.modules["transforms3d"], sub, sys.modules["transforms3d." + sub])
    from niantic.modules import criterion, posenet
    return posenet, criterion


if __name__ == "__main__":
    print("staged:", stage())
