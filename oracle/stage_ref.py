"""Stage the UNMODIFIED reference hot-path files for the CPU arm of bench.py (TEST / BASELINE INFRASTRUCTURE ONLY).

The reference is pure Python, so there is nothing to compile: what travels to the GPU box is one opaque archive of the
files the path consists of, exactly as they lie under /root/reference --

    Env/1D/DMP_Env_1D_static.py                      Env/1D/DMP_Env_1D_dynamic_usedata_plan.py
    Env/2D/DMP_Env_2D_static.py                      Env/2D/DMP_Env_2D_dynamic_usedata_plan.py
    Env/3D/DMP_simulator_3d_static_circle.py         Env/3D/DMP_simulator_3d_dynamic_triangle_usedata.py
    multiprocess.py                                  Env/*/data_*_envplan_500_train.pkl (the plan datasets)

-- written to the git-ignored ``oracle/_ref/snac_reference_hotpath.tar.gz`` (never into history; gpurun ships ignored
files).  ``unpack()`` extracts it into a fresh temporary directory and returns that directory, which
``oracle.refload.use_root()`` then imports from through its gym / matplotlib stubs.  Nothing under snac_b200/ touches
any of this; only ``bench.py --impl reference`` / the ``cpu_baseline`` leg do.

    python oracle/stage_ref.py            # (re)build the archive; needs /root/reference
"""
from __future__ import annotations

import io
import os
import tarfile
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
REF_SRC = os.environ.get("SNAC_REFERENCE", "/root/reference")
ARCHIVE = os.path.join(HERE, "_ref", "snac_reference_hotpath.tar.gz")

FILES = [
    "Env/1D/DMP_Env_1D_static.py", "Env/1D/DMP_Env_1D_dynamic_usedata_plan.py",
    "Env/2D/DMP_Env_2D_static.py", "Env/2D/DMP_Env_2D_dynamic_usedata_plan.py",
    "Env/3D/DMP_simulator_3d_static_circle.py", "Env/3D/DMP_simulator_3d_dynamic_triangle_usedata.py",
    "multiprocess.py",
    "Env/1D/data_1d_dynamic_sin_envplan_500_train.pkl",
    "Env/2D/data_2d_dynamic_dense_envplan_500_train.pkl", "Env/2D/data_2d_dynamic_sparse_envplan_500_train.pkl",
    "Env/3D/data_3d_dynamic_dense_envplan_500_train.pkl", "Env/3D/data_3d_dynamic_sparse_envplan_500_train.pkl",
]


def source_available() -> bool:
    return all(os.path.exists(os.path.join(REF_SRC, f)) for f in FILES)


def staged() -> bool:
    return os.path.exists(ARCHIVE)


def stage(force: bool = False) -> str:
    """Build the archive from /root/reference (build container only).  Returns its path."""
    if staged() and not force:
        return ARCHIVE
    if not source_available():
        raise RuntimeError("reference tree not present at %s" % REF_SRC)
    os.makedirs(os.path.dirname(ARCHIVE), exist_ok=True)
    tmp = ARCHIVE + ".tmp"
    with tarfile.open(tmp, "w:gz") as tar:
        for f in FILES:
            with open(os.path.join(REF_SRC, f), "rb") as fh:
                data = fh.read()
            info = tarfile.TarInfo(f)
            info.size, info.mtime, info.mode = len(data), 0, 0o644
            tar.addfile(info, io.BytesIO(data))
    os.replace(tmp, ARCHIVE)
    return ARCHIVE


def unpack() -> str:
    """Extract the staged archive into a fresh temporary directory (the caller's process owns it)."""
    if not staged():
        raise RuntimeError("no staged reference archive at %s (run oracle/stage_ref.py in the build container)" % ARCHIVE)
    root = tempfile.mkdtemp(prefix="snac_ref_")
    with tarfile.open(ARCHIVE, "r:gz") as tar:
        for m in tar.getmembers():
            if m.name not in FILES:
                raise RuntimeError("unexpected member %r in %s" % (m.name, ARCHIVE))
        tar.extractall(root)
    return root


if __name__ == "__main__":
    print(stage(force=True))
