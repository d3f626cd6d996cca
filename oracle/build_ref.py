"""
Populates ``oracle/_ref/`` with the UNMODIFIED upstream package so that the CPU arm of ``bench.py``
(``--impl reference`` and the ``cpu_baseline`` leg) times the reference itself on the GPU box
(geosss/mcmc.py:335-401 run through ``gs.ShrinkageSphericalSliceSampler`` as shipped).

    python oracle/build_ref.py            # needs /root/reference (build container only)

``oracle/_ref/`` is git-ignored (no reference source enters the history) but NOT gpurun-ignored, so it
travels to the GPU box with the snapshot like the built ``.so`` files.  The reference is pure Python with a
poetry-core build backend, which the offline wheelhouse does not hold, so ``pip install --target`` cannot
build it (tried first, outcome recorded in ``oracle/_ref/INSTALL.json``); the package directory and the
registration data file are then copied byte for byte.  Test infrastructure: only ``bench.py``'s CPU legs use it.
"""

import hashlib
import json
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
DEST = os.path.join(HERE, "_ref")
SRC = os.environ.get("GEOSSS_REFERENCE", "/root/reference")


def _sha(path):
    h = hashlib.sha256()
    with open(path, "rb") as f:
        h.update(f.read())
    return h.hexdigest()[:16]


def build(force=False):
    """Returns the path of oracle/_ref when the upstream package is available there, else None."""
    marker = os.path.join(DEST, "INSTALL.json")
    if os.path.exists(marker) and os.path.isdir(os.path.join(DEST, "geosss")) and not force:
        return DEST
    if not os.path.isdir(os.path.join(SRC, "geosss")):
        return None  # GPU box: only a prebuilt oracle/_ref can be used
    shutil.rmtree(DEST, ignore_errors=True)
    os.makedirs(DEST)
    record = {"source": SRC}
    # the contract's route first: an offline pip install of a writable copy of the tree
    tmp = os.path.join("/tmp", "geosss_ref_src")
    shutil.rmtree(tmp, ignore_errors=True)
    shutil.copytree(SRC, tmp, ignore=shutil.ignore_patterns(".git", "assets", "*.ipynb"))
    pip = subprocess.run([sys.executable, "-m", "pip", "install", "--no-index", "--no-build-isolation", "--no-deps",
                          "--find-links", "/opt/wheelhouse", "--target", DEST, tmp], capture_output=True, text=True)
    record["pip_install_rc"] = pip.returncode
    record["pip_install_tail"] = (pip.stderr or pip.stdout).strip().splitlines()[-1:] if pip.returncode else []
    if not os.path.isdir(os.path.join(DEST, "geosss")):
        shutil.copytree(os.path.join(SRC, "geosss"), os.path.join(DEST, "geosss"), ignore=shutil.ignore_patterns("__pycache__"))
        record["method"] = "copied package directory (pure Python)"
    else:
        record["method"] = "pip install --target"
    os.makedirs(os.path.join(DEST, "data"), exist_ok=True)
    shutil.copy2(os.path.join(SRC, "data", "protein_registration.npz"), os.path.join(DEST, "data", "protein_registration.npz"))
    shutil.rmtree(tmp, ignore_errors=True)
    record["files"] = {f: _sha(os.path.join(DEST, "geosss", f)) for f in sorted(os.listdir(os.path.join(DEST, "geosss"))) if f.endswith(".py")}
    for f, digest in record["files"].items():  # byte-identical to the reference tree
        assert digest == _sha(os.path.join(SRC, "geosss", f)), f
    with open(marker, "w") as fh:
        json.dump(record, fh, indent=1)
    return DEST


def available():
    return os.path.isdir(os.path.join(DEST, "geosss"))


if __name__ == "__main__":
    print(build(force="--force" in sys.argv))
