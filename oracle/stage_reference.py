"""Recipe: stage the UNMODIFIED reference modules of the hot path under baseline/_ref/.

  python oracle/stage_reference.py            (run by __graft_entry__.build() when /root/reference exists)

The reference is a flat Python repo without packaging metadata (no setup.py / pyproject.toml,
`pip install /root/reference` has nothing to build), so "installing" it means placing the two
modules the hot path lives in — progan_modules.py and mnist_pggan.py — where `bench.py --impl
reference` can import them on the GPU box, which has no /root/reference.  baseline/_ref/ is
git-ignored (the reference's sources never enter this repository's history) but travels with
gpurun snapshots.  A MANIFEST with the sha256 of each staged file is written next to them;
tests/test_reference_staging.py checks the staged copies are byte-identical to the reference
whenever both are present.

TEST/BENCH INFRASTRUCTURE ONLY: the product package never imports baseline/_ref."""
import hashlib
import json
import os
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"
DST = os.path.join(ROOT, "baseline", "_ref")
FILES = ["progan_modules.py", "mnist_pggan.py"]


def sha256(path):
    h = hashlib.sha256()
    with open(path, "rb") as f:
        h.update(f.read())
    return h.hexdigest()


def stage(verbose=False):
    """Copy the hot-path modules; returns the destination directory or None if there is no
    reference checkout on this machine."""
    if not os.path.isdir(REF):
        return None
    os.makedirs(DST, exist_ok=True)
    manifest = {}
    for name in FILES:
        src, dst = os.path.join(REF, name), os.path.join(DST, name)
        if not os.path.exists(dst) or sha256(dst) != sha256(src):
            shutil.copyfile(src, dst)
        manifest[name] = sha256(dst)
    with open(os.path.join(DST, "MANIFEST.json"), "w") as f:
        json.dump({"source": REF, "sha256": manifest}, f, indent=1)
    if verbose:
        print("[stage_reference] %s -> %s" % (", ".join(FILES), DST))
    return DST


def staged_dir():
    """baseline/_ref if the reference modules are staged there, else None."""
    return DST if all(os.path.exists(os.path.join(DST, f)) for f in FILES) else None


if __name__ == "__main__":
    d = stage(verbose=True)
    if d is None:
        print("no %s on this machine; nothing staged" % REF)
        sys.exit(0)
