"""Recipe for ``oracle/_ref/``: a byte-for-byte staging of the reference's own Python package so that
``bench.py --impl reference`` (and the ``cpu_baseline`` leg) can time the UNMODIFIED reference classes on the GPU box's
host cores, where ``/root/reference`` does not exist.

TEST / BASELINE INFRASTRUCTURE ONLY (see oracle/__init__.py): nothing under ``ursabench_b200/`` imports it.  The
reference is pure Python (no native code to compile), so "building" it is copying its ``*.py`` files -- and the
hyper-parameter JSONs its classes read -- from where they lie under ``/root/reference`` into ``oracle/_ref/URSABench``.
``oracle/_ref/`` is git-ignored (no reference source enters the history) but NOT gpurun-ignored, so it travels to the
box next to the built ``.so``.  ``__graft_entry__.build()`` runs this when ``/root/reference`` is present.

    python oracle/make_ref.py            # -> oracle/_ref/URSABench + oracle/_ref/MANIFEST.json (sha256 per file)
"""
import hashlib
import json
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC_ROOT = os.environ.get("URSA_REFERENCE_ROOT", "/root/reference")
DST_ROOT = os.path.join(HERE, "_ref")
KEEP_EXT = (".py", ".json")
SKIP_DIRS = {"examples", "trtprof", "__pycache__"}        # notebooks / Jetson profiling: not on the path, not needed to import


def staged_root():
    """Path to put on sys.path for ``import URSABench`` from the staged copy, or None if it was never built."""
    return DST_ROOT if os.path.isfile(os.path.join(DST_ROOT, "URSABench", "__init__.py")) else None


def make(verbose=False):
    src_pkg = os.path.join(SRC_ROOT, "URSABench")
    if not os.path.isdir(src_pkg):
        return None
    dst_pkg = os.path.join(DST_ROOT, "URSABench")
    if os.path.isdir(dst_pkg):
        shutil.rmtree(dst_pkg)
    manifest = {}
    for d, dirs, files in os.walk(src_pkg):
        dirs[:] = sorted(x for x in dirs if x not in SKIP_DIRS)
        rel = os.path.relpath(d, src_pkg)
        out_dir = os.path.normpath(os.path.join(dst_pkg, rel))
        os.makedirs(out_dir, exist_ok=True)
        for f in sorted(files):
            if not f.endswith(KEEP_EXT):
                continue
            s, t = os.path.join(d, f), os.path.join(out_dir, f)
            shutil.copyfile(s, t)
            manifest[os.path.normpath(os.path.join(rel, f))] = hashlib.sha256(open(t, "rb").read()).hexdigest()
    json.dump({"source": src_pkg, "files": manifest}, open(os.path.join(DST_ROOT, "MANIFEST.json"), "w"), indent=1, sort_keys=True)
    if verbose:
        print("oracle/_ref: staged %d files of the unmodified reference from %s" % (len(manifest), src_pkg))
    return DST_ROOT


if __name__ == "__main__":
    sys.exit(0 if make(verbose=True) else 1)
