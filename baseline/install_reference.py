#!/usr/bin/env python
"""Recipe that ships the UNMODIFIED reference to the GPU box: /root/reference -> baseline/_ref/.

The reference (sabarim/STEm-Seg) has no setup.py / pyproject, so `pip install --target baseline/_ref /root/reference`
cannot work; it is pure Python, so "installing" it is copying its package tree.  baseline/_ref/ is git-ignored (no
reference source ever enters this repository's history) but NOT gpurun-ignored, so it travels with the snapshot and
`bench.py --impl reference` / the `-m gpu` reference-integration tests can import `stemseg` there.

    python baseline/install_reference.py            # copy if /root/reference is present, no-op otherwise

Called by ``__graft_entry__.build()``.  Nothing is edited: a manifest with the sha256 of every copied file is written
next to the tree (baseline/_ref/MANIFEST.json) and `verify()` re-hashes it, so "unmodified" is checkable on the box.
"""
import hashlib
import json
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
DEST = os.path.join(HERE, "_ref")
SOURCE = os.environ.get("STEMSEG_REFERENCE_SOURCE", "/root/reference")


def _sha(path):
    h = hashlib.sha256()
    with open(path, "rb") as f:
        h.update(f.read())
    return h.hexdigest()


def _tree(root):
    out = {}
    for base, dirs, files in os.walk(root):
        dirs[:] = sorted(d for d in dirs if d != "__pycache__" and not d.startswith("."))
        for name in sorted(files):
            if name.endswith((".pyc", ".pyo")) or name == "MANIFEST.json":
                continue
            full = os.path.join(base, name)
            out[os.path.relpath(full, root)] = _sha(full)
    return out


def install(source=SOURCE, dest=DEST):
    """Copy the reference tree; returns the destination or None when the source is absent (GPU box)."""
    if not os.path.isdir(os.path.join(source, "stemseg")):
        return dest if os.path.isdir(os.path.join(dest, "stemseg")) else None
    want = _tree(source)
    manifest_path = os.path.join(dest, "MANIFEST.json")
    if os.path.exists(manifest_path):
        try:
            if json.load(open(manifest_path))["files"] == want and _tree(dest) == want:
                return dest
        except Exception:
            pass
    if os.path.isdir(dest):
        shutil.rmtree(dest)
    shutil.copytree(source, dest, ignore=shutil.ignore_patterns("__pycache__", "*.pyc", ".git"))
    with open(manifest_path, "w") as f:
        json.dump({"source": source, "files": want}, f, indent=0, sort_keys=True)
    return dest


def verify(dest=DEST):
    """True when every file under baseline/_ref still has the sha256 recorded at install time."""
    manifest_path = os.path.join(dest, "MANIFEST.json")
    if not os.path.exists(manifest_path):
        return False
    return json.load(open(manifest_path))["files"] == _tree(dest)


if __name__ == "__main__":
    path = install()
    print(path if path else "reference source not found at %s and no previous install" % SOURCE)
    sys.exit(0 if path else 1)
