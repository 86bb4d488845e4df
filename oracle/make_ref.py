"""Recipe for `oracle/_ref/`: packs the UNMODIFIED reference package (`/root/reference/lhotse`, pure Python) into
`oracle/_ref/lhotse_ref.zip` so that the GPU box — where `/root/reference` does not exist — can import and run the
real reference (zipimport) as the checker in `-m gpu` tests and as the timed CPU arm of `bench.py --impl reference`.

TEST INFRASTRUCTURE ONLY: `oracle/_ref/` is git-ignored (no reference source enters the history), it is not listed in
`.gpurunignore` (so it travels with the snapshot like the built `.so`), and nothing under `lhotse_b200/` imports it.
Run by `__graft_entry__.build()` whenever the reference tree is present; a no-op otherwise.

    python oracle/make_ref.py            # (re)builds the archive, prints its path
"""
from __future__ import annotations

import hashlib
import json
import os
import sys
import zipfile

HERE = os.path.dirname(os.path.abspath(__file__))
REF_ROOT = os.environ.get("LHOTSE_REFERENCE_ROOT", "/root/reference")
OUT_DIR = os.path.join(HERE, "_ref")
ZIP_PATH = os.path.join(OUT_DIR, "lhotse_ref.zip")
INFO_PATH = os.path.join(OUT_DIR, "lhotse_ref.json")


def _files():
    pkg = os.path.join(REF_ROOT, "lhotse")
    for d, dirs, fs in os.walk(pkg):
        dirs[:] = sorted(x for x in dirs if x != "__pycache__")
        for f in sorted(fs):
            if not f.endswith(".pyc"):
                yield os.path.join(d, f)


def build(force: bool = False) -> str | None:
    """Returns the archive path, or None when there is no reference tree to pack (and no archive from before)."""
    if not os.path.isdir(os.path.join(REF_ROOT, "lhotse")):
        return ZIP_PATH if os.path.exists(ZIP_PATH) else None
    files = list(_files())
    h = hashlib.sha256()
    for p in files:
        h.update(os.path.relpath(p, REF_ROOT).encode())
        with open(p, "rb") as f:
            h.update(f.read())
    digest = h.hexdigest()
    if not force and os.path.exists(ZIP_PATH) and os.path.exists(INFO_PATH):
        try:
            if json.load(open(INFO_PATH)).get("sha256") == digest:
                return ZIP_PATH
        except Exception:
            pass
    os.makedirs(OUT_DIR, exist_ok=True)
    tmp = ZIP_PATH + ".tmp"
    with zipfile.ZipFile(tmp, "w", zipfile.ZIP_DEFLATED) as z:
        for p in files:
            z.write(p, os.path.relpath(p, REF_ROOT))
    os.replace(tmp, ZIP_PATH)
    with open(INFO_PATH, "w") as f:
        json.dump({"source": REF_ROOT, "files": len(files), "sha256": digest}, f)
    return ZIP_PATH


if __name__ == "__main__":
    print(build(force="--force" in sys.argv))
