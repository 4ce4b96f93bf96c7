"""Build oracle/_ref/: the UNMODIFIED reference package, staged so that it travels to the GPU box.

The reference (speechLabBcCuny/onssen @ 179cff9) is pure Python: there is nothing to compile.
"Building" it means copying its own `onssen/` package tree byte for byte from /root/reference into the
git-ignored directory oracle/_ref/ (outputs only; never committed, but shipped by gpurun like the built .so), plus
two EMPTY placeholder packages for the third-party imports that are not installed in this image and that the
reference needs only to get past its import lines:

  * `librosa`   (onssen/data/feature_utils.py:1)  -- no arithmetic; anything that would call into it (get_stft,
                 librosa.core.istft) is monkey-patched by the caller to the oracle's restatement, and says so.
  * `attrdict`  (onssen/utils/train.py:1)         -- `AttrDict = dict`; never used by the tests / bench.

A MANIFEST with the sha256 of every copied file is written next to the copy so that a test can assert the staged
files are unmodified.  Test infrastructure only: the product (onssen_b200/) never imports anything from here.

Usage:  python oracle/make_ref.py            (called by __graft_entry__.build() when /root/reference exists)
"""
import hashlib
import json
import os
import shutil
import sys

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "_ref")

STUBS = {
    "librosa/__init__.py": '"""EMPTY placeholder for the uninstalled librosa (see oracle/make_ref.py)."""\n'
                           "from . import core, feature  # noqa: F401\n",
    "librosa/core.py": '"""EMPTY placeholder (no arithmetic)."""\n',
    "librosa/feature.py": '"""EMPTY placeholder (no arithmetic)."""\n',
    "attrdict.py": '"""EMPTY placeholder for the uninstalled attrdict (see oracle/make_ref.py)."""\nAttrDict = dict\n',
}


def build(ref=REF, out=OUT):
    if not os.path.isdir(os.path.join(ref, "onssen")):
        print(f"make_ref: {ref}/onssen not present; keeping the existing {out} (if any)")
        return False
    if os.path.isdir(out):
        shutil.rmtree(out)
    manifest = {}
    for sub in ("onssen", "egs"):
        for d, _, files in os.walk(os.path.join(ref, sub)):
            for f in files:
                if not (f.endswith(".py") or f.endswith(".json") or f == "RESULT"):
                    continue
                src = os.path.join(d, f)
                rel = os.path.relpath(src, ref)
                dst = os.path.join(out, rel)
                os.makedirs(os.path.dirname(dst), exist_ok=True)
                shutil.copyfile(src, dst)
                manifest[rel] = hashlib.sha256(open(src, "rb").read()).hexdigest()
    for rel, text in STUBS.items():
        dst = os.path.join(out, "_stubs", rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        with open(dst, "w") as fh:
            fh.write(text)
    with open(os.path.join(out, "MANIFEST.json"), "w") as fh:
        json.dump({"source": "speechLabBcCuny/onssen @ 179cff9 (/root/reference)", "sha256": manifest}, fh, indent=1,
                  sort_keys=True)
    print(f"make_ref: staged {len(manifest)} reference files under {out}")
    return True


if __name__ == "__main__":
    sys.exit(0 if build() or os.path.isdir(OUT) else 1)
