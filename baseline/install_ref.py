"""Places the UNMODIFIED reference files the hot path's end-to-end harness needs under baseline/_ref/.

    python baseline/install_ref.py [/root/reference]

baseline/_ref/ is git-ignored (nothing of the reference enters this repository's history) but NOT
gpurun-ignored, so it travels to the GPU box, where /root/reference does not exist.  The reference has no
setup.py / pyproject, so `pip install --target baseline/_ref /root/reference` has nothing
to build; this script is the install: a byte-for-byte copy of the eleven Python files that TCStereo.forward
imports (core/tc_stereo.py:1-8 and their own imports) plus train_stereo.py for its init_loss().  The drivers (evaluate_stereo.py,
train_stereo.py) hard-require wandb / skimage / pykitti and are never imported.

Used by: oracle/ref_model.py (tests, bench.py --impl reference-gpu, the cpu arm).  Never by the product path.
"""
import filecmp
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
DEST = os.path.join(HERE, "_ref")
FILES = [
    "core/__init__.py",
    "core/corr.py",
    "core/tc_stereo.py",
    "core/update.py",
    "core/extractor.py",
    "core/utils/__init__.py",
    "core/utils/utils.py",
    "core/utils/geo_utils.py",
    "core/utils/basic_layers.py",
    "core/utils/splatting/__init__.py",
    "core/utils/splatting/softsplat.py",
    "train_stereo.py",          # never imported (it needs wandb): oracle/ref_model.load_init_loss() extracts init_loss() from its source
]


def install(src="/root/reference", dest=DEST, quiet=False):
    """Returns dest, or None when the reference checkout is absent (the GPU box: it uses what travelled)."""
    if not os.path.isdir(os.path.join(src, "core")):
        return dest if os.path.exists(os.path.join(dest, "core", "tc_stereo.py")) else None
    for rel in FILES:
        s, d = os.path.join(src, rel), os.path.join(dest, rel)
        os.makedirs(os.path.dirname(d), exist_ok=True)
        if not (os.path.exists(d) and filecmp.cmp(s, d, shallow=False)):
            shutil.copyfile(s, d)
    if not quiet:
        print("reference installed under", dest, "(%d files, unmodified)" % len(FILES))
    return dest


if __name__ == "__main__":
    out = install(sys.argv[1] if len(sys.argv) > 1 else "/root/reference")
    if out is None:
        raise SystemExit("no reference checkout and no previous install")
