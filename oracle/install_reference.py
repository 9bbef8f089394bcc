"""Install the UNMODIFIED reference python modules of the hot path under baseline/_ref/ (git-ignored; travels to the GPU box
with the repository snapshot like the built .so files) so that `bench.py --impl reference` can time the reference's own
modules under oracle/shim.py on the box's host cores.  TEST / BASELINE INFRASTRUCTURE: nothing under lang2seg_b200/ reads
it, and nothing is committed.

    python -m oracle.install_reference [--src /root/reference]

Copies only `lib/**.py` and `pyutils/mask-faster-rcnn/lib/**.py` (1.4 MB of sources; no data, no prebuilt binaries).
The reference is not pip-installable (no setup.py, Python 2.7 / PyTorch 0.3 API), so this copy + the compatibility shim
 is the install.
"""
import os
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DST = os.path.join(ROOT, "baseline", "_ref")
TREES = ("lib", os.path.join("pyutils", "mask-faster-rcnn", "lib"))


def installed():
    return all(os.path.isdir(os.path.join(DST, t)) for t in TREES)


def install(src="/root/reference", quiet=False):
    if not all(os.path.isdir(os.path.join(src, t)) for t in TREES):
        return False
    n = 0
    for t in TREES:
        for dirpath, _, files in os.walk(os.path.join(src, t)):
            rel = os.path.relpath(dirpath, src)
            for f in files:
                if f.endswith(".py"):
                    os.makedirs(os.path.join(DST, rel), exist_ok=True)
                    shutil.copyfile(os.path.join(dirpath, f), os.path.join(DST, rel, f))
                    n += 1
    if not quiet:
        print("installed %d reference modules under %s" % (n, DST))
    return True


if __name__ == "__main__":
    src = sys.argv[sys.argv.index("--src") + 1] if "--src" in sys.argv else "/root/reference"
    sys.exit(0 if install(src) else 1)
