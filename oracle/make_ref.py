"""Recipe for oracle/_ref: byte-for-byte copies of the three reference source files on the hot path, so that the
UNMODIFIED reference can be timed (bench.py --impl reference / cpu_baseline, kind "reference") and used as a checker
on the GPU box, where /root/reference does not exist.  TEST / BENCH INFRASTRUCTURE ONLY.

    python oracle/make_ref.py            # needs the reference checkout (MIL_REFERENCE_ROOT or /root/reference)

Outputs go to oracle/_ref/ only.  That directory is git-ignored (reference sources never enter this repository's
history) but not gpurun-ignored, so it travels with the snapshot like the built .so files.  __graft_entry__.build()
runs this whenever the checkout is present.  Nothing is edited: the three-piece shim of oracle/ref_shim.py (stub
PyTorchHelpers, .cuda() -> identity, pass-through DataParallel) is applied at import time, from outside."""
import hashlib
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
DEST = os.path.join(HERE, "_ref")
FILES = ("gbm/model.py", "nnBlocks.py", "alt_resnet.py")   # Attention / ResNet / ContextLayer; BasicResBlock / CrossEntropyWithProbs; the wide extractor


def make(reference_root=None, quiet=False) -> bool:
    root = reference_root or os.environ.get("MIL_REFERENCE_ROOT", "/root/reference")
    if not os.path.isfile(os.path.join(root, FILES[0])):
        if not quiet:
            print(f"make_ref: no reference checkout under {root}; oracle/_ref left as it is")
        return False
    lines = []
    for rel in FILES:
        dst = os.path.join(DEST, rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        shutil.copyfile(os.path.join(root, rel), dst)
        lines.append(f"{hashlib.sha256(open(dst, 'rb').read()).hexdigest()}  {rel}")
    with open(os.path.join(DEST, "SOURCE.txt"), "w") as f:
        f.write("unmodified copies from the reference checkout (oracle/make_ref.py); sha256:\n" + "\n".join(lines) + "\n")
    if not quiet:
        print("make_ref: wrote", DEST)
    return True


if __name__ == "__main__":
    sys.exit(0 if make() else 1)
