"""TEST / BASELINE INFRASTRUCTURE ONLY.  Recipe for oracle/_ref/: an UNMODIFIED copy of the reference's pure-Python hot
path (pjh5672/llm-compressor, /root/reference/llm_compressor), so that the GPU box -- where /root/reference does not
exist -- can time the reference's own code (bench.py --impl reference, bench.py's reference_eager_b200 block) and run
the drop-in boundary test (tests/test_reference_dropin_gpu.py).

oracle/_ref/ is git-ignored (reference sources never enter this repository's history) but not gpurun-ignored, so it
travels with the snapshot like the built .so files.  `__graft_entry__.build()` runs this whenever /root/reference is
present.  Copied: utils/, modules/, pruning/, quantization/ minus the SpinQuant tree (2 MB of literal Hadamard tables,
off the timed path).  Nothing is edited; the import shim (oracle/ref_shim.py) supplies the three missing import-time
dependencies exactly as it does for /root/reference."""
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.environ.get("LC_REFERENCE_SRC", "/root/reference")
DST = os.path.join(HERE, "_ref")
KEEP = ("utils", "modules", "pruning", "quantization")
SKIP_DIRS = ("spinquant", "__pycache__")


def make(verbose=False):
    src_pkg = os.path.join(SRC, "llm_compressor")
    if not os.path.isdir(src_pkg):
        return False
    dst_pkg = os.path.join(DST, "llm_compressor")
    if os.path.isdir(dst_pkg):
        shutil.rmtree(dst_pkg)
    os.makedirs(dst_pkg)
    n = 0
    for f in os.listdir(src_pkg):
        if f.endswith(".py"):
            shutil.copy2(os.path.join(src_pkg, f), os.path.join(dst_pkg, f))
    for sub in KEEP:
        for root, dirs, files in os.walk(os.path.join(src_pkg, sub)):
            dirs[:] = [d for d in dirs if d not in SKIP_DIRS]
            rel = os.path.relpath(root, src_pkg)
            os.makedirs(os.path.join(dst_pkg, rel), exist_ok=True)
            for f in files:
                if f.endswith(".py"):
                    shutil.copy2(os.path.join(root, f), os.path.join(dst_pkg, rel, f))
                    n += 1
    with open(os.path.join(DST, "PROVENANCE.txt"), "w") as fh:
        fh.write("unmodified copy of %s (%d files) made by oracle/make_ref.py; not part of the repository\n" % (src_pkg, n))
    if verbose:
        print("oracle/_ref: %d files copied from %s" % (n, src_pkg))
    return True


if __name__ == "__main__":
    ok = make(verbose=True)
    sys.exit(0 if ok else 1)
