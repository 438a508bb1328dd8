"""TEST / BASELINE INFRASTRUCTURE ONLY.  Recipe for oracle/_ref/: an UNMODIFIED copy of the reference's pure-Python hot
path (pjh5672/llm-compressor, /root/reference/llm_compressor), so that the GPU box -- where /root/reference does not
exist -- can time the reference's own code (bench.py --impl reference, bench.py's reference_eager_b200 block) and run
the drop-in boundary test (tests/test_reference_dropin_gpu.py).

oracle/_ref/ holds ONE archive, llm_compressor_ref.zip (git-ignored: reference sources never enter this repository's
history or tree; not gpurun-ignored, so it travels with the snapshot like the built .so files); Python imports the
package straight out of the archive (zipimport).  `__graft_entry__.build()` runs this whenever /root/reference is
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
    """Pack the reference's hot-path modules, unmodified, into oracle/_ref/llm_compressor_ref.zip (one binary artefact,
    imported in place through zipimport by oracle/ref_shim.py -- nothing is unpacked into the tree)."""
    import zipfile

    src_pkg = os.path.join(SRC, "llm_compressor")
    if not os.path.isdir(src_pkg):
        return False
    os.makedirs(DST, exist_ok=True)
    old = os.path.join(DST, "llm_compressor")
    if os.path.isdir(old):
        shutil.rmtree(old)
    zpath = os.path.join(DST, "llm_compressor_ref.zip")
    n = 0
    dirs_done = set()

    def put(z, src, arc):
        # the reference's packages are namespace packages (no __init__.py): zipimport finds those only through explicit
        # directory entries
        parts = arc.split("/")[:-1]
        for i in range(1, len(parts) + 1):
            d = "/".join(parts[:i]) + "/"
            if d not in dirs_done:
                dirs_done.add(d)
                z.writestr(zipfile.ZipInfo(d), "")
        z.write(src, arc)

    with zipfile.ZipFile(zpath, "w", zipfile.ZIP_DEFLATED) as z:
        for f in sorted(os.listdir(src_pkg)):
            if f.endswith(".py"):
                put(z, os.path.join(src_pkg, f), "llm_compressor/" + f)
        for sub in KEEP:
            for root, dirs, files in os.walk(os.path.join(src_pkg, sub)):
                dirs[:] = sorted(d for d in dirs if d not in SKIP_DIRS)
                rel = os.path.relpath(root, src_pkg)
                for f in sorted(files):
                    if f.endswith(".py"):
                        put(z, os.path.join(root, f), "llm_compressor/" + rel + "/" + f)
                        n += 1
    with open(os.path.join(DST, "PROVENANCE.txt"), "w") as fh:
        fh.write("llm_compressor_ref.zip: unmodified copy of %s (%d files) packed by oracle/make_ref.py; not part of the repository\n"
                 % (src_pkg, n))
    if verbose:
        print("oracle/_ref/llm_compressor_ref.zip: %d files from %s" % (n, src_pkg))
    return True


if __name__ == "__main__":
    ok = make(verbose=True)
    sys.exit(0 if ok else 1)
