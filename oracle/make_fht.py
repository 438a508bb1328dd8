"""TEST / BASELINE INFRASTRUCTURE ONLY.  Recipe for oracle/_ref/fast_hadamard_transform_cuda.so: the reference's vendored
third-party FWHT extension (third_party/fast-hadamard-transform, the kernel its spinquant/hadamard_utils.py:3-15 imports),
compiled UNMODIFIED from the sources where they lie under /root/reference, for sm_100a, against this image's torch headers.

It is only the same-box comparison bar for lcb_hadamard_rows (bench.py `hadamard.tri_dao_fwht`, SURVEY 8f-1 / VERDICT
item 9); nothing in llm_compressor_b200/ loads it.  Output goes to the git-ignored oracle/_ref/ only; no source is copied.

    python oracle/make_fht.py          # needs /root/reference (or LC_REFERENCE_ROOT) and nvcc; a few minutes
"""
import os
import subprocess
import sys
import sysconfig

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(os.environ.get("LC_REFERENCE_ROOT", "/root/reference"), "third_party", "fast-hadamard-transform", "csrc")
DST = os.path.join(HERE, "_ref")
NAME = "fast_hadamard_transform_cuda"


def make(verbose=False, force=False):
    out = os.path.join(DST, NAME + ".so")
    if not os.path.isdir(SRC):
        return os.path.exists(out)
    if os.path.exists(out) and not force:
        return True
    from torch.utils import cpp_extension as ce

    os.makedirs(DST, exist_ok=True)
    tmp = os.path.join(DST, "_build_fht")
    os.makedirs(tmp, exist_ok=True)
    inc = ["-I" + p for p in ce.include_paths("cuda")] + ["-I" + sysconfig.get_paths()["include"], "-I" + SRC]
    common = ["-DTORCH_EXTENSION_NAME=" + NAME, "-DTORCH_API_INCLUDE_EXTENSION_H", "-D_GLIBCXX_USE_CXX11_ABI=1", "-std=c++17", "-O3"]
    nvcc = os.path.join(os.environ.get("CUDA_HOME", "/usr/local/cuda"), "bin", "nvcc")
    objs = []
    jobs = [
        ([nvcc, "-c", os.path.join(SRC, "fast_hadamard_transform_cuda.cu"), "-o", os.path.join(tmp, "fht_cuda.o"),
          "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "--use_fast_math", "--expt-relaxed-constexpr",
          "--expt-extended-lambda", "-U__CUDA_NO_HALF_OPERATORS__", "-U__CUDA_NO_HALF_CONVERSIONS__",
          "-U__CUDA_NO_BFLOAT16_OPERATORS__", "-U__CUDA_NO_BFLOAT16_CONVERSIONS__", "-U__CUDA_NO_BFLOAT162_OPERATORS__",
          "-U__CUDA_NO_BFLOAT162_CONVERSIONS__", "-Xcompiler", "-fPIC"] + common + inc),
        (["g++", "-c", os.path.join(SRC, "fast_hadamard_transform.cpp"), "-o", os.path.join(tmp, "fht.o"), "-fPIC"] + common + inc),
    ]
    for cmd in jobs:
        if verbose:
            print(" ".join(cmd[:6]), "...")
        subprocess.run(cmd, check=True)
        objs.append(cmd[cmd.index("-o") + 1])
    libs = ["-L" + p for p in ce.library_paths("cuda")] + ["-lc10", "-ltorch", "-ltorch_cpu", "-ltorch_python", "-lc10_cuda",
                                                            "-ltorch_cuda", "-lcudart"]
    subprocess.run(["g++", "-shared", "-o", out] + objs + libs, check=True)
    for o in objs:
        os.remove(o)
    os.rmdir(tmp)
    if verbose:
        print("built", out)
    return True


if __name__ == "__main__":
    sys.exit(0 if make(verbose=True, force="--force" in sys.argv) else 1)
