"""Build tests/hostsim/_hostsim.so (TEST INFRASTRUCTURE: host replay of the PCR_HD functions)."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
SO = os.path.join(HERE, "_hostsim.so")


def build(force=False):
    src = os.path.join(HERE, "hostsim.cu")
    deps = [src] + [os.path.join(HERE, "..", "..", "point_cloud_registration_b200", "csrc", f)
                    for f in ("pcr_common.cuh", "pcr_grid.cuh", "pcr_linalg.cuh", "pcr_terms.cuh", "pcr_tile.cuh")]
    if not force and os.path.exists(SO) and all(os.path.getmtime(SO) >= os.path.getmtime(d) for d in deps):
        return SO
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    subprocess.check_call([nvcc, "-O2", "-std=c++17", "-shared", "-Xcompiler", "-fPIC", "-o", SO, src])
    return SO


if __name__ == "__main__":
    print(build(force=True))
