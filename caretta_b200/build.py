"""Builds caretta_b200/libcaretta_b200.so (hand-written sm_100a CUDA + C ABI) in-tree with nvcc."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libcaretta_b200.so")
SOURCES = ["crt_api.cu", "crt_kernels.cuh", "crt_fill_f32.cuh", "crt_fill1_v2.cuh", "crt_fill1_v4.cuh", "crt_multi.inl", "crt_fill2_v3.cuh", "crt_dp_batch.cuh", "crt_nj.cuh", "crt_node.cuh", "crt_consumers.cuh", "crt_consumers_api.inl", "crt_level_api.inl", os.path.join("..", "..", "include", "caretta_b200.h"), "Makefile"]


def needs_build() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(os.path.join(CSRC, s)) > t for s in SOURCES)


def build(force: bool = False, verbose: bool = False) -> str:
    if force or needs_build():
        r = subprocess.run(["make", "-C", CSRC] + (["-B"] if force else []), capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("nvcc build of libcaretta_b200.so failed:\n" + r.stdout + r.stderr)
        if verbose:
            print(r.stdout)
    return LIB


if __name__ == "__main__":
    print(build(force=True, verbose=True))
