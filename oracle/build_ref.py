"""TEST / BENCH INFRASTRUCTURE ONLY.  Recipe for oracle/_ref/: the UNMODIFIED reference package (TurtleTools/caretta, pure Python
+ numba) copied file by file from where it lies under /root/reference, so that the GPU box -- where /root/reference does not
exist -- can time the reference's own numba path beside the CUDA engine (bench.py: cpu_baseline.kind = "reference", --impl
reference) and the parity gate of the bench can ask the reference itself.

oracle/_ref/ is a build output: it is listed in .gitignore (no reference source ever enters the history) and NOT in
.gpurunignore (it travels to the GPU box like the built .so files).  Nothing under caretta_b200/ imports it.

    python -m oracle.build_ref        # run by __graft_entry__.build() when /root/reference is present
"""
import filecmp
import os
import shutil

HERE = os.path.dirname(os.path.abspath(__file__))
REF_SRC = os.environ.get("CARETTA_REFERENCE", "/root/reference")
REF_OUT = os.path.join(HERE, "_ref")
# the modules the pair path imports (multiple_alignment.py:18-25); app/ (dash GUI) is not on any path we time
FILES = ["__init__.py", "dynamic_time_warping.py", "score_functions.py", "superposition_functions.py", "helper.py",
         "neighbor_joining.py", "multiple_alignment.py", "feature_extraction.py"]


def build() -> str:
    """Copies the reference's modules into oracle/_ref/caretta/ (idempotent).  Returns the directory to put on sys.path, or ""
    when neither the reference nor an earlier copy is present."""
    src = os.path.join(REF_SRC, "caretta")
    dst = os.path.join(REF_OUT, "caretta")
    if not os.path.isdir(src):
        return REF_OUT if os.path.isdir(dst) else ""
    os.makedirs(dst, exist_ok=True)
    for f in FILES:
        a, b = os.path.join(src, f), os.path.join(dst, f)
        if not os.path.exists(b) or not filecmp.cmp(a, b, shallow=False):
            shutil.copyfile(a, b)
    lic = os.path.join(REF_SRC, "LICENSE")
    if os.path.exists(lic):
        shutil.copyfile(lic, os.path.join(REF_OUT, "LICENSE"))
    with open(os.path.join(REF_OUT, "README"), "w") as fh:
        fh.write("Unmodified copy of TurtleTools/caretta's hot-path modules, made by oracle/build_ref.py from /root/reference.\n"
                 "Build output (git-ignored); used only by bench.py's CPU legs and the tests as the reference itself.\n")
    return REF_OUT


def root() -> str:
    """Where the reference can be imported from: /root/reference in the build container, oracle/_ref on the GPU box."""
    if os.path.isdir(os.path.join(REF_SRC, "caretta")):
        return REF_SRC
    if os.path.isdir(os.path.join(REF_OUT, "caretta")):
        return REF_OUT
    return ""


if __name__ == "__main__":
    print(build() or "reference not found")
