"""Stage the reference files the drop-in proof needs under the git-ignored `baseline/_ref/` (it travels to the GPU box
with gpurun; nothing under it is tracked or shipped):  libs/*.py, examples/{encoder_memory_profile,libs_path,__init__}.py,
config.yml.  Run in the build container, where /root/reference is mounted."""
import os
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("GALERKIN_REFERENCE_SRC", "/root/reference")
DST = os.path.join(ROOT, "baseline", "_ref")
FILES = ["config.yml", "examples/encoder_memory_profile.py", "examples/libs_path.py", "examples/__init__.py"]


def main():
    if not os.path.isdir(os.path.join(REF, "libs")):
        print(f"{REF}/libs not found: nothing staged")
        return 0
    files = FILES + ["libs/" + f for f in os.listdir(os.path.join(REF, "libs")) if f.endswith(".py")]
    for rel in files:
        src, dst = os.path.join(REF, rel), os.path.join(DST, rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        shutil.copyfile(src, dst)
    print(f"staged {len(files)} reference files under {DST}")
    return 0


if __name__ == "__main__":
    sys.exit(main())
