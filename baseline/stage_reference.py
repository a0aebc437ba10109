"""Stages a verbatim copy of the reference's Python sources for the hot path under baseline/_ref/ (git-ignored, NOT
gpurun-ignored: the GPU box has no /root/reference, the copy is what `bench.py --impl reference` runs there).

    python baseline/stage_reference.py            # copies from /root/reference (or $MILLIEYE_REFERENCE)

`pip install --target baseline/_ref /root/reference` is not possible (the reference has neither setup.py nor
pyproject.toml), so the files are copied as they are: module3_our_dataset/{my_models.py, yolov3/, utils/, config/*.cfg}
and module2_mixed/my_models.py.  Nothing under baseline/_ref is imported by the product or by the tests; only
baseline/ref_runner.py (the reference arm of bench.py) loads it."""
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
DST = os.path.join(HERE, "_ref")
WANTED = [
    ("module3_our_dataset/my_models.py", "module3_our_dataset/my_models.py"),
    ("module3_our_dataset/yolov3", "module3_our_dataset/yolov3"),
    ("module3_our_dataset/utils", "module3_our_dataset/utils"),
    ("module3_our_dataset/config", "module3_our_dataset/config"),
    ("module2_mixed/my_models.py", "module2_mixed/my_models.py"),
    ("LICENSE", "LICENSE"),
]


def stage(src_root=None, quiet=False):
    src_root = src_root or os.environ.get("MILLIEYE_REFERENCE", "/root/reference")
    if not os.path.isdir(src_root):
        return False
    for rel_src, rel_dst in WANTED:
        src, dst = os.path.join(src_root, rel_src), os.path.join(DST, rel_dst)
        if not os.path.exists(src):
            continue
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        if os.path.isdir(src):
            shutil.copytree(src, dst, dirs_exist_ok=True,
                            ignore=shutil.ignore_patterns("__pycache__", "*.pyc", "*.weights", "*.pth", "*.pt"))
        else:
            shutil.copy2(src, dst)
    with open(os.path.join(DST, "PROVENANCE.txt"), "w") as fh:
        fh.write(f"verbatim copy of {src_root} (sxontheway/milliEye) made by baseline/stage_reference.py; unmodified\n")
    if not quiet:
        print(f"staged reference under {DST}")
    return True


if __name__ == "__main__":
    sys.exit(0 if stage() else 1)
