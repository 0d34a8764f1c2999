"""Build a tuning variant of libblipgpu.so:  python tools/build_variant.py NAME [-DFLAG ...]
The library lands in pyblip_b200/_C/variants/NAME.so; the in-tree library is left alone."""
import os
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pyblip_b200 import build  # noqa: E402

name, flags = sys.argv[1], tuple(sys.argv[2:])
keep = build.LIB + ".keep"
have = os.path.exists(build.LIB)
if have:
    shutil.copy(build.LIB, keep)
stamp = os.path.join(build.OUT_DIR, "build.stamp")
old_stamp = open(stamp).read() if os.path.exists(stamp) else None
build.build(force=True, extra_flags=flags)
os.makedirs(os.path.join(build.OUT_DIR, "variants"), exist_ok=True)
dst = os.path.join(build.OUT_DIR, "variants", name + ".so")
shutil.move(build.LIB, dst)
if have:
    shutil.move(keep, build.LIB)
if old_stamp is not None:
    open(stamp, "w").write(old_stamp)
print(dst)
