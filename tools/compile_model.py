"""Regenerate mycobotgym_b200/assets/mycobot280_{joint,mocap}.json from the reference MJCF trees.

Runs only where /root/reference is mounted (this container).  The GPU box loads the
committed JSON.  Usage: python tools/compile_model.py [/root/reference]
"""
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from mycobotgym_b200 import mjcf  # noqa: E402

ref = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
for xml, out in (("mycobot280.xml", mjcf.COMPILED_JOINT), ("mycobot280_mocap.xml", mjcf.COMPILED_MOCAP)):
    log = []
    m = mjcf.compile_mjcf(os.path.join(ref, "mycobotgym/envs/assets", xml), log)
    m.to_json(out)
    print("\n".join(log))
    print("wrote", out, os.path.getsize(out), "bytes; nbody", m.nbody, "nu", m.nu, "neq", m.neq, "nmocap", m.nmocap)
