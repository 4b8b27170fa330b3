"""Regenerate mycobotgym_b200/assets/mycobot280_joint.json from the reference MJCF tree.

Runs only where /root/reference is mounted (this container).  The GPU box loads the
committed JSON.  Usage: python tools/compile_model.py [/root/reference]
"""
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from mycobotgym_b200 import mjcf  # noqa: E402

ref = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
log = []
m = mjcf.compile_mjcf(os.path.join(ref, "mycobotgym/envs/assets/mycobot280.xml"), log)
m.to_json(mjcf.COMPILED_JOINT)
print("\n".join(log))
print("wrote", mjcf.COMPILED_JOINT, os.path.getsize(mjcf.COMPILED_JOINT), "bytes")
