#!/usr/bin/env python
"""Pack reference scenes (JSON + PLY + PNG/HDR) into single `.jtscene` files so benchmarks and GPU
tests can run where /root/reference does not exist. Usage:
    python tools/pack_scene.py [--ref /root/reference] cornellbox classroom ...
Only input DATA is packed (geometry, textures, materials); no reference source code."""
import argparse
import importlib
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
jt = importlib.import_module("julia-raytracer_b200")

ap = argparse.ArgumentParser()
ap.add_argument("--ref", default="/root/reference")
ap.add_argument("--out", default=os.path.join(ROOT, "assets", "scenes"))
ap.add_argument("scenes", nargs="+")
args = ap.parse_args()
os.makedirs(args.out, exist_ok=True)
for name in args.scenes:
    src = os.path.join(args.ref, "scenes", name, f"{name}.json")
    scene = jt.load_scene(src)
    dst = os.path.join(args.out, f"{name}.jtscene")
    jt.save_packed(scene, dst)
    back = jt.load_packed(dst)
    assert len(back.instances) == len(scene.instances) and len(back.shapes) == len(scene.shapes)
    print(f"{name}: {os.path.getsize(dst) / 1e6:.2f} MB  notes={scene.notes}")
