#!/usr/bin/env python
"""Generate the golden fixtures of tests/golden/ from the reference checkout (run in the build
container, where /root/reference exists; the GPU box only ever reads the committed outputs).

1. ref_<scene>_<sampler>.png : the reference's own shipped renders (images/*.png, 8-bit sRGB, spp
   unknown, full asset set, unseeded RNG -> LOOSE goldens), box-downsampled to 160 px wide.
2. scene_facts.json : structural facts read from the reference's scene files (counts of instances,
   shapes, elements per kind, textures) that the loader + BVH restatements must reproduce.
"""
import json
import os
import sys

import numpy as np
from PIL import Image

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
REF = "/root/reference"
OUT = os.path.join(ROOT, "tests", "golden")
os.makedirs(OUT, exist_ok=True)

for scene, sampler in [("cornellbox", "path"), ("materials1", "path"), ("materials1", "naive"),
                       ("features1", "path"), ("features1", "naive"), ("classroom", "path"), ("ecosys", "path")]:
    src = os.path.join(REF, "images", f"{scene}_{sampler}.png")
    im = Image.open(src).convert("RGB")
    w = 160
    h = round(im.height * w / im.width)
    im.resize((w, h), Image.BOX).save(os.path.join(OUT, f"ref_{scene}_{sampler}.png"))
    print("golden", scene, sampler, im.size, "->", (w, h))

import importlib
jt = importlib.import_module("julia-raytracer_b200")
facts = {}
for scene in ["cornellbox", "materials1", "features1", "classroom", "ecosys"]:
    js = json.load(open(os.path.join(REF, "scenes", scene, f"{scene}.json")))
    sc = jt.load_scene(os.path.join(REF, "scenes", scene, f"{scene}.json"))
    facts[scene] = dict(
        json_instances=len(js.get("instances", [])), json_shapes=len(js.get("shapes", [])),
        json_materials=len(js.get("materials", [])), json_textures=len(js.get("textures", [])),
        instances=len(sc.instances), triangles=int(sum(len(s.triangles) for s in sc.shapes)),
        quads=int(sum(len(s.quads) for s in sc.shapes)),
        texture_sizes=[[t.width, t.height] for t in sc.textures], notes=sc.notes)
json.dump(facts, open(os.path.join(OUT, "scene_facts.json"), "w"), indent=1)
print("wrote scene_facts.json")
