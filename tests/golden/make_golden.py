#!/usr/bin/env python
"""tests/golden/make_golden.py -- regenerates the committed fixtures from the REFERENCE ITSELF.

Run in the build container (needs /root/reference and `make -C oracle ref`):
    python tests/golden/make_golden.py

Writes
  scenes/<name>.npz            flat snapshot of each reference scene, exactly as the reference's own parser
                               (src/scene.cpp, compiled in place into oracle/_ref/libskr_ref.so) produced it --
                               including spheres2's fog record, whose fields are whatever the reference's broken
                               sscanf left on the stack in this build: both sides of every parity test
                               consume the same numbers.
  scenes/spheres2_nofog.npz    spheres2 with the fog record removed (deterministic variant).
  testcpu_dragon_640x480.npz   the reference's only golden render, renders/testcpu.ppm, stored losslessly as
                               (sha256 of the file, its two colours, packed bit mask) -- the file is a 2-colour image.
  ref_images.npz               float32 images rendered by the reference's own shade() (ref_driver.cpp) for a grid of
                               scenes x flag combinations at small resolutions, deterministic AND stochastic (serial,
                               srand(seed)), so that the C port stays pinned where oracle/_ref cannot be rebuilt.
"""
import hashlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import oracle_lib as O  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
REF_ROOT = "/root/reference"

# (key, scene, Options kwargs, seed)
CASES = []
for scene in ["spheres1", "spheres2", "spheres2_nofog", "bear", "test"]:
    CASES.append((f"{scene}/det", scene, dict(width=160, height=90, max_depth=1), 0))
    CASES.append((f"{scene}/det_shadow", scene, dict(width=160, height=90, max_depth=3, use_shadows=True), 0))
    CASES.append((f"{scene}/jsample3_shadow", scene, dict(width=96, height=54, grid_size=3, use_shadows=True), 11))
    if scene != "test":
        CASES.append((f"{scene}/gillum4_d3_shadow", scene,
                      dict(width=64, height=36, max_depth=3, monte_carlo=True, num_path_traces=4, use_shadows=True), 12))
    CASES.append((f"{scene}/gillum3_js2_d2", scene,
                  dict(width=48, height=27, max_depth=2, monte_carlo=True, num_path_traces=3, grid_size=2), 13))
CASES.append(("dragon/det", "dragon", dict(width=96, height=72, max_depth=1), 0))
CASES.append(("spheres1/fov90_square", "spheres1", dict(width=80, height=80, fov=90.0, max_depth=2, use_shadows=True), 0))
CASES.append(("bear/depth0", "bear", dict(width=32, height=18, max_depth=0), 0))


def main():
    ref = O.Ref()
    scenes = {}
    for name in ["spheres1", "spheres2", "bear", "dragon", "test"]:
        s = ref.parse(os.path.join(REF_ROOT, "scenes", name + ".scn"))
        scenes[name] = s
        s.save(os.path.join(HERE, "scenes", name + ".npz"))
        print(name, len(s.spheres), "spheres", len(s.tris), "tris", len(s.plights), "plights", len(s.fogs), "fogs")
    nf = O.Scene.load(os.path.join(HERE, "scenes", "spheres2.npz"))
    nf.fogs = np.zeros((0, 9), np.float32)
    nf.save(os.path.join(HERE, "scenes", "spheres2_nofog.npz"))
    scenes["spheres2_nofog"] = nf

    ppm = open(os.path.join(REF_ROOT, "renders", "testcpu.ppm"), "rb").read()
    header = b"P6\n640 480\n255\n"
    assert ppm.startswith(header)
    px = np.frombuffer(ppm[len(header):], np.uint8).reshape(480, 640, 3)
    colours = np.unique(px.reshape(-1, 3), axis=0)
    assert len(colours) == 2, colours
    mask = (px == colours[1]).all(axis=2)
    np.savez_compressed(os.path.join(HERE, "testcpu_dragon_640x480.npz"), sha256=hashlib.sha256(ppm).hexdigest(),
                        colours=colours, mask=np.packbits(mask))
    print("testcpu.ppm", hashlib.sha256(ppm).hexdigest(), colours.tolist(), int(mask.sum()))

    out = {}
    for key, scene, kw, seed in CASES:
        opt = O.Options(**kw)
        rgb32, rgb8, secs = ref.render(scenes[scene], opt, seed=seed, threads=1)
        out[key] = rgb32
        print(f"{key:34s} {secs:6.2f}s mean={rgb32.mean():.5f}")
    np.savez_compressed(os.path.join(HERE, "ref_images.npz"), **out)


if __name__ == "__main__":
    main()
