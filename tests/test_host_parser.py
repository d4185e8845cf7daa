"""host/scene_parser.cpp (the product's .scn reader) against the reference parser's output and the grammar's quirks
(reference src/scene.cpp:12-227)."""
import os
import textwrap

import numpy as np
import pytest

import skele_raytracer_b200 as S
from conftest import scn_dir


@pytest.mark.parametrize("name", ["spheres1", "spheres2", "bear", "dragon", "test"])
def test_matches_reference_parser_snapshot(scenes, name):
    d = scn_dir()
    if d is None:
        pytest.skip("no .scn files here (they live in /root/reference or oracle/_ref/scenes)")
    s = S.parseScene(os.path.join(d, name + ".scn"))
    g = scenes[name]
    for f in ("spheres", "tris", "plights", "dlights", "camera", "ambient", "background"):
        assert np.array_equal(getattr(s, f), getattr(g, f)), (name, f)
    assert len(s.fogs) == len(g.fogs)


def write(tmp_path, text, newline="\n"):
    p = tmp_path / "t.scn"
    p.write_bytes(textwrap.dedent(text).replace("\n", newline).encode())
    return str(p)


def test_material_state_machine_and_field_order(tmp_path):
    p = write(tmp_path, """\
        #comment in column 0
        sphere 1 2 3 4
        material .1 .2 .3 .4 .5 .6 .7 .8 .9 32 .11 .12 .13 1.5
        sphere 5 6 7 8
        point_light 10 20 30 1 2 3
        ambient_light .25 .25 .25
        ambient_light .1 0 0
        background .05 .06 .07
        directional_light 2 .5 .5 0 -1 0
        camera 0 2 -10 0 -.1 .9 0 1 0 26
        film_resolution 800 600
        max_depth 5
        output_image foo.bmp
        bogus_command 1 2 3
        """)
    s = S.parseScene(p)
    # first sphere: the default Material (src/material.h:9-26)
    assert s.spheres[0].tolist() == [1, 2, 3, 4] + [0] * 12 + [1, 1]
    np.testing.assert_allclose(s.spheres[1], [5, 6, 7, 8, .1, .2, .3, .4, .5, .6, .7, .8, .9, .11, .12, .13, 32, 1.5], rtol=1e-6)
    assert s.plights.tolist() == [[1, 2, 3, 10, 20, 30]]          # colour first on the line, position stored first
    np.testing.assert_allclose(s.ambient, [.35, .25, .25], rtol=1e-6)  # ambient accumulates
    np.testing.assert_allclose(s.background, [.05, .06, .07], rtol=1e-6)
    assert len(s.dlights) == 0                                     # parsed, never stored
    # right = cross(-direction, up), nothing normalised
    np.testing.assert_allclose(s.camera, [0, 2, -10, 0, -.1, .9, 0, 1, 0, .9, 0, 0], atol=1e-7)
    assert s.film_resolution == (800, 600) and s.max_depth == 5 and s.unknown_commands == 1
    kd = S.parseScene(p, keep_directional=True)
    assert kd.dlights.tolist() == [[0, -1, 0, 1, .5, .5]]          # colour clamped to <= 1


def test_crlf_vertices_triangles(tmp_path):
    p = write(tmp_path, """\
        vertex 0 0 0
        vertex 1 0 0
        vertex 0 1 0
        vertex 0 0 1
        triangle 0 1 2
        triangle 3 2 1
        """, newline="\r\n")
    s = S.parseScene(p)
    assert s.tris.tolist() == [[0, 0, 0, 1, 0, 0, 0, 1, 0], [0, 0, 1, 0, 1, 0, 1, 0, 0]]


def test_triangle_index_out_of_range_is_an_error(tmp_path):
    p = write(tmp_path, "vertex 0 0 0\ntriangle 0 1 2\n")
    with pytest.raises(S.SkrError, match="vertex"):
        S.parseScene(p)


def test_missing_file(tmp_path):
    with pytest.raises(S.SkrError, match="Can't open file"):
        S.parseScene(str(tmp_path / "nope.scn"))


def test_spherical_fog_parses_the_intended_fields(tmp_path):
    p = write(tmp_path, "spherical_fog 0 -50 0 100 1 .5 .25 .5\n")
    s = S.parseScene(p)
    assert s.fogs.tolist() == [[.5, 0, 1, .5, .25, 100, 0, -50, 0]]
    assert len(S.parseScene(p, fog=False).fogs) == 0


def test_write_ppm_bytes(tmp_path):
    img = np.arange(2 * 3 * 3, dtype=np.uint8).reshape(2, 3, 3)
    S.write_ppm(str(tmp_path / "o.ppm"), img)
    assert (tmp_path / "o.ppm").read_bytes() == b"P6\n3 2\n255\n" + img.tobytes()
