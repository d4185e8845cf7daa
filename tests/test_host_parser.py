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


def test_number_conversion_is_strtof_bit_for_bit():
    """The reader converts plain decimals on a fast exact path and hands everything else to strtof; either way the float
    must be the one the reference's sscanf("%f") produces."""
    import ctypes as C
    import random
    import struct

    from skele_raytracer_b200 import api
    L = api._host_lib()
    L.skr_host_read_floats.argtypes = [C.c_char_p, C.POINTER(C.c_float), C.c_int]
    libc = C.CDLL("libc.so.6")
    libc.strtof.restype = C.c_float
    libc.strtof.argtypes = [C.c_char_p, C.c_void_p]
    rnd = random.Random(7)
    buf = (C.c_float * 4)()
    cases = ["0", "-0", "0.0", ".5", "5.", "+3", "1e10", "1E-10", "1e38", "3.5e38", "1e-45", "16777217", "0.1", "1.0000000596046448",
             "1.00000005960464478", "8.5e-46", "1e23", "123456789012345678", "0.000000000000000000001", "1e", "1e+", "0x10", "12abc"]
    for _ in range(60000):
        k = rnd.random()
        if k < 0.4:
            cases.append("%.*f" % (rnd.randint(0, 9), rnd.uniform(-100, 100)))
        elif k < 0.6:
            cases.append("%.*e" % (rnd.randint(0, 14), rnd.uniform(-1, 1) * 10 ** rnd.randint(-30, 30)))
        elif k < 0.8:
            cases.append(str(rnd.randint(-10 ** rnd.randint(1, 15), 10 ** rnd.randint(1, 15))))
        else:
            cases.append("%s%s.%s" % (rnd.choice(["", "-", "+"]), rnd.choice(["", "0", "12", "000"]),
                                      "".join(rnd.choice("0123456789") for _ in range(rnd.randint(1, 12)))))
    for sv in cases:
        n = L.skr_host_read_floats(sv.encode(), buf, 1)
        ref = libc.strtof(sv.encode(), None)
        assert n == 1 and struct.pack("f", buf[0]) == struct.pack("f", ref), sv
    # several numbers on a line, CRLF, trailing junk stops the scan like sscanf
    assert L.skr_host_read_floats(b" 1 -2.5\t3e1 .25 x 7\r\n", buf, 4) == 4 and list(buf) == [1.0, -2.5, 30.0, 0.25]
    assert L.skr_host_read_floats(b" 1 2 abc 3", buf, 4) == 2
