"""Pin the CPU oracle against every exact-value known-answer test the reference holds for the
draw path.  Each test names the reference test it restates
(paths relative to /root/reference/)."""
import ctypes as C

import numpy as np
import pytest

from oracle import pyoracle

L = pyoracle.lib()


def _u32(*v):
    return (C.c_uint32 * len(v))(*v)


def _f32(*v):
    return (C.c_float * len(v))(*v)


# ---- wgpu-cpu/src/util/sort.rs:57-135 ----
SORT_TRIPLES = [
    [[847, 895, 730], [730, 847, 895]], [[407, 622, 50], [50, 407, 622]], [[498, 698, 242], [242, 498, 698]],
    [[729, 619, 845], [619, 729, 845]], [[78, 484, 855], [78, 484, 855]], [[856, 128, 44], [44, 128, 856]],
    [[64, 618, 512], [64, 512, 618]], [[132, 858, 663], [132, 663, 858]], [[122, 182, 486], [122, 182, 486]],
    [[521, 930, 464], [464, 521, 930]], [[536, 956, 417], [417, 536, 956]], [[440, 364, 860], [364, 440, 860]],
    [[46, 855, 920], [46, 855, 920]], [[226, 163, 389], [163, 226, 389]], [[369, 992, 285], [285, 369, 992]],
    [[823, 464, 856], [464, 823, 856]], [[847, 1000, 179], [179, 847, 1000]], [[144, 7, 213], [7, 144, 213]],
    [[839, 82, 108], [82, 108, 839]], [[36, 649, 856], [36, 649, 856]], [[508, 363, 603], [363, 508, 603]],
    [[252, 34, 29], [29, 34, 252]], [[274, 629, 663], [274, 629, 663]], [[900, 439, 406], [406, 439, 900]],
    [[231, 941, 884], [231, 884, 941]], [[98, 524, 25], [25, 98, 524]], [[714, 63, 201], [63, 201, 714]],
    [[587, 748, 250], [250, 587, 748]], [[569, 602, 123], [123, 569, 602]], [[896, 104, 31], [31, 104, 896]],
    [[525, 463, 510], [463, 510, 525]], [[290, 324, 18], [18, 290, 324]], [[822, 770, 850], [770, 822, 850]],
    [[330, 592, 965], [330, 592, 965]], [[489, 70, 491], [70, 489, 491]], [[166, 434, 779], [166, 434, 779]],
    [[430, 435, 367], [367, 430, 435]], [[580, 220, 810], [220, 580, 810]], [[807, 391, 254], [254, 391, 807]],
    [[214, 297, 524], [214, 297, 524]], [[494, 777, 964], [494, 777, 964]], [[452, 409, 62], [62, 409, 452]],
    [[676, 631, 494], [494, 631, 676]], [[297, 97, 938], [97, 297, 938]], [[381, 788, 572], [381, 572, 788]],
    [[330, 192, 138], [138, 192, 330]], [[354, 229, 575], [229, 354, 575]], [[290, 110, 807], [110, 290, 807]],
    [[258, 313, 183], [183, 258, 313]], [[253, 103, 815], [103, 253, 815]],
]


def _sort3(v):
    a = (C.c_int64 * 3)(*v)
    L.orc_bubblesort3_i64(a)
    return list(a)


def test_bubblesort3_bug():  # sort.rs:61-71
    assert _sort3([230, 84, 75]) == [75, 84, 230]


@pytest.mark.parametrize("unsorted,expected", SORT_TRIPLES)
def test_random_bubblesorts(unsorted, expected):  # sort.rs:73-134 (the reference sorts the 2nd column; both must hold)
    assert _sort3(unsorted) == expected
    assert _sort3(expected) == expected


# ---- wgpu-cpu/src/util/bresenham.rs:163-219 ----
def _bres(s, e):
    pts = (C.c_uint32 * 128)()
    ts = (C.c_float * 64)()
    n = L.orc_bresenham(_u32(*s), _u32(*e), pts, ts, 64)
    return [(pts[2 * i], pts[2 * i + 1]) for i in range(n)], [ts[i] for i in range(n)]


def test_bresenham_wp_example():
    assert _bres((0, 1), (6, 4))[0] == [(0, 1), (1, 1), (2, 2), (3, 2), (4, 3), (5, 3)]


def test_bresenham_inverse_wp():
    assert _bres((6, 4), (0, 1))[0] == [(6, 4), (5, 4), (4, 3), (3, 3), (2, 2), (1, 2)]


def test_bresenham_straight_hline():
    assert _bres((2, 3), (5, 3))[0] == [(2, 3), (3, 3), (4, 3)]


def test_bresenham_straight_vline():
    assert _bres((2, 3), (2, 6))[0] == [(2, 3), (2, 4), (2, 5)]


def test_bresenham_t_accumulates():  # bresenham.rs:121,144-145: t += 1/dx, summed not multiplied
    _, ts = _bres((0, 0), (7, 3))
    acc, dt = np.float32(0), np.float32(1) / np.float32(7)
    for t in ts:
        assert np.float32(t) == acc
        acc = np.float32(acc + dt)


# ---- wgpu-cpu/src/render_pass/clipper.rs:419-562 (Cohen-Sutherland) ----
def _clip_line(a, b):
    out = (C.c_float * 8)()
    al = (C.c_float * 2)()
    ok = L.orc_clip_line(_f32(*a, *b), out, al)
    return (None if not ok else (list(out[0:4]), list(out[4:8]), list(al)))


def test_line_trivially_inside():
    a, b = [-0.5, -0.5, 0.5, 1.0], [0.5, 0.5, 0.5, 1.0]
    r = _clip_line(a, b)
    assert r[0] == a and r[1] == b and r[2] == [0.0, 1.0]


@pytest.mark.parametrize("a,b", [
    ([-1.5, -1.5, 0.5, 1.0], [-1.5, 1.5, 0.5, 1.0]), ([1.5, -1.5, 0.5, 1.0], [1.5, 1.5, 0.5, 1.0]),
    ([-1.5, -1.5, 0.5, 1.0], [1.5, -1.5, 0.5, 1.0]), ([-1.5, 1.5, 0.5, 1.0], [1.5, 1.5, 0.5, 1.0]),
    # line_non_trivially_outside
    ([-10.0, 0.0, 0.5, 1.0], [0.0, -10.0, 0.5, 1.0]), ([10.0, 0.0, 0.5, 1.0], [0.0, -10.0, 0.5, 1.0]),
    ([-10.0, 0.0, 0.5, 1.0], [0.0, 10.0, 0.5, 1.0]), ([10.0, 0.0, 0.5, 1.0], [0.0, 10.0, 0.5, 1.0]),
])
def test_line_outside(a, b):
    assert _clip_line(a, b) is None


@pytest.mark.parametrize("a,b,c", [
    ([-1.5, -0.5, 0.5, 1.0], [0.5, 0.5, 0.5, 1.0], [-1.0, -0.25, 0.5, 1.0]),   # left
    ([1.5, -0.5, 0.5, 1.0], [0.5, 0.5, 0.5, 1.0], [1.0, 0.0, 0.5, 1.0]),       # right
    ([-0.5, -1.5, 0.5, 1.0], [0.5, 0.5, 0.5, 1.0], [-0.25, -1.0, 0.5, 1.0]),   # bottom
    ([-0.5, 1.5, 0.5, 1.0], [0.5, 0.5, 0.5, 1.0], [0.0, 1.0, 0.5, 1.0]),       # top
    ([0.5, -0.5, 1.5, 1.0], [0.5, 0.5, 0.5, 1.0], [0.5, 0.0, 1.0, 1.0]),       # back
    ([0.5, -0.5, -0.5, 1.0], [0.5, 0.5, 0.5, 1.0], [0.5, 0.0, 0.0, 1.0]),      # front
])
def test_clip_line_against_plane(a, b, c):
    r = _clip_line(a, b)
    np.testing.assert_allclose(r[0], c, atol=1e-6)
    np.testing.assert_allclose(r[1], b, atol=0)
    r = _clip_line(b, a)
    np.testing.assert_allclose(r[0], b, atol=0)
    np.testing.assert_allclose(r[1], c, atol=1e-6)


# ---- wgpu-cpu/src/render_pass/clipper.rs:823-892 (triangle clipper) ----
def _clip_tri(tri):
    pos = (C.c_float * (64 * 12))()
    bary = (C.c_float * (64 * 9))()
    n = L.orc_clip_triangle(_f32(*[x for v in tri for x in v]), pos, bary, 64)
    p = np.array(pos[: n * 12], dtype=np.float32).reshape(n, 3, 4)
    b = np.array(bary[: n * 9], dtype=np.float32).reshape(n, 3, 3)
    return p, b


def test_tri_fully_inside_passes_through():
    tri = [[0.0, 0.5, 0.0, 1.0], [-0.5, -0.5, 0.0, 1.0], [0.5, -0.5, 0.0, 1.0]]
    p, b = _clip_tri(tri)
    assert p.shape[0] == 1
    np.testing.assert_array_equal(p[0], np.array(tri, dtype=np.float32))
    np.testing.assert_array_equal(b[0], np.eye(3, dtype=np.float32))


def test_tri_fully_outside_rejected():
    p, _ = _clip_tri([[-2.0, 0.5, 0.0, 1.0], [-2.5, -0.5, 0.0, 1.0], [-1.5, -0.5, 0.0, 1.0]])
    assert p.shape[0] == 0


def test_tri_one_vertex_outside():
    p, b = _clip_tri([[0.0, 1.5, 0.0, 1.0], [-0.5, 0.5, 0.0, 1.0], [0.5, 0.5, 0.0, 1.0]])
    expected = np.array([
        [[-0.25, 1.0, 0.0, 1.0], [-0.5, 0.5, 0.0, 1.0], [0.25, 1.0, 0.0, 1.0]],
        [[-0.5, 0.5, 0.0, 1.0], [0.5, 0.5, 0.0, 1.0], [0.25, 1.0, 0.0, 1.0]]], dtype=np.float32)
    np.testing.assert_allclose(p, expected, atol=1e-6)
    # the carried barycentrics reproduce the clipped positions from the unclipped ones
    src = np.array([[0.0, 1.5, 0.0, 1.0], [-0.5, 0.5, 0.0, 1.0], [0.5, 0.5, 0.0, 1.0]], dtype=np.float32)
    np.testing.assert_allclose(b @ src, p, atol=1e-6)


def test_tri_one_vertex_inside():
    p, _ = _clip_tri([[0.0, -0.5, 0.0, 1.0], [-0.5, -1.5, 0.0, 1.0], [0.5, -1.5, 0.0, 1.0]])
    expected = np.array([[[0.0, -0.5, 0.0, 1.0], [-0.25, -1.0, 0.0, 1.0], [0.25, -1.0, 0.0, 1.0]]], dtype=np.float32)
    np.testing.assert_allclose(p, expected, atol=1e-6)


# ---- wgpu-cpu/src/render_pass/primitive.rs:501-531 ----
def test_triangle_strip_order():
    out = (C.c_uint32 * 30)()
    n = L.orc_tri_strip(_u32(0, 1, 2, 3, 4, 5), 6, 0, out, 10)
    assert n == 4
    assert [list(out[3 * i: 3 * i + 3]) for i in range(n)] == [[0, 1, 2], [2, 1, 3], [2, 3, 4], [4, 3, 5]]


def test_triangle_strip_restart():  # primitive.rs:407-487: separators reset the generator, lone vertices are dropped
    S = 0xFFFFFFFF
    out = (C.c_uint32 * 30)()
    n = L.orc_tri_strip(_u32(0, 1, 2, 3, S, 4, S, 5, 6, 7, 8), 11, 1, out, 10)
    assert [list(out[3 * i: 3 * i + 3]) for i in range(n)] == [[0, 1, 2], [2, 1, 3], [5, 6, 7], [7, 6, 8]]


def test_triangle_front_face():
    cw = [[-1.0, 0.0, 0.0, 1.0], [0.0, 1.0, 0.0, 1.0], [1.0, 0.0, 0.0, 1.0]]
    ccw = [[1.0, 0.0, 0.0, 1.0], [0.0, 1.0, 0.0, 1.0], [-1.0, 0.0, 0.0, 1.0]]
    assert L.orc_front_face_ccw(_f32(*[x for v in cw for x in v])) == 0
    assert L.orc_front_face_ccw(_f32(*[x for v in ccw for x in v])) == 1


# ---- wgpu-cpu/src/render_pass/index.rs:144-207 ----
@pytest.mark.parametrize("dtype,fmt", [(np.uint16, 1), (np.uint32, 2)])
def test_indirect_indices_without_stops(dtype, fmt):
    indices = np.array([2, 4, 9, 0, 8, 3, 1, 5, 6, 7], dtype=dtype)
    for i in range(10):
        v, sep = C.c_uint32(), C.c_int()
        e = L.orc_resolve_index(indices.ctypes.data_as(C.c_void_p), C.c_uint64(indices.nbytes), fmt, 10, i, 0,
                                C.byref(v), C.byref(sep))
        assert e == 0 and sep.value == 0 and v.value == 10 + int(indices[i])


def test_index_separator_and_bounds():
    indices = np.array([1, 0xFFFF, 3], dtype=np.uint16)
    v, sep = C.c_uint32(), C.c_int()
    args = (indices.ctypes.data_as(C.c_void_p), C.c_uint64(indices.nbytes), 1, 0)
    assert L.orc_resolve_index(*args, 1, 1, C.byref(v), C.byref(sep)) == 0 and sep.value == 1
    assert L.orc_resolve_index(*args, 1, 0, C.byref(v), C.byref(sep)) == 0 and sep.value == 0 and v.value == 0xFFFF
    assert L.orc_resolve_index(*args, 3, 0, C.byref(v), C.byref(sep)) != 0          # slice index out of range panics
    assert L.orc_resolve_index(indices.ctypes.data_as(C.c_void_p), C.c_uint64(indices.nbytes), 1, -2, 0, 0,
                               C.byref(v), C.byref(sep)) != 0                          # strict_add_signed overflow panics


# ---- wgpu-cpu/src/texture.rs:515-549 ----
def test_texture_data_layout():
    assert L.orc_texture_byte_size(pyoracle.FORMAT["rgba8unorm"], 3, 5, 1) == 60
    offs = [L.orc_texel_byte_offset(pyoracle.FORMAT["rgba8unorm"], 3, 5, x, y, 0) for y in range(5) for x in range(3)]
    assert offs == [4 * i for i in range(15)]


# ---- wgpu-cpu/src/util/scanline.rs:206-275: two regression triangles (the reference only checks "no panic") ----
def _rows(tri):
    rows = (C.c_uint32 * (3 * 4096))()
    n = L.orc_scanlines(_u32(*tri), rows, 4096)
    return [(rows[3 * i], rows[3 * i + 1], rows[3 * i + 2]) for i in range(n)]


def test_scanline_bug_negative_coordinates_generated():
    rows = _rows((0, 0, 120, 0, 119, 26))
    assert len(rows) == 27 and rows[0] == (0, 0, 120) and rows[-1] == (26, 119, 119)
    assert all(x1 <= x2 for _, x1, x2 in rows)


def test_scanline_bug_switch_to_second_half():
    rows = _rows((256, 230, 257, 84, 358, 75))
    assert [r[0] for r in rows] == list(range(75, 231))


def test_scanline_rules():
    # flat-bottom triangles lose their last row (scanline.rs:120-122); flat-top start in the second half
    assert [r[0] for r in _rows((10, 10, 5, 20, 15, 20))] == list(range(10, 20))
    assert [r[0] for r in _rows((5, 10, 15, 10, 10, 20))] == list(range(10, 21))
    # degenerate (all one row): min..max x
    assert _rows((7, 3, 2, 3, 9, 3)) == [(3, 2, 9)]
    # truncating division toward zero for negative slopes
    assert _rows((10, 0, 0, 3, 10, 3))[1] == (1, 7, 10)


# ---- naga-cranelift/src/tests.rs:1287-1315, :47-110 ----
def test_matrix_vector_product():
    m = _f32(1, 1, 1, 1, 0, 2, 0, 2, 3, 2, 1, 0, 1, 2, 3, 4)   # columns
    out = (C.c_float * 4)()
    L.orc_mat4_mul_vec4(m, _f32(1, 2, 3, 4), out)
    assert list(out) == [14.0, 19.0, 16.0, 21.0]


def test_vertex_index_triangle_positions():
    exp = [[-1.0, -1.0], [0.0, 1.0], [1.0, -1.0]]
    for vi in range(6):
        pos = (C.c_float * 4)()
        var = (C.c_float * 16)()
        assert L.orc_run_vertex_shader(pyoracle.SHADER["colored_triangle"], vi, 0, None, pos, var) == 0
        assert [pos[0], pos[1]] == exp[vi % 3] and pos[2] == 0.0 and pos[3] == 1.0
        assert list(var[0:4]) == [float(vi % 3 == 0), float(vi % 3 == 1), float(vi % 3 == 2), 1.0]


# ---- source-text restatements without a reference test ("parity unpinned", see oracle.h) ----
def test_f32_to_u8_truncates():  # texture.rs:376-379
    assert L.orc_f32_to_u8(1.0) == 255 and L.orc_f32_to_u8(0.999) == 254 and L.orc_f32_to_u8(0.5) == 127
    assert L.orc_f32_to_u8(-3.0) == 0 and L.orc_f32_to_u8(7.0) == 255 and L.orc_f32_to_u8(float("nan")) == 0


def test_texel_coordinate():  # binding.rs:151-164
    A = pyoracle.ADDRESS
    assert L.orc_texel_coordinate(0.5, A["repeat"], 512) == 256            # round(0.5*511 = 255.5) away from zero
    assert L.orc_texel_coordinate(1.25, A["repeat"], 512) == 128
    assert L.orc_texel_coordinate(-0.25, A["repeat"], 512) == 383
    assert L.orc_texel_coordinate(1.5, A["clamp-to-edge"], 512) == 511
    assert L.orc_texel_coordinate(1.25, A["mirror-repeat"], 512) == 383
    assert L.orc_texel_coordinate(-5.0, A["clamp-to-edge"], 512) == 0


def test_to_raster():  # raster.rs:129-160: no +0.5 pixel centre, y flipped, max(0), truncation
    rs = pyoracle.RasterState()
    L.orc_default_raster_state(512, 512, C.byref(rs))
    fb = (C.c_uint32 * 2)()
    frag = (C.c_float * 4)()
    assert L.orc_to_raster(C.byref(rs), _f32(-1.0, -1.0, 0.0, 1.0), fb, frag) == 0
    assert list(fb) == [0, 512] and list(frag) == [0.0, 512.0, 0.0, 1.0]
    assert L.orc_to_raster(C.byref(rs), _f32(1.0, 2.0, 1.0, 2.0), fb, frag) == 0
    assert list(fb) == [384, 0] and list(frag) == [384.0, 0.0, 0.5, 0.5]
    assert L.orc_to_raster(C.byref(rs), _f32(1.0, 2.0, 1.0, 0.0), fb, frag) != 0     # expect("w=0")
