"""LAS / PLY adapters (more3d_b200/io_formats.py; reference: levelsets_func.py:43-84 via laspy / plyfile).
laspy and plyfile are not available offline, so the readers are checked against files assembled byte by byte
from the format specifications and against round trips through the writers.  CPU only."""
import struct

import numpy as np
import pytest

from more3d_b200 import io_formats as io


def _cloud(n=500, seed=0):
    rng = np.random.default_rng(seed)
    xyz = np.column_stack([rng.uniform(5000, 5100, n), rng.uniform(-200, -100, n), rng.uniform(300, 320, n)])
    nrm = rng.normal(size=(n, 3))
    nrm /= np.linalg.norm(nrm, axis=1)[:, None]
    return xyz, nrm, rng.random(n), rng.random(n).astype(np.float32)


@pytest.mark.parametrize("version,fmt", [((1, 4), 6), ((1, 2), 1), ((1, 2), 0)])
def test_las_round_trip(tmp_path, version, fmt):
    xyz, nrm, curv, sal = _cloud()
    path = str(tmp_path / "a.las")
    extra = {"Curvature": curv, "NormalX": nrm[:, 0], "NormalY": nrm[:, 1], "NormalZ": nrm[:, 2], "saliency": sal,
             "counts": np.arange(len(xyz), dtype=np.int16)}
    io.write_las(path, xyz, extra, version=version, point_format=fmt)
    las = io.parse_las(path)
    assert las.header["version"] == version and las.header["point_format"] == fmt and las.header["n"] == len(xyz)
    assert np.abs(np.column_stack((las.x, las.y, las.z)) - xyz).max() <= 0.0005 + 1e-9      # 1 mm quantisation
    d = io.read_las(path, ["counts"])
    back = d["xyz"] + d["origin"]
    assert np.abs(back - xyz).max() <= 0.0005 + 1e-9
    np.testing.assert_allclose(d["xyz"].mean(axis=0), 0.0, atol=1e-9)                      # centred, :63-64
    assert np.array_equal(d["curvature"], curv) and d["curvature"].dtype == np.float64
    assert np.array_equal(d["normals"], nrm)
    assert np.array_equal(d["saliency"], sal) and d["saliency"].dtype == np.float32
    assert np.array_equal(d["counts"], np.arange(len(xyz)))
    with pytest.raises(AttributeError):
        io.read_las(path, ["missing_dimension"])


def test_las_hand_built_file(tmp_path):
    """LAS 1.2, point format 0, one VLR describing two extra-bytes dimensions (a scaled uint16 and a float64),
    assembled with struct.pack straight from the specification -- independent of write_las."""
    X = np.array([[1000, 2000, 3000], [-500, 0, 12], [2147483647, -2147483648, 7]], dtype=np.int64)
    scale, offset = (0.01, 0.01, 0.001), (100.0, -50.0, 10.0)
    h = bytearray(227)
    h[0:4] = b"LASF"
    h[24], h[25] = 1, 2
    struct.pack_into("<H", h, 94, 227)
    descr = bytearray(384)
    descr[2], descr[3] = 3, 0x18                      # unsigned short, scale + offset relevant
    descr[4:9] = b"range"
    struct.pack_into("<d", descr, 112, 0.5)
    struct.pack_into("<d", descr, 136, 2.0)
    descr[192 + 2], descr[192 + 3] = 10, 0           # double
    descr[192 + 4:192 + 13] = b"Curvature"
    vlr = struct.pack("<H16sHH32s", 0, b"LASF_Spec", 4, len(descr), b"") + bytes(descr)
    other = struct.pack("<H16sHH32s", 0, b"someone", 77, 5, b"") + b"hello"       # an unrelated VLR in front
    struct.pack_into("<II", h, 96, 227 + len(other) + len(vlr), 2)
    h[104] = 0
    struct.pack_into("<H", h, 105, 20 + 2 + 8)
    struct.pack_into("<I", h, 107, 3)
    struct.pack_into("<3d", h, 131, *scale)
    struct.pack_into("<3d", h, 155, *offset)
    body = b""
    for i in range(3):
        body += struct.pack("<iiiHBBbBH", int(X[i, 0]), int(X[i, 1]), int(X[i, 2]), 7 + i, 0x11, 2, -3, 9, 42)
        body += struct.pack("<Hd", 10 * (i + 1), 0.25 * i)
    path = str(tmp_path / "h.las")
    with open(path, "wb") as f:
        f.write(bytes(h) + other + vlr + body)
    las = io.parse_las(path)
    assert np.array_equal(las.X, X[:, 0]) and np.array_equal(las.Z, X[:, 2])
    np.testing.assert_array_equal(las.x, X[:, 0] * 0.01 + 100.0)
    np.testing.assert_array_equal(las.z, X[:, 2] * 0.001 + 10.0)
    assert np.array_equal(las.intensity, [7, 8, 9]) and np.array_equal(las.classification, [2, 2, 2])
    assert np.array_equal(las.scan_angle, [-3, -3, -3]) and np.array_equal(las.point_source_id, [42, 42, 42])
    np.testing.assert_array_equal(getattr(las, "range"), np.array([10, 20, 30]) * 0.5 + 2.0)
    np.testing.assert_array_equal(las.Curvature, [0.0, 0.25, 0.5])
    d = io.read_las(path)
    np.testing.assert_array_equal(d["curvature"], [0.0, 0.25, 0.5])
    assert "normals" not in d and "saliency" not in d


def test_las_rejects_bad_input(tmp_path):
    p = str(tmp_path / "x.las")
    open(p, "wb").write(b"NOPE" + bytes(300))
    with pytest.raises(ValueError):
        io.parse_las(p)
    xyz = _cloud(10)[0]
    io.write_las(p, xyz)
    raw = bytearray(open(p, "rb").read())
    raw[104] |= 0x80                                   # LAZ marker
    open(p, "wb").write(bytes(raw))
    with pytest.raises(NotImplementedError):
        io.parse_las(p)
    io.write_las(p, xyz)
    raw = open(p, "rb").read()
    open(p, "wb").write(raw[:-5])                      # truncated
    with pytest.raises(ValueError):
        io.parse_las(p)
    io.write_las(p, xyz[:0])                           # empty cloud
    assert io.parse_las(p).header["n"] == 0


@pytest.mark.parametrize("binary", [True, False])
def test_ply_round_trip(tmp_path, binary):
    xyz, nrm, _, sal = _cloud(300, seed=3)
    path = str(tmp_path / "a.ply")
    cols = {"x": xyz[:, 0].astype(np.float32), "y": xyz[:, 1].astype(np.float32), "z": xyz[:, 2].astype(np.float32),
            "nx": nrm[:, 0], "ny": nrm[:, 1], "nz": nrm[:, 2], "scalar_saliency": sal,
            "label": np.arange(300, dtype=np.uint8)}
    io.write_ply(path, cols, binary=binary)
    v = io.parse_ply_vertices(path)
    assert v.dtype.names == tuple(cols) and np.array_equal(v["label"], cols["label"])
    d = io.read_ply(path)
    want = np.column_stack([cols["x"], cols["y"], cols["z"]]).astype(np.float64)
    np.testing.assert_array_equal(d["origin"], want.mean(axis=0))
    np.testing.assert_array_equal(d["xyz"], want - want.mean(axis=0))
    np.testing.assert_array_equal(d["normals"], nrm)
    np.testing.assert_array_equal(d["saliency"], sal.astype(np.float64))


def test_ply_hand_written_with_faces(tmp_path):
    txt = ("ply\r\nformat ascii 1.0\r\ncomment made by hand\r\nelement vertex 3\r\nproperty float x\r\n"
           "property float y\r\nproperty float z\r\nproperty float nx\r\nproperty float ny\r\nproperty float nz\r\n"
           "property double scalar_saliency\r\nelement face 1\r\nproperty list uchar int vertex_indices\r\n"
           "end_header\r\n0 0 0 0 0 1 0.5\r\n1 0 0 0 0 1 0.25\r\n0 2 1 0 1 0 1e-3\r\n3 0 1 2\r\n")
    p = str(tmp_path / "h.ply")
    open(p, "wb").write(txt.encode())
    d = io.read_ply(p)
    np.testing.assert_allclose(d["origin"], [1 / 3, 2 / 3, 1 / 3])
    np.testing.assert_allclose(d["xyz"] + d["origin"], [[0, 0, 0], [1, 0, 0], [0, 2, 1]], atol=1e-15)
    np.testing.assert_array_equal(d["saliency"], [0.5, 0.25, 1e-3])
    np.testing.assert_array_equal(d["normals"][2], [0, 1, 0])
    open(p, "wb").write(b"plx\nend_header\n")
    with pytest.raises(ValueError):
        io.read_ply(p)


def test_levelsets_func_exports_readers():
    from more3d_b200 import levelsets_func as ls
    assert ls.read_las is io.read_las and ls.read_ply is io.read_ply
