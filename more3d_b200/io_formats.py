"""Point-cloud file adapters either side of the hot path: LAS (1.2-1.4, uncompressed) and PLY (ascii / binary).

The reference reads its inputs with ``laspy`` and ``plyfile`` (levelsets_func.py:43-84, run_levelsets.py:129,
timeseries/rateLAS.py:19-42); neither package is available offline, so these are self-contained numpy readers
written from the published format specifications (ASPRS LAS 1.4 R15; Bourke's PLY description).  They return
exactly the dictionaries ``read_las`` / ``read_ply`` of the reference return -- ``xyz`` centred on its mean,
``origin``, and the optional ``curvature`` / ``normals`` / ``saliency`` / requested attribute columns.
Parity is UNPINNED against laspy / plyfile (they cannot run here); tests/test_io_formats.py checks the readers
against files built byte by byte from the specifications and against round trips through the writers below.
Host-side code: file parsing is not on the GPU path.
"""
import struct

import numpy as np

# ---------------------------------------------------------------------------------------------------------
# LAS
# ---------------------------------------------------------------------------------------------------------
# size of the standard part of a point record per point data format (LAS 1.4 R15, tables 7-17)
_LAS_BASE = {0: 20, 1: 28, 2: 26, 3: 34, 4: 57, 5: 63, 6: 30, 7: 36, 8: 38, 9: 59, 10: 67}
# extra-bytes data types (table 24): code -> numpy dtype
_LAS_EB = {1: "u1", 2: "i1", 3: "<u2", 4: "<i2", 5: "<u4", 6: "<i4", 7: "<u8", 8: "<i8", 9: "<f4", 10: "<f8"}
_LAS_EB_CODE = {np.dtype(v).str.lstrip("|"): k for k, v in _LAS_EB.items()}


def _cstr(b):
    return b.split(b"\0", 1)[0].decode("latin-1")


class LasFile:
    """Parsed LAS file: ``x``/``y``/``z`` (scaled float64), ``X``/``Y``/``Z`` (raw int32), the standard fields as a
    structured array (``points``) and the extra-bytes dimensions by name (attribute access, like laspy)."""

    def __init__(self, header, points, extra):
        self.header, self.points, self.extra = header, points, extra

    def __getattr__(self, name):
        extra = self.__dict__.get("extra", {})
        if name in extra:
            return extra[name]
        if name in ("X", "Y", "Z"):
            return self.points[name]
        if name in ("x", "y", "z"):
            i = "xyz".index(name)
            return self.points[name.upper()] * self.header["scale"][i] + self.header["offset"][i]
        pts = self.__dict__.get("points")
        if pts is not None and pts.dtype.names and name in pts.dtype.names:
            return pts[name]
        raise AttributeError(name)


def parse_las(filename):
    """Read an uncompressed LAS 1.0-1.4 file (LAZ needs a LASzip decoder and is rejected)."""
    with open(filename, "rb") as f:
        raw = f.read()
    if raw[:4] != b"LASF":
        raise ValueError("%s: not a LAS file (signature %r)" % (filename, raw[:4]))
    major, minor = raw[24], raw[25]
    header_size, = struct.unpack_from("<H", raw, 94)
    offset_to_points, n_vlr = struct.unpack_from("<II", raw, 96)
    fmt_byte = raw[104]
    rec_len, = struct.unpack_from("<H", raw, 105)
    n_legacy, = struct.unpack_from("<I", raw, 107)
    scale = struct.unpack_from("<3d", raw, 131)
    offset = struct.unpack_from("<3d", raw, 155)
    if fmt_byte & 0x80 or fmt_byte & 0x40:
        raise NotImplementedError("%s: compressed (LAZ) point data is not supported" % filename)
    fmt = fmt_byte & 0x3f
    if fmt not in _LAS_BASE:
        raise ValueError("%s: unknown point data format %d" % (filename, fmt))
    n = n_legacy
    if (major, minor) >= (1, 4) and header_size >= 375:
        n64, = struct.unpack_from("<Q", raw, 247)
        if n64:
            n = n64
    base = _LAS_BASE[fmt]
    if rec_len < base:
        raise ValueError("%s: record length %d is shorter than format %d needs (%d)" % (filename, rec_len, fmt, base))
    # variable length records: only the extra-bytes description (LASF_Spec / 4) matters here
    eb = []
    pos = header_size
    for _ in range(n_vlr):
        user = _cstr(raw[pos + 2:pos + 18])
        rec_id, rlen = struct.unpack_from("<HH", raw, pos + 18)
        body = raw[pos + 54:pos + 54 + rlen]
        if user == "LASF_Spec" and rec_id == 4:
            for k in range(len(body) // 192):
                d = body[192 * k:192 * (k + 1)]
                dtype, options = d[2], d[3]
                name = _cstr(d[4:36])
                sc = struct.unpack_from("<3d", d, 112)
                of = struct.unpack_from("<3d", d, 136)
                eb.append(dict(name=name, type=dtype, options=options, scale=sc[0], offset=of[0]))
        pos += 54 + rlen
    fields = [("X", "<i4"), ("Y", "<i4"), ("Z", "<i4"), ("intensity", "<u2")]
    used = 14
    if fmt >= 6:
        fields += [("flags", "<u2"), ("classification", "u1"), ("user_data", "u1"), ("scan_angle", "<i2"),
                   ("point_source_id", "<u2"), ("gps_time", "<f8")]
        used = 30
    else:
        fields += [("flags", "u1"), ("classification", "u1"), ("scan_angle", "i1"), ("user_data", "u1"),
                   ("point_source_id", "<u2")]
        used = 20
        if fmt in (1, 3, 4, 5):
            fields.append(("gps_time", "<f8"))
            used += 8
    if used < base:
        fields.append(("_rest", "V%d" % (base - used)))      # colour / NIR / wave packets: carried, not decoded
    extra_off = base
    extra_fields = []
    for e in eb:
        if e["type"] == 0:
            size = e["options"]
            dt = "V%d" % size if size else None
        elif e["type"] in _LAS_EB:
            dt = _LAS_EB[e["type"]]
            size = np.dtype(dt).itemsize
        else:
            raise NotImplementedError("%s: extra-bytes data type %d (deprecated tuple types) is not supported"
                                      % (filename, e["type"]))
        if dt is not None and extra_off + size <= rec_len:
            extra_fields.append((e, "_eb%d" % len(extra_fields), dt))
            fields.append((extra_fields[-1][1], dt))
        extra_off += size
    if extra_off < rec_len:
        fields.append(("_pad", "V%d" % (rec_len - extra_off)))
    dt = np.dtype(fields)
    assert dt.itemsize == rec_len, (dt.itemsize, rec_len)
    if offset_to_points + n * rec_len > len(raw):
        raise ValueError("%s: truncated point data (%d records of %d bytes expected)" % (filename, n, rec_len))
    pts = np.frombuffer(raw, dtype=dt, count=n, offset=offset_to_points)
    extra = {}
    for e, key, _ in extra_fields:
        v = pts[key]
        if e["type"] and (e["options"] & 0x18):               # scale / offset relevant (bits 3, 4)
            s = e["scale"] if e["options"] & 0x08 else 1.0
            o = e["offset"] if e["options"] & 0x10 else 0.0
            v = v * s + o
        extra[e["name"]] = np.array(v)
    header = dict(version=(major, minor), point_format=fmt, record_length=rec_len, n=int(n), scale=scale,
                  offset=offset, extra_dims=[e["name"] for e, _, _ in extra_fields])
    return LasFile(header, pts, extra)


def read_las(filename, attributes=None):
    """``levelsets_func.read_las`` (levelsets_func.py:59-84): centred xyz, origin and the optional columns."""
    las = parse_las(filename)
    data = {}
    data['xyz'] = np.column_stack((las.x, las.y, las.z))
    data['origin'] = data['xyz'].mean(axis=0)
    data['xyz'] -= data['origin']
    if attributes:
        for attribute in attributes:
            data[str(attribute)] = getattr(las, attribute)        # AttributeError for a missing one, like laspy
    if 'Curvature' in las.extra:
        data['curvature'] = las.extra['Curvature']
    if all(k in las.extra for k in ('NormalX', 'NormalY', 'NormalZ')):
        data['normals'] = np.column_stack((las.extra['NormalX'], las.extra['NormalY'], las.extra['NormalZ']))
    if 'saliency' in las.extra:
        data['saliency'] = las.extra['saliency']
    return data


def write_las(filename, xyz, extra=None, scale=(0.001, 0.001, 0.001), offset=None, version=(1, 4), point_format=6):
    """Minimal LAS writer (uncompressed; formats 0, 1, 6): coordinates plus float / integer extra-bytes dimensions.
    What timeseries/rateLAS.py:19-42 does with laspy (LAS 1.4, point format 6, float64 extra dimensions)."""
    xyz = np.asarray(xyz, dtype=np.float64)
    n = len(xyz)
    extra = dict(extra or {})
    if point_format not in (0, 1, 6):
        raise NotImplementedError("writer supports point formats 0, 1 and 6")
    if offset is None:
        offset = np.floor(xyz.min(axis=0)) if n else np.zeros(3)
    scale, offset = np.asarray(scale, float), np.asarray(offset, float)
    base = _LAS_BASE[point_format]
    fields = [("X", "<i4"), ("Y", "<i4"), ("Z", "<i4"), ("_std", "V%d" % (base - 12))]
    descr = b""
    for name, col in extra.items():
        col = np.asarray(col)
        code = _LAS_EB_CODE.get(col.dtype.newbyteorder("<").str.lstrip("|"))
        if code is None or len(col) != n:
            raise ValueError("extra dimension %r: unsupported dtype %s or wrong length" % (name, col.dtype))
        fields.append((name, _LAS_EB[code]))
        d = bytearray(192)
        d[2], d[3] = code, 0
        nm = name.encode("latin-1")[:31]
        d[4:4 + len(nm)] = nm
        descr += bytes(d)
    dt = np.dtype(fields)
    pts = np.zeros(n, dtype=dt)
    q = np.rint((xyz - offset) / scale)
    if n and (np.abs(q).max() >= 2 ** 31):
        raise ValueError("coordinates do not fit int32 with this scale / offset")
    pts["X"], pts["Y"], pts["Z"] = q[:, 0].astype("<i4"), q[:, 1].astype("<i4"), q[:, 2].astype("<i4")
    for name, col in extra.items():
        pts[name] = col
    header_size = {0: 227, 1: 227, 2: 227, 3: 235, 4: 375}[version[1]]
    vlr = b""
    if descr:
        vlr = struct.pack("<H16sHH32s", 0, b"LASF_Spec", 4, len(descr), b"extra bytes") + descr
    h = bytearray(header_size)
    h[0:4] = b"LASF"
    h[24], h[25] = version
    h[26:26 + 13] = b"more3d_b200\0\0"
    h[58:58 + 11] = b"more3d_b200"
    struct.pack_into("<H", h, 94, header_size)
    struct.pack_into("<II", h, 96, header_size + len(vlr), 1 if vlr else 0)
    h[104] = point_format
    struct.pack_into("<H", h, 105, dt.itemsize)
    struct.pack_into("<I", h, 107, n if (version[1] < 4 and n < 2 ** 32) or point_format < 6 else 0)
    struct.pack_into("<3d", h, 131, *scale)
    struct.pack_into("<3d", h, 155, *offset)
    if n:
        back = q * scale + offset
        mx, mn = back.max(axis=0), back.min(axis=0)
        struct.pack_into("<6d", h, 179, mx[0], mn[0], mx[1], mn[1], mx[2], mn[2])
    if version[1] >= 4:
        struct.pack_into("<Q", h, 247, n)
    with open(filename, "wb") as f:
        f.write(bytes(h))
        f.write(vlr)
        f.write(pts.tobytes())


# ---------------------------------------------------------------------------------------------------------
# PLY
# ---------------------------------------------------------------------------------------------------------
_PLY_T = {"char": "i1", "int8": "i1", "uchar": "u1", "uint8": "u1", "short": "i2", "int16": "i2", "ushort": "u2",
          "uint16": "u2", "int": "i4", "int32": "i4", "uint": "u4", "uint32": "u4", "float": "f4", "float32": "f4",
          "double": "f8", "float64": "f8"}


def parse_ply_vertices(filename):
    """The ``vertex`` element of a PLY file as a structured numpy array (what ``PlyData.read(f)['vertex'].data`` is)."""
    with open(filename, "rb") as f:
        raw = f.read()
    end = raw.find(b"end_header")
    if not raw.startswith(b"ply") or end < 0:
        raise ValueError("%s: not a PLY file" % filename)
    nl = raw.find(b"\n", end)
    head = raw[:end].decode("latin-1").replace("\r", "").split("\n")
    body = raw[nl + 1:]
    fmt = None
    elements = []            # (name, count, [(prop, type) | (prop, ('list', ctype, itype))])
    for line in head[1:]:
        tok = line.split()
        if not tok or tok[0] in ("comment", "obj_info"):
            continue
        if tok[0] == "format":
            fmt = tok[1]
        elif tok[0] == "element":
            elements.append((tok[1], int(tok[2]), []))
        elif tok[0] == "property":
            if tok[1] == "list":
                elements[-1][2].append((tok[4], ("list", tok[2], tok[3])))
            else:
                if tok[1] not in _PLY_T:
                    raise ValueError("%s: unknown property type %r" % (filename, tok[1]))
                elements[-1][2].append((tok[2], _PLY_T[tok[1]]))
    if fmt not in ("ascii", "binary_little_endian", "binary_big_endian"):
        raise ValueError("%s: unknown PLY format %r" % (filename, fmt))
    if not elements or elements[0][0] != "vertex":
        raise NotImplementedError("%s: the vertex element must come first" % filename)
    _, n, props = elements[0]
    if any(isinstance(t, tuple) for _, t in props):
        raise NotImplementedError("%s: list properties on the vertex element are not supported" % filename)
    if fmt == "ascii":
        dt = np.dtype([(p, "<" + t if t[1] != "1" else t) for p, t in props])
        toks = body.split(None, n * len(props))[:n * len(props)]
        if len(toks) < n * len(props):
            raise ValueError("%s: truncated vertex data" % filename)
        table = np.array(toks, dtype="S").reshape(n, len(props))
        out = np.zeros(n, dtype=dt)
        for c, (p, _) in enumerate(props):
            out[p] = table[:, c].astype(np.float64).astype(dt[p])
        return out
    bo = "<" if fmt == "binary_little_endian" else ">"
    dt = np.dtype([(p, bo + t if t[1] != "1" else t) for p, t in props])
    if len(body) < n * dt.itemsize:
        raise ValueError("%s: truncated vertex data" % filename)
    return np.frombuffer(body, dtype=dt, count=n)


def read_ply(filename):
    """``levelsets_func.read_ply`` (levelsets_func.py:43-56): centred xyz, origin, normals, saliency."""
    v = parse_ply_vertices(filename)
    data = {}
    data['xyz'] = np.column_stack([v[_] for _ in ('x', 'y', 'z')]).view().astype('float64')
    data['origin'] = data['xyz'].mean(axis=0)
    data['xyz'] -= data['origin']
    data['normals'] = np.column_stack([v[_] for _ in ('nx', 'ny', 'nz')]).view().astype('float64')
    data['saliency'] = np.array(v['scalar_saliency']).astype('float64')
    return data


def write_ply(filename, columns, binary=True):
    """Vertex-only PLY from {property: 1-D array} (float32 / float64 / integer columns)."""
    names = list(columns)
    cols = [np.asarray(columns[k]) for k in names]
    n = len(cols[0]) if cols else 0
    inv = {"i1": "char", "u1": "uchar", "i2": "short", "u2": "ushort", "i4": "int", "u4": "uint", "f4": "float",
           "f8": "double"}
    types = []
    for k, c in zip(names, cols):
        key = c.dtype.str.lstrip("<>|=")
        if key not in inv or len(c) != n:
            raise ValueError("column %r: unsupported dtype %s or wrong length" % (k, c.dtype))
        types.append(key)
    head = "ply\nformat %s 1.0\ncomment written by more3d_b200\nelement vertex %d\n" % (
        "binary_little_endian" if binary else "ascii", n)
    head += "".join("property %s %s\n" % (inv[t], k) for k, t in zip(names, types)) + "end_header\n"
    with open(filename, "wb") as f:
        f.write(head.encode("latin-1"))
        if binary:
            rec = np.zeros(n, dtype=[(k, "<" + t if t[1] != "1" else t) for k, t in zip(names, types)])
            for k, c in zip(names, cols):
                rec[k] = c
            f.write(rec.tobytes())
        else:
            for i in range(n):
                f.write((" ".join(repr(c[i].item()) for c in cols) + "\n").encode("latin-1"))
