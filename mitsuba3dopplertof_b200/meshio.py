"""PLY / OBJ / serialized readers for the host-side scene ingest.

Behaviour follows the reference loaders for the parts the hot path observes:
  * src/shapes/ply.cpp: vertex x/y/z, optional nx/ny/nz, optional u/v (or s/t), faces as lists, polygons fan-
    triangulated; when the file has no normals and `face_normals` is false the reference computes smooth
    vertex normals (Mesh::recompute_vertex_normals, src/render/mesh.cpp:283-345) -- done here with the
    same angle weighting.
  * src/shapes/obj.cpp: v / vt / vn / f with index triplets, vertices de-duplicated per (v,vt,vn) key.
  * src/shapes/serialized.cpp:229-392: the `.serialized` container (header 0x041C, version 3 or 4, one zlib stream per
    sub-mesh, end-of-file offset dictionary selected by `shape_index`); single or double precision payload narrowed to
    float32, optional normals / texture coordinates, vertex colours skipped, uint32 indices.
"""
from __future__ import annotations

import struct
import zlib
from typing import Optional, Sequence, Tuple

import numpy as np

__all__ = ["load_mesh", "load_ply", "load_obj", "load_serialized", "write_serialized", "vertex_normals"]

_PLY_TYPES = {
    "char": "b", "int8": "b", "uchar": "B", "uint8": "B", "short": "h", "int16": "h", "ushort": "H", "uint16": "H",
    "int": "i", "int32": "i", "uint": "I", "uint32": "I", "float": "f", "float32": "f", "double": "d", "float64": "d",
}


def vertex_normals(pos: np.ndarray, faces: np.ndarray) -> np.ndarray:
    """Angle-weighted smooth normals ("Computing Vertex Normals from Polygonal Facets", Thuermer & Wuethrich),
    as Mesh::recompute_vertex_normals does."""
    p = pos.astype(np.float64)
    n = np.zeros_like(p)
    v0, v1, v2 = p[faces[:, 0]], p[faces[:, 1]], p[faces[:, 2]]
    fn = np.cross(v1 - v0, v2 - v0)
    ln = np.linalg.norm(fn, axis=1, keepdims=True)
    fn = np.where(ln > 0, fn / np.maximum(ln, 1e-300), 0)
    for k in range(3):
        a = p[faces[:, k]]
        d0 = p[faces[:, (k + 1) % 3]] - a
        d1 = p[faces[:, (k + 2) % 3]] - a
        d0 /= np.maximum(np.linalg.norm(d0, axis=1, keepdims=True), 1e-300)
        d1 /= np.maximum(np.linalg.norm(d1, axis=1, keepdims=True), 1e-300)
        ang = np.arccos(np.clip((d0 * d1).sum(1), -1, 1))
        np.add.at(n, faces[:, k], fn * ang[:, None])
    ln = np.linalg.norm(n, axis=1, keepdims=True)
    n = np.where(ln > 0, n / np.maximum(ln, 1e-300), np.array([1.0, 0.0, 0.0]))
    return n.astype(np.float32)


def load_ply(path: str):
    with open(path, "rb") as f:
        if f.readline().strip() != b"ply":
            raise ValueError(f"{path}: not a PLY file")
        fmt = None
        elements = []
        while True:
            line = f.readline()
            if not line:
                raise ValueError(f"{path}: truncated PLY header")
            tok = line.decode("ascii", "replace").split()
            if not tok or tok[0] == "comment":
                continue
            if tok[0] == "format":
                fmt = tok[1]
            elif tok[0] == "element":
                elements.append([tok[1], int(tok[2]), []])
            elif tok[0] == "property":
                elements[-1][2].append(tok[1:])
            elif tok[0] == "end_header":
                break
        endian = {"ascii": None, "binary_little_endian": "<", "binary_big_endian": ">"}[fmt]
        verts = {}
        faces = []
        for name, count, props in elements:
            scalar = all(p[0] != "list" for p in props)
            if endian is None:
                rows = [f.readline().split() for _ in range(count)]
                if name == "vertex":
                    arr = np.array(rows, dtype=np.float64)
                    for i, p in enumerate(props):
                        verts[p[-1]] = arr[:, i]
                elif name == "face":
                    for r in rows:
                        k = int(r[0])
                        idx = [int(x) for x in r[1:1 + k]]
                        faces += [(idx[0], idx[i], idx[i + 1]) for i in range(1, k - 1)]
            else:
                if scalar:
                    dt = np.dtype([(p[-1], endian + _PLY_TYPES[p[0]]) for p in props])
                    arr = np.frombuffer(f.read(dt.itemsize * count), dtype=dt)
                    if name == "vertex":
                        for p in props:
                            verts[p[-1]] = arr[p[-1]].astype(np.float64)
                else:
                    for _ in range(count):
                        for p in props:
                            if p[0] == "list":
                                ct, it = _PLY_TYPES[p[1]], _PLY_TYPES[p[2]]
                                k = struct.unpack(endian + ct, f.read(struct.calcsize(ct)))[0]
                                idx = struct.unpack(endian + it * k, f.read(struct.calcsize(it) * k))
                                if name == "face" and p[-1] in ("vertex_indices", "vertex_index"):
                                    faces += [(idx[0], idx[i], idx[i + 1]) for i in range(1, k - 1)]
                            else:
                                f.read(struct.calcsize(_PLY_TYPES[p[0]]))
    pos = np.stack([verts["x"], verts["y"], verts["z"]], 1).astype(np.float32)
    nrm = np.stack([verts["nx"], verts["ny"], verts["nz"]], 1).astype(np.float32) if "nx" in verts else None
    uv = None
    for a, b in (("u", "v"), ("s", "t"), ("texture_u", "texture_v")):
        if a in verts:
            uv = np.stack([verts[a], verts[b]], 1).astype(np.float32)
            break
    return pos, np.asarray(faces, np.uint32).reshape(-1, 3), nrm, uv


def load_obj(path: str, flip_tex_coords: bool = True):
    """`flip_tex_coords` (default true, obj.cpp:151,266-267): v = 1 - v in single precision."""
    v, vt, vn, keys, faces = [], [], [], {}, []
    out_p, out_t, out_n = [], [], []
    with open(path, "r") as f:
        for line in f:
            tok = line.split()
            if not tok:
                continue
            if tok[0] == "v":
                v.append([float(x) for x in tok[1:4]])
            elif tok[0] == "vt":
                tv = np.float32(tok[2])
                vt.append([float(tok[1]), float(np.float32(1.0) - tv if flip_tex_coords else tv)])
            elif tok[0] == "vn":
                vn.append([float(x) for x in tok[1:4]])
            elif tok[0] == "f":
                idx = []
                for t in tok[1:]:
                    parts = (t.split("/") + ["", ""])[:3]
                    key = tuple((int(p) if p else 0) for p in parts)
                    key = tuple((k + (len(a) + 1) if k < 0 else k) for k, a in zip(key, (v, vt, vn)))
                    if key not in keys:
                        keys[key] = len(out_p)
                        out_p.append(v[key[0] - 1])
                        out_t.append(vt[key[1] - 1] if key[1] else None)
                        out_n.append(vn[key[2] - 1] if key[2] else None)
                    idx.append(keys[key])
                faces += [(idx[0], idx[i], idx[i + 1]) for i in range(1, len(idx) - 1)]
    pos = np.asarray(out_p, np.float32)
    uv = np.asarray(out_t, np.float32) if out_t and all(t is not None for t in out_t) else None
    nrm = np.asarray(out_n, np.float32) if out_n and all(n is not None for n in out_n) else None
    return pos, np.asarray(faces, np.uint32).reshape(-1, 3), nrm, uv


_SER_HEADER, _SER_V3, _SER_V4 = 0x041C, 0x0003, 0x0004
_SER_NORMALS, _SER_TEXCOORDS, _SER_COLORS, _SER_FACE_NORMALS, _SER_SINGLE, _SER_DOUBLE = 0x1, 0x2, 0x8, 0x10, 0x1000, 0x2000


def load_serialized(path: str, shape_index: int = 0):
    """One sub-mesh of a Mitsuba `.serialized` file (src/shapes/serialized.cpp:229-392). Errors carry the reference's
    messages. Returns (positions, faces, normals | None, texcoords | None) as float32 / uint32."""
    def fail(descr):
        raise ValueError(f'Error while loading serialized file "{path}": {descr}!')

    if shape_index < 0:
        fail("shape index must be nonnegative")
    try:
        data = open(path, "rb").read()
    except OSError:
        fail("file not found")
    if len(data) < 4:
        fail("encountered an invalid file format")
    fmt, version = struct.unpack_from("<HH", data, 0)
    if fmt != _SER_HEADER:
        fail("encountered an invalid file format")
    if version not in (_SER_V3, _SER_V4):
        fail("encountered an incompatible file version")
    offset = 0
    if shape_index != 0:
        count = struct.unpack_from("<I", data, len(data) - 4)[0]
        if shape_index >= count:
            fail(f"Unable to unserialize mesh, shape index is out of range! (requested {shape_index} out of 0..{count - 1})")
        if version == _SER_V4:
            offset = struct.unpack_from("<Q", data, len(data) - 8 * (count - shape_index) - 4)[0]
        else:
            offset = struct.unpack_from("<I", data, len(data) - 4 * (count - shape_index + 1))[0]
    z = zlib.decompressobj()
    raw = z.decompress(data[offset + 4:])       # the sub-mesh's own header (4 bytes) precedes its zlib stream
    pos_in = [0]

    def take(n):
        if pos_in[0] + n > len(raw):
            fail("unexpected end of the compressed stream")
        out = raw[pos_in[0]:pos_in[0] + n]
        pos_in[0] += n
        return out

    flags = struct.unpack("<I", take(4))[0]
    if version == _SER_V4:
        end = raw.index(b"\0", pos_in[0])      # null-terminated shape name
        pos_in[0] = end + 1
    n_vertices, n_faces = struct.unpack("<QQ", take(16))
    dtype = np.dtype("<f8") if flags & _SER_DOUBLE else np.dtype("<f4")

    def array(dim):
        a = np.frombuffer(take(n_vertices * dim * dtype.itemsize), dtype=dtype).reshape(n_vertices, dim)
        return a.astype(np.float32)              # read_helper: (float) values[i]

    pos = array(3)
    nrm = array(3) if flags & _SER_NORMALS else None
    uv = array(2) if flags & _SER_TEXCOORDS else None
    if flags & _SER_COLORS:
        array(3)                                  # advance_helper: vertex colours are skipped
    faces = np.frombuffer(take(n_faces * 12), dtype="<u4").reshape(n_faces, 3).astype(np.uint32)
    return pos, faces, nrm, uv


def write_serialized(path: str, meshes: Sequence[dict], version: int = 4, double_precision: bool = False) -> None:
    """Writes sub-meshes {positions, faces, normals?, texcoords?, colors?, name?} as a `.serialized` file (format of
    src/shapes/serialized.cpp:98-176) -- used to produce test assets the reference's own loader reads."""
    assert version in (_SER_V3, _SER_V4)
    ft = "<f8" if double_precision else "<f4"
    blob, offsets = b"", []
    for m in meshes:
        offsets.append(len(blob))
        flags = _SER_DOUBLE if double_precision else _SER_SINGLE
        body = b""
        if version == _SER_V4:
            body += m.get("name", "mesh").encode("utf-8") + b"\0"
        pos = np.asarray(m["positions"], np.float64)
        faces = np.asarray(m["faces"], np.uint32)
        body += struct.pack("<QQ", len(pos), len(faces)) + pos.astype(ft).tobytes()
        for key, flag in (("normals", _SER_NORMALS), ("texcoords", _SER_TEXCOORDS), ("colors", _SER_COLORS)):
            if m.get(key) is not None:
                flags |= flag
                body += np.asarray(m[key], np.float64).astype(ft).tobytes()
        body += faces.astype("<u4").tobytes()
        blob += struct.pack("<HH", _SER_HEADER, version) + zlib.compress(struct.pack("<I", flags) + body)
    if version == _SER_V4:
        blob += b"".join(struct.pack("<Q", o) for o in offsets)
    else:
        blob += b"".join(struct.pack("<I", o) for o in offsets)
    blob += struct.pack("<I", len(meshes))
    with open(path, "wb") as f:
        f.write(blob)


def load_mesh(path: str, face_normals: bool = False, shape_index: int = 0,
              compute_missing: bool = True, flip_tex_coords: bool = True) -> Tuple[np.ndarray, np.ndarray, Optional[np.ndarray], Optional[np.ndarray]]:
    """`compute_missing=False` leaves the normals None when the file has none, so that the caller can run
    recompute_vertex_normals where the reference runs it: on the positions AFTER to_world (obj.cpp:237,401-403,
    ply.cpp:288,435-437) -- angle weights are not preserved under non-uniform scale or shear."""
    low = path.lower()
    if low.endswith(".serialized"):
        pos, faces, nrm, uv = load_serialized(path, shape_index)
        if face_normals:
            nrm = None                            # serialized.cpp:341-346: normals in the file are skipped
    else:
        pos, faces, nrm, uv = load_ply(path) if low.endswith(".ply") else load_obj(path, flip_tex_coords)
    if nrm is None and not face_normals and compute_missing:
        nrm = vertex_normals(pos, faces)
    return pos, faces, nrm, uv
