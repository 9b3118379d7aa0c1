"""VTK-free reader / writer for VTK XML rectilinear grids (``.vtr``).

The reference reads models with the python ``vtk`` wheel
(``vtkXMLRectilinearGridReader``, src/ttcrpy/rgrid.pyx:1344-1372) and writes traveltime
fields with ``Grid3Drn::saveTT`` format 2 (ttcr/Grid3Drn.h:2696-2746).  This module handles
the subset those files use -- inline ``DataArray`` elements with ``format="ascii"`` or
``format="binary"`` (base64, optional ``vtkZLibDataCompressor``, ``header_type`` UInt32 or
UInt64) -- with ``re`` + ``base64`` + ``zlib`` + numpy only.

Arrays are returned flat in VTK order (x fastest), exactly what ``vtk_to_numpy`` gives.
"""
from __future__ import annotations

import base64
import re
import zlib

import numpy as np

_VTK_TYPES = {
    "Float64": np.float64, "Float32": np.float32, "Int32": np.int32, "UInt32": np.uint32,
    "Int64": np.int64, "UInt64": np.uint64, "Int8": np.int8, "UInt8": np.uint8,
    "Int16": np.int16, "UInt16": np.uint16,
}
_NP_TO_VTK = {np.dtype(v): k for k, v in _VTK_TYPES.items()}


def _b64_block(text: str, start: int, nbytes: int):
    """decode the base64 block that begins at ``start`` and holds ``nbytes`` raw bytes"""
    nchar = (nbytes + 2) // 3 * 4
    return base64.b64decode(text[start:start + nchar]), start + nchar


def _decode_binary(text: str, dtype, header_dtype, compressed: bool) -> np.ndarray:
    text = "".join(text.split())
    hsz = np.dtype(header_dtype).itemsize
    if not compressed:
        raw, pos = _b64_block(text, 0, hsz)
        n = int(np.frombuffer(raw[:hsz], dtype=header_dtype)[0])
        # the length word is either encoded together with the data or as a block of its own
        data = base64.b64decode(text[:(hsz + n + 2) // 3 * 4])[hsz:hsz + n]
        if len(data) < n:
            data = base64.b64decode(text[pos:])[:n]
        return np.frombuffer(data, dtype=dtype).copy()
    raw, _ = _b64_block(text, 0, 3 * hsz)
    nblocks = int(np.frombuffer(raw, dtype=header_dtype)[0])
    raw, pos = _b64_block(text, 0, (3 + nblocks) * hsz)
    head = np.frombuffer(raw, dtype=header_dtype)
    sizes = head[3:3 + nblocks].astype(np.int64)
    comp = base64.b64decode(text[pos:])
    out, off = [], 0
    for s in sizes:
        out.append(zlib.decompress(comp[off:off + int(s)]))
        off += int(s)
    return np.frombuffer(b"".join(out), dtype=dtype).copy()


def read_vtr(filename: str) -> dict:
    """Read a ``.vtr`` file.

    Returns ``{"x","y","z": coordinates, "point_data": {name: flat array}, "cell_data": {...}}``.
    """
    with open(filename, "r", encoding="latin-1") as f:
        txt = f.read()
    m = re.search(r"<VTKFile([^>]*)>", txt)
    if not m or 'type="RectilinearGrid"' not in m.group(1):
        raise ValueError(f"{filename}: not a VTK XML RectilinearGrid file")
    attrs = dict(re.findall(r'(\w+)="([^"]*)"', m.group(1)))
    if attrs.get("byte_order", "LittleEndian") != "LittleEndian":
        raise ValueError("only LittleEndian .vtr files are supported")
    header_dtype = np.uint64 if attrs.get("header_type", "UInt32") == "UInt64" else np.uint32
    compressed = "compressor" in attrs
    if compressed and attrs["compressor"] != "vtkZLibDataCompressor":
        raise ValueError(f"unsupported compressor {attrs['compressor']}")

    def arrays(section: str) -> dict:
        sm = re.search(r"<%s[^>]*?(?:/>|>(.*?)</%s>)" % (section, section), txt, re.S)
        out = {}
        if not sm or sm.group(1) is None:
            return out
        for am in re.finditer(r"<DataArray([^>]*?)(?:/>|>(.*?)</DataArray>)", sm.group(1), re.S):
            a = dict(re.findall(r'(\w+)="([^"]*)"', am.group(1)))
            dtype = _VTK_TYPES[a["type"]]
            body = am.group(2) or ""
            fmt = a.get("format", "ascii")
            if fmt == "ascii":
                arr = np.array(body.split(), dtype=dtype)
            elif fmt == "binary":
                arr = _decode_binary(body, dtype, header_dtype, compressed)
            else:
                raise ValueError(f"unsupported DataArray format '{fmt}'")
            ncomp = int(a.get("NumberOfComponents", "1"))
            if ncomp > 1:
                arr = arr.reshape(-1, ncomp)
            out[a.get("Name", f"array{len(out)}")] = arr
        return out

    coords = list(arrays("Coordinates").values())
    if len(coords) != 3:
        raise ValueError(f"{filename}: expected 3 coordinate arrays")
    return {"x": coords[0], "y": coords[1], "z": coords[2],
            "point_data": arrays("PointData"), "cell_data": arrays("CellData")}


def _encode(arr: np.ndarray, compress: bool) -> str:
    raw = np.ascontiguousarray(arr).tobytes()
    if not compress:
        return base64.b64encode(np.uint32(len(raw)).tobytes() + raw).decode()
    block = 1 << 15
    chunks = [raw[i:i + block] for i in range(0, len(raw), block)] or [b""]
    comp = [zlib.compress(c) for c in chunks]
    last = len(chunks[-1]) if len(chunks[-1]) != block else 0
    head = np.array([len(chunks), block, last] + [len(c) for c in comp], dtype=np.uint32)
    return base64.b64encode(head.tobytes()).decode() + base64.b64encode(b"".join(comp)).decode()


def write_vtr(filename: str, x, y, z, point_data: dict | None = None, cell_data: dict | None = None,
              compress: bool = True) -> None:
    """Write a ``.vtr`` file readable by VTK/ParaView and by :func:`read_vtr`.

    Data arrays must be flat in VTK order (x fastest) -- the order ``Grid3Drn::saveTT`` format 2
    writes its "Travel time" point array in (ttcr/Grid3Drn.h:2716-2735).
    """
    x, y, z = (np.asarray(a) for a in (x, y, z))
    ext = f"0 {x.size - 1} 0 {y.size - 1} 0 {z.size - 1}"
    comp_attr = ' compressor="vtkZLibDataCompressor"' if compress else ""

    def da(name, arr):
        arr = np.asarray(arr)
        vt = _NP_TO_VTK[arr.dtype]
        return (f'      <DataArray type="{vt}" Name="{name}" format="binary">\n'
                f"        {_encode(arr.ravel(), compress)}\n      </DataArray>\n")

    s = ['<?xml version="1.0"?>\n',
         f'<VTKFile type="RectilinearGrid" version="0.1" byte_order="LittleEndian" header_type="UInt32"{comp_attr}>\n',
         f'  <RectilinearGrid WholeExtent="{ext}">\n', f'  <Piece Extent="{ext}">\n']
    for section, data, n in (("PointData", point_data, x.size * y.size * z.size),
                             ("CellData", cell_data, (x.size - 1) * (y.size - 1) * (z.size - 1))):
        s.append(f"    <{section}>\n")
        for name, arr in (data or {}).items():
            if np.asarray(arr).size != n:
                raise ValueError(f"{section} array '{name}' has {np.asarray(arr).size} values, expected {n}")
            s.append(da(name, arr))
        s.append(f"    </{section}>\n")
    s.append("    <Coordinates>\n")
    for name, arr in (("x", x), ("y", y), ("z", z)):
        s.append(da(name, arr))
    s.append("    </Coordinates>\n  </Piece>\n  </RectilinearGrid>\n</VTKFile>\n")
    with open(filename, "w") as f:
        f.write("".join(s))
