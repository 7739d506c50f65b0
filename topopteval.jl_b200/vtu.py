"""Minimal VTU (VTK XML UnstructuredGrid) reader / writer for the harness.

The reference keeps its own Julia loader (`src/MeshImport/MeshImport.jl:20-164`,
`extract_cell_density` at `:177-215`) and writer (`src/ResultsExport/ResultsExport.jl:25-37`);
the north star leaves those in Julia.  This module exists only so that the Python host
mirror, the tests and `bench.py` can read the two fixtures of the reference and dump results
in a format the reference's tools can diff.  Dialect handled (the one both fixtures use and
the one WriteVTK emits by default): `header_type="UInt64"`, `vtkZLibDataCompressor`,
`<AppendedData encoding="raw">`; also uncompressed appended-raw and inline ascii.

Connectivity in the file is 0-based; `read_vtu` returns it **1-based** exactly as ReadVTK
hands it to the reference (`MeshImport.jl:45-86`).
"""
from __future__ import annotations

import re
import struct
import zlib
from dataclasses import dataclass, field

import numpy as np

_VTK_DTYPES = {
    "Float64": "<f8", "Float32": "<f4", "Int64": "<i8", "Int32": "<i4",
    "UInt64": "<u8", "UInt32": "<u4", "UInt8": "u1", "Int8": "i1",
    "Int16": "<i2", "UInt16": "<u2",
}
# names the reference accepts for the density cell array (MeshImport.jl:193-194)
DENSITY_NAMES = ("density", "rho", "Density", "DENSITY", "volfrac", "VolFrac", "vol_frac")
VTK_TETRA, VTK_HEXAHEDRON = 10, 12


@dataclass
class VtuMesh:
    points: np.ndarray            # (nn, 3) float64
    cells: np.ndarray             # (ne, npc) int64, 1-based
    cell_type: int                # 10 tet / 12 hex (dominant type, MeshImport.jl:97-125)
    cell_data: dict = field(default_factory=dict)
    point_data: dict = field(default_factory=dict)


def _attrs(tag: str) -> dict:
    return dict(re.findall(r'(\w+)="([^"]*)"', tag))


def _read_appended(blob: bytes, offset: int, dtype: str, compressed: bool, hdr: str) -> np.ndarray:
    hsz = struct.calcsize(hdr)
    p = offset
    if compressed:
        nblocks, _bs, _last = struct.unpack_from("<3" + hdr[-1], blob, p)
        p += 3 * hsz
        csizes = struct.unpack_from("<%d%s" % (nblocks, hdr[-1]), blob, p)
        p += nblocks * hsz
        out = bytearray()
        for cs in csizes:
            out += zlib.decompress(blob[p:p + cs])
            p += cs
        return np.frombuffer(bytes(out), dtype=dtype)
    (nbytes,) = struct.unpack_from(hdr, blob, p)
    p += hsz
    return np.frombuffer(blob[p:p + nbytes], dtype=dtype)


_NODES_PER_CELL = {5: 3, 9: 4, VTK_TETRA: 4, VTK_HEXAHEDRON: 8}


def read_vtu(path: str, cell_types=(VTK_TETRA, VTK_HEXAHEDRON)) -> VtuMesh:
    """`cell_types`: VTK types accepted as the dominant type — volume meshes by default (`import_mesh`); pass (5, 9) to read back
    the triangle / quad file `export_boundary_conditions` writes."""
    raw = open(path, "rb").read()
    m = re.search(rb"<AppendedData[^>]*>\s*_", raw)
    xml = raw[: m.start()] if m else raw
    blob = raw[m.end():] if m else b""
    text = xml.decode("utf-8", errors="replace")
    vf = _attrs(re.search(r"<VTKFile[^>]*>", text).group(0))
    hdr = "<Q" if vf.get("header_type", "UInt32") == "UInt64" else "<I"
    compressed = "compressor" in vf

    def arrays(section: str) -> dict:
        sec = re.search(r"<%s[^>]*>(.*?)</%s>" % (section, section), text, re.S)
        out = {}
        if not sec:
            return out
        for tag, body in re.findall(r"(<DataArray[^>]*?)(?:/>|>(.*?)</DataArray>)", sec.group(1), re.S):
            a = _attrs(tag)
            dt = _VTK_DTYPES[a["type"]]
            if a.get("format") == "appended":
                arr = _read_appended(blob, int(a["offset"]), dt, compressed, hdr)
            elif a.get("format") == "ascii":
                arr = np.array(body.split(), dtype=np.dtype(dt).newbyteorder("="))
            else:
                raise ValueError("unsupported DataArray format %r" % a.get("format"))
            nc = int(a.get("NumberOfComponents", "1"))
            if nc > 1:
                arr = arr.reshape(-1, nc)
            out[a.get("Name", section)] = np.ascontiguousarray(arr)
        return out

    pts = list(arrays("Points").values())[0].astype(np.float64).reshape(-1, 3)
    cells = arrays("Cells")
    conn = cells["connectivity"].astype(np.int64)
    offs = cells["offsets"].astype(np.int64)
    types = cells["types"].astype(np.int64)
    # dominant-type selection, as MeshImport.jl:97-125 does
    vals, counts = np.unique(types, return_counts=True)
    dom = int(vals[np.argmax(counts)])
    if dom not in cell_types or dom not in _NODES_PER_CELL:
        raise ValueError("only Tet4 (10) / Hex8 (12) meshes are on the hot path, got VTK type %d" % dom)
    npc = _NODES_PER_CELL[dom]
    starts = np.concatenate(([0], offs[:-1]))
    sel = np.nonzero(types == dom)[0]
    idx = starts[sel][:, None] + np.arange(npc)[None, :]
    cell_nodes = conn[idx] + 1                       # → 1-based (ReadVTK does this shift)
    cd = {k: v[sel] if len(v) == len(types) else v for k, v in arrays("CellData").items()}
    return VtuMesh(pts, np.ascontiguousarray(cell_nodes), dom, cd, arrays("PointData"))


def extract_cell_density(path: str) -> np.ndarray:
    """Mirror of `extract_cell_density` (MeshImport.jl:177-215): first matching name wins."""
    mesh = read_vtu(path)
    for name in DENSITY_NAMES:
        if name in mesh.cell_data:
            return np.asarray(mesh.cell_data[name], dtype=np.float64)
    raise KeyError("No density data found in %s (looked for %s)" % (path, ", ".join(DENSITY_NAMES)))


def write_vtu(path: str, points: np.ndarray, cells_1based: np.ndarray, cell_type: int,
              point_data: dict | None = None, cell_data: dict | None = None) -> str:
    """Appended-raw + zlib writer (WriteVTK's default dialect, ResultsExport.jl:29-35).
    `point_data["u"]` of shape (nn,3) is what `export_results(u, dh, file)` produces."""
    if not path.endswith(".vtu"):
        path += ".vtu"
    points = np.ascontiguousarray(points, dtype="<f8")
    cells0 = np.ascontiguousarray(cells_1based, dtype="<i8") - 1
    ne, npc = cells0.shape
    chunks, xml_arrays = [], {"Points": [], "Cells": [], "PointData": [], "CellData": []}
    off = 0

    def add(section, name, arr, vtktype, ncomp):
        nonlocal off
        data = np.ascontiguousarray(arr).tobytes()
        comp = zlib.compress(data)
        header = struct.pack("<4Q", 1, len(data), len(data), len(comp)) if data else struct.pack("<3Q", 0, 0, 0)
        chunks.append(header + (comp if data else b""))
        xml_arrays[section].append(
            '<DataArray type="%s" Name="%s" NumberOfComponents="%d" format="appended" offset="%d"/>'
            % (vtktype, name, ncomp, off))
        off += len(chunks[-1])

    add("Points", "Points", points, "Float64", 3)
    add("Cells", "connectivity", cells0.reshape(-1), "Int64", 1)
    add("Cells", "offsets", (np.arange(1, ne + 1, dtype="<i8") * npc), "Int64", 1)
    add("Cells", "types", np.full(ne, cell_type, dtype="u1"), "UInt8", 1)
    for sec, dd in (("PointData", point_data or {}), ("CellData", cell_data or {})):
        for k, v in dd.items():
            v = np.asarray(v)
            if np.issubdtype(v.dtype, np.integer):         # WriteVTK keeps Int data as Int64 (boundary_type, ResultsExport.jl:189)
                v = np.ascontiguousarray(v, dtype="<i8"); vt = "Int64"
            else:
                v = np.ascontiguousarray(v, dtype="<f8"); vt = "Float64"
            add(sec, k, v, vt, 1 if v.ndim == 1 else v.shape[1])
    lines = ['<?xml version="1.0" encoding="utf-8"?>',
             '<VTKFile type="UnstructuredGrid" version="1.0" byte_order="LittleEndian" header_type="UInt64" '
             'compressor="vtkZLibDataCompressor">', "  <UnstructuredGrid>",
             '    <Piece NumberOfPoints="%d" NumberOfCells="%d">' % (len(points), ne)]
    for sec in ("Points", "Cells", "PointData", "CellData"):
        if xml_arrays[sec]:
            lines.append("      <%s>" % sec)
            lines += ["        " + a for a in xml_arrays[sec]]
            lines.append("      </%s>" % sec)
    lines += ["    </Piece>", "  </UnstructuredGrid>", '  <AppendedData encoding="raw">']
    with open(path, "wb") as fh:
        fh.write(("\n".join(lines) + "\n_").encode())
        for c in chunks:
            fh.write(c)
        fh.write(b"\n  </AppendedData>\n</VTKFile>\n")
    return path
