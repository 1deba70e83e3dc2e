"""Stdlib reader for the reference's mock .vtu fixtures (TEST INFRASTRUCTURE, build container only).

Format: XML VTKFile/UnstructuredGrid/Piece; every DataArray format="binary" is
base64(header) + base64(zlib blocks), header = [nblocks, blocksize, lastblocksize, csize x nblocks]
as uint32 (compressor vtkZLibDataCompressor).  meshio is not installable here."""
from __future__ import annotations

import base64
import xml.etree.ElementTree as ET
import zlib

import numpy as np

_DT = {"Float32": np.float32, "Float64": np.float64, "Int32": np.int32, "Int64": np.int64, "UInt8": np.uint8,
       "UInt32": np.uint32, "UInt64": np.uint64, "Int8": np.int8}


def _decode(text: str, dtype, header_dtype=np.uint32) -> np.ndarray:
    text = "".join(text.split())
    hsz = np.dtype(header_dtype).itemsize
    first = base64.b64decode(text[: ((3 * hsz + 2) // 3) * 4])
    nblocks = int(np.frombuffer(first[:hsz], header_dtype)[0])
    hbytes = (3 + nblocks) * hsz
    hchars = ((hbytes + 2) // 3) * 4
    header = np.frombuffer(base64.b64decode(text[:hchars])[:hbytes], header_dtype)
    csizes = header[3:3 + nblocks].astype(np.int64)
    data = base64.b64decode(text[hchars:])
    out, off = [], 0
    for cs in csizes:
        out.append(zlib.decompress(data[off:off + cs]))
        off += cs
    return np.frombuffer(b"".join(out), dtype)


def read_vtu(path: str):
    """Returns dict(points (N,3), cells (C,k) with k = 3 triangles / 4 tetrahedra, point_data {name: array})."""
    root = ET.parse(path).getroot()
    hdt = _DT.get(root.attrib.get("header_type", "UInt32"), np.uint32)
    piece = root.find("UnstructuredGrid/Piece")
    n_points = int(piece.attrib["NumberOfPoints"])

    def arr(el):
        a = _decode(el.text, _DT[el.attrib["type"]], hdt)
        nc = int(el.attrib.get("NumberOfComponents", "1"))
        return a.reshape(-1, nc) if nc > 1 else a

    points = arr(piece.find("Points/DataArray")).reshape(n_points, 3)
    cells = {d.attrib["Name"]: arr(d) for d in piece.find("Cells")}
    k = int(cells["offsets"][0])
    conn = cells["connectivity"].reshape(-1, k).astype(np.int64)
    pdata = {d.attrib["Name"]: arr(d) for d in (piece.find("PointData") or [])}
    return {"points": points, "cells": conn, "point_data": pdata}
