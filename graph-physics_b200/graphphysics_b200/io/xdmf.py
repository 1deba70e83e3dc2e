"""XDMF time series (+ HDF5 heavy data) and the reference's two dataset layouts without meshio / h5py.

  TimeSeriesReader / TimeSeriesWriter   the subset of meshio.xdmf's classes the reference uses
                                        (graphphysics/dataset/xdmf_dataset.py:69-110, utils/meshio_mesh.py:119-233)
  meshes_to_xdmf, append_frame_to_xdmf  prediction archives (meshio_mesh.py:119-233: one mesh, point data per time step)
  meshdata_to_graph                     utils/torch_graph.py:137-221: node features = hstack(point data, time), tetrahedra ->
                                        the four triangles of `face`
  XDMFTrajectory, H5Trajectories        frame (t, t+1) -> Data, the indexing of XDMFDataset.__getitem__ (xdmf_dataset.py:69-163)
                                        and of H5Dataset via get_traj_as_meshes / get_frame_as_graph (utils/hierarchical.py:51-170)

File formats handled: XDMF 3 as meshio writes it (a "mesh" grid with Geometry / Topology and a temporal collection of
grids with Time + Attribute items; or one uniform grid), DataItems in HDF ("file.h5:/dataN"), XML (inline text) or Binary
format.  The graph construction that follows (FaceToEdge, edge features, world edges, noise) is
graphphysics_b200.preprocessing, on the device.
"""
from __future__ import annotations

import json
import os
import xml.etree.ElementTree as ET
from typing import Any, Dict, List, Optional, Tuple

import numpy as np
import torch

from ..graph import Data
from .hdf5 import H5File, write_h5

_XDMF_TO_CELL = {"Triangle": ("triangle", 3), "Tetrahedron": ("tetra", 4), "Quadrilateral": ("quad", 4), "Polyline": ("line", 2),
                 "Polyvertex": ("vertex", 1)}
_CELL_TO_XDMF = {v[0]: (k, v[1]) for k, v in _XDMF_TO_CELL.items()}
_NP_TO_XDMF = {"f": "Float", "i": "Int", "u": "UInt"}


def _local(tag: str) -> str:
    return tag.rsplit("}", 1)[-1]


class TimeSeriesReader:
    """meshio.xdmf.TimeSeriesReader: `read_points_cells()`, `num_steps`, `read_data(k) -> (time, point_data, cell_data)`."""

    def __init__(self, filename: str):
        self.filename = filename
        self._dir = os.path.dirname(os.path.abspath(filename))
        self._h5: Dict[str, H5File] = {}
        root = ET.parse(filename).getroot()
        if _local(root.tag) != "Xdmf":
            raise ValueError(f"{filename}: not an XDMF file")
        domain = next(c for c in root if _local(c.tag) == "Domain")
        grids = [c for c in domain if _local(c.tag) == "Grid"]
        self._mesh_grid, self._steps = None, []
        for g in grids:
            if g.get("GridType") == "Collection" and g.get("CollectionType") == "Temporal":
                self._steps = [c for c in g if _local(c.tag) == "Grid"]
            elif self._mesh_grid is None:
                self._mesh_grid = g
        if self._mesh_grid is None:                              # a collection whose steps carry their own mesh
            self._mesh_grid = self._steps[0]
        if not self._steps:                                      # one uniform grid: a single "time step"
            self._steps = [self._mesh_grid]
        self.num_steps = len(self._steps)

    def __enter__(self):
        return self

    def __exit__(self, *a):
        for f in self._h5.values():
            f.close()
        self._h5 = {}

    # ---------------------------------------------------------------- DataItem
    def _item(self, item: ET.Element) -> np.ndarray:
        dims = [int(d) for d in item.get("Dimensions", "").split()]
        fmt = item.get("Format", "XML")
        kind = {"Float": "f", "Int": "i", "UInt": "u"}[item.get("DataType", item.get("NumberType", "Float"))]
        dtype = np.dtype(f"{kind}{item.get('Precision', '4')}")
        text = (item.text or "").strip()
        if fmt == "HDF":
            fname, path = text.split(":", 1)
            if fname not in self._h5:
                self._h5[fname] = H5File(os.path.join(self._dir, fname))
            a = self._h5[fname][path][()]
            return a.reshape(dims) if dims else a
        if fmt == "XML":
            return np.array(text.split(), dtype=dtype).reshape(dims)
        if fmt == "Binary":
            return np.fromfile(os.path.join(self._dir, text), dtype=dtype).reshape(dims)
        raise NotImplementedError(f"XDMF DataItem format {fmt!r}")

    def read_points_cells(self) -> Tuple[np.ndarray, List[Tuple[str, np.ndarray]]]:
        points, cells = None, []
        for c in self._mesh_grid:
            tag = _local(c.tag)
            if tag == "Geometry":
                points = self._item(next(x for x in c if _local(x.tag) == "DataItem"))
                if c.get("GeometryType", "XYZ") == "XY" and points.shape[1] == 2:
                    points = np.concatenate([points, np.zeros((len(points), 1), points.dtype)], 1)
            elif tag == "Topology":
                tt = c.get("TopologyType") or c.get("Type")
                if tt not in _XDMF_TO_CELL:
                    raise NotImplementedError(f"XDMF topology type {tt!r}")
                cells.append((_XDMF_TO_CELL[tt][0], self._item(next(x for x in c if _local(x.tag) == "DataItem"))))
        if points is None:
            raise ValueError(f"{self.filename}: no Geometry in the mesh grid")
        return points, cells

    def read_data(self, k: int) -> Tuple[float, Dict[str, np.ndarray], Dict[str, np.ndarray]]:
        grid = self._steps[k]
        t, point_data, cell_data = float(k), {}, {}
        for c in grid:
            tag = _local(c.tag)
            if tag == "Time":
                t = float(c.get("Value"))
            elif tag == "Attribute":
                a = self._item(next(x for x in c if _local(x.tag) == "DataItem"))
                (point_data if c.get("Center", "Node") == "Node" else cell_data)[c.get("Name")] = a
        return t, point_data, cell_data


class TimeSeriesWriter:
    """meshio.xdmf.TimeSeriesWriter: `write_points_cells(points, cells)`, `write_data(t, point_data=...)`; the heavy data goes
    to `<name>.h5` next to the .xdmf file as /data0, /data1, ... (written on close)."""

    def __init__(self, filename: str):
        self.filename = filename
        self.h5_filename = os.path.splitext(filename)[0] + ".h5"
        self._data: Dict[str, np.ndarray] = {}
        self._root = ET.Element("Xdmf", Version="3.0")
        self._domain = ET.SubElement(self._root, "Domain")
        self._collection = ET.SubElement(self._domain, "Grid", Name="TimeSeries_meshio", GridType="Collection", CollectionType="Temporal")
        self._has_mesh = False

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def _put(self, parent: ET.Element, arr: np.ndarray) -> None:
        arr = np.asarray(arr)
        key = f"data{len(self._data)}"
        self._data[key] = arr
        item = ET.SubElement(parent, "DataItem", DataType=_NP_TO_XDMF[arr.dtype.kind], Dimensions=" ".join(str(d) for d in arr.shape),
                             Format="HDF", Precision=str(arr.dtype.itemsize))
        item.text = f"{os.path.basename(self.h5_filename)}:/{key}"

    def write_points_cells(self, points: np.ndarray, cells) -> None:
        grid = ET.SubElement(self._domain, "Grid", Name="mesh", GridType="Uniform")
        geo = ET.SubElement(grid, "Geometry", GeometryType="XYZ" if np.asarray(points).shape[1] == 3 else "XY")
        self._put(geo, points)
        for ctype, conn in (cells.items() if isinstance(cells, dict) else cells):
            conn = np.asarray(conn)
            topo = ET.SubElement(grid, "Topology", TopologyType=_CELL_TO_XDMF[ctype][0], NumberOfElements=str(len(conn)))
            self._put(topo, conn)
        self._has_mesh = True

    def write_data(self, t: float, point_data: Optional[Dict[str, np.ndarray]] = None, cell_data: Optional[Dict[str, np.ndarray]] = None) -> None:
        if not self._has_mesh:
            raise RuntimeError("write_points_cells() must come first")
        grid = ET.SubElement(self._collection, "Grid")
        ET.SubElement(grid, "{http://www.w3.org/2003/XInclude}include",
                      xpointer='xpointer(//Grid[@Name="mesh"]/*[self::Topology or self::Geometry])')
        ET.SubElement(grid, "Time", Value=str(t))
        for center, data in (("Node", point_data or {}), ("Cell", cell_data or {})):
            for name, arr in data.items():
                arr = np.asarray(arr)
                kind = "Scalar" if arr.ndim == 1 or arr.shape[1] == 1 else ("Vector" if arr.shape[1] in (2, 3) else "Matrix")
                att = ET.SubElement(grid, "Attribute", Name=name, AttributeType=kind, Center=center)
                self._put(att, arr)

    def close(self) -> None:
        if self._root is None:
            return
        write_h5(self.h5_filename, self._data)
        ET.ElementTree(self._root).write(self.filename)
        self._root = None


def meshes_to_xdmf(filename: str, points: np.ndarray, cells, frames: List[Dict[str, np.ndarray]], timestep: float = 1) -> None:
    """utils/meshio_mesh.py:119-160: a time series over ONE mesh -> `<filename>.xdmf` + `<filename>.h5` (the heavy file is
    written next to the .xdmf directly; the reference has to move it there from the working directory)."""
    with TimeSeriesWriter(f"{filename}.xdmf") as w:
        w.write_points_cells(points, cells)
        t = 0
        for point_data in frames:
            w.write_data(t, point_data=point_data)
            t += timestep


def append_frame_to_xdmf(filename: str, point_data: Dict[str, np.ndarray], timestep: float = 1.0) -> None:
    """utils/meshio_mesh.py:163-233: add one time step to an existing archive (same mesh).  The archive is re-written: this
    writer keeps no free-space bookkeeping to extend an HDF5 file in place."""
    with TimeSeriesReader(f"{filename}.xdmf") as r:
        points, cells = r.read_points_cells()
        steps = [r.read_data(k) for k in range(r.num_steps)]
    with TimeSeriesWriter(f"{filename}.xdmf") as w:
        w.write_points_cells(points, cells)
        for t, pd, _ in steps:
            w.write_data(t, point_data=pd)
        w.write_data(steps[-1][0] + timestep, point_data=point_data)


# ------------------------------------------------------------------------------------------------ mesh -> graph
def meshdata_to_graph(points: np.ndarray, cells: np.ndarray, point_data: Optional[Dict[str, np.ndarray]], time: float = 1,
                      target: Optional[Dict[str, np.ndarray]] = None, id: Optional[str] = None, next_data=None) -> Data:
    """utils/torch_graph.py:137-221, field for field (x = hstack(point data in dict order, time); `face` (3, F) -- for
    tetrahedra the four triangles of every cell -- and `tetra` (4, T); y = hstack(target fields); pos fp32)."""
    n = len(points)
    if point_data is not None:
        if any(np.asarray(d).ndim > 1 for d in point_data.values()):
            x = np.hstack([np.asarray(d) for d in point_data.values()] + [np.full((n,), time).reshape(-1, 1)])
        else:
            x = np.vstack([np.asarray(d) for d in point_data.values()] + [np.full((n,), time)]).T
        x = torch.tensor(x, dtype=torch.float32)
    else:
        x = torch.zeros((n, 1), dtype=torch.float32)
    y = None
    if target is not None and len(target):
        if any(np.asarray(d).ndim > 1 for d in target.values()):
            y = torch.tensor(np.hstack([np.asarray(d) for d in target.values()]), dtype=torch.float32)
        else:
            y = torch.tensor(np.vstack([np.asarray(d) for d in target.values()]).T, dtype=torch.float32)
    c = torch.as_tensor(np.asarray(cells)).T.long()
    tetra, face = None, None
    if c.shape[0] == 4:
        tetra = c
        face = torch.cat([c[0:3], c[1:4], torch.stack([c[2], c[3], c[0]]), torch.stack([c[3], c[0], c[1]])], dim=1)
    elif c.shape[0] == 3:
        face = c
    else:
        raise ValueError("Unsupported cell type. Only 'triangle' and 'tetra' cells are supported.")
    return Data(x=x, face=face, tetra=tetra, y=y, pos=torch.tensor(np.asarray(points), dtype=torch.float32), id=id, next_data=next_data)


def _col(a: np.ndarray) -> np.ndarray:
    return a.reshape(-1, 1) if a.ndim == 1 else a


class XDMFTrajectory:
    """One .xdmf time series as (frame t -> Data with the frame t+1 targets): XDMFDataset.__getitem__ (xdmf_dataset.py:69-163)
    for one file.  `meta` is the dataset's meta JSON (features -> dtype / type), `targets` the target field names."""

    def __init__(self, xdmf_file: str, meta: Dict[str, Any], targets: List[str]):
        self.file, self.meta, self.targets = xdmf_file, meta, list(targets)
        with TimeSeriesReader(xdmf_file) as r:
            self.points, cells = r.read_points_cells()
            self.num_steps = r.num_steps
            self._frames = [r.read_data(k) for k in range(r.num_steps)]
        cd = dict(cells)
        if "triangle" in cd:
            self.cells = cd["triangle"]
        elif "tetra" in cd:
            self.cells = cd["tetra"]
        else:
            raise ValueError("Unsupported cell type. Only 'triangle' and 'tetra' cells are supported.")
        self.mesh_id = os.path.splitext(os.path.basename(xdmf_file))[0].rsplit("_", 1)[-1]

    def __len__(self) -> int:
        return self.num_steps - 1

    def __getitem__(self, frame: int) -> Data:
        if frame >= self.num_steps - 1 or frame < 0:
            raise IndexError(f"Frame index {frame} out of bounds for a trajectory with {self.num_steps} frames.")
        time, pd, _ = self._frames[frame]
        _, nxt, _ = self._frames[frame + 1]
        feats = self.meta["features"]
        point_data = {k: _col(np.asarray(pd[k]).astype(feats[k]["dtype"])) for k in feats if k in pd}
        target = {k: _col(np.asarray(nxt[k]).astype(feats[k]["dtype"])) for k in feats if k in self.targets}
        next_data = {k: np.asarray(nxt[k]).astype(feats[k]["dtype"]) for k in feats
                     if k not in self.targets and k in nxt and feats[k]["type"] == "dynamic"}
        return meshdata_to_graph(self.points.astype(np.float32), self.cells, point_data, time=time, target=target, id=self.mesh_id,
                                 next_data=next_data)


class H5Trajectories:
    """The trajectory file of H5Dataset: `file[trajectory][feature]` arrays cast / reshaped by the meta JSON
    (utils/hierarchical.py:51-86), frames as graphs like get_frame_as_graph (hierarchical.py:89-170)."""

    def __init__(self, h5_path: str, meta_path: str, targets: List[str]):
        self.file = H5File(h5_path)
        self.meta = json.load(open(meta_path))
        self.targets = list(targets)
        self.keys = self.file.keys()
        self._cache: Dict[str, Dict[str, np.ndarray]] = {}

    def __len__(self) -> int:
        return len(self.keys)

    def trajectory(self, key: str) -> Dict[str, np.ndarray]:
        if key not in self._cache:
            grp = self.file[key]
            self._cache[key] = {k: grp[k][()].astype(f["dtype"]).reshape(f["shape"]) for k, f in self.meta["features"].items()}
        return self._cache[key]

    def frame(self, key: str, frame: int) -> Data:
        traj = self.trajectory(key)
        static = ("mesh_pos", "cells", "node_type")
        point_data = {k: traj[k][frame] for k in traj if k not in static}
        point_data["node_type"] = traj["node_type"][0]
        target = {k: traj[k][frame + 1] for k in self.targets}
        next_data = {k: traj[k][frame + 1] for k in traj if k not in static and k not in self.targets}
        mesh_pos = traj["mesh_pos"][frame] if traj["mesh_pos"].shape[0] > 1 else traj["mesh_pos"][0]
        cells = traj["cells"][frame] if traj["cells"].shape[0] > 1 else traj["cells"][0]
        return meshdata_to_graph(mesh_pos, cells, point_data, time=frame * self.meta.get("dt", 1), target=target, next_data=next_data)
