"""Stage the I/O ground-truth fixture (build container only; TEST INFRASTRUCTURE):

    python oracle/make_golden_io.py      ->  tests/golden/mock_xdmf/{mock.h5, mock.xdmf, PROVENANCE.txt}

These two files are DATA written by meshio + h5py (the reference's own test archive, tests/mock_xdmf/): an HDF5 / XDMF
reader can only be pinned against files produced by the real libraries, and neither library is installed here.  They are
copied byte for byte (sha256 recorded); tests/test_io_cpu.py checks that what graphphysics_b200.io reads from them equals
the VTU-derived tests/golden/cylinder_mesh.npz (same mesh and velocity frames, read through a different code path)."""
import hashlib
import os
import shutil

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = "/root/reference/tests/mock_xdmf"
DST = os.path.join(ROOT, "tests", "golden", "mock_xdmf")


def main():
    os.makedirs(DST, exist_ok=True)
    lines = ["Data fixture: the reference's test archive tests/mock_xdmf/ (meshio TimeSeriesWriter + h5py), copied byte for byte by",
             "oracle/make_golden_io.py.  1923 nodes, 3612 triangles, 6 time steps of velocity_x / velocity_y.", ""]
    for name in ("mock.h5", "mock.xdmf"):
        shutil.copyfile(os.path.join(SRC, name), os.path.join(DST, name))
        lines.append(f"{hashlib.sha256(open(os.path.join(DST, name), 'rb').read()).hexdigest()}  {name}")
    open(os.path.join(DST, "PROVENANCE.txt"), "w").write("\n".join(lines) + "\n")
    print("\n".join(lines))


if __name__ == "__main__":
    main()
