"""CPU: the drop-in boundary.  include/gp_b200.h is the contract; the ctypes structs of
graphphysics_b200/_lib.py, the struct printed in INTEGRATION.md and the sizes compiled into
libgp_b200.so must all agree with it (field names, order, kinds, sizeof)."""
import ctypes as C
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def parse_header_structs():
    """{struct name: [(field name, kind, array length)]}; kind in {'ptr', 'i32', 'i64', 'f32', 'u8'}."""
    text = open(os.path.join(ROOT, "include", "gp_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    out = {}
    for m in re.finditer(r"typedef\s+struct\s+(\w+)\s*\{(.*?)\}\s*(\w+)\s*;", text, flags=re.S):
        assert m.group(1) == m.group(3)
        fields = []
        for decl in m.group(2).split(";"):
            decl = " ".join(decl.split())
            if not decl:
                continue
            first, *rest = [d.strip() for d in decl.split(",")]
            tm = re.match(r"(.*?)(\**)\s*(\w+)(\[(\d+)\])?$", first)
            base, stars = tm.group(1).strip(), tm.group(2)
            names = [(stars, tm.group(3), tm.group(5))]
            for r in rest:
                rm = re.match(r"(\**)\s*(\w+)(\[(\d+)\])?$", r)
                names.append((rm.group(1), rm.group(2), rm.group(4)))
            for st, name, dim in names:
                if st:
                    kind = "ptr"
                else:
                    b = base.replace("const ", "").strip()
                    kind = {"int32_t": "i32", "int64_t": "i64", "float": "f32", "uint8_t": "u8", "int": "i32"}[b]
                fields.append((name, kind, int(dim) if dim else 0))
        out[m.group(1)] = fields
    return out


def ctypes_fields(cls):
    res = []
    for name, ty in cls._fields_:
        dim = 0
        if hasattr(ty, "_length_"):
            dim, ty = ty._length_, ty._type_
        kind = {C.c_void_p: "ptr", C.c_int32: "i32", C.c_int64: "i64", C.c_float: "f32", C.c_uint8: "u8"}[ty]
        res.append((name, kind, dim))
    return res


def test_ctypes_structs_match_header_and_library():
    from graphphysics_b200 import _lib
    structs = parse_header_structs()
    table = {"gp_mlp_fwd_args": _lib.MlpFwdArgs, "gp_mlp_bwd_args": _lib.MlpBwdArgs, "gp_linear_bwd_args": _lib.LinearBwdArgs,
             "gp_pack_entry": _lib.PackEntry, "gp_reduce_seg": _lib.ReduceSeg, "gp_attention_args": _lib.AttentionArgs}
    table.update(getattr(_lib, "EXTRA_STRUCTS", {}))
    assert set(structs) == set(table), (sorted(structs), sorted(table))
    lib = _lib.lib()
    lib.gp_sizeof_struct.restype = C.c_int
    for name, cls in table.items():
        assert ctypes_fields(cls) == structs[name], f"{name}: ctypes layout differs from include/gp_b200.h"
        assert C.sizeof(cls) == lib.gp_sizeof_struct(name.encode()), name
    assert lib.gp_sizeof_struct(b"no_such_struct") == -1


def test_integration_md_struct_is_the_header_struct():
    """The reference-side stub printed in INTEGRATION.md must list gp_mlp_fwd_args field by field."""
    md = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    m = re.search(r"class MlpFwdArgs\(C\.Structure\):.*?_fields_ = \[(.*?)\]\n", md, flags=re.S)
    assert m, "INTEGRATION.md lost its ctypes stub"
    doc = re.findall(r'\("(\w+)",\s*C\.(\w+)(?:\s*\*\s*(\d+))?\)', m.group(1))
    kinds = {"c_void_p": "ptr", "c_int32": "i32", "c_int64": "i64", "c_float": "f32"}
    doc = [(n, kinds[t], int(d) if d else 0) for n, t, d in doc]
    assert doc == parse_header_structs()["gp_mlp_fwd_args"]
