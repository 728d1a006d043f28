"""The Rust side of the boundary cannot be compiled in this image (no cargo / rustc), so it is checked structurally:
rust/ffi.rs must declare every function and struct of include/thermite_gpu.h with the same arity, field order and widths,
and rust/build.rs must compile every source the Makefile links and watch every header."""
import ctypes as C
import os
import re
import subprocess
import sys

import thermite_b200.api as api

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
RUST_SIZES = {"u8": 1, "i8": 1, "u16": 2, "i16": 2, "u32": 4, "i32": 4, "f32": 4, "u64": 8, "i64": 8, "f64": 8, "usize": 8,
              "tg_status": 4, "c_int": 4}


def _ffi():
    return open(os.path.join(ROOT, "rust", "ffi.rs")).read()


def _header_functions():
    """name -> number of parameters, parsed independently of tools/gen_rust_ffi.py."""
    hdr = re.sub(r"/\*.*?\*/", "", open(os.path.join(ROOT, "include", "thermite_gpu.h")).read(), flags=re.S)
    out = {}
    for m in re.finditer(r"\b(tg_[a-z0-9_]+)\s*\(([^)]*)\)\s*;", hdr):
        args = m.group(2).strip()
        out[m.group(1)] = 0 if args in ("", "void") else args.count(",") + 1
    return out


def test_ffi_rs_is_generated_from_the_header_and_up_to_date():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "gen_rust_ffi.py"), "--check"], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr


def test_every_abi_function_is_declared_with_the_same_arity():
    want = _header_functions()
    assert set(want) == set(api.ABI_SYMBOLS)
    got = {}
    for m in re.finditer(r"pub fn (tg_\w+)\(([^)]*)\)", _ffi()):
        args = m.group(2).strip()
        got[m.group(1)] = 0 if not args else args.count(",") + 1
    assert got == want, {k: (want.get(k), got.get(k)) for k in set(want) | set(got) if want.get(k) != got.get(k)}


def _rust_struct_layout(name):
    m = re.search(r"pub struct %s \{(.*?)\n\}" % name, _ffi(), flags=re.S)
    assert m, name
    off, align, fields = 0, 1, []
    for f in re.finditer(r"pub (\w+): ([^,]+),", m.group(1)):
        t = f.group(2).strip()
        size = 8 if t.startswith("*") else RUST_SIZES[t]
        off = (off + size - 1) // size * size
        fields.append((f.group(1), off, size))
        off += size
        align = max(align, size)
    return fields, (off + align - 1) // align * align


def test_struct_layouts_match_the_c_abi():
    # numpy dtypes (packed, checked against the C library's records by the GPU tests) and ctypes structures of api.py
    for name, dtype in (("tg_aln", api.ALN_DTYPE), ("tg_aln_c", api.ALN_C_DTYPE), ("tg_seed", api.SEED_DTYPE)):
        fields, size = _rust_struct_layout(name)
        assert size == dtype.itemsize, name
        assert [(n, o) for n, o, _ in fields] == [(n, dtype.fields[n][1]) for n in dtype.names], name
    for name, st in (("tg_opts", api._Opts), ("tg_result", api._Result), ("tg_result_c", api._ResultC),
                     ("tg_seed_result", api._SeedResult), ("tg_read_alns", api._ReadAlns)):
        fields, size = _rust_struct_layout(name)
        assert size == C.sizeof(st), name
        assert [(n, o) for n, o, _ in fields] == [(n, getattr(st, n).offset) for n, _ in st._fields_], name


def test_build_rs_compiles_what_the_makefile_links_and_watches_every_header():
    mk = open(os.path.join(ROOT, "thermite_b200", "csrc", "Makefile")).read()
    rule = re.search(r"^libthermite_gpu\.so:(.*)$", mk, flags=re.M).group(1).split()
    sources = sorted(f for f in rule if f.endswith((".cu", ".cpp")))
    headers = sorted(os.path.basename(f) for f in rule if f.endswith(".h"))
    rs = open(os.path.join(ROOT, "rust", "build.rs")).read()
    objs = sorted(re.findall(r'"([\w]+\.(?:cu|cpp))"', rs))
    assert objs == sources
    watched = sorted(re.findall(r'"([\w]+\.h)"', rs))
    assert watched == headers
    assert "arch=compute_100a,code=sm_100a" in rs
