"""CPU checks of the drop-in boundary: the shared library builds for sm_100a, loads, and exports every symbol that
include/topopt_b200.h declares; without a GPU the product path fails loudly instead of falling back."""
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    hdr = open(os.path.join(ROOT, "include", "topopt_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    return sorted(set(re.findall(r"\b(toe_[a-z0-9_]+)\s*\(", hdr)))


def test_header_symbols_are_exported_and_bound(pkg):
    import __graft_entry__ as graft
    graft.build()
    lib = pkg._lib.load()
    names = _declared()
    assert len(names) >= 35
    for n in names:
        assert hasattr(lib, n), "libtopopt_b200.so does not export %s" % n
        assert n in pkg._lib.SIGNATURES, "ctypes binding lacks %s" % n
    assert sorted(pkg._lib.SIGNATURES) == names
    assert lib.toe_version() == 100


def test_library_is_sm100a_only():
    so = os.path.join(ROOT, "topopteval.jl_b200", "libtopopt_b200.so")
    out = subprocess.run(["cuobjdump", "-lelf", so], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_(\d+a?)", out))
    assert archs == {"100a"}, archs


def test_no_cpu_fallback(pkg, have_gpu):
    if have_gpu:
        pytest.skip("GPU present")
    with pytest.raises(pkg.TopOptError, match="no CPU fallback|no CUDA device"):
        pkg.Context(0)


def test_product_does_not_import_oracle():
    pdir = os.path.join(ROOT, "topopteval.jl_b200")
    for dp, _, files in os.walk(pdir):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".jl")):
                src = open(os.path.join(dp, f)).read()
                assert "fea_oracle" not in src and "oracle/" not in src and "import oracle" not in src, f


def test_product_library_holds_no_emulation_code():
    """the host-side emulation of the CUDA sources (tests/cuda_emu) is compiled from `#ifdef TOE_EMU` sections that the product
    build never defines: libtopopt_b200.so exports no emu_* symbol, and neither bench.py nor tools/ nor the package can load the
    emulated library (only tests/emu_support.py does)"""
    so = os.path.join(ROOT, "topopteval.jl_b200", "libtopopt_b200.so")
    syms = subprocess.run(["nm", "-D", "--defined-only", so], capture_output=True, text=True).stdout
    assert "emu_" not in syms and " T toe_create" in syms
    for path in [os.path.join(ROOT, "bench.py"), os.path.join(ROOT, "__graft_entry__.py")] + \
                [os.path.join(dp, f) for top in ("tools", "topopteval.jl_b200") for dp, _, fs in os.walk(os.path.join(ROOT, top)) for f in fs
                 if f.endswith((".py", ".sh", ".jl"))]:
        src = open(path).read()
        assert "emu_support" not in src and "libtopopt_emu" not in src, path


def test_header_is_plain_c_and_the_c_example_builds(tmp_path, have_gpu):
    """include/topopt_b200.h is the drop-in boundary for hosts in any language: it must compile as strict C99 and as C++, and
    examples/cantilever.c (the hot path through the C ABI alone) must build against the product library; without a GPU the
    program refuses loudly."""
    hdr = os.path.join(ROOT, "include", "topopt_b200.h")
    src = tmp_path / "hdr.c"
    src.write_text('#include "%s"\nint main(void) { return toe_version() == 0; }\n' % hdr)
    for cmd in (["gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-fsyntax-only", str(src)],
                ["g++", "-std=c++11", "-Wall", "-Wextra", "-Werror", "-fsyntax-only", "-x", "c++", str(src)]):
        r = subprocess.run(cmd, capture_output=True, text=True)
        assert r.returncode == 0, r.stderr
    libdir = os.path.join(ROOT, "topopteval.jl_b200")
    exe = tmp_path / "cantilever"
    r = subprocess.run(["gcc", "-std=c99", "-Wall", "-Wextra", "-Werror", "-O1", os.path.join(ROOT, "examples", "cantilever.c"), "-I" + os.path.join(ROOT, "include"),
                        "-L" + libdir, "-ltopopt_b200", "-Wl,-rpath," + libdir, "-lm", "-o", str(exe)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    if not have_gpu:
        r = subprocess.run([str(exe), "4", "2", "1"], capture_output=True, text=True)
        assert r.returncode == 1 and "no CPU fallback" in r.stderr


def test_struct_layouts_agree_across_c_python_julia(pkg):
    """toe_pcg_stats / toe_timings cross the ABI by pointer: the field lists of the header, the ctypes mirror and the Julia shim
    must be the same names and types in the same order."""
    hdr = open(os.path.join(ROOT, "include", "topopt_b200.h")).read()
    jl = open(os.path.join(ROOT, "topopteval.jl_b200", "julia", "TopOptEvalB200.jl")).read()
    ctype = {"int64_t": "c_long", "int32_t": "c_int", "double": "c_double"}
    jtype = {"int64_t": "Int64", "int32_t": "Int32", "double": "Float64"}

    def c_fields(name):
        body = re.search(r"typedef struct %s \{(.*?)\} %s;" % (name, name), hdr, re.S).group(1)
        body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
        out = []
        for t, names in re.findall(r"\b(int64_t|int32_t|double)\s+([\w\s,]+?)\s*;", body):      # `double a, b, c;` declares three fields
            out += [(t, n.strip()) for n in names.split(",")]
        return out

    for cname, pycls in (("toe_pcg_stats", pkg._lib.PcgStats), ("toe_timings", pkg._lib.Timings)):
        cf = c_fields(cname)
        assert len(cf) >= 9
        assert [n for _, n in cf] == [n for n, _ in pycls._fields_], cname
        assert [ctype[t] for t, _ in cf] == [t.__name__ for _, t in pycls._fields_], cname
    # Julia: struct PcgStats ... end with `name::Type` entries
    jbody = re.search(r"struct PcgStats\n(.*?)\nend", jl, re.S).group(1)
    jf = re.findall(r"(\w+)::(\w+)", jbody)
    cf = c_fields("toe_pcg_stats")
    assert [(n, jtype[t]) for t, n in cf] == jf


def test_julia_ccall_signatures_match_the_header():
    """Julia is not installed here, so the shim cannot be run: every `ccall((:toe_x, LIB), ret, (types…), …)` in TopOptEvalB200.jl is
    checked statically against the prototype of toe_x in include/topopt_b200.h — return type, number of arguments and the class of
    each one (ctx, pointer-to-double / -int64 / -int, scalar int64 / double / int, struct pointer, byte string)."""
    hdr = re.sub(r"/\*.*?\*/", "", open(os.path.join(ROOT, "include", "topopt_b200.h")).read(), flags=re.S)
    jl = open(os.path.join(ROOT, "topopteval.jl_b200", "julia", "TopOptEvalB200.jl")).read()
    protos = {}
    for ret, name, args in re.findall(r"\b(int|void|const char\*)\s+(toe_\w+)\s*\(([^;{]*?)\)\s*;", hdr, re.S):
        protos[name] = (ret, [a.strip() for a in args.replace("\n", " ").split(",")] if args.strip() not in ("", "void") else [])

    def cclass(a):
        a = re.sub(r"\s+", " ", a)
        for pat, cls in ((r"toe_ctx\s*\*\*", "ctx**"), (r"toe_ctx", "ctx*"), (r"toe_pcg_stats", "stats*"), (r"toe_timings", "timings*"),
                         (r"double\s*\*|double \w+\[", "f64*"), (r"int64_t\s*\*", "i64*"), (r"int32_t\s*\*", "i32*"), (r"\bint\s*\*", "int*"),
                         (r"char", "bytes"), (r"\bint64_t\b", "i64"), (r"\bdouble\b", "f64"), (r"\bint\b", "int")):
            if re.search(pat, a):
                return cls
        raise AssertionError("unclassified C argument %r" % a)

    jmap = {"Ptr{Cvoid}": "ctx*", "Ref{Ptr{Cvoid}}": "ctx**", "Ptr{Float64}": "f64*", "Ref{Float64}": "f64*", "Ptr{Int64}": "i64*", "Ref{Int64}": "i64*",
            "Ptr{Int32}": "i32*", "Ref{Cint}": "int*", "Ptr{Cint}": "int*", "Int64": "i64", "Float64": "f64", "Cint": "int", "Ref{PcgStats}": "stats*",
            "Ptr{UInt8}": "bytes", "Cstring": "bytes"}
    rmap = {"int": "Cint", "void": "Cvoid", "const char*": "Cstring"}
    calls = re.findall(r"ccall\(\(:(toe_\w+), LIB\),\s*(\w+),\s*\(([^)]*)\)", jl)
    assert len(calls) >= 40
    for name, ret, types in calls:
        assert name in protos, "the shim calls %s, which the header does not declare" % name
        cret, cargs = protos[name]
        got = [jmap[t.strip()] for t in types.split(",") if t.strip()]
        assert ret == rmap[cret] and got == [cclass(a) for a in cargs], (name, got, [cclass(a) for a in cargs])
    # the path's entry points are all bound
    used = {c[0] for c in calls}
    for need in ("toe_create", "toe_destroy", "toe_set_mesh", "toe_build_dofs", "toe_build_pattern", "toe_assemble_lame", "toe_assemble_simp",
                 "toe_add_nodal_force", "toe_add_volume_force", "toe_apply_dirichlet", "toe_solve_pcg", "toe_get_solution", "toe_energy", "toe_stresses",
                 "toe_calculate_stresses", "toe_calculate_stresses_simp", "toe_comm_init", "toe_set_mesh_distributed"):
        assert need in used, need


def test_shims_cover_the_reference_export_list():
    """every name FiniteElementAnalysis exports (FiniteElementAnalysis.jl:11-24, :70-87 — frozen here, the reference tree does not travel)
    is defined by the Julia shim and by the Python mirror; `get_face_nodes` dispatches on Ferrite cell types and stays the reference's
    own in Julia."""
    exports = ["create_material_model", "setup_problem", "assemble_stiffness_matrix!", "get_node_dofs", "apply_fixed_boundary!", "apply_sliding_boundary!",
               "apply_force!", "solve_system", "calculate_stresses", "create_simp_material_model", "assemble_stiffness_matrix_simp!",
               "calculate_stresses_simp", "solve_system_simp", "solve_system_adaptive", "get_face_nodes", "select_nodes_by_plane", "select_nodes_by_circle",
               "apply_volume_force!", "apply_gravity!", "apply_acceleration!", "apply_variable_density_volume_force!", "solve_system_robust",
               "solve_system_robust_simp", "SolverConfig", "get_boundary_facets", "apply_surface_traction!", "apply_uniform_surface_traction!",
               "compute_boundary_area"]
    jl = open(os.path.join(ROOT, "topopteval.jl_b200", "julia", "TopOptEvalB200.jl")).read()
    import __graft_entry__ as graft
    pkg = graft.load_package()
    for name in exports:
        assert hasattr(pkg, name.rstrip("!")), "Python mirror lacks %s" % name
        if name == "get_face_nodes":
            continue
        pat = r"(?m)^(?:function\s+|Base\.@kwdef\s+struct\s+|struct\s+)?%s(?=[\s(])" % re.escape(name)
        assert re.search(pat, jl), "Julia shim lacks %s" % name
        assert re.search(r"export[^#]*?[\s,]%s(?=[\s,])" % re.escape(name), jl, re.S), "Julia shim does not export %s" % name
    if os.path.exists("/root/reference/src/FiniteElementAnalysis/FiniteElementAnalysis.jl"):      # build container only: the frozen list is complete
        ref = open("/root/reference/src/FiniteElementAnalysis/FiniteElementAnalysis.jl").read()
        found, lines = set(), ref.splitlines()
        for i, ln in enumerate(lines):
            if not ln.startswith("export "):
                continue
            blk, j = ln[len("export "):], i
            while blk.rstrip().endswith(","):                       # the list continues on the next line
                j += 1
                blk += lines[j]
            found |= set(re.findall(r"[A-Za-z_]\w*!?", blk))
        assert found == set(exports), (found - set(exports), set(exports) - found)


def test_julia_ccalls_pass_as_many_arguments_as_they_declare():
    """second static check of the shim: in every ccall the number of values after the type tuple equals the length of the tuple
    (balanced-bracket parse), and brackets balance over the whole file"""
    jl = open(os.path.join(ROOT, "topopteval.jl_b200", "julia", "TopOptEvalB200.jl")).read()

    def split_top(s):
        out, depth, cur = [], 0, ""
        for ch in s:
            depth += ch in "([{"
            depth -= ch in ")]}"
            if ch == "," and depth == 0:
                out.append(cur.strip()); cur = ""
            else:
                cur += ch
        return out + ([cur.strip()] if cur.strip() else [])

    n = 0
    for m in re.finditer(r"ccall\(", jl):
        i = j = m.end()
        depth = 1
        while depth:
            depth += jl[j] in "([{"
            depth -= jl[j] in ")]}"
            j += 1
        parts = split_top(jl[i:j - 1])
        types = parts[2]
        assert types.startswith("(") and types.endswith(")"), parts[0]
        assert len([t for t in split_top(types[1:-1]) if t]) == len(parts) - 3, parts[0]
        n += 1
    assert n >= 40
    code = re.sub(r'"(?:\\.|[^"\\])*"', '""', re.sub(r'"""(.*?)"""', '""', jl, flags=re.S))
    code = re.sub(r"#.*", "", code)
    for a, b in ("()", "[]", "{}"):
        assert code.count(a) == code.count(b), (a, code.count(a), code.count(b))
