"""CPU-side checks of the product library: it loads, exports every symbol the header declares, refuses
to run without a GPU (no CPU fallback), and its host layer (standard form, phase builders, MPS reader)
agrees with the oracle stage by stage.  No kernel is launched here."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import problems as P
from ellp_b200 import _native as N
from ellp_b200.problem import Bound, ConstraintOp, EllPError, Problem
from ellp_b200.solver import parse_mps
from oracle import binding as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _has_cuda():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "ellp_b200.h")).read()
    declared = sorted(set(re.findall(r"\b(ellp_b200_[a-z0-9_]+)\s*\(", hdr)))
    assert len(declared) >= 30
    lib = C.CDLL(N.LIB_PATH)
    missing = [s for s in declared if not hasattr(lib, s)]
    assert not missing, f"declared in include/ellp_b200.h but not exported: {missing}"
    assert N.lib.ellp_b200_version().decode().startswith("ellp_b200")


def test_no_cpu_fallback_without_a_device():
    if _has_cuda():
        pytest.skip("a CUDA device is present")
    h = C.c_void_p()
    assert N.lib.ellp_b200_create(0, C.byref(h)) == N.E_CUDA
    with pytest.raises(N.NativeError):
        N.Context(0)


def _stage(problem, which):
    arr = problem.to_arrays()
    desc, _keep = N.problem_desc(arr)
    h = C.c_void_p()
    infeasible = C.c_int(0)
    err = C.create_string_buffer(256)
    rc = N.lib.ellp_b200_stage_new(C.byref(desc), which, C.byref(h), C.byref(infeasible), err)
    assert rc == N.OK, err.value
    if infeasible.value:
        return None
    try:
        dims = [C.c_int32() for _ in range(7)]
        N.lib.ellp_b200_stage_dims(h, *[C.byref(v) for v in dims])
        m, n, nx, nB, nN, lc, lb_ = [v.value for v in dims]
        A = np.zeros((m, n), order="F"); c = np.zeros(max(lc, 1)); b = np.zeros(max(m, 1))
        kind = np.zeros(max(lb_, 1), dtype=np.uint8); lb = np.zeros(max(lb_, 1)); ub = np.zeros(max(lb_, 1))
        x = np.zeros(max(nx, 1)); B = np.zeros(max(nB, 1), dtype=np.int32); Nn = np.zeros(max(nN, 1), dtype=np.int32)
        Ns = np.zeros(max(nN, 1), dtype=np.uint8); y = np.zeros(max(m, 1)); d = np.zeros(max(n, 1))
        N.lib.ellp_b200_stage_copy(h, N.ptr(A), N.ptr(c), N.ptr(b), N.ptr(kind), N.ptr(lb), N.ptr(ub), N.ptr(x), N.ptr(B),
                                   N.ptr(Nn), N.ptr(Ns), N.ptr(y) if which == 2 else None, N.ptr(d) if which == 2 else None)
        return dict(m=m, n=n, A=A, c=c[:lc], b=b[:m], kind=kind[:lb_], lb=lb[:lb_], ub=ub[:lb_], x=x[:nx], B=B[:nB],
                    N=Nn[:nN], N_side=Ns[:nN], y=y[:m], d=d[:n])
    finally:
        N.lib.ellp_b200_stage_free(h)


ALL = [(f.__name__, f) for f in P.GOLDEN] + [(n, (lambda n=n: P.netlib(n))) for n in P.NETLIB]


@pytest.mark.parametrize("which", [0, 1, 2], ids=["standard_form", "primal_phase1", "dual_phase1"])
@pytest.mark.parametrize("name,make", ALL, ids=[a for a, _ in ALL])
def test_host_stages_match_oracle(name, make, which):
    prob, _ = make()
    try:
        ref = O.stage(prob, which)
    except O.OracleError:
        pytest.skip("reference panics while building this stage")
    got = _stage(prob, which)
    if ref is None:
        assert got is None
        return
    assert got is not None
    assert (got["m"], got["n"]) == (ref.m, ref.n)
    np.testing.assert_array_equal(got["A"], ref.A)      # same arithmetic, same order => bit-exact
    np.testing.assert_array_equal(got["c"], ref.c)
    np.testing.assert_array_equal(got["b"], ref.b)
    np.testing.assert_array_equal(got["kind"], ref.kind)
    # lb/ub are only meaningful for the kinds that carry them
    for j, k in enumerate(ref.kind):
        if k in (1, 3, 4):
            assert got["lb"][j] == ref.lb[j]
        if k in (2, 3):
            assert got["ub"][j] == ref.ub[j]
    if which:
        np.testing.assert_array_equal(got["B"], ref.B)
        np.testing.assert_array_equal(got["N"], ref.N)
        np.testing.assert_array_equal(got["N_side"], ref.N_side)
        np.testing.assert_allclose(got["x"], ref.x, rtol=0, atol=0)
    if which == 2 and ref.m:
        np.testing.assert_array_equal(got["y"], ref.y)
        np.testing.assert_array_equal(got["d"], ref.d)


@pytest.mark.parametrize("name", P.NETLIB)
def test_mps_reader_round_trips_netlib_fixtures(name):
    text = P.netlib_mps_text(name)
    got = parse_mps(text)
    want, _ = P.netlib(name)
    a, b = got.to_arrays(), want.to_arrays()
    for k in ("nvars", "ncons"):
        assert a[k] == b[k]
    for k in ("obj", "kind", "lb", "ub", "row_ptr", "col_id", "coef", "op", "rhs"):
        np.testing.assert_array_equal(a[k], b[k])


def test_mps_reader_small_example_and_errors():
    # same shape as the reference's unit test (src/parse_mps.rs:565-643): 3 variables, 3 rows, bounds
    text = """NAME TEST
ROWS
 N COST
 L LIM1
 G LIM2
 E MYEQN
COLUMNS
 X COST 1.0
 X LIM1 1.0
 X LIM2 1.0
 Y COST 2.0
 Y LIM1 1.0
 Y MYEQN -1.0
 Z COST -1.0
 Z MYEQN 1.0
RHS
 RHS LIM1 4.0
 RHS LIM2 1.0
 MYEQN 7.0
BOUNDS
 UP BND X 4.0
 LO BND Y -1.0
 UP BND Y 1.0
 FR BND Z
ENDATA
"""
    p = parse_mps(text)
    assert [v.obj_coeff for v in p.variables] == [1.0, 2.0, -1.0]
    assert p.variables[0].bound == Bound.Upper(4.0)
    assert p.variables[1].bound == Bound.TwoSided(-1.0, 1.0)
    assert p.variables[2].bound == Bound.Free()
    assert [(int(c.op), c.rhs) for c in p.constraints] == [(0, 4.0), (2, 1.0), (1, 7.0)]
    assert [(int(v), a) for v, a in p.constraints[2].coeffs] == [(1, -1.0), (2, 1.0)]
    with pytest.raises(EllPError, match="expected 'ROWS'"):
        parse_mps("NAME T\nROWZ\n")
    with pytest.raises(EllPError, match="could not find the row"):
        parse_mps("NAME T\nROWS\n N C\nCOLUMNS\n X NOPE 1.0\nRHS\nENDATA\n")
    with pytest.raises(EllPError, match="should not specify rhs value for the objective"):
        parse_mps("NAME T\nROWS\n N C\n L R\nCOLUMNS\n X R 1.0\nRHS\n B C 1.0\nENDATA\n")


def test_problem_builder_matches_reference_unit_tests():
    # src/problem.rs:372-429
    p = Problem.new()
    x = p.add_var(1.0, Bound.TwoSided(0.0, 1.0), "x")
    assert p.variables[0].name == "x" and p.variables[0].bound == Bound.TwoSided(0.0, 1.0)
    with pytest.raises(EllPError):
        p.add_var(1.0, Bound.TwoSided(1.0, 0.0), "bad")
    with pytest.raises(EllPError):
        p.add_var(1.0, Bound.Lower(float("inf")), "inf")
    with pytest.raises(EllPError):
        p.add_var(1.0, Bound.Free(), "x")  # duplicate name
    with pytest.raises(EllPError):
        from ellp_b200.problem import VariableId
        p.add_constraint([(VariableId(99), 1.0)], ConstraintOp.Eq, 0.0)
    p.add_constraint([(x, 2.0)], ConstraintOp.Lte, 1.0)
    assert p.is_feasible([0.5]) and not p.is_feasible([0.75]) and not p.is_feasible([0.1, 0.2])


# ---------------------------------------------------------------- the header is valid C and the ctypes mirrors match it
def _build_c_abi_smoke():
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = os.path.join(root, "tests", "c_abi", "c_abi_smoke")
    src = os.path.join(root, "tests", "c_abi", "c_abi_smoke.c")
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Wextra", "-Werror", "-pedantic", "-I" + os.path.join(root, "include"), src, "-o", exe,
                           "-L" + os.path.join(root, "ellp_b200"), "-lellp_b200", "-Wl,-rpath," + os.path.join(root, "ellp_b200"), "-lm"])
    return exe


def test_header_is_c99_and_struct_layouts_equal_the_ctypes_mirrors():
    import json
    import subprocess
    from ellp_b200 import _native as N
    exe = _build_c_abi_smoke()
    lay = json.loads(subprocess.run([exe], capture_output=True, text=True, check=True).stdout)
    pairs = {"ellp_std_form": N.StdForm, "ellp_point": N.Point, "ellp_opts": N.Opts, "ellp_result": N.Result, "ellp_batch": N.Batch,
             "ellp_batch_result": N.BatchResult, "ellp_problem_desc": N.ProblemDesc, "ellp_solution": N.Solution}
    checked = 0
    for cname, cls in pairs.items():
        assert lay["sizeof " + cname] == C.sizeof(cls), cname
        names = [f[0] for f in cls._fields_]
        cfields = [k.split(".", 1)[1] for k in lay if k.startswith(cname + ".")]
        assert cfields == names, (cname, cfields, names)   # same fields in the same order
        for f in names:
            off, size = lay[f"{cname}.{f}"]
            d = getattr(cls, f)
            assert (d.offset, d.size) == (off, size), (cname, f)
            checked += 1
    assert lay["sizeof ellp_trace_rec"] == N.TRACE_DTYPE.itemsize
    for f in N.TRACE_DTYPE.names:
        off, size = lay[f"ellp_trace_rec.{f}"]
        assert N.TRACE_DTYPE.fields[f][1] == off and N.TRACE_DTYPE.fields[f][0].itemsize == size
    assert checked > 60


# ---------------------------------------------------------------- Display impls (src/problem.rs:199-223,305-360; standard_form.rs:223-237; solver.rs:14-25)
LIB_RS_EXAMPLE = """minimize
+ 2 x1 + 10 x2 + 1 x4

subject to
+ 2.5 x1 + 3.5 x2 \u2265 5
+ 2.5 x2 + 4.5 x1 \u2264 1
- 1 x3 - 3 x4 - 4 x5 = 2

with the bounds
-1 \u2264 x1 \u2264 1
x2 \u2264 6
x3 \u2265 0
x4 = 0
x5 free
"""
LIB_RS_POINT = """
  \u250c                     \u2510
  \u2502 -0.9473684210526313 \u2502
  \u2502  2.1052631578947367 \u2502
  \u2502                   0 \u2502
  \u2502                   0 \u2502
  \u2502                -0.5 \u2502
  \u2514                     \u2518

"""


def _lib_rs_problem():
    p = Problem.new()
    x1 = p.add_var(2., Bound.TwoSided(-1., 1.), "x1"); x2 = p.add_var(10., Bound.Upper(6.), "x2")
    x3 = p.add_var(0., Bound.Lower(0.), "x3"); x4 = p.add_var(1., Bound.Fixed(0.), "x4"); x5 = p.add_var(0., Bound.Free(), "x5")
    p.add_constraint([(x1, 2.5), (x2, 3.5)], ConstraintOp.Gte, 5.)
    p.add_constraint([(x2, 2.5), (x1, 4.5)], ConstraintOp.Lte, 1.)
    p.add_constraint([(x3, -1.), (x4, -3.), (x5, -4.)], ConstraintOp.Eq, 2.)
    return p


def test_display_of_problem_and_vectors_reproduces_the_output_documented_in_lib_rs():
    """Golden text: the console output the reference documents for its own example (src/lib.rs:62-98).  The Display impl
    (problem.rs:305-351) also prints a '{} variables and {} constraints' header and a blank after every objective term,
    which the doc block drops; both are checked separately."""
    from ellp_b200.standard_form import nalgebra_display
    from ellp_b200.problem import rust_f64
    text = str(_lib_rs_problem())
    assert text.startswith("5 variables and 3 constraints\n\nminimize\n+ 2 x1 + 10 x2 + 1 x4 \n\nsubject to\n")
    body = text.split("\n", 2)[2]
    assert [ln.rstrip() for ln in body.splitlines()] == LIB_RS_EXAMPLE.splitlines()
    assert nalgebra_display([-0.9473684210526313, 2.1052631578947367, 0., 0., -0.5]) == LIB_RS_POINT
    assert nalgebra_display(np.zeros((0, 3))) == "[ ]"
    assert nalgebra_display(np.array([[1.5, -2.], [10., 0.25]])) == "\n  \u250c           \u2510\n  \u2502  1.5   -2 \u2502\n  \u2502   10 0.25 \u2502\n  \u2514           \u2518\n\n"
    # f64 Display of Rust: integers without '.0', never an exponent
    assert [rust_f64(v) for v in (2.0, -0.5, 1e21, 1e-7, float("inf"), 19.157894736842103)] == \
        ["2", "-0.5", "1000000000000000000000", "0.0000001", "inf", "19.157894736842103"]
    # unnamed variables print as id[k] (problem.rs:353-360); Bound Display (:213-223)
    q = Problem.new(); v = q.add_var(1., Bound.Lower(0.))
    assert "id[0] \u2265 0" in str(q)
    assert [str(b) for b in (Bound.Free(), Bound.Lower(1.), Bound.Upper(2.5), Bound.TwoSided(-1., 1.), Bound.Fixed(3.))] == \
        ["(-\u221e, \u221e)", "[1, \u221e)", "(-\u221e, 2.5]", "[-1, 1]", "[3, 3]"]


def test_display_of_standard_form_and_solver_result():
    from ellp_b200.standard_form import StandardForm
    from ellp_b200.solver import Solution, SolverResult
    sf = StandardForm.from_problem(_lib_rs_problem())
    ref = O.stage(_lib_rs_problem(), 0)
    np.testing.assert_array_equal(sf.A, ref.A)
    assert (sf.rows(), sf.cols()) == (ref.m, ref.n)
    text = str(sf)
    assert text.startswith("c:\n  \u250c") and "\n\nA:\n  \u250c" in text and "bounds:\n\nx0: [-1, 1]\nx1: (-\u221e, 6]\nx2: [0, \u221e)\nx3: [0, 0]\nx4: (-\u221e, \u221e)\n" in text
    assert text.count("\n  \u2502") == len(sf.c) + sf.rows() + len(sf.b)
    assert str(SolverResult("Optimal", Solution(19.157894736842103, np.zeros(1)))) == "found optimal point with objective 19.157894736842103"
    assert str(SolverResult("Infeasible")) == "problem is infeasible" and str(SolverResult("Unbounded")) == "problem is unbounded"
    assert str(SolverResult("MaxIter", obj=float("inf"))) == "reached max iterations, current objective = inf"


# ---------------------------------------------------------------- rust_shim/src/solvers/gpu/ffi.rs mirrors the header field by field
def test_rust_ffi_structs_mirror_the_header():
    """No Rust toolchain here: the #[repr(C)] structs of rust_shim/.../ffi.rs are checked STATICALLY against include/ellp_b200.h --
    same structs, same number of fields, same order, compatible types -- and every extern "C" fn it declares is exported by the .so."""
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    hdr = open(os.path.join(root, "include", "ellp_b200.h")).read()
    ffi = open(os.path.join(root, "rust_shim", "src", "solvers", "gpu", "ffi.rs")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    ffi_nc = re.sub(r"//[^\n]*", "", ffi)

    def c_fields(name):
        body = re.search(r"typedef struct \{([^{}]*)\}\s*" + name + r"\s*;", hdr, flags=re.S).group(1)
        out = []
        for decl in body.split(";"):
            decl = " ".join(decl.split())
            if not decl:
                continue
            m_ = re.match(r"(const\s+)?(\w+)\s*(\*?)\s*(.*)$", decl)
            const, ty, star0, rest = m_.groups()
            for nm in rest.split(","):
                nm = nm.strip()
                star = star0 or ("*" if nm.startswith("*") else "")
                out.append((ty, bool(star), bool(const)))
        return out

    def rs_fields(name):
        body = re.search(r"pub struct " + name + r"\s*\{(.*?)\}", ffi_nc, flags=re.S).group(1)
        out = []
        for decl in body.split(","):
            decl = " ".join(decl.split())
            if not decl:
                continue
            ty = decl.split(":", 1)[1].strip()
            ptr = ty.startswith("*")
            const = ty.startswith("*const")
            base = ty.replace("*const", "").replace("*mut", "").strip()
            out.append((base, ptr, const))
        return out

    cmap = {"int32_t": "i32", "uint32_t": "u32", "int64_t": "i64", "uint64_t": "u64", "double": "f64", "uint8_t": "u8", "ellp_trace_rec": "ellp_trace_rec"}
    for name in ("ellp_std_form", "ellp_point", "ellp_trace_rec", "ellp_opts", "ellp_result"):
        cf, rf = c_fields(name), rs_fields(name)
        assert len(cf) == len(rf), (name, cf, rf)
        for (cty, cptr, cconst), (rty, rptr, rconst) in zip(cf, rf):
            assert cmap[cty] == rty and cptr == rptr and (not cptr or cconst == rconst), (name, cty, rty)
    lib = C.CDLL(N.LIB_PATH)
    fns = re.findall(r"pub fn (ellp_b200_\w+)\(", ffi_nc)
    assert len(fns) >= 6
    for fn in fns:
        assert hasattr(lib, fn), fn
        assert re.search(r"\b" + fn + r"\s*\(", hdr), fn


def test_every_tuning_key_of_the_engine_is_documented_in_the_header():
    """ellp_b200_set_tuning accepts string keys; the header comment is their only documentation -- keep the two in step."""
    import re
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    src = open(os.path.join(root, "ellp_b200", "csrc", "engine.cu")).read()
    hdr = open(os.path.join(root, "include", "ellp_b200.h")).read()
    keys = sorted(set(re.findall(r'strcmp\(key, "([a-z_0-9]+)"\)', src)))
    assert len(keys) >= 20
    missing = [k for k in keys if f'"{k}"' not in hdr]
    assert not missing, f"tuning keys missing from include/ellp_b200.h: {missing}"
