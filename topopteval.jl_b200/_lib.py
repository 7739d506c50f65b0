"""ctypes binding of libtopopt_b200.so (include/topopt_b200.h).  Product path: no CPU fallback — if the
shared library is missing or no B200 is present, calls fail loudly."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libtopopt_b200.so")
CSRC = os.path.join(HERE, "csrc")

ASM_AUTO, ASM_ATOMIC, ASM_GATHER, ASM_ROWS = 0, 1, 2, 3
PCG_MATRIX_FREE, PCG_NO_GRAPH, PCG_TWO_LEVEL, PCG_L2_NORM = 1, 2, 4, 8


class PcgStats(C.Structure):
    _fields_ = [("niter", C.c_int64), ("converged", C.c_int32), ("breakdown", C.c_int32),
                ("res0_M", C.c_double), ("res_M", C.c_double), ("rel_res_l2", C.c_double),
                ("solve_seconds", C.c_double), ("spmv_seconds", C.c_double), ("spmv_bytes", C.c_double),
                ("kernel_launches", C.c_int64), ("restarts", C.c_int64), ("coarse_dofs", C.c_int64), ("precond_seconds", C.c_double), ("true_res", C.c_double)]

    def asdict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


class Timings(C.Structure):
    _fields_ = [(k, C.c_double) for k in ("set_mesh", "build_dofs", "build_pattern", "assemble", "loads",
                                          "dirichlet", "solve", "energy")] + [("kernel_launches", C.c_int64)]

    def asdict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


class TopOptError(RuntimeError):
    """Raised for any non-zero status — the analogue of the reference's `error(...)` exceptions."""


_P = C.c_void_p
_D = C.POINTER(C.c_double)
_I64 = C.POINTER(C.c_int64)
_I32 = C.POINTER(C.c_int32)

# name -> (restype, argtypes); every symbol include/topopt_b200.h declares
SIGNATURES = {
    "toe_version": (C.c_int, []),
    "toe_create": (C.c_int, [C.c_int, C.POINTER(_P)]),
    "toe_destroy": (None, [_P]),
    "toe_last_error": (C.c_char_p, [_P]),
    "toe_get_timings": (C.c_int, [_P, C.POINTER(Timings)]),
    "toe_debug_stale_cuda_errors": (C.c_int, [_P, _I64, C.POINTER(C.c_char_p)]),
    "toe_timer_start": (C.c_int, [_P]),
    "toe_timer_stop": (C.c_int, [_P, _D]),
    "toe_set_mesh": (C.c_int, [_P, C.c_int64, _D, C.c_int64, C.c_int, _I64]),
    "toe_build_dofs": (C.c_int, [_P, _I64]),
    "toe_get_node_dofs": (C.c_int, [_P, _I64]),
    "toe_get_cell_dofs": (C.c_int, [_P, C.c_int64, C.c_int64, _I64]),
    "toe_build_pattern": (C.c_int, [_P, _I64]),
    "toe_get_pattern": (C.c_int, [_P, _I64, _I64]),
    "toe_assemble_lame": (C.c_int, [_P, C.c_double, C.c_double, C.c_int]),
    "toe_assemble_simp": (C.c_int, [_P, C.c_double, C.c_double, C.c_double, C.c_double, _D, C.c_int]),
    "toe_assemble_lame_per_cell": (C.c_int, [_P, _D, _D, C.c_int]),
    "toe_set_material_lame": (C.c_int, [_P, C.c_double, C.c_double]),
    "toe_set_material_simp": (C.c_int, [_P, C.c_double, C.c_double, C.c_double, C.c_double, _D]),
    "toe_ke_batch": (C.c_int, [_P, C.c_int64, C.c_int64, _D]),
    "toe_get_values": (C.c_int, [_P, _D]),
    "toe_get_diagonal": (C.c_int, [_P, _D]),
    "toe_get_rhs": (C.c_int, [_P, _D]),
    "toe_set_rhs": (C.c_int, [_P, _D]),
    "toe_add_nodal_force": (C.c_int, [_P, _I64, C.c_int64, _D]),
    "toe_add_volume_force": (C.c_int, [_P, _D, C.c_double, _D, C.c_double, _D]),
    "toe_surface_nodes": (C.c_int, [_P, _I64, _I64]),
    "toe_select_nodes_by_plane": (C.c_int, [_P, _D, _D, C.c_double, _I64, _I64]),
    "toe_select_nodes_by_circle": (C.c_int, [_P, _D, _D, C.c_double, C.c_double, _I64, _I64]),
    "toe_boundary_facets": (C.c_int, [_P, _I64, C.c_int64, _I64, C.c_int64, _I64]),
    "toe_boundary_area": (C.c_int, [_P, _I64, C.c_int64, _D]),
    "toe_facet_quadrature": (C.c_int, [_P, _I64, C.c_int64, _D, _D]),
    "toe_add_surface_traction": (C.c_int, [_P, _I64, C.c_int64, _D, _D, _D, _D]),
    "toe_apply_dirichlet": (C.c_int, [_P, _I64, C.c_int64, _D]),
    "toe_solve_pcg": (C.c_int, [_P, C.c_double, C.c_double, C.c_int64, C.c_int, C.POINTER(PcgStats), _D, C.c_int64]),
    "toe_get_solution": (C.c_int, [_P, _D]),
    "toe_set_solution": (C.c_int, [_P, _D]),
    "toe_energy": (C.c_int, [_P, _D, _D, _D]),
    "toe_energy_assembled": (C.c_int, [_P, _D]),
    "toe_stresses": (C.c_int, [_P, _D, _D, _D, _I64]),
    "toe_calculate_stresses": (C.c_int, [_P, _D, C.c_double, C.c_double, _D, _D, _D, _I64]),
    "toe_calculate_stresses_simp": (C.c_int, [_P, _D, C.c_double, C.c_double, C.c_double, C.c_double, _D, _D, _D, _D, _I64]),
    "toe_calculate_stresses_lame_per_cell": (C.c_int, [_P, _D, _D, _D, _D, _D, _D, _I64]),
    "toe_spmv": (C.c_int, [_P, _D, _D, C.c_int]),
    "toe_time_spmv": (C.c_int, [_P, C.c_int, C.c_int, _D, _D]),
    "toe_debug_cg_trace": (C.c_int, [_P, _D, C.c_int64]),
    "toe_spmv_soak": (C.c_int, [_P, C.c_int, C.c_int, C.c_int64, _I64, _I64]),
    "toe_comm_unique_id": (C.c_int, [C.c_char_p]),
    "toe_comm_init": (C.c_int, [_P, C.c_int, C.c_int, C.c_char_p]),
    "toe_set_mesh_distributed": (C.c_int, [_P, C.c_int64, _D, C.c_int64, C.c_int, _I64]),
    "toe_get_partition": (C.c_int, [_P, _I32]),
    "toe_local_sizes": (C.c_int, [_P, _I64, _I64, _I64, _I64]),
    "toe_comm_info": (C.c_int, [_P, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int)]),
}

_lib = None


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile csrc/*.cu for sm_100a into libtopopt_b200.so (in-tree)."""
    cmd = ["make", "-C", CSRC, "-j8"] + (["-B"] if force else [])
    res = subprocess.run(cmd, capture_output=True, text=True)
    if verbose or res.returncode != 0:
        print(res.stdout[-4000:])
        print(res.stderr[-8000:])
    if res.returncode != 0:
        raise RuntimeError("building libtopopt_b200.so failed")
    return LIB_PATH


def load():
    """dlopen the library and declare every prototype. Raises if the .so has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise TopOptError("libtopopt_b200.so not found at %s — run `python -c 'import __graft_entry__ as g; g.build()'` "
                          "(there is no CPU fallback)" % LIB_PATH)
    lib = C.CDLL(LIB_PATH, mode=C.RTLD_GLOBAL)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def _dp(a):
    return None if a is None else a.ctypes.data_as(_D)


def _ip(a):
    return None if a is None else a.ctypes.data_as(_I64)


def f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def i64(a):
    return np.ascontiguousarray(a, dtype=np.int64)


class Context:
    """One toe_ctx (= one GPU).  Thin, explicit wrappers; all arrays are numpy, indices 1-based int64."""

    def __init__(self, device: int = 0):
        self.lib = load()
        h = _P()
        st = self.lib.toe_create(int(device), C.byref(h))
        if st != 0:
            raise TopOptError("toe_create failed: %s" % self.lib.toe_last_error(None).decode())
        self.h = h
        self.device = device
        self.nn = self.ne = self.npc = 0
        self.ndofs = self.nnz = 0

    def close(self):
        if getattr(self, "h", None):
            self.lib.toe_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, st):
        if st != 0:
            raise TopOptError(self.lib.toe_last_error(self.h).decode())

    # -- setup_problem ---------------------------------------------------------------------------------
    def set_mesh(self, points, cells, distributed=False):
        points = f64(points); cells = i64(cells)
        self.nn, self.ne, self.npc = points.shape[0], cells.shape[0], cells.shape[1]
        fn = self.lib.toe_set_mesh_distributed if distributed else self.lib.toe_set_mesh
        self._ck(fn(self.h, self.nn, _dp(points), self.ne, self.npc, _ip(cells)))

    def build_dofs(self):
        n = C.c_int64()
        self._ck(self.lib.toe_build_dofs(self.h, C.byref(n)))
        self.ndofs = n.value
        return self.ndofs

    def build_pattern(self):
        n = C.c_int64()
        self._ck(self.lib.toe_build_pattern(self.h, C.byref(n)))
        self.nnz = n.value
        if not self.ndofs:
            self.build_dofs()
        return self.nnz

    def node_dofs(self):
        out = np.empty(self.nn, dtype=np.int64)
        self._ck(self.lib.toe_get_node_dofs(self.h, _ip(out)))
        return out

    def cell_dofs(self, first=1, count=None):
        count = self.ne - first + 1 if count is None else count
        out = np.empty((count, 3 * self.npc), dtype=np.int64)
        self._ck(self.lib.toe_get_cell_dofs(self.h, first, count, _ip(out)))
        return out

    def pattern(self):
        colptr = np.empty(self.ndofs + 1, dtype=np.int64)
        rowval = np.empty(self.nnz, dtype=np.int64)
        self._ck(self.lib.toe_get_pattern(self.h, _ip(colptr), _ip(rowval)))
        return colptr, rowval

    # -- assembly --------------------------------------------------------------------------------------
    def assemble_lame(self, lam, mu, variant=ASM_AUTO):
        self._ck(self.lib.toe_assemble_lame(self.h, float(lam), float(mu), variant))

    def assemble_simp(self, E0, nu, Emin, p, density, variant=ASM_AUTO):
        density = f64(density)
        if density.shape[0] != self.ne:
            raise TopOptError("density_data has %d entries, the grid has %d cells" % (density.shape[0], self.ne))
        self._ck(self.lib.toe_assemble_simp(self.h, E0, nu, Emin, p, _dp(density), variant))

    def assemble_lame_per_cell(self, lam_e, mu_e, variant=ASM_AUTO):
        lam_e = f64(lam_e); mu_e = f64(mu_e)
        self._ck(self.lib.toe_assemble_lame_per_cell(self.h, _dp(lam_e), _dp(mu_e), variant))

    def set_material_lame(self, lam, mu):
        self._ck(self.lib.toe_set_material_lame(self.h, float(lam), float(mu)))

    def set_material_simp(self, E0, nu, Emin, p, density):
        density = f64(density)
        self._ck(self.lib.toe_set_material_simp(self.h, E0, nu, Emin, p, _dp(density)))

    def ke_batch(self, first, count):
        nb = 3 * self.npc
        out = np.empty((count, nb, nb), dtype=np.float64)
        self._ck(self.lib.toe_ke_batch(self.h, first, count, _dp(out)))
        return out.transpose(0, 2, 1)          # column-major per cell → ke[e, i, j]

    def values(self):
        out = np.empty(self.nnz, dtype=np.float64)
        self._ck(self.lib.toe_get_values(self.h, _dp(out)))
        return out

    def diagonal(self):
        out = np.empty(self.ndofs, dtype=np.float64)
        self._ck(self.lib.toe_get_diagonal(self.h, _dp(out)))
        return out

    # -- loads / constraints -----------------------------------------------------------------------------
    def rhs(self):
        out = np.empty(self.ndofs, dtype=np.float64)
        self._ck(self.lib.toe_get_rhs(self.h, _dp(out)))
        return out

    def set_rhs(self, f):
        f = f64(f)
        self._ck(self.lib.toe_set_rhs(self.h, _dp(f)))

    def add_nodal_force(self, nodes, F):
        nodes = i64(list(nodes) if isinstance(nodes, (set, frozenset)) else nodes)
        F = f64(F)
        self._ck(self.lib.toe_add_nodal_force(self.h, _ip(nodes) if nodes.size else None, nodes.size, _dp(F)))

    def add_volume_force(self, b, rho_uniform=1.0, density=None, skip_below=0.0):
        b = f64(b)
        tot = np.zeros(3)
        d = None if density is None else f64(density)
        self._ck(self.lib.toe_add_volume_force(self.h, _dp(b), float(rho_uniform), _dp(d), float(skip_below), _dp(tot)))
        return tot

    # -- boundary-node selection / surface traction (SelectNodesForBC.jl, SurfaceTraction.jl) ------------------
    def _select(self, fn, *args):
        n = C.c_int64()
        self._ck(fn(self.h, *args, None, C.byref(n)))
        out = np.empty(n.value, dtype=np.int64)
        if n.value:
            self._ck(fn(self.h, *args, _ip(out), C.byref(n)))
        return out

    def surface_nodes(self):
        return self._select(self.lib.toe_surface_nodes)

    def select_nodes_by_plane(self, point, normal, tolerance=1.0):
        point = f64(point); normal = f64(normal)
        return self._select(self.lib.toe_select_nodes_by_plane, _dp(point), _dp(normal), float(tolerance))

    def select_nodes_by_circle(self, center, normal, radius, tolerance=1.0):
        center = f64(center); normal = f64(normal)
        return self._select(self.lib.toe_select_nodes_by_circle, _dp(center), _dp(normal), float(radius), float(tolerance))

    def boundary_facets(self, nodes):
        nodes = i64(sorted(nodes) if isinstance(nodes, (set, frozenset)) else nodes)
        n = C.c_int64()
        self._ck(self.lib.toe_boundary_facets(self.h, _ip(nodes) if nodes.size else None, nodes.size, None, 0, C.byref(n)))
        out = np.empty((n.value, 2), dtype=np.int64)
        if n.value:
            self._ck(self.lib.toe_boundary_facets(self.h, _ip(nodes), nodes.size, _ip(out), n.value, C.byref(n)))
        return out

    @staticmethod
    def _facets(facets):
        if isinstance(facets, (set, frozenset)):
            facets = sorted(facets)
        return i64(np.asarray(facets, dtype=np.int64).reshape(-1, 2))

    def boundary_area(self, facets):
        facets = self._facets(facets)
        a = C.c_double()
        self._ck(self.lib.toe_boundary_area(self.h, _ip(facets) if facets.size else None, facets.shape[0], C.byref(a)))
        return a.value

    def facet_quadrature(self, facets):
        facets = self._facets(facets)
        nqp = 3 if self.npc == 4 else 4
        xq = np.empty((facets.shape[0], nqp, 3)); dg = np.empty((facets.shape[0], nqp))
        self._ck(self.lib.toe_facet_quadrature(self.h, _ip(facets) if facets.size else None, facets.shape[0], _dp(xq), _dp(dg)))
        return xq, dg

    def add_surface_traction(self, facets, traction_qp=None, traction_uniform=None):
        facets = self._facets(facets)
        tq = None if traction_qp is None else f64(traction_qp)
        tu = None if traction_uniform is None else f64(traction_uniform)
        if tq is not None and tq.size != 3 * (3 if self.npc == 4 else 4) * facets.shape[0]:
            raise TopOptError("traction_qp must hold 3 values per quadrature point of every facet")
        a = C.c_double(); tot = np.zeros(3)
        self._ck(self.lib.toe_add_surface_traction(self.h, _ip(facets) if facets.size else None, facets.shape[0], _dp(tq), _dp(tu), C.byref(a), _dp(tot)))
        return a.value, tot

    def apply_dirichlet(self, dofs):
        dofs = i64(dofs)
        m = C.c_double()
        self._ck(self.lib.toe_apply_dirichlet(self.h, _ip(dofs) if dofs.size else None, dofs.size, C.byref(m)))
        return m.value

    # -- solve / post -------------------------------------------------------------------------------------
    def solve_pcg(self, atol=1e-8, rtol=1e-8, itmax=10000, matrix_free=False, graph=True, history=False, two_level=False, l2_norm=False):
        st = PcgStats()
        flags = ((PCG_MATRIX_FREE if matrix_free else 0) | (0 if graph else PCG_NO_GRAPH) | (PCG_TWO_LEVEL if two_level else 0)
                 | (PCG_L2_NORM if l2_norm else 0))
        hist = np.zeros(min(itmax + 1, 1 << 20)) if history else None
        self._ck(self.lib.toe_solve_pcg(self.h, atol, rtol, itmax, flags, C.byref(st), _dp(hist), 0 if hist is None else hist.size))
        out = st.asdict()
        if history:
            out["residuals"] = hist[: min(st.niter + 1, hist.size)]
        return out

    def solution(self):
        out = np.empty(self.ndofs, dtype=np.float64)
        self._ck(self.lib.toe_get_solution(self.h, _dp(out)))
        return out

    def set_solution(self, u):
        u = f64(u)
        self._ck(self.lib.toe_set_solution(self.h, _dp(u)))

    def energy(self, per_element=False):
        e = C.c_double(); c = C.c_double()
        pe = np.empty(self.ne) if per_element else None
        self._ck(self.lib.toe_energy(self.h, C.byref(e), C.byref(c), _dp(pe)))
        return e.value, c.value, pe

    def energy_assembled(self):
        e = C.c_double()
        self._ck(self.lib.toe_energy_assembled(self.h, C.byref(e)))
        return e.value

    def stresses(self, want_sigma=False, want_vm=False):
        nq = 4 if self.npc == 4 else 8
        sig = np.empty((self.ne, nq, 6)) if want_sigma else None
        vm = np.empty(self.ne) if want_vm else None
        mx = C.c_double(); arg = C.c_int64()
        self._ck(self.lib.toe_stresses(self.h, _dp(sig), _dp(vm), C.byref(mx), C.byref(arg)))
        return sig, vm, mx.value, arg.value

    def calculate_stresses(self, u=None, lame=None, simp=None, lame_per_cell=None, want_sigma=False, want_vm=False):
        """calculate_stresses(u, …) / calculate_stresses_simp(u, …) for ANY u (None = the stored solution) and material; exactly one
        of lame=(λ, μ), simp=(E0, ν, Emin, p, density), lame_per_cell=(λₑ, μₑ).  Leaves K, constraints, material, solution alone."""
        nq = 4 if self.npc == 4 else 8
        sig = np.empty((self.ne, nq, 6)) if want_sigma else None
        vm = np.empty(self.ne) if want_vm else None
        mx = C.c_double(); arg = C.c_int64()
        u = None if u is None else f64(u)
        if u is not None and u.shape != (self.ndofs,):
            raise TopOptError("calculate_stresses: u has %d entries, the problem has %d DOFs" % (u.size, self.ndofs))
        out = (_dp(sig), _dp(vm), C.byref(mx), C.byref(arg))
        if lame is not None:
            self._ck(self.lib.toe_calculate_stresses(self.h, _dp(u), float(lame[0]), float(lame[1]), *out))
        elif simp is not None:
            rho = f64(simp[4])
            if rho.shape != (self.ne,):
                raise TopOptError("calculate_stresses_simp: density_data has %d entries, the mesh has %d cells" % (rho.size, self.ne))
            self._ck(self.lib.toe_calculate_stresses_simp(self.h, _dp(u), float(simp[0]), float(simp[1]), float(simp[2]), float(simp[3]), _dp(rho), *out))
        elif lame_per_cell is not None:
            le, me = f64(lame_per_cell[0]), f64(lame_per_cell[1])
            if le.shape != (self.ne,) or me.shape != (self.ne,):
                raise TopOptError("calculate_stresses: per-cell material arrays must have one entry per cell")
            self._ck(self.lib.toe_calculate_stresses_lame_per_cell(self.h, _dp(u), _dp(le), _dp(me), *out))
        else:
            raise TopOptError("calculate_stresses: no material given")
        return sig, vm, mx.value, arg.value

    def spmv(self, x, matrix_free=False):
        x = f64(x); y = np.empty_like(x)
        self._ck(self.lib.toe_spmv(self.h, _dp(x), _dp(y), 1 if matrix_free else 0))
        return y

    def time_spmv(self, matrix_free=False, reps=20):
        s = C.c_double(); b = C.c_double()
        self._ck(self.lib.toe_time_spmv(self.h, 1 if matrix_free else 0, reps, C.byref(s), C.byref(b)))
        return s.value, b.value

    def cg_trace(self, iterations):
        out = np.zeros((int(iterations), 4))
        self._ck(self.lib.toe_debug_cg_trace(self.h, _dp(out), int(iterations)))
        return out

    def spmv_soak(self, reps, what=1, matrix_free=False):
        """(batches with a mismatch, mismatching entries) of `reps` repeated operator applications against the first one;
        what: 1 = local product, 2 = interface exchange, 3 = both"""
        a, b = C.c_int64(), C.c_int64()
        self._ck(self.lib.toe_spmv_soak(self.h, 1 if matrix_free else 0, int(what), int(reps), C.byref(a), C.byref(b)))
        return a.value, b.value

    def timer_start(self):
        self._ck(self.lib.toe_timer_start(self.h))

    def timer_stop(self):
        s = C.c_double()
        self._ck(self.lib.toe_timer_stop(self.h, C.byref(s)))
        return s.value

    def stale_cuda_errors(self):
        n = C.c_int64(); msg = C.c_char_p()
        self._ck(self.lib.toe_debug_stale_cuda_errors(self.h, C.byref(n), C.byref(msg)))
        return n.value, (msg.value or b"").decode()

    def timings(self):
        t = Timings()
        self._ck(self.lib.toe_get_timings(self.h, C.byref(t)))
        return t.asdict()

    # -- multi-GPU ------------------------------------------------------------------------------------------
    @staticmethod
    def comm_unique_id() -> bytes:
        lib = load()
        buf = C.create_string_buffer(128)
        if lib.toe_comm_unique_id(buf) != 0:
            raise TopOptError(lib.toe_last_error(None).decode())
        return buf.raw

    def comm_init(self, nranks, rank, uid: bytes):
        self._ck(self.lib.toe_comm_init(self.h, nranks, rank, uid))

    def partition(self):
        out = np.empty(self.ne, dtype=np.int32)
        self._ck(self.lib.toe_get_partition(self.h, out.ctypes.data_as(_I32)))
        return out

    def comm_info(self):
        a, b, c = C.c_int(), C.c_int(), C.c_int()
        self._ck(self.lib.toe_comm_info(self.h, C.byref(a), C.byref(b), C.byref(c)))
        return {"nranks": a.value, "rank": b.value, "transport": {0: "single-gpu", 1: "nccl", 2: "peer-memory", 3: "nccl-allgather"}[c.value]}

    def local_sizes(self):
        a, b, c, d = C.c_int64(), C.c_int64(), C.c_int64(), C.c_int64()
        self._ck(self.lib.toe_local_sizes(self.h, C.byref(a), C.byref(b), C.byref(c), C.byref(d)))
        return {"ne_local": a.value, "ndofs_local": b.value, "nnz_local": c.value, "n_interface_dofs": d.value}
