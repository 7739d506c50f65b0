"""Generates tests/golden/*.npz from the reference's own fixtures and recipes.

Run in the build container (where /root/reference exists):  python tests/golden/make_golden.py
The GPU box has no /root/reference, so the decoded meshes travel inside the .npz files together with
the oracle's results for the reference's two test recipes (test/runtests.jl:21-49 and :51-89), the
BASELINE config-2 load (variable-density volume force) and a few synthetic cantilevers.

The numbers are outputs of oracle/fea_oracle.py (a restatement — Julia is not installed here), frozen
so that an accidental change of the oracle is caught by tests/test_oracle.py.
"""
import hashlib
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "topopteval.jl_b200"))

import meshgen  # noqa: E402
import vtu  # noqa: E402
from oracle import fea_oracle as fo  # noqa: E402

REF_DATA = "/root/reference/data"


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def run_case(points, cells, *, density=None, simp=None, E=1.0, nu=0.3, load="tip", body=(0.0, 0.0, -1.0)):
    prob = fo.setup_problem(points, cells)
    if density is None:
        lam, mu = fo.create_material_model(E, nu)
        ke = fo.assemble_stiffness_matrix(prob, lam, mu)
    else:
        mm = fo.create_simp_material_model(*simp)
        ke = fo.assemble_stiffness_matrix_simp(prob, mm, density)
        lam, mu = mm(np.asarray(density))
    K_unconstrained = prob.nzval.copy()
    fixed = meshgen.nodes_at_plane(points, 0, 0.0)
    loadn = meshgen.nodes_at_plane(points, 0, points[:, 0].max())
    pres = fo.fixed_boundary_dofs(prob, fixed)
    if load == "tip":
        fo.apply_force(prob, loadn, [0.0, 0.0, -1.0])
    else:
        fo.apply_variable_density_volume_force(prob, body, density)
    f_loaded = prob.f.copy()
    m = fo.apply_dirichlet(prob, pres)
    u = fo.solve_direct(prob)
    energy = fo.deformation_energy(prob, u)
    upcg, st = fo.solve_pcg(prob, 1e-8, 100000)
    ee = fo.element_energies(prob, u, ke)
    _, vm, maxvm, argvm = fo.calculate_stresses(prob, u, lam, mu)
    sample = np.unique(np.linspace(0, prob.ne - 1, 16).astype(np.int64))
    return dict(
        ndofs=prob.ndofs, nnz=prob.nnz, node_first_dof=prob.node_first_dof.astype(np.int64),
        colptr_sha=sha(prob.colptr.astype(np.int64)), rowval_sha=sha(prob.rowval.astype(np.int64)),
        fixed_nodes=fixed, load_nodes=loadn, prescribed=pres, mean_diag=m,
        f_loaded=f_loaded, K_unconstrained_sum=K_unconstrained.sum(), K_unconstrained_abs_sum=np.abs(K_unconstrained).sum(),
        u=u, energy=energy, compliance=float(prob.f @ u), max_abs_u=np.abs(u).max(),
        pcg_niter=st["niter"], pcg_rel_err=np.linalg.norm(upcg - u) / np.linalg.norm(u),
        elem_energy=ee, ke_sample_ids=sample + 1, ke_sample=ke[sample],
        von_mises=vm, max_von_mises=maxvm, max_stress_cell=argvm,
    )


def main():
    out = {}
    # C1: test/runtests.jl:21-49
    m1 = vtu.read_vtu(os.path.join(REF_DATA, "beam_linear_volume_mesh.vtu"))
    r1 = run_case(m1.points, m1.cells)
    np.savez_compressed(os.path.join(HERE, "c1_tet_beam.npz"), points=m1.points, cells=m1.cells.astype(np.int32),
                        cell_type=m1.cell_type, **r1)
    out["c1"] = r1
    # C2: test/runtests.jl:51-89 (tip load) and BASELINE config 2 (variable-density volume force)
    path2 = os.path.join(REF_DATA, "beam_vfrac_04_Raw.vtu")
    m2 = vtu.read_vtu(path2)
    rho = vtu.extract_cell_density(path2)
    r2 = run_case(m2.points, m2.cells, density=rho, simp=(1.0, 0.3, 1e-8, 3.0))
    r2b = run_case(m2.points, m2.cells, density=rho, simp=(1.0, 0.3, 1e-8, 3.0), load="volume")
    np.savez_compressed(os.path.join(HERE, "c2_hex_simp.npz"), points=m2.points, cells=m2.cells.astype(np.int32),
                        cell_type=m2.cell_type, density=rho, **r2,
                        **{"vf_" + k: v for k, v in r2b.items() if k in ("f_loaded", "u", "energy", "compliance", "max_abs_u", "pcg_niter", "elem_energy", "mean_diag")})
    out["c2"] = r2; out["c2_vf"] = r2b
    # synthetic cantilevers (SURVEY §8(c) candidates)
    syn = {}
    for dims in ((12, 4, 2), (24, 8, 4)):
        p, c = meshgen.cantilever(*dims)
        r = run_case(p, c)
        tag = "x".join(map(str, dims))
        for k in ("ndofs", "nnz", "energy", "compliance", "max_abs_u", "pcg_niter", "mean_diag", "colptr_sha", "rowval_sha"):
            syn[tag + "_" + k] = r[k]
        syn[tag + "_u"] = r["u"]
        out["syn" + tag] = r
    np.savez_compressed(os.path.join(HERE, "synthetic_tet.npz"), **syn)
    for k, r in out.items():
        print(k, {kk: r[kk] for kk in ("ndofs", "nnz", "mean_diag", "energy", "compliance", "max_abs_u", "pcg_niter", "pcg_rel_err", "max_von_mises", "max_stress_cell")})
        print("   sum(e_e)/energy-1 =", r["elem_energy"].sum() / r["energy"] - 1)


if __name__ == "__main__":
    main()
