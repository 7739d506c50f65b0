# baseline/run_reference.jl — times the genuine reference (jezekon/TopOptEval.jl) on its own CPU path.
# Needs Julia + the packages of the reference's Project.toml; NOT runnable in the build image (no Julia there), which is
# why bench.py's CPU arm is the C restatement oracle/oracle.c (kind "port").  Recipe = test/runtests.jl:21-49 on a
# synthetic cantilever of the same family bench.py uses.
using TopOptEval, Ferrite, LinearAlgebra

function cantilever(nx, ny, nz)
    generate_grid(Tetrahedron, (nx, ny, nz), Vec((0.0, 0.0, 0.0)), Vec((60.0, 20.0, 4.0)))
end

function nodes_at_plane(grid, axis, value; tol = 1e-6)
    Set(n for n in 1:getnnodes(grid) if abs(grid.nodes[n].x[axis] - value) < tol)
end

function run(nx = 60, ny = 20, nz = 8)
    grid = cantilever(nx, ny, nz)
    λ, μ = create_material_model(1.0, 0.3)
    t_setup = @elapsed ((dh, cv, K, f) = setup_problem(grid))
    t_asm = @elapsed assemble_stiffness_matrix!(K, f, dh, cv, λ, μ)
    ch = apply_fixed_boundary!(K, f, dh, nodes_at_plane(grid, 1, 0.0))
    apply_force!(f, dh, collect(nodes_at_plane(grid, 1, 60.0)), [0.0, 0.0, -1.0])
    cfg = SolverConfig(method = :cg, preconditioner = :diagonal, tolerance = 1e-8, max_iterations = 100000, verbose = false)
    t_solve = @elapsed ((u, energy, _, _, _) = solve_system_robust(K, f, dh, cv, λ, μ, ch; config = cfg))
    ne = getncells(grid)
    println("cells $ne  setup $(t_setup)s  assemble $(t_asm)s ($(ne / t_asm) el/s)  solve+stress $(t_solve)s  energy $energy")
    println("elements/s through the full path: ", ne / (t_setup + t_asm + t_solve), "  threads: ", Threads.nthreads())
end

run()
