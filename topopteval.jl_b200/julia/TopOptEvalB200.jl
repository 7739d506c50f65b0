# TopOptEvalB200.jl — thin `ccall` shim over libtopopt_b200.so (include/topopt_b200.h).
#
# Drop-in for the bodies of TopOptEval.FiniteElementAnalysis' hot-path entry points (same names, argument order and
# return tuples as src/FiniteElementAnalysis/FiniteElementAnalysis.jl:11-24,75-87 of the reference); MeshImport,
# ResultsExport and Utils stay the reference's own Julia code.  No CUDA.jl, no kernel generation: every numeric step
# is a call into the C ABI.  Julia is not installed in the build image, so this file is written mechanically against
# the header and is exercised there only through its Python twin (topopteval.jl_b200/api.py), which makes the
# identical sequence of C calls.
#
# Usage inside the reference package (see INTEGRATION.md):
#     include("TopOptEvalB200.jl"); using .TopOptEvalB200
#     dh, cv, K, f = TopOptEvalB200.setup_problem(grid)          # K, f are device handles
#     TopOptEvalB200.assemble_stiffness_matrix!(K, f, dh, cv, λ, μ)
#     ch = TopOptEvalB200.apply_fixed_boundary!(K, f, dh, fixed_nodes)
#     TopOptEvalB200.apply_force!(f, dh, collect(load_nodes), [0.0, 0.0, -1.0])
#     u, energy, stress_field, max_vm, max_cell = TopOptEvalB200.solve_system(K, f, dh, cv, λ, μ, ch)
module TopOptEvalB200

using Ferrite
using SparseArrays
using Tensors

export setup_problem, create_material_model, create_simp_material_model,
       assemble_stiffness_matrix!, assemble_stiffness_matrix_simp!,
       apply_fixed_boundary!, apply_sliding_boundary!, apply_force!,
       apply_volume_force!, apply_gravity!, apply_acceleration!, apply_variable_density_volume_force!,
       solve_system, solve_system_simp, solve_system_robust, solve_system_robust_simp, solve_system_adaptive,
       calculate_stresses, calculate_stresses_simp, ferrite_dofhandler, comm_unique_id,
       SolverConfig, element_energies, compliance,
       select_nodes_by_plane, select_nodes_by_circle, get_node_dofs, get_boundary_facets, compute_boundary_area,
       apply_surface_traction!, apply_uniform_surface_traction!

const LIB = get(ENV, "TOPOPT_B200_LIB", joinpath(@__DIR__, "..", "libtopopt_b200.so"))

# ---- handles ---------------------------------------------------------------------------------------------------
mutable struct Ctx
    ptr::Ptr{Cvoid}
    function Ctx(device::Integer = 0)
        out = Ref{Ptr{Cvoid}}(C_NULL)
        st = ccall((:toe_create, LIB), Cint, (Cint, Ref{Ptr{Cvoid}}), device, out)
        st == 0 || error(unsafe_string(ccall((:toe_last_error, LIB), Cstring, (Ptr{Cvoid},), C_NULL)))
        c = new(out[])
        finalizer(x -> (x.ptr == C_NULL || ccall((:toe_destroy, LIB), Cvoid, (Ptr{Cvoid},), x.ptr); x.ptr = C_NULL), c)
        return c
    end
end

check(c::Ctx, st) = st == 0 ? nothing : error(unsafe_string(ccall((:toe_last_error, LIB), Cstring, (Ptr{Cvoid},), c.ptr)))

"DofHandler stand-in: keeps the grid, the ctx and the Ferrite-ordered node → first-dof map."
struct B200DofHandler
    ctx::Ctx
    grid::Grid
    ndofs::Int
    node_first_dof::Vector{Int}      # 1-based, 0 = node in no cell  (what get_node_dofs :265-293 rebuilds per call)
    ferrite::Base.RefValue{Any}      # the genuine Ferrite.DofHandler, built on first request (ferrite_dofhandler)
end
B200DofHandler(ctx, grid, ndofs, nfd) = B200DofHandler(ctx, grid, ndofs, nfd, Ref{Any}(nothing))
Ferrite.ndofs(dh::B200DofHandler) = dh.ndofs

"""
    ferrite_dofhandler(dh::B200DofHandler) -> Ferrite.DofHandler

The reference's own `DofHandler` for the same grid (`setup_problem`, FiniteElementAnalysis.jl:173-176), built lazily and cached.
The GPU path numbers DOFs exactly as Ferrite's `close!` does (first touch in cell order), so `u` and `prescribed_dofs` index it
directly.  Needed only where reference code dispatches on `Ferrite.DofHandler` — `ResultsExport.export_results(u, dh, file)`
(ResultsExport.jl:25) and `export_boundary_conditions` — never on the hot path.
"""
function ferrite_dofhandler(dh::B200DofHandler)
    if dh.ferrite[] === nothing
        hex = typeof(getcells(dh.grid, 1)) <: Ferrite.Hexahedron
        ip = hex ? Lagrange{RefHexahedron,1}()^3 : Lagrange{RefTetrahedron,1}()^3
        fdh = DofHandler(dh.grid)
        add!(fdh, :u, ip)
        close!(fdh)
        ndofs(fdh) == dh.ndofs || error("Ferrite numbers $(ndofs(fdh)) DOFs, the GPU path $(dh.ndofs)")
        dh.ferrite[] = fdh
    end
    return dh.ferrite[]::DofHandler
end

struct B200CellValues; npc::Int; nqp::Int; end
struct B200Matrix; ctx::Ctx; n::Int; nnz::Int; end          # K lives in HBM; materialise with sparse(K)
struct B200Vector; ctx::Ctx; n::Int; end                     # f lives in HBM; materialise with Vector(f)
struct B200Constraint; prescribed_dofs::Vector{Int}; end     # ConstraintHandler stand-in (zero-valued Dirichlet)

Base.size(K::B200Matrix) = (K.n, K.n)
Base.size(K::B200Matrix, i) = K.n
function Base.Vector(f::B200Vector)
    out = Vector{Float64}(undef, f.n)
    check(f.ctx, ccall((:toe_get_rhs, LIB), Cint, (Ptr{Cvoid}, Ptr{Float64}), f.ctx.ptr, out)); out
end
"Copy K out as the SparseMatrixCSC the reference's `allocate_matrix(dh)` + assembly would hold."
function sparse_copy(K::B200Matrix)
    colptr = Vector{Int}(undef, K.n + 1); rowval = Vector{Int}(undef, K.nnz); nzval = Vector{Float64}(undef, K.nnz)
    check(K.ctx, ccall((:toe_get_pattern, LIB), Cint, (Ptr{Cvoid}, Ptr{Int64}, Ptr{Int64}), K.ctx.ptr, colptr, rowval))
    check(K.ctx, ccall((:toe_get_values, LIB), Cint, (Ptr{Cvoid}, Ptr{Float64}), K.ctx.ptr, nzval))
    return SparseArrays.SparseMatrixCSC(K.n, K.n, colptr, rowval, nzval)
end

# ---- material models (FiniteElementAnalysis.jl:103-109, 616-634) -----------------------------------------------------
function create_material_model(E::Float64, ν::Float64)
    λ = E * ν / ((1 + ν) * (1 - 2 * ν)); μ = E / (2 * (1 + ν)); return λ, μ
end
"Callable like the reference's closure, but carrying its parameters so the kernel evaluates E(ρ) itself."
struct SimpModel; E0::Float64; nu::Float64; Emin::Float64; p::Float64; end
function (m::SimpModel)(ρ::Float64)
    E = m.Emin + (m.E0 - m.Emin) * ρ^m.p
    return E * m.nu / ((1 + m.nu) * (1 - 2 * m.nu)), E / (2 * (1 + m.nu))
end
create_simp_material_model(E0::Float64, nu::Float64, Emin::Float64 = 1e-6, p::Float64 = 1.0) = SimpModel(E0, nu, Emin, p)

# ---- setup_problem (:151-185) ------------------------------------------------------------------------------------------
"128-byte NCCL unique id: call on ONE process (rank 0) and hand the bytes to the others (MPI.Bcast!, Distributed.remotecall, a file …)."
function comm_unique_id()
    id = Vector{UInt8}(undef, 128)
    st = ccall((:toe_comm_unique_id, LIB), Cint, (Ptr{UInt8},), id)
    st == 0 || error(unsafe_string(ccall((:toe_last_error, LIB), Cstring, (Ptr{Cvoid},), C_NULL)))
    return id
end

# `comm = (nranks, rank, id)`: one Julia process per GPU, every process passes the SAME global grid; the library partitions it,
# and every later call is collective and takes / returns global data (INTEGRATION.md §5).  Default: a single GPU.
function setup_problem(grid::Grid, interpolation_order::Int = 1; device::Integer = 0, comm = nothing)
    interpolation_order == 1 || error("only linear Lagrange interpolation is on the GPU path")
    cell_type = typeof(getcells(grid, 1))                                   # :157
    npc = cell_type <: Ferrite.Hexahedron ? 8 : 4
    println(npc == 8 ? "Setting up problem with hexahedral elements" : "Setting up problem with tetrahedral elements")
    nn, ne = getnnodes(grid), getncells(grid)
    xyz = Matrix{Float64}(undef, 3, nn)
    for (i, n) in enumerate(grid.nodes); xyz[:, i] .= n.x; end
    conn = Matrix{Int64}(undef, npc, ne)
    for (e, c) in enumerate(grid.cells); conn[:, e] .= c.nodes; end
    ctx = Ctx(device)
    GC.@preserve xyz conn begin
        if comm === nothing
            check(ctx, ccall((:toe_set_mesh, LIB), Cint, (Ptr{Cvoid}, Int64, Ptr{Float64}, Int64, Cint, Ptr{Int64}), ctx.ptr, nn, xyz, ne, npc, conn))
        else
            nranks, rank, id = comm
            length(id) == 128 || error("comm id must be the 128 bytes of comm_unique_id()")
            idb = convert(Vector{UInt8}, id)
            check(ctx, ccall((:toe_comm_init, LIB), Cint, (Ptr{Cvoid}, Cint, Cint, Ptr{UInt8}), ctx.ptr, nranks, rank, idb))
            check(ctx, ccall((:toe_set_mesh_distributed, LIB), Cint, (Ptr{Cvoid}, Int64, Ptr{Float64}, Int64, Cint, Ptr{Int64}), ctx.ptr, nn, xyz, ne, npc, conn))
        end
    end
    nd = Ref{Int64}(0); nnz = Ref{Int64}(0)
    check(ctx, ccall((:toe_build_dofs, LIB), Cint, (Ptr{Cvoid}, Ref{Int64}), ctx.ptr, nd))
    println("Number of DOFs: $(nd[])")
    check(ctx, ccall((:toe_build_pattern, LIB), Cint, (Ptr{Cvoid}, Ref{Int64}), ctx.ptr, nnz))
    nfd = Vector{Int}(undef, nn)
    check(ctx, ccall((:toe_get_node_dofs, LIB), Cint, (Ptr{Cvoid}, Ptr{Int64}), ctx.ptr, nfd))
    dh = B200DofHandler(ctx, grid, nd[], nfd)
    GRID_CTX[grid] = ctx                                          # boundary-node selection on this grid reuses the device mesh
    return dh, B200CellValues(npc, npc == 4 ? 4 : 8), B200Matrix(ctx, nd[], nnz[]), B200Vector(ctx, nd[])
end

# ---- assembly (:204-250, :654-707) -----------------------------------------------------------------------------------------
function assemble_stiffness_matrix!(K::B200Matrix, f, dh::B200DofHandler, cellvalues, λ, μ; variant::Integer = 0)
    check(dh.ctx, ccall((:toe_assemble_lame, LIB), Cint, (Ptr{Cvoid}, Float64, Float64, Cint), dh.ctx.ptr, λ, μ, variant))
    println("Stiffness matrix assembled successfully")
end
function assemble_stiffness_matrix_simp!(K::B200Matrix, f, dh::B200DofHandler, cellvalues, material_model, density_data; variant::Integer = 0)
    ρ = convert(Vector{Float64}, density_data)
    length(ρ) == getncells(dh.grid) || error("density_data has $(length(ρ)) entries, the grid has $(getncells(dh.grid)) cells")
    if material_model isa SimpModel
        m = material_model
        check(dh.ctx, ccall((:toe_assemble_simp, LIB), Cint, (Ptr{Cvoid}, Float64, Float64, Float64, Float64, Ptr{Float64}, Cint),
                            dh.ctx.ptr, m.E0, m.nu, m.Emin, m.p, ρ, variant))
    else                                              # arbitrary closure: evaluated on the host, per cell (:670-674)
        lm = material_model.(ρ); λe = first.(lm); μe = last.(lm)
        check(dh.ctx, ccall((:toe_assemble_lame_per_cell, LIB), Cint, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Cint), dh.ctx.ptr, λe, μe, variant))
    end
    println("Stiffness matrix assembled successfully with variable material properties")
end

# ---- constraints: defined here, applied once inside solve_* (:314-333, :356-374) ------------------------------------------------
function _constraint(dh::B200DofHandler, nodes, comps)
    base = [dh.node_first_dof[n] for n in sort!(unique(collect(nodes))) if dh.node_first_dof[n] > 0]
    return B200Constraint(sort!(unique(vec([b + d - 1 for d in comps, b in base]))))
end
function apply_fixed_boundary!(K, f, dh::B200DofHandler, nodes)
    ch = _constraint(dh, nodes, 1:3); println("Defined fixed boundary conditions for $(length(nodes)) nodes"); ch
end
function apply_sliding_boundary!(K, f, dh::B200DofHandler, nodes, fixed_dofs)
    ch = _constraint(dh, nodes, fixed_dofs)
    println("Defined sliding boundary conditions for $(length(nodes)) nodes, fixing DOFs: $fixed_dofs"); ch
end

# ---- loads (:392-418, VolumeForce.jl) -----------------------------------------------------------------------------------------
function apply_force!(f, dh::B200DofHandler, nodes, force_vector)
    isempty(nodes) && error("No nodes provided for force application.")
    ids = convert(Vector{Int64}, collect(nodes)); F = convert(Vector{Float64}, force_vector)
    check(dh.ctx, ccall((:toe_add_nodal_force, LIB), Cint, (Ptr{Cvoid}, Ptr{Int64}, Int64, Ptr{Float64}), dh.ctx.ptr, ids, length(ids), F))
    println("Applied force $force_vector distributed over $(length(nodes)) nodes")
end
function _volume(dh, b, ρu, density, skip)
    tot = zeros(3); bb = convert(Vector{Float64}, b)
    dptr = density === nothing ? Ptr{Float64}(C_NULL) : pointer(density)
    GC.@preserve density check(dh.ctx, ccall((:toe_add_volume_force, LIB), Cint, (Ptr{Cvoid}, Ptr{Float64}, Float64, Ptr{Float64}, Float64, Ptr{Float64}),
                                             dh.ctx.ptr, bb, ρu, dptr, skip, tot))
    return tot
end
function apply_volume_force!(f, dh::B200DofHandler, cellvalues, body_force_vector, density = 1.0)
    tot = _volume(dh, body_force_vector, Float64(density), nothing, 0.0)
    println("Applied volume force: $body_force_vector N/m³"); println("Total force applied: $tot N")
end
apply_gravity!(f, dh, cv, density = 1.0, g = 9.81, direction = [0.0, 0.0, -1.0]) =
    apply_volume_force!(f, dh, cv, density * g .* (direction ./ sqrt(sum(abs2, direction))), 1.0)
apply_acceleration!(f, dh, cv, a, density = 1.0) = apply_volume_force!(f, dh, cv, density .* a, 1.0)
function apply_variable_density_volume_force!(f, dh::B200DofHandler, cellvalues, body_force_vector, density_data)
    tot = _volume(dh, body_force_vector, 1.0, convert(Vector{Float64}, density_data), 1e-6)       # VolumeForce.jl:199
    println("Applied variable density volume force"); println("Total force applied: $tot N")
end

# ---- solves (:538-598, :831-862; RobustSolver.jl) -----------------------------------------------------------------------------
Base.@kwdef struct SolverConfig
    method::Symbol = :auto
    preconditioner::Symbol = :diagonal
    tolerance::Float64 = 1e-8
    max_iterations::Int = 10000
    memory_limit::Float64 = 0.0
    verbose::Bool = true
    restart::Int = 30
    drop_tolerance::Float64 = 1e-4
    history::Bool = false
    matrix_free::Bool = false        # extension: element-by-element operator
    l2_norm::Bool = false            # extension: stop on ||r||_2 <= tol*(1 + ||r0||_2) (TOE_PCG_L2_NORM) instead of Krylov.jl's M-norm rule
end

struct PcgStats
    niter::Int64; converged::Int32; breakdown::Int32
    res0_M::Float64; res_M::Float64; rel_res_l2::Float64
    solve_seconds::Float64; spmv_seconds::Float64; spmv_bytes::Float64; kernel_launches::Int64; restarts::Int64
    coarse_dofs::Int64; precond_seconds::Float64; true_res::Float64
end

"Lazy stand-in for Dict{Int,Vector{SymmetricTensor}}: nothing leaves the GPU until indexed (`fetch!` fills a 6 × nqp × ne array)."
mutable struct StressField; ctx::Ctx; ne::Int; nqp::Int; sigma::Union{Nothing,Array{Float64,3}}; fetch!::Function; end
_fetch_stored(c::Ctx) = sig -> begin
    mx = Ref(0.0); arg = Ref{Int64}(0)
    check(c, ccall((:toe_stresses, LIB), Cint, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Ref{Float64}, Ref{Int64}), c.ptr, sig, C_NULL, mx, arg))
end
StressField(c::Ctx, ne::Int, nqp::Int, sigma) = StressField(c, ne, nqp, sigma, _fetch_stored(c))
function Base.getindex(s::StressField, cell::Int)
    if s.sigma === nothing
        sig = Array{Float64,3}(undef, 6, s.nqp, s.ne)
        s.fetch!(sig)
        s.sigma = sig
    end
    return [SymmetricTensor{2,3}((v[1], v[4], v[6], v[2], v[5], v[3])) for v in eachcol(@view s.sigma[:, :, cell])]
end
"The reference's container: `Dict{Int64,Vector{SymmetricTensor{2,3}}}` (what `export_results(stress_field, dh, file)`, ResultsExport.jl:55, dispatches on)."
Base.Dict(s::StressField) = Dict{Int64,Vector{SymmetricTensor{2,3,Float64,6}}}(c => s[c] for c in 1:s.ne)
Base.length(s::StressField) = s.ne
Base.haskey(s::StressField, cell::Int) = 1 <= cell <= s.ne
Base.keys(s::StressField) = 1:s.ne

# ---- calculate_stresses(u, dh, cv, λ, μ) (:440-509) and calculate_stresses_simp(u, dh, cv, material_model, density_data) (:730-801) ----
# free functions in the reference: ANY displacement vector, ANY material; the ctx's K, constraints, material and solution stay untouched
function _stress_call(dh::B200DofHandler, u::Vector{Float64}, kind::Symbol, a, b)
    c = dh.ctx
    length(u) == dh.ndofs || error("calculate_stresses: u has $(length(u)) entries, the problem has $(dh.ndofs) DOFs")
    run = sig -> begin
        mx = Ref(0.0); arg = Ref{Int64}(0)
        sp = sig === nothing ? Ptr{Float64}(C_NULL) : pointer(sig)
        GC.@preserve sig u a b begin
            if kind === :lame
                check(c, ccall((:toe_calculate_stresses, LIB), Cint, (Ptr{Cvoid}, Ptr{Float64}, Float64, Float64, Ptr{Float64}, Ptr{Float64}, Ref{Float64}, Ref{Int64}),
                               c.ptr, u, a, b, sp, C_NULL, mx, arg))
            elseif kind === :simp
                check(c, ccall((:toe_calculate_stresses_simp, LIB), Cint, (Ptr{Cvoid}, Ptr{Float64}, Float64, Float64, Float64, Float64, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ref{Float64}, Ref{Int64}),
                               c.ptr, u, a.E0, a.nu, a.Emin, a.p, b, sp, C_NULL, mx, arg))
            else
                check(c, ccall((:toe_calculate_stresses_lame_per_cell, LIB), Cint, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ref{Float64}, Ref{Int64}),
                               c.ptr, u, a, b, sp, C_NULL, mx, arg))
            end
        end
        return mx[], Int(arg[])
    end
    max_vm, max_cell = run(nothing)
    ne = getncells(dh.grid); nqp = length(dh.grid.cells[1].nodes) == 4 ? 4 : 8
    return StressField(c, ne, nqp, nothing, sig -> run(sig)), max_vm, max_cell
end
calculate_stresses(u, dh::B200DofHandler, cellvalues, λ, μ) = _stress_call(dh, convert(Vector{Float64}, copy(u)), :lame, Float64(λ), Float64(μ))
function calculate_stresses_simp(u, dh::B200DofHandler, cellvalues, material_model, density_data)
    ρ = convert(Vector{Float64}, density_data)
    length(ρ) == getncells(dh.grid) || error("density_data has $(length(ρ)) entries, the grid has $(getncells(dh.grid)) cells")
    uu = convert(Vector{Float64}, copy(u))
    material_model isa SimpModel && return _stress_call(dh, uu, :simp, material_model, ρ)
    lm = material_model.(ρ)                                  # arbitrary closure: evaluated on the host, per cell (:744-745)
    return _stress_call(dh, uu, :percell, first.(lm), last.(lm))
end

const DIRECT_EQUIVALENT_TOL = 1e-10      # `K \ f` entry points run PCG to the accuracy the reference's direct solve reaches

# `stress` = closure u -> (stress_field, max_von_mises, max_stress_cell) built from the CALLER's material arguments: the reference hands
# λ, μ (or material_model, density_data) of the solve call to calculate_stresses (FiniteElementAnalysis.jl:553 / :854), and they may
# legally differ from what K was assembled with
function _solve(dh::B200DofHandler, constraints, tol, itmax, matrix_free, verbose, two_level, stress, l2_norm = false)
    c = dh.ctx
    for ch in constraints                                  # SINGLE APPLICATION POINT (:540-542)
        m = Ref(0.0)
        check(c, ccall((:toe_apply_dirichlet, LIB), Cint, (Ptr{Cvoid}, Ptr{Int64}, Int64, Ref{Float64}), c.ptr, ch.prescribed_dofs, length(ch.prescribed_dofs), m))
    end
    verbose && println("Solving linear system...")
    st = Ref{PcgStats}()
    check(c, ccall((:toe_solve_pcg, LIB), Cint, (Ptr{Cvoid}, Float64, Float64, Int64, Cint, Ref{PcgStats}, Ptr{Float64}, Int64),
                   c.ptr, tol, tol, itmax, (matrix_free ? 1 : 0) | (two_level ? 4 : 0) | (l2_norm ? 8 : 0), st, C_NULL, 0))      # TOE_PCG_MATRIX_FREE | TOE_PCG_TWO_LEVEL | TOE_PCG_L2_NORM
    st[].breakdown != 0 && error("CG breakdown: p'Ap <= 0 (matrix not positive definite)")
    st[].converged == 0 && @warn "PCG did not converge in $(st[].niter) iterations (residual $(st[].res_M))"
    u = Vector{Float64}(undef, dh.ndofs)
    check(c, ccall((:toe_get_solution, LIB), Cint, (Ptr{Cvoid}, Ptr{Float64}), c.ptr, u))
    e = Ref(0.0); comp = Ref(0.0)
    check(c, ccall((:toe_energy, LIB), Cint, (Ptr{Cvoid}, Ref{Float64}, Ref{Float64}, Ptr{Float64}), c.ptr, e, comp, C_NULL))
    sf, mx, arg = stress(u)
    if verbose
        println("Analysis complete"); println("Deformation energy: $(e[]) J")
        println("Maximum von Mises stress: $(mx) at cell $(arg)")
    end
    return u, e[], sf, mx, arg
end
_direct_itmax(dh) = max(100000, 4 * dh.ndofs)

solve_system(K, f, dh, cv, λ, μ, constraints...) =
    _solve(dh, constraints, DIRECT_EQUIVALENT_TOL, _direct_itmax(dh), false, true, false, u -> calculate_stresses(u, dh, cv, λ, μ))
solve_system_simp(K, f, dh, cv, material_model, density_data, constraints...) =
    _solve(dh, constraints, DIRECT_EQUIVALENT_TOL, _direct_itmax(dh), false, true, false, u -> calculate_stresses_simp(u, dh, cv, material_model, density_data))
function _robust(dh, constraints, config::SolverConfig, stress)
    config.method in (:auto, :cg, :direct) || error("method :$(config.method) is not on the GPU path (SPD system: :cg only)")
    config.preconditioner in (:diagonal, :two_level) || error("preconditioner :$(config.preconditioner) is not on the GPU path (:diagonal or :two_level)")
    tl = config.preconditioner == :two_level             # Jacobi + rigid-body coarse space (TOE_PCG_TWO_LEVEL)
    # :direct, and :auto below 50 000 DOFs (select_solver_method, RobustSolver.jl:206, picks the factorisation there): no factorisation on
    # the GPU path, so PCG runs to the accuracy a direct solve delivers
    (config.method == :direct || (config.method == :auto && dh.ndofs < 50000)) &&
        return _solve(dh, constraints, DIRECT_EQUIVALENT_TOL, _direct_itmax(dh), config.matrix_free, config.verbose, tl, stress)
    return _solve(dh, constraints, config.tolerance, config.max_iterations, config.matrix_free, config.verbose, tl, stress, config.l2_norm)
end
solve_system_robust(K, f, dh, cv, λ, μ, constraints...; config::SolverConfig = SolverConfig()) =
    _robust(dh, constraints, config, u -> calculate_stresses(u, dh, cv, λ, μ))
solve_system_robust_simp(K, f, dh, cv, material_model, density_data, constraints...; config::SolverConfig = SolverConfig()) =
    _robust(dh, constraints, config, u -> calculate_stresses_simp(u, dh, cv, material_model, density_data))
function solve_system_adaptive(K, f, dh, cv, λ, μ, constraints...)
    n = dh.ndofs
    n < 50000 && return solve_system(K, f, dh, cv, λ, μ, constraints...)                     # :574-575
    cfg = SolverConfig(method = :auto, tolerance = 1e-7, max_iterations = min(max(n ÷ 10, 5000), 50000), history = true)
    return solve_system_robust(K, f, dh, cv, λ, μ, constraints...; config = cfg)
end

# ---- boundary-node selection (SelectNodesForBC.jl) and surface traction (SurfaceTraction.jl) ----------------------------------
# The reference caches the surface nodes per grid (GRID_CACHE_STORAGE, SelectNodesForBC.jl:271-301); here the ctx that holds the
# grid's mesh is remembered per grid object (set by setup_problem).
# Weak keys: the entry (and with it the ctx — mesh, pattern, K in HBM — once its dof handler is gone too) disappears with the grid; a grid
# that is set up again replaces its entry, and the replaced ctx is released by its finalizer (toe_destroy).
const GRID_CTX = WeakKeyDict{Grid,Ctx}()
function _grid_ctx(grid::Grid)
    haskey(GRID_CTX, grid) && return GRID_CTX[grid]
    dh, _, _, _ = setup_problem(grid)
    return dh.ctx
end
function select_nodes_by_plane(grid::Grid, point::Vector{Float64}, normal::Vector{Float64}, tolerance::Float64 = 1.0)     # :325-335
    c = _grid_ctx(grid); n = Ref{Int64}(0)
    check(c, ccall((:toe_select_nodes_by_plane, LIB), Cint, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Float64, Ptr{Int64}, Ref{Int64}), c.ptr, point, normal, tolerance, C_NULL, n))
    out = Vector{Int}(undef, n[])
    n[] > 0 && check(c, ccall((:toe_select_nodes_by_plane, LIB), Cint, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Float64, Ptr{Int64}, Ref{Int64}), c.ptr, point, normal, tolerance, out, n))
    println("Selected $(length(out)) surface nodes on the specified plane")
    return Set{Int}(out)
end
function select_nodes_by_circle(grid::Grid, center::Vector{Float64}, normal::Vector{Float64}, radius::Float64, tolerance::Float64 = 1.0)   # :357-368
    c = _grid_ctx(grid); n = Ref{Int64}(0)
    check(c, ccall((:toe_select_nodes_by_circle, LIB), Cint, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Float64, Float64, Ptr{Int64}, Ref{Int64}), c.ptr, center, normal, radius, tolerance, C_NULL, n))
    out = Vector{Int}(undef, n[])
    n[] > 0 && check(c, ccall((:toe_select_nodes_by_circle, LIB), Cint, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Float64, Float64, Ptr{Int64}, Ref{Int64}), c.ptr, center, normal, radius, tolerance, out, n))
    println("Selected $(length(out)) surface nodes in the circular region")
    return Set{Int}(out)
end
"get_node_dofs(dh) (:265-293): node id → its 3 DOFs, from the device-built map instead of a sweep over all cells."
get_node_dofs(dh::B200DofHandler) = Dict{Int,Vector{Int}}(g => [d, d + 1, d + 2] for (g, d) in enumerate(dh.node_first_dof) if d > 0)
function get_boundary_facets(grid::Grid, nodes::Set{Int})                                                  # SurfaceTraction.jl:45-66
    c = _grid_ctx(grid); v = sort!(collect(nodes)); n = Ref{Int64}(0)
    check(c, ccall((:toe_boundary_facets, LIB), Cint, (Ptr{Cvoid}, Ptr{Int64}, Int64, Ptr{Int64}, Int64, Ref{Int64}), c.ptr, v, length(v), C_NULL, 0, n))
    out = Matrix{Int64}(undef, 2, n[])
    n[] > 0 && check(c, ccall((:toe_boundary_facets, LIB), Cint, (Ptr{Cvoid}, Ptr{Int64}, Int64, Ptr{Int64}, Int64, Ref{Int64}), c.ptr, v, length(v), out, n[], n))
    println("Found $(n[]) boundary facets")
    return Set{Tuple{Int,Int}}((out[1, i], out[2, i]) for i in 1:n[])
end
_facet_matrix(facets) = (v = sort!(collect(facets)); m = Matrix{Int64}(undef, 2, length(v)); for (i, (c, f)) in enumerate(v); m[1, i] = c; m[2, i] = f; end; m)
function compute_boundary_area(grid::Grid, dh::B200DofHandler, boundary_facets)                            # :88-122
    m = _facet_matrix(boundary_facets); a = Ref(0.0)
    check(dh.ctx, ccall((:toe_boundary_area, LIB), Cint, (Ptr{Cvoid}, Ptr{Int64}, Int64, Ref{Float64}), dh.ctx.ptr, m, size(m, 2), a)); a[]
end
function apply_surface_traction!(f, dh::B200DofHandler, grid::Grid, boundary_facets, traction_function::Function)   # :160-225
    m = _facet_matrix(boundary_facets); nf = size(m, 2)
    nqp = typeof(getcells(grid, 1)) <: Ferrite.Hexahedron ? 4 : 3
    xq = Array{Float64}(undef, 3, nqp, nf)
    check(dh.ctx, ccall((:toe_facet_quadrature, LIB), Cint, (Ptr{Cvoid}, Ptr{Int64}, Int64, Ptr{Float64}, Ptr{Float64}), dh.ctx.ptr, m, nf, xq, C_NULL))
    tq = similar(xq)
    for i in 1:nf, q in 1:nqp; tq[:, q, i] .= traction_function(xq[1, q, i], xq[2, q, i], xq[3, q, i]); end    # the callback stays on the host
    area = Ref(0.0); total = zeros(3)
    check(dh.ctx, ccall((:toe_add_surface_traction, LIB), Cint, (Ptr{Cvoid}, Ptr{Int64}, Int64, Ptr{Float64}, Ptr{Float64}, Ref{Float64}, Ptr{Float64}),
                        dh.ctx.ptr, m, nf, tq, C_NULL, area, total))
    println("Applied surface traction over $(nf) facets")
    println("  Total boundary area: $(round(area[], digits=6))")
    println("  Total applied force: [$(round(total[1], digits=6)), $(round(total[2], digits=6)), $(round(total[3], digits=6))]")
end
function apply_uniform_surface_traction!(f, dh::B200DofHandler, grid::Grid, boundary_facets, total_force_vector::Vector{Float64})   # :261-287
    area = compute_boundary_area(grid, dh, boundary_facets)
    area < 1e-12 && error("Boundary area is effectively zero. Check facet selection.")
    traction = total_force_vector ./ area
    println("Uniform surface traction:"); println("  Boundary area: $(round(area, digits=6))"); println("  Traction magnitude: $(round(sqrt(sum(abs2, traction)), digits=6))")
    m = _facet_matrix(boundary_facets); a = Ref(0.0); total = zeros(3)
    check(dh.ctx, ccall((:toe_add_surface_traction, LIB), Cint, (Ptr{Cvoid}, Ptr{Int64}, Int64, Ptr{Float64}, Ptr{Float64}, Ref{Float64}, Ptr{Float64}),
                        dh.ctx.ptr, m, size(m, 2), C_NULL, traction, a, total))
    println("Applied surface traction over $(size(m, 2)) facets")
end

# ---- north-star outputs the reference does not have: per-element energies and compliance ------------------------------------------
function element_energies(dh::B200DofHandler)
    ee = Vector{Float64}(undef, getncells(dh.grid)); e = Ref(0.0); comp = Ref(0.0)
    check(dh.ctx, ccall((:toe_energy, LIB), Cint, (Ptr{Cvoid}, Ref{Float64}, Ref{Float64}, Ptr{Float64}), dh.ctx.ptr, e, comp, ee)); ee
end
function compliance(dh::B200DofHandler)
    e = Ref(0.0); comp = Ref(0.0)
    check(dh.ctx, ccall((:toe_energy, LIB), Cint, (Ptr{Cvoid}, Ref{Float64}, Ref{Float64}, Ptr{Float64}), dh.ctx.ptr, e, comp, C_NULL)); comp[]
end

end # module
