# dump_reference.jl -- export one residual evaluation of the UNMODIFIED reference
# (StableSpectralElements.jl) for array-level parity checks at 1e-12 (SURVEY.md §8c).
#
# WRITTEN, NOT RUN: julia is not available in the build container or on the GPU box.  A
# maintainer runs it once where the reference is installed:
#
#     julia --project=<StableSpectralElements.jl checkout> tools/dump_reference.jl OUTDIR [p] [M]
#
# and copies OUTDIR to tests/golden/reference_dump/; tests/test_reference_dump.py then checks the
# oracle (CPU) and the CUDA path (GPU) against `dudt` on exactly these operators, geometry,
# connectivity and state.  Everything is written as raw little-endian Float64 / Int64 arrays in
# Julia's own (column-major) memory order plus a plain-text manifest "name dtype dims...".
#
# Problem: the north-star configuration (BASELINE.json configs[3]) at a small size: 3-D Euler,
# ModalTensor(p) tetrahedra, ChanWarping(1/16), FluxDifferencingForm with the entropy-conservative
# two-point flux and Lax-Friedrichs interface flux, weight-adjusted mass solver, ReferenceOperator.
using StableSpectralElements, LinearAlgebra

outdir = length(ARGS) >= 1 ? ARGS[1] : "reference_dump"
p = length(ARGS) >= 2 ? parse(Int, ARGS[2]) : 4
M = length(ARGS) >= 3 ? parse(Int, ARGS[3]) : 2
mkpath(outdir)

L = 2π
conservation_law = EulerEquations{3}(1.4)
initial_data = EulerPeriodicTest(conservation_law, 0.2, L)          # euler_navierstokes.jl:289-294
reference_approximation = ReferenceApproximation(ModalTensor(p), Tet(), mapping_degree = min(p, 3))
mesh = warp_mesh(uniform_periodic_mesh(reference_approximation, ((0.0, L), (0.0, L), (0.0, L)),
        (M, M, M)), reference_approximation, ChanWarping(1 / 16, (L, L, L)))
spatial_discretization = SpatialDiscretization(mesh, reference_approximation, ChanWilcoxMetrics())
form = FluxDifferencingForm(inviscid_numerical_flux = LaxFriedrichsNumericalFlux())
ode = semidiscretize(conservation_law, spatial_discretization, initial_data, form, (0.0, 1.0),
    ReferenceOperator())
solver = ode.p
u = copy(ode.u0)
# a rough, still admissible state so that both log-mean branches are exercised (deterministic)
for idx in eachindex(u)
    u[idx] *= 1.0 + 0.05 * (mod(0.6180339887498949 * idx, 1.0) - 0.5)
end
dudt = similar(u)
semi_discrete_residual!(dudt, u, solver, 0.0)

manifest = IOBuffer()
function dump(name, A)
    T = eltype(A) <: Integer ? Int64 : Float64
    B = Array{T}(A)
    open(joinpath(outdir, name * ".bin"), "w") do io
        write(io, B)
    end
    println(manifest, name, " ", T == Int64 ? "i8" : "f8", " ", join(size(B), " "))
end

(; V, R, W, B, D, reference_mapping) = reference_approximation
gf = spatial_discretization.geometric_factors
dump("u", u); dump("dudt", dudt)                                   # (N_p, N_c, N_e)
dump("V", Matrix(V)); dump("R", Matrix(R))                         # dense views of the operators
for (m, Dm) in enumerate(D)
    dump("D$m", Matrix(Dm))
end
dump("W", diag(W)); dump("B", diag(B))
dump("Lambda_ref", reference_mapping.Λ_ref); dump("J_ref", reference_mapping.J_ref)
dump("J_q", gf.J_q); dump("Lambda_q", gf.Λ_q); dump("J_f", gf.J_f); dump("nJf", gf.nJf)
dump("mapP", mesh.mapP)                                            # 1-based linear indices
dump("n_ref", hcat(reference_approximation.reference_element.nrstJ...))  # (N_f, d) scaled reference normals
open(joinpath(outdir, "manifest.txt"), "w") do io
    println(io, "# StableSpectralElements.jl residual dump: Euler 3-D, ModalTensor($p) Tet, M=$M, ",
        "FluxDifferencingForm(EC volume, LF interface), WeightAdjustedSolver, gamma=1.4")
    println(io, "p $p"); println(io, "M $M"); println(io, "gamma 1.4")
    write(io, take!(manifest))
end
println("wrote ", outdir)
