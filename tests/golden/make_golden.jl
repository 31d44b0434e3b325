# make_golden.jl — emits the OUTPUT keys of tests/golden/hotpath_golden.npz from the real TensorTrainNumerics.jl package.
#
# STATUS: never executed in the build image (no `julia` binary, no registry access).  It exists so that a maintainer with a
# Julia installation can pin parity against raw reference outputs:
#
#     julia --project=/path/to/TensorTrainNumerics.jl tests/golden/make_golden.jl
#         (needs NPZ.jl in the environment:  ] add NPZ)
#
# It READS the seeded inputs from hotpath_golden.npz (so both sides consume bit-identical arrays; the Julia RNG stream is
# never involved) and WRITES tests/golden/hotpath_golden_julia.npz with the same output keys.  tests/_golden.py prefers
# the Julia-made file when it is present and reports `provenance = "julia"`; until then every parity statement of this
# repository is "pinned on the NumPy restatement + the reference's known-answer tests", i.e. unpinned for raw outputs.
#
# Key by key (reference entry point in brackets):
#   cfg1_x        [als_linsolve, src/solvers/als.jl:161]            dense solution of README.md:82-102 from cfg1_x0_core*
#   cmp_out, cmp_out_rks, cmp_sigma
#                 [tt_compress!, src/tt_tools.jl:772; _svdtrunc]    dense result, ranks, per-bond singular values
#   orth_dense    [orthogonalize(x; i=4), src/tt_tools.jl:511]
#   apply_dense   [Δ(8) * x, src/tt_operations.jl:101]
#   heis_e0       [heisenberg_xyz_tto + eigvals of the dense matrix, examples/heisenberg_xyz_dmrg.jl:9-19]
#   svd_s         [_svdtrunc(svd_A), src/tt_cross_interpolation.jl:149]
#   had_out, had_out_rks   [hadamard_ttm(x, y; tol=1e-12), src/tt_operations.jl:399]
#   qtt_out, qtt_out_rks   [to_qtt, src/qtt_tools.jl:254]
#   gen_E         [als_gen_eigsolv, src/solvers/als.jl:344]
#   dft_spec      [fourier_qtto * x, tt_compress!, matricize — examples/dft.jl:5-25]
# Not emitted (no public entry point in the reference): mv_Y (K_matfree is a closure inside K_eigmin, dmrg.jl:239-244),
# reo_out (reorder works on QTTvector metadata; the test uses the generic swap list) — these keys stay oracle-made.

using TensorTrainNumerics
using LinearAlgebra
using NPZ

const HERE = @__DIR__
g = npzread(joinpath(HERE, "hotpath_golden.npz"))
out = Dict{String, Any}()

"TTvector from the `prefix_core<k>` / `prefix_rks` entries (cores are stored (n, r_l, r_r), 0-based k)"
function load_tt(prefix; dims = nothing)
    rks = Int.(vec(g[prefix * "_rks"]))
    d = length(rks) - 1
    cores = [Array{Float64, 3}(g["$(prefix)_core$(k - 1)"]) for k in 1:d]
    nd = dims === nothing ? Tuple(size(c, 1) for c in cores) : Tuple(Int.(vec(dims)))
    return TTvector{Float64, d}(d, cores, nd, rks, zeros(Int64, d))
end

# NumPy flattens in C order: tensor[s_1, ..., s_d] with s_d fastest
dense(x) = vec(permutedims(ttv_to_tensor(x), reverse(1:x.N)))

# 1. cfg1
let d = 6
    A = id_tto(d)
    b = qtt_sin(d; λ = π)
    x0 = load_tt("cfg1_x0")
    x = als_linsolve(A, b, x0; sweep_count = 4)
    out["cfg1_x"] = dense(x)
    out["cfg1_b"] = dense(b)
end

# 2. tt_compress! with the per-bond singular values (same order as the sweep: L->R then R->L)
let y = load_tt("cmp_in")
    sig = Vector{Vector{Float64}}()
    ψ = copy(y)
    N = ψ.N
    for k in vcat(1:(N - 1), (N - 1):-1:1)
        A = permutedims(ψ.ttv_vec[k], (2, 1, 3)); B = permutedims(ψ.ttv_vec[k + 1], (2, 1, 3))
        Θ = reshape(reshape(A, :, size(A, 3)) * reshape(B, size(B, 1), :), size(A, 1) * size(A, 2), :)
        s = svdvals(Θ)
        push!(sig, s[1:min(5, length(s))])
        TensorTrainNumerics._tt_bond_truncate!(ψ, k; max_bond = 5, truncerr = 0.0)
    end
    z = tt_compress!(copy(y), 5)
    out["cmp_out"] = dense(z)
    out["cmp_out_rks"] = Int64.(z.ttv_rks)
    smax = maximum(length.(sig))
    out["cmp_sigma"] = permutedims(hcat([vcat(s, zeros(smax - length(s))) for s in sig]...))
    # 3. orthogonalize, 4. apply
    out["orth_dense"] = dense(orthogonalize(y; i = 4))
    out["apply_dense"] = dense(Δ(8) * y)
end

# 6. Heisenberg ground-state energy at d = 10
let d = 10
    H = heisenberg_xyz_tto(d; jx = 1.1, jy = 0.8, jz = 1.2, λ = 0.0)
    M = reshape(permutedims(tto_to_tensor(H), vcat(reverse(1:d), reverse((d + 1):(2d)))), 2^d, 2^d)
    out["heis_e0"] = minimum(eigvals(Hermitian(real.(M))))
end

# 7. _svdtrunc on the prescribed spectrum
let A = Matrix{Float64}(g["svd_A"])
    U, S, Vt = TensorTrainNumerics._svdtrunc(A; max_bond = 10, truncerr = 0.0)
    out["svd_s"] = diag(S)
end

# 8. hadamard_ttm
let hx = load_tt("had_x"; dims = g["had_dims"]), hy = load_tt("had_y"; dims = g["had_dims"])
    hz = hadamard_ttm(hx, hy; tol = 1.0e-12)
    out["had_out"] = dense(hz)
    out["had_out_rks"] = Int64.(hz.ttv_rks)
end

# 10. to_qtt
let qx = load_tt("qtt_x"; dims = g["qtt_dims"])
    qq = to_qtt(qx, [[2, 2, 2], [4], [3, 2]])
    out["qtt_out"] = dense(qq)
    out["qtt_out_rks"] = Int64.(qq.ttv_rks)
end

# 11. als_gen_eigsolv
let d = 5
    Ag = Δ(d) + 2.0 * id_tto(d)
    Sg = id_tto(d) + (-0.15) * (Δ(d) + (-2.0) * id_tto(d))
    x0 = load_tt("gen_x0")
    E, _ = als_gen_eigsolv(Ag, Sg, x0; sweep_schedule = [4], rmax_schedule = [2])
    out["gen_E"] = Float64.(E)
end

# 12. examples/dft.jl
let d = 10, K = 50
    coeffs = ComplexF64.(vec(g["dft_coeffs"]))
    r = length(coeffs)
    f(x) = sum(coeffs .* exp.(2im * π .* (0:(r - 1)) .* x))
    F = fourier_qtto(d; K = K, sign = -1.0, normalize = true)
    fx = function_to_qtt_uniform(f, d)
    y = tt_compress!(F * fx, 100)
    out["dft_spec"] = matricize(y, d)
end

npzwrite(joinpath(HERE, "hotpath_golden_julia.npz"), out)
println("wrote hotpath_golden_julia.npz with keys: ", join(sort(collect(keys(out))), ", "))
