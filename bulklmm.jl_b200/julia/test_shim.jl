# test_shim.jl — executes the Julia shim (BulkLMMB200.jl) end to end and, when the reference package is installed in the
# active environment, compares every entry point with BulkLMM.jl itself on the same inputs.
#
#     BLMM_B200_LIB=/path/to/libblmm_b200.so julia --project=<env> bulklmm.jl_b200/julia/test_shim.jl
#
# STATUS: like the shim, never executed in the build image (no Julia there).  It is the script a maintainer runs once
# on a B200 box with Julia: it turns "shim unexecuted" and (with BulkLMM installed) "parity unpinned" into test results.
using Test, Random, LinearAlgebra, Statistics
include(joinpath(@__DIR__, "BulkLMMB200.jl"))
const B = BulkLMMB200
const HAVE_REF = try
    @eval import BulkLMM
    true
catch
    false
end

rel(a, b) = maximum(abs.(a .- b) ./ max.(1.0, abs.(b)))

Random.seed!(1)
n, p, m = 79, 300, 40
G = Float64.(rand(n, p) .< 0.5)
G[rand(n, p) .< 0.02] .= 0.37
K = B.calcKinship(G)
Y = 11.0 .+ 0.5 .* randn(n, m) .+ 0.4 .* G[:, 7] .* randn(1, m)
Z = [Float64.(rand(n) .< 0.5) randn(n)]
grid = collect(0.0:0.1:0.9)

@testset "shim runs and is self-consistent" begin
    @test size(K) == (n, n) && K ≈ K' && all(diag(K) .== 1.0)
    r = B.bulkscan(Y, G, K)                                  # defaults: null-grid
    @test size(r.L) == (p, m) && length(r.h2_null_list) == m && all(isfinite, r.L)
    a = B.bulkscan(Y, G, K; method = "alt-grid")
    @test all(a.L .>= r.L .- 1e-9) && all(in(grid), a.h2_panel)
    e = B.bulkscan(Y, G, K; method = "null-exact", reml = true, prior_variance = 0.0)
    s = B.scan(Y[:, 3], G, K; reml = true)                   # vector trait method
    @test rel(e.L[:, 3], s.lod) < 1e-9 && e.h2_null_list[3] == s.h2_null
    sp = B.scan(reshape(Y[:, 3], :, 1), G, K; reml = true, permutation_test = true, nperms = 64, rndseed = 0)
    @test size(sp.L_perms) == (p, 64) && rel(sp.lod, s.lod) < 1e-9
    t = B.get_thresholds(sp.L_perms, [0.10, 0.05])
    @test t.thrs ≈ [quantile(vec(maximum(sp.L_perms, dims = 1)), q) for q in (0.90, 0.95)]
    pv = B.scan(Y[:, 3], G, Z, K; output_pvals = true)
    @test pv.log10pvals ≈ B.lod2log10p(pv.lod, 1) && B.lod2log10p(pv.lod[1], 1) ≈ pv.log10pvals[1]
    @test_throws ErrorException B.bulkscan(Y, G, K; h2_grid = [0.5, 1.0])          # "Heritability of 1 is not allowed."
    @test_throws ErrorException B.scan(Y[:, 1:2], G, K; permutation_test = true)   # "Can only handle one trait."
    if B.device_count(B.default_context()) >= 1 && get(ENV, "BLMM_B200_NDEV", "1") != "1"
        nd = parse(Int, ENV["BLMM_B200_NDEV"])
        a8 = B.bulkscan(Y, G, K; method = "alt-grid", ndev = nd)
        @test a8.L == a.L && a8.h2_panel == a.h2_panel        # bit-identical on several GPUs
    end
end

if HAVE_REF
    @testset "against BulkLMM.jl itself" begin
        for (meth, kw) in (("null-grid", ()), ("alt-grid", ()), ("null-exact", (reml = true, prior_variance = 0.0)))
            ref = BulkLMM.bulkscan(Y, G, K; method = meth, kw...)
            got = B.bulkscan(Y, G, K; method = meth, kw...)
            @test rel(got.L, ref.L) < (meth == "null-exact" ? 1e-5 : 1e-8)
            meth == "null-grid" && @test got.h2_null_list == ref.h2_null_list
            meth == "alt-grid" && @test mean(got.h2_panel .!= ref.h2_panel) < 1e-3
            meth == "null-exact" && @test maximum(abs.(got.h2_null_list .- ref.h2_null_list)) < 2e-7
        end
        ref = BulkLMM.bulkscan(Y, G, Z, K; reml = true)
        got = B.bulkscan(Y, G, Z, K; reml = true)
        @test rel(got.L, ref.L) < 1e-8 && got.h2_null_list == ref.h2_null_list
        # permutations: same MersenneTwister(rndseed) shuffles on both sides; eigenvector signs differ between LAPACK
        # and cuSOLVER, which moves permuted LODs (not the un-permuted column), so compare distributions and the scan
        y = reshape(Y[:, 5], :, 1)
        ref = BulkLMM.scan(y, G, K; permutation_test = true, nperms = 256, rndseed = 7)
        got = B.scan(y, G, K; permutation_test = true, nperms = 256, rndseed = 7)
        @test rel(got.lod, ref.lod) < 1e-5 && abs(got.h2_null - ref.h2_null) < 2e-7
        @test abs(median(vec(maximum(got.L_perms, dims = 1))) - median(vec(maximum(ref.L_perms, dims = 1)))) < 0.5
        @test B.calcKinship(G) ≈ BulkLMM.calcKinship(G) atol = 1e-13
    end
else
    @info "BulkLMM.jl is not installed in this environment: reference comparison skipped"
end
