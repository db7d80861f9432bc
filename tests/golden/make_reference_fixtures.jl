# make_reference_fixtures.jl — pins the parity tests to the REAL reference.
#
#     julia --project=<env with BulkLMM.jl v1.2.0> tests/golden/make_reference_fixtures.jl
#
# Reads tests/golden/ref_inputs/*.csv (written by make_reference_inputs.py, committed), runs senresearch/BulkLMM.jl
# itself on them — bulkscan x 3 methods, scan null / permutations / alt, calcKinship, get_thresholds, lod2log10p —
# and writes every result to tests/golden/ref_outputs/*.csv with round-trip float formatting.  The permutation
# indices the reference's RNG draws are dumped too (0-based), so that the engine can be fed the same shuffles, and
# so is eigen(K), because permutation LODs depend on eigenvector signs (SURVEY section 7, hard part 3e).
# tests/test_reference_fixtures.py consumes the directory when it exists: the CPU oracle (-m "not gpu") and the
# CUDA engine (-m gpu) are then checked against the reference's own numbers.
#
# STATUS: Julia is not installed in the build image, so this script has not been executed there; the ref_outputs
# directory is therefore absent and the consuming tests skip (parity stays "unpinned" until someone runs this once).
using BulkLMM, DelimitedFiles, LinearAlgebra, Random, Statistics

const HERE = @__DIR__
const INP = joinpath(HERE, "ref_inputs")
const OUT = joinpath(HERE, "ref_outputs")
mkpath(OUT)

rd(name) = readdlm(joinpath(INP, name * ".csv"), ',', Float64)
function wr(name, A)
    open(joinpath(OUT, name * ".csv"), "w") do io
        writedlm(io, A, ',')     # Julia prints the shortest representation that round-trips
    end
end

G = rd("G"); K = rd("K"); Y = rd("Y"); Z = rd("Covar"); w = vec(rd("weights"))
(n, p) = size(G); m = size(Y, 2)
grid = collect(0.0:0.1:0.9)

# versions
wr("versions", reshape([string(VERSION), string(pkgversion(BulkLMM))], :, 1))

# calcKinship (src/kinship.jl:4-14) and the decomposition inside transform_rotation (src/transform_helpers.jl:21-34)
wr("kinship", calcKinship(G))
EF = eigen(K)
wr("eig_U", EF.vectors); wr("eig_lambda", EF.values)
(Y0, X0, lambda0) = transform_rotation(Y, G, K)
wr("rot_Y0", Y0); wr("rot_X0", X0)

# bulkscan, three methods (src/bulkscan.jl:81-526)
r = bulkscan_null_grid(Y, G, K, grid);                         wr("nullgrid_L", r.L);      wr("nullgrid_h2", r.h2_null_list)
r = bulkscan_null_grid(Y, G, Z, K, grid; reml = true);         wr("nullgrid_cov_reml_L", r.L); wr("nullgrid_cov_reml_h2", r.h2_null_list)
r = bulkscan_null_grid(Y, G, K, grid; weights = w);            wr("nullgrid_weights_L", r.L);  wr("nullgrid_weights_h2", r.h2_null_list)
r = bulkscan_alt_grid(Y, G, K, grid);                          wr("altgrid_L", r.L);       wr("altgrid_h2panel", r.h2_panel)
r = bulkscan_alt_grid(Y, G, K, grid; reml = true);             wr("altgrid_reml_L", r.L);  wr("altgrid_reml_h2panel", r.h2_panel)
r = bulkscan_null(Y, G, K; nb = 1, reml = true, prior_variance = 0.0);   wr("nullexact_reml_L", r.L); wr("nullexact_reml_h2", r.h2_null_list)
r = bulkscan_null(Y, G, Z, K; nb = 1, optim_interval = 4);     wr("nullexact_cov_oi4_L", r.L); wr("nullexact_cov_oi4_h2", r.h2_null_list)
r = bulkscan(Y, G, K; method = "null-grid", output_pvals = true); wr("bulkscan_pvals", r.log10Pvals_mat)

# scan: single trait (src/scan.jl:94-360)
y = reshape(Y[:, 3], :, 1)
for (tag, reml) in (("ml", false), ("reml", true))
    s = scan(y, G, K; reml = reml)
    wr("scan_null_$(tag)_lod", s.lod); wr("scan_null_$(tag)_scalars", [s.sigma2_e, s.h2_null])
end
s = scan(y, G, Z, K; reml = true)
wr("scan_null_cov_reml_lod", s.lod); wr("scan_null_cov_reml_scalars", [s.sigma2_e, s.h2_null])
s = scan(y, G, K; assumption = "alt")
wr("scan_alt_lod", s.lod); wr("scan_alt_h2_each", s.h2_each_marker); wr("scan_alt_scalars", [s.sigma2_e, s.h2_null])
wr("lod2log10p_df1", lod2log10p.(scan(y, G, K).lod, 1)); wr("lod2log10p_df3", lod2log10p.(scan(y, G, K).lod, 3))

# scan with permutations (src/scan.jl:485-557).  The shuffles: MersenneTwister(rndseed) + shuffle per column
# (src/transform_helpers.jl:94-102, src/util.jl:162-179); shuffle(rng, x) permutes positions identically for any x,
# so shuffling 1:n with an identically seeded generator yields the indices the reference applies to r0.
nperms = 64; rndseed = 0
s = scan(y, G, K; permutation_test = true, nperms = nperms, rndseed = rndseed)
wr("perms_lod", s.lod); wr("perms_L", s.L_perms); wr("perms_scalars", [s.sigma2_e, s.h2_null])
rng = MersenneTwister(rndseed)
idx = zeros(Int64, n, nperms)
for k in 1:nperms
    idx[:, k] = shuffle(rng, collect(1:n)) .- 1
end
wr("perms_idx0", idx)
t = get_thresholds(s.L_perms, [0.10, 0.05])
wr("perms_thresholds", t.thrs)
println("wrote ", OUT)
