# The reference's OWN search-time scoring path, timed on a fixture written by bench/export_fixture.py.
# NOT executed in this repo's image (no Julia there): for machines that have Julia + ColBERT.jl.
#
#   julia --project=/path/to/ColBERT.jl bench/julia_baseline.jl FIXTURE_DIR
#
# Runs, per query, exactly what `search` runs after the encoder (src/searching.jl:103-127):
# `retrieve` -> `_collect_compressed_embs_for_pids` -> `decompress` -> `maxsim` -> `sortperm`,
# prints queries/s (BLAS threads = Sys.CPU_THREADS, everything else single-threaded as upstream)
# and checks the top-k against the oracle's (expected.json): pids identical, scores within 1e-3.
using ColBERT, JSON, LinearAlgebra

dir = ARGS[1]
meta = JSON.parsefile(joinpath(dir, "meta.json"))
dim, nbits, K, Np, Ne, R = meta["dim"], meta["nbits"], meta["K"], meta["N_p"], meta["N_e"], meta["R"]
nq, T, nprobe, k = meta["nq"], meta["T"], meta["nprobe"], meta["k"]
rd(name, ::Type{Ty}, dims...) where {Ty} = (a = Array{Ty}(undef, dims...); read!(joinpath(dir, name), a); a)

centroids = rd("centroids.f32", Float32, dim, K)
bucket_weights = rd("bucket_weights.f32", Float32, 2^nbits)
codes = rd("codes.u32", UInt32, Ne)
residuals = rd("residuals.u8", UInt8, R, Ne)
doclens = Vector{Int}(rd("doclens.i64", Int64, Np))
ivf = Vector{Int}(rd("ivf.i64", Int64, Ne))
ivf_lengths = Vector{Int}(rd("ivf_lengths.i64", Int64, K))
Qs = rd("queries.f32", Float32, dim, T, nq)
emb2pid = ColBERT._build_emb2pid(doclens)                      # src/searching.jl:82-91
BLAS.set_num_threads(Sys.CPU_THREADS)

function search_one(Q)
    pids = ColBERT.retrieve(ivf, ivf_lengths, centroids, emb2pid, nprobe, Q)
    codes_packed, residuals_packed = ColBERT._collect_compressed_embs_for_pids(doclens, codes, residuals, pids)
    D = ColBERT.decompress(dim, nbits, centroids, bucket_weights, codes_packed, residuals_packed)
    scores = ColBERT.maxsim(Q, D, pids, doclens)
    idx = sortperm(scores, rev = true)
    pids[idx][1:min(k, end)], scores[idx][1:min(k, end)]
end

search_one(Qs[:, :, 1])                                        # compile
t = @elapsed results = [search_one(Qs[:, :, q]) for q in 1:nq]
println("ColBERT.jl search minus encoder: ", round(nq / t, digits = 3), " queries/s over ", nq, " queries (",
    Sys.CPU_THREADS, " BLAS threads, JULIA_NUM_THREADS=", Threads.nthreads(), ")")

expected = JSON.parsefile(joinpath(dir, "expected.json"))
ok = true
for q in 1:nq
    p, s = results[q]
    ep, es = Vector{Int}(expected[q]["pids"]), Vector{Float32}(expected[q]["scores"])
    global ok &= (p == ep) && all(abs.(s .- es) .<= 1f-3 .* abs.(es))
end
println(ok ? "parity with the oracle: top-k pids identical, scores within 1e-3" : "PARITY MISMATCH with the oracle's expected.json")
