# bench/julia_write_index.jl -- for machines WITH Julia (none in the build container): writes a small synthetic index
# with the reference's own savers (real JLD2.jl bytes) plus every array as raw little-endian binary, so that
#   python tools/check_jld2_reader.py <dir>
# can pin the native reader (colbert.jl_b200/csrc/jld2.cpp, cb_index_open) against files JLD2 itself produced.
# usage: julia --project=<ColBERT.jl checkout> bench/julia_write_index.jl <out_dir> [n_passages] [K]
# Written blind (never executed here); it only calls the reference's public savers (src/savers.jl:16-29, 52-84).
using ColBERT, JLD2, JSON, Random

out = ARGS[1]
n_passages = length(ARGS) >= 2 ? parse(Int, ARGS[2]) : 2000
K = length(ARGS) >= 3 ? parse(Int, ARGS[3]) : 512
dim, nbits, n_chunks = 128, 2, 3
rng = MersenneTwister(1234)
mkpath(out)

centroids = randn(rng, Float32, dim, K)
centroids ./= sqrt.(sum(abs2, centroids, dims = 1))
bucket_weights = Float32[-0.041035336, -0.009812315, 0.008938393, 0.039779153]
bucket_cutoffs = Float32[-0.02, 0.0, 0.02]
doclens = rand(rng, 8:120, n_passages)
n_e = sum(doclens)
codes = rand(rng, UInt32(1):UInt32(K), n_e)
residuals = rand(rng, UInt8, div(dim, 8) * nbits, n_e)

config = ColBERTConfig(index_path = out, dim = dim, nbits = nbits)
ColBERT.save(config)
ColBERT.save_codec(out, centroids, bucket_cutoffs, bucket_weights, Float32(0.0123))
bounds = round.(Int, range(0, n_passages, length = n_chunks + 1))
cs = cumsum([0; doclens])
for c in 1:n_chunks
    p = (bounds[c] + 1):bounds[c + 1]
    e = (cs[bounds[c] + 1] + 1):cs[bounds[c + 1] + 1]
    ColBERT.save_chunk(out, codes[e], residuals[:, e], c, bounds[c] + 1, doclens[p])
end
ivf, ivf_lengths = ColBERT._build_ivf(codes, K)
JLD2.save_object(joinpath(out, "ivf.jld2"), ivf)
JLD2.save_object(joinpath(out, "ivf_lengths.jld2"), ivf_lengths)
open(joinpath(out, "plan.json"), "w") do io
    JSON.print(io, Dict("num_chunks" => n_chunks, "num_embeddings" => n_e, "num_partitions" => K, "num_documents" => n_passages), 4)
end

# raw twins: column-major Julia memory == the C layout the library takes
raw = joinpath(out, "raw")
mkpath(raw)
for (name, a) in ("centroids" => centroids, "bucket_weights" => bucket_weights, "doclens" => doclens, "codes" => codes,
    "residuals" => residuals, "ivf" => ivf, "ivf_lengths" => ivf_lengths)
    write(joinpath(raw, name * ".bin"), a)
end
open(joinpath(raw, "shapes.json"), "w") do io
    JSON.print(io, Dict("dim" => dim, "nbits" => nbits, "K" => K, "n_passages" => n_passages, "n_embeddings" => n_e, "n_chunks" => n_chunks))
end
println("wrote ", out)
