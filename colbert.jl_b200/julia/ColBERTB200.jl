# ColBERTB200.jl -- the reference-side binding of libcolbert_b200.so (include/colbert_b200.h).
#
# This is the stub a ColBERT.jl maintainer adds: it keeps the public API (`ColBERTConfig`,
# `Searcher`, `search(searcher, query, k)`) and the on-disk index layout exactly as they are and
# replaces what runs underneath `search` (src/searching.jl:103-127) by `ccall`s.  Julia is not
# installed in the image this repo is developed in, so the file is written against the header and
# mirrored 1:1 by the ctypes binding the tests use (colbert.jl_b200/_lib.py, searcher.py).
#
# Memory layouts: a Julia `Matrix{T}(a, b)` is column-major, i.e. the C array `T[b][a]` -- every
# array of `struct Searcher` (src/searching.jl:1-16) is handed over as it is, no copies or
# transposes on the Julia side.  All ids crossing the ABI are 1-based like the reference's.
module ColBERTB200

using ColBERT
using ColBERT: Searcher, ColBERTConfig, encode_queries

const LIB = get(ENV, "COLBERT_B200_LIB", "libcolbert_b200.so")

# ---- status codes -> the reference's exception types --------------------------------------------
const CB_OK, CB_ERR_BAD_ARG, CB_ERR_DOMAIN, CB_ERR_CUDA, CB_ERR_OOM, CB_ERR_UNSUPPORTED, CB_ERR_BOUNDS = 0:6

last_error() = unsafe_string(ccall((:cb_last_error, LIB), Cstring, ()))

function check(status::Int32)
    status == CB_OK && return nothing
    msg = last_error()
    status == CB_ERR_BAD_ARG && throw(DimensionMismatch(msg))      # ranking.jl:9-12, 71-74
    status == CB_ERR_DOMAIN && throw(DomainError(msg))             # residual.jl:701-706, 763-768: DomainError("...") -- the message IS the payload
    status == CB_ERR_OOM && throw(OutOfMemoryError())
    status == CB_ERR_BOUNDS && throw(ErrorException(msg))          # (a pid outside 1:N handed to a hook; `search` raises its own BoundsError below)
    error("libcolbert_b200 (status $status): $msg")                # CUDA / unsupported
end

version() = unsafe_string(ccall((:cb_version, LIB), Cstring, ()))
device_count() = Int(ccall((:cb_device_count, LIB), Int32, ()))

# ---- the resident index --------------------------------------------------------------------------
"""
One index (or one passage-range shard of it) resident in the HBM of one GPU.  Built once from the
host arrays of a `Searcher`; the hot path never touches host memory again.
"""
mutable struct ResidentIndex
    handle::Ptr{Cvoid}
    device::Int
    pid_base::Int
    function ResidentIndex(handle, device, pid_base)
        ix = new(handle, device, pid_base)
        finalizer(ix) do x
            if x.handle != C_NULL
                ccall((:cb_index_destroy, LIB), Int32, (Ptr{Cvoid},), x.handle)
                x.handle = C_NULL
            end
        end
        ix
    end
end

"""
    ResidentIndex(searcher; device = 0)

Uploads `searcher`'s codec, IVF and compressed embeddings (src/searching.jl:44-59).
`centroids` / `bucket_weights` are fetched back to the host once if the user moved them to the
GPU with `Flux.gpu` (src/searching.jl:45-47).
"""
function ResidentIndex(s::Searcher; device::Integer = 0, pid_base::Integer = 0)
    centroids = Array{Float32}(s.centroids)             # (dim, K)   == C float[K][dim]
    weights = Array{Float32}(s.bucket_weights)          # (2^nbits,)
    dim, K = size(centroids)
    handle = Ref{Ptr{Cvoid}}(C_NULL)
    GC.@preserve centroids weights s begin
        check(ccall((:cb_index_create, LIB), Int32,
            (Ref{Ptr{Cvoid}}, Int32, Int32, Int32, Int64, Int64, Int64, Ptr{Float32}, Ptr{Float32},
                Ptr{UInt32}, Ptr{UInt8}, Ptr{Int64}, Ptr{Int64}, Ptr{Int64}, Int64, Int32),
            handle, device, dim, s.config.nbits, K, length(s.doclens), length(s.codes),
            centroids, weights, s.codes, s.residuals, s.doclens, s.ivf, s.ivf_lengths,
            pid_base, 0))
    end
    ResidentIndex(handle[], device, pid_base)
end

# one resident index per Searcher, created lazily on the first search
const _RESIDENT = IdDict{Searcher, ResidentIndex}()
resident(s::Searcher) = get!(() -> ResidentIndex(s), _RESIDENT, s)

# ---- the hot path ---------------------------------------------------------------------------------
"""
    search_batch(ix, Q::Array{Float32, 3}, nprobe, k) -> (pids (k, nq), scores (k, nq), counts (nq,))

`Q` is `encode_queries`' output `(dim, query_maxlen, nq)` (src/modelling/checkpoint.jl:300), which
is exactly the C layout `float[nq][T][dim]`.  Slots beyond `counts[q]` hold pid 0 / -Inf.
"""
function search_batch(ix::ResidentIndex, Q::Array{Float32, 3}, nprobe::Integer, k::Integer)
    dim, T, nq = size(Q)
    pids = zeros(Int64, k, nq)
    scores = fill(-Inf32, k, nq)
    counts = zeros(Int32, nq)
    GC.@preserve Q pids scores counts begin
        check(ccall((:cb_search_batch, LIB), Int32,
            (Ptr{Cvoid}, Ptr{Float32}, Int32, Int32, Int32, Int32, Ptr{Int64}, Ptr{Float32}, Ptr{Int32}),
            ix.handle, Q, nq, T, nprobe, k, pids, scores, counts))
    end
    pids, scores, counts
end

"""
    search_batch_plaid(ix, Q::Array{Float32, 3}, k; ncells = 4, centroid_score_threshold = 0.4f0, ndocs = 1000)

PLAID-style pruned search (the reference's own roadmap item, README.md:187; semantics in
oracle/oracle.py `plaid_search`): candidates from `ncells` cells per token, centroid-score threshold
pruning, approximate centroid-only scores, the best `ndocs` re-scored exactly, stable top-k.
`counts[q]` = number of exactly scored passages.
"""
function search_batch_plaid(ix::ResidentIndex, Q::Array{Float32, 3}, k::Integer; ncells::Integer = 4,
        centroid_score_threshold::Real = 0.4f0, ndocs::Integer = 1000)
    dim, T, nq = size(Q)
    pids = zeros(Int64, k, nq)
    scores = fill(-Inf32, k, nq)
    counts = zeros(Int32, nq)
    GC.@preserve Q pids scores counts begin
        check(ccall((:cb_search_batch_plaid, LIB), Int32,
            (Ptr{Cvoid}, Ptr{Float32}, Int32, Int32, Int32, Float32, Int32, Int32, Ptr{Int64}, Ptr{Float32}, Ptr{Int32}),
            ix.handle, Q, nq, T, ncells, Float32(centroid_score_threshold), ndocs, k, pids, scores, counts))
    end
    pids, scores, counts
end

"""
    ColBERTB200.search(searcher, query, k)    # drop-in for ColBERT.search, src/searching.jl:93-128

Same return value and error behaviour as the reference: `(pids[1:k], scores[1:k])`, ties in
ascending pid, `BoundsError` when fewer than `k` candidates exist (searching.jl:127).

This module does NOT overwrite `ColBERT.search(::Searcher, ::String, ::Int)`: redefining a method another package
owns is an error during precompilation since Julia 1.10 ("Method overwriting is not permitted during Module
precompilation").  Either call `ColBERTB200.search`, or apply the one-line hook of INTEGRATION.md to
src/searching.jl (`search(s, q, k) = ColBERTB200.search(s, q, k)` behind a `use_b200` flag of `ColBERTConfig`).
"""
function search(searcher::Searcher, query::String, k::Int)
    pids, scores = search(searcher, [query], k)
    pids[1], scores[1]
end

"Batched extension (the reference asserts one query per call, searching.jl:98)."
function search(searcher::Searcher, queries::Vector{String}, k::Int)
    Q = encode_queries(searcher.bert, searcher.linear, searcher.tokenizer, queries, searcher.config.dim,
        searcher.config.index_bsize, searcher.config.query_token, searcher.config.attend_to_mask_tokens,
        searcher.skiplist)
    @assert(isequal(size(Q, 2), searcher.config.query_maxlen),
        "size(Q): $(size(Q)), query_maxlen: $(searcher.config.query_maxlen)")
    search(searcher, Array{Float32, 3}(Q), k)
end

function search(searcher::Searcher, Q::Array{Float32, 3}, k::Int)
    pids, scores, counts = search_batch(resident(searcher), Q, searcher.config.nprobe, k)
    for q in eachindex(counts)
        # what `pids[indices][1:k]` throws upstream: BoundsError(<the candidate vector>, (1:k,))
        counts[q] < k && throw(BoundsError(pids[1:counts[q], q], (1:k,)))
    end
    [pids[:, q] for q in axes(pids, 2)], [scores[:, q] for q in axes(scores, 2)]
end

# ---- stage-level entry points with the reference's own signatures ---------------------------------
"`retrieve` (src/search/ranking.jl:23-44) on a resident index: sorted unique candidate pids of one query."
function retrieve(ix::ResidentIndex, nprobe::Int, Q::AbstractMatrix{Float32})
    Qh = Array{Float32}(Q)
    n = Ref{Int64}(0)
    GC.@preserve Qh check(ccall((:cb_retrieve, LIB), Int32,
        (Ptr{Cvoid}, Ptr{Float32}, Int32, Int32, Ptr{Int64}, Int64, Ref{Int64}),
        ix.handle, Qh, size(Qh, 2), nprobe, C_NULL, 0, n))
    pids = zeros(Int64, n[])
    n[] == 0 && return pids
    GC.@preserve Qh pids check(ccall((:cb_retrieve, LIB), Int32,
        (Ptr{Cvoid}, Ptr{Float32}, Int32, Int32, Ptr{Int64}, Int64, Ref{Int64}),
        ix.handle, Qh, size(Qh, 2), nprobe, pids, length(pids), n))
    pids
end

"`decompress` (src/indexing/codecs/residual.jl:759-784); `bsize` is accepted and ignored."
function decompress(dim::Int, nbits::Int, centroids::Matrix{Float32}, bucket_weights::Vector{Float32},
        codes::Vector{UInt32}, residuals::Matrix{UInt8}; bsize::Int = 10000, device::Integer = 0)
    length(codes) == size(residuals, 2) ||
        throw(DomainError("The number of codes should be equal to the number of residual embeddings!"))
    out = zeros(Float32, dim, length(codes))
    GC.@preserve centroids bucket_weights codes residuals out check(ccall((:cb_decompress, LIB), Int32,
        (Int32, Int32, Int32, Int64, Ptr{Float32}, Ptr{Float32}, Ptr{UInt32}, Ptr{UInt8}, Int64,
            Ptr{Float32}, Ptr{UInt8}, Ptr{Float32}),
        device, dim, nbits, size(centroids, 2), centroids, bucket_weights, codes, residuals, length(codes),
        out, C_NULL, C_NULL))
    out
end

"`maxsim(Q, D, pids, doclens)` (src/search/ranking.jl:69-86)."
function maxsim(Q::Matrix{Float32}, D::Matrix{Float32}, pids::Vector{Int}, doclens::Vector{Int};
        device::Integer = 0)
    out = zeros(Float32, length(pids))
    GC.@preserve Q D pids doclens out check(ccall((:cb_maxsim, LIB), Int32,
        (Int32, Int32, Int32, Ptr{Float32}, Ptr{Float32}, Int64, Ptr{Int64}, Int64, Ptr{Int64}, Int64, Ptr{Float32}),
        device, size(Q, 1), size(Q, 2), Q, D, size(D, 2), pids, length(pids), doclens, length(doclens), out))
    out
end

"`compress` (src/indexing/codecs/residual.jl:586-604); `bsize` is accepted and ignored."
function compress(centroids::Matrix{Float32}, bucket_cutoffs::Vector{Float32}, dim::Int, nbits::Int,
        embs::Matrix{Float32}; bsize::Int = 10000, device::Integer = 0)
    n = size(embs, 2)
    codes = zeros(UInt32, n)
    residuals = Matrix{UInt8}(undef, div(dim, 8) * nbits, n)
    GC.@preserve centroids bucket_cutoffs embs codes residuals check(ccall((:cb_compress, LIB), Int32,
        (Int32, Int32, Int32, Int64, Ptr{Float32}, Ptr{Float32}, Ptr{Float32}, Int64, Ptr{UInt32}, Ptr{UInt8}),
        device, dim, nbits, size(centroids, 2), centroids, bucket_cutoffs, embs, n, codes, residuals))
    codes, residuals
end

"""
    ResidentIndex(index_path::String; device = 0, shard = 0, n_shards = 1)

Opens the index directory the Indexer wrote (src/savers.jl) natively: the JLD2 files are mapped and uploaded by the
library (cb_index_open), no host copy of the 20+ GB arrays is ever built (what `Searcher(index_path)` does at
src/searching.jl:50-55).
"""
function ResidentIndex(index_path::String; device::Integer = 0, shard::Integer = 0, n_shards::Integer = 1)
    handle = Ref{Ptr{Cvoid}}(C_NULL)
    base = Ref{Int64}(0)
    check(ccall((:cb_index_open, LIB), Int32, (Ref{Ptr{Cvoid}}, Cstring, Int32, Int32, Int32, Ref{Int64}),
        handle, index_path, device, shard, n_shards, base))
    ResidentIndex(handle[], device, base[])
end

# ---- passage-sharded search over the GPUs of one box ----------------------------------------------
"""
    ShardedIndex(index_path, n_gpus)      # opens the shards from disk (cb_multi_open)
    ShardedIndex(searcher, n_gpus)        # from the host arrays of a loaded Searcher

The passages are split into `n_gpus` contiguous ranges balanced by embedding count, one shard per device (each
shard's IVF is rebuilt on its device from its codes, `_build_ivf` semantics).  `search_batch` is ONE ccall
(cb_multi_search_batch): this Julia thread drives all devices -- queries uploaded once and forwarded over NVLink,
stage 1 split by query, every device scoring its range concurrently, the per-shard top-k lists merged on the first
device by (score desc, pid asc), the order of the reference's stable `sortperm`.  No process per GPU, no NCCL.
"""
mutable struct ShardedIndex
    handle::Ptr{Cvoid}
    shards::Vector{ResidentIndex}        # kept alive (borrowed by the group); empty when the library owns them
    function ShardedIndex(handle, shards)
        sx = new(handle, shards)
        finalizer(sx) do x
            if x.handle != C_NULL
                ccall((:cb_multi_destroy, LIB), Int32, (Ptr{Cvoid},), x.handle)
                x.handle = C_NULL
            end
        end
        sx
    end
end

function ShardedIndex(index_path::String, n_gpus::Integer)
    handle = Ref{Ptr{Cvoid}}(C_NULL)
    check(ccall((:cb_multi_open, LIB), Int32, (Ref{Ptr{Cvoid}}, Cstring, Int32, Ptr{Int32}), handle, index_path, n_gpus, C_NULL))
    ShardedIndex(handle[], ResidentIndex[])
end

function ShardedIndex(s::Searcher, n_gpus::Integer)
    csum = cumsum(s.doclens)
    total = csum[end]
    cuts = [searchsortedfirst(csum, total * r ÷ n_gpus) for r in 1:(n_gpus - 1)]
    bounds = [0; cuts; length(s.doclens)]
    centroids = Array{Float32}(s.centroids)
    weights = Array{Float32}(s.bucket_weights)
    dim, K = size(centroids)
    shards = ResidentIndex[]
    for r in 1:n_gpus
        lo, hi = bounds[r], bounds[r + 1]                # passages lo+1 .. hi
        e_lo = lo == 0 ? 0 : csum[lo]
        e_hi = hi == 0 ? 0 : csum[hi]
        codes = view(s.codes, (e_lo + 1):e_hi)
        residuals = view(s.residuals, :, (e_lo + 1):e_hi)
        doclens = view(s.doclens, (lo + 1):hi)
        handle = Ref{Ptr{Cvoid}}(C_NULL)
        GC.@preserve centroids weights s check(ccall((:cb_index_create, LIB), Int32,
            (Ref{Ptr{Cvoid}}, Int32, Int32, Int32, Int64, Int64, Int64, Ptr{Float32}, Ptr{Float32},
                Ptr{UInt32}, Ptr{UInt8}, Ptr{Int64}, Ptr{Int64}, Ptr{Int64}, Int64, Int32),
            handle, r - 1, dim, s.config.nbits, K, hi - lo, e_hi - e_lo, centroids, weights,
            pointer(codes), pointer(residuals), pointer(doclens), C_NULL, C_NULL, lo, 0))
        push!(shards, ResidentIndex(handle[], r - 1, lo))
    end
    group = Ref{Ptr{Cvoid}}(C_NULL)
    handles = [ix.handle for ix in shards]
    GC.@preserve handles check(ccall((:cb_multi_create, LIB), Int32, (Ref{Ptr{Cvoid}}, Int32, Ptr{Ptr{Cvoid}}), group, n_gpus, handles))
    ShardedIndex(group[], shards)
end

function search_batch(sx::ShardedIndex, Q::Array{Float32, 3}, nprobe::Integer, k::Integer)
    dim, T, nq = size(Q)
    pids = zeros(Int64, k, nq)
    scores = fill(-Inf32, k, nq)
    counts = zeros(Int32, nq)
    GC.@preserve Q pids scores counts check(ccall((:cb_multi_search_batch, LIB), Int32,
        (Ptr{Cvoid}, Ptr{Float32}, Int32, Int32, Int32, Int32, Ptr{Int64}, Ptr{Float32}, Ptr{Int32}),
        sx.handle, Q, nq, T, nprobe, k, pids, scores, counts))
    pids, scores, counts
end

end # module
