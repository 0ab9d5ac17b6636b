"""A minimal writer of JLD2 0.4 containers, for tests only (Julia / JLD2.jl are not installed here).

`JLD2.save_object(path, x)` for the array types ColBERT.jl stores (src/savers.jl:16-29, 52-84) produces an
HDF5-subset file: a 512-byte text header, a version-2 superblock at offset 512 with base address 512, a
version-2 object header per object, the root group's members as Link messages, and one dataset named
"single_stored_object" with Dataspace / Datatype / Fill-value / Data-layout messages -- compact layout below
8 KB, contiguous above (JLD2 datasets.jl), dimensions reversed, checksums = Jenkins lookup3 (HDF5 spec).
This module emits exactly that structure from numpy arrays.  It follows the published format, not bytes
captured from a real JLD2 run, so the reader's parity against real files stays UNPINNED here
(bench/julia_write_index.jl produces real ones on a machine with Julia)."""
import json
import os
import struct

import numpy as np

UNDEF = 0xFFFFFFFFFFFFFFFF
BASE = 512


def _rot(x, k):
    return ((x << k) | (x >> (32 - k))) & 0xFFFFFFFF


def lookup3(data: bytes, initval: int = 0) -> int:
    """Bob Jenkins' lookup3 hashlittle, the checksum of HDF5 metadata (H5_checksum_lookup3)."""
    M = 0xFFFFFFFF
    n = len(data)
    a = b = c = (0xDEADBEEF + n + initval) & M
    i = 0
    while n > 12:
        a = (a + int.from_bytes(data[i:i + 4], "little")) & M
        b = (b + int.from_bytes(data[i + 4:i + 8], "little")) & M
        c = (c + int.from_bytes(data[i + 8:i + 12], "little")) & M
        a = (a - c) & M; a ^= _rot(c, 4); c = (c + b) & M
        b = (b - a) & M; b ^= _rot(a, 6); a = (a + c) & M
        c = (c - b) & M; c ^= _rot(b, 8); b = (b + a) & M
        a = (a - c) & M; a ^= _rot(c, 16); c = (c + b) & M
        b = (b - a) & M; b ^= _rot(a, 19); a = (a + c) & M
        c = (c - b) & M; c ^= _rot(b, 4); b = (b + a) & M
        i += 12
        n -= 12
    if n == 0:
        return c
    tail = data[i:] + b"\0" * (12 - n)
    a = (a + int.from_bytes(tail[0:4], "little")) & M
    b = (b + int.from_bytes(tail[4:8], "little")) & M
    c = (c + int.from_bytes(tail[8:12], "little")) & M
    c ^= b; c = (c - _rot(b, 14)) & M
    a ^= c; a = (a - _rot(c, 11)) & M
    b ^= a; b = (b - _rot(a, 25)) & M
    c ^= b; c = (c - _rot(b, 16)) & M
    a ^= c; a = (a - _rot(c, 4)) & M
    b ^= a; b = (b - _rot(a, 14)) & M
    c ^= b; c = (c - _rot(b, 24)) & M
    return c


def _msg(mtype, body, flags=0):
    return struct.pack("<BHB", mtype, len(body), flags) + body


def _object_header(messages: bytes) -> bytes:
    """version-2 object header, no timestamps; chunk-0 size field as narrow as it fits (flags bits 0-1)."""
    n = len(messages)
    sz_flag = 0 if n < 256 else 1 if n < 65536 else 2
    head = b"OHDR" + struct.pack("<BB", 2, sz_flag) + n.to_bytes(1 << sz_flag, "little")
    body = head + messages
    return body + struct.pack("<I", lookup3(body))


def _datatype(dt: np.dtype) -> bytes:
    dt = np.dtype(dt)
    if dt.kind == "f":
        sz = dt.itemsize
        if sz == 4:
            props = struct.pack("<HHBBBBI", 0, 32, 23, 8, 0, 23, 127)
            bits = bytes([0x20, 31, 0])
        else:
            props = struct.pack("<HHBBBBI", 0, 64, 52, 11, 0, 52, 1023)
            bits = bytes([0x20, 63, 0])
        return bytes([0x11]) + bits + struct.pack("<I", sz) + props          # version 1, class 1 (floating point)
    if dt.kind in "iu":
        bits = bytes([0x08 if dt.kind == "i" else 0x00, 0, 0])
        return bytes([0x10]) + bits + struct.pack("<I", dt.itemsize) + struct.pack("<HH", 0, 8 * dt.itemsize)   # class 0
    raise TypeError(dt)


def save_object(path: str, x, compact_below: int = 8192):
    """`JLD2.save_object(path, x)` for a numpy array or scalar.  A Julia array of size (a, b) is passed as the
    numpy array of shape (b, a) (same memory); the dataspace then holds (b, a) -- Julia's dims reversed."""
    x = np.asarray(x)
    data = np.ascontiguousarray(x).tobytes()
    if x.ndim == 0:
        space = struct.pack("<BBBB", 2, 0, 0, 0)                             # version 2, scalar dataspace
    else:
        space = struct.pack("<BBBB", 2, x.ndim, 0, 1) + b"".join(struct.pack("<Q", d) for d in x.shape)
    fill = bytes([3, 0x09])                                                   # fill value v3: allocate early, undefined
    pos = BASE + 48                                                           # first byte after the superblock
    blobs = []
    if len(data) < compact_below:
        layout = struct.pack("<BBH", 3, 0, len(data)) + data                  # compact
    else:
        data_addr = pos
        blobs.append(data)
        pos += len(data)
        pos += (-pos) % 8
        blobs.append(b"\0" * ((-len(data)) % 8))
        layout = struct.pack("<BBQQ", 3, 1, data_addr - BASE, len(data))      # contiguous, address relative to base
    ds_hdr = _object_header(_msg(0x01, space) + _msg(0x03, _datatype(x.dtype), flags=1) + _msg(0x05, fill) + _msg(0x08, layout))
    ds_addr = pos
    pos += len(ds_hdr)
    name = b"single_stored_object"
    link = struct.pack("<BB", 1, 0x10) + bytes([1]) + struct.pack("<B", len(name)) + name + struct.pack("<Q", ds_addr - BASE)
    link_info = struct.pack("<BB", 0, 0) + struct.pack("<QQ", UNDEF, UNDEF)
    group_info = struct.pack("<BB", 0, 0)
    root_hdr = _object_header(_msg(0x02, link_info) + _msg(0x0A, group_info) + _msg(0x06, link) + _msg(0x00, b"\0" * 24))
    root_addr = pos
    pos += len(root_hdr)
    sb = b"\x89HDF\r\n\x1a\n" + struct.pack("<BBBB", 2, 8, 8, 0) + struct.pack("<QQQQ", BASE, UNDEF, pos - BASE, root_addr - BASE)
    sb += struct.pack("<I", lookup3(sb))
    text = b"HDF5-based Julia Data Format, version 0.1.1\x00 (Julia 1.10.0 x86_64-linux-gnu, written by tests/jld2_writer.py)"
    with open(path, "wb") as f:
        f.write(text + b"\0" * (BASE - len(text)))
        f.write(sb)
        for b in blobs:
            f.write(b)
        f.write(ds_hdr)
        f.write(root_hdr)


def write_index(path: str, ix: dict, n_chunks: int = 3, nprobe: int = 2, compact_below: int = 8192):
    """The directory `index(indexer)` leaves behind (src/indexing.jl:63-147, src/savers.jl): config.json, plan.json,
    the codec files, <chunk>.codes / .residuals / .metadata.json, doclens.<chunk>, ivf, ivf_lengths.
    `ix` is a colbert_jl_b200.synthetic index (C layouts)."""
    os.makedirs(path, exist_ok=True)
    dim, nbits = ix["dim"], ix["nbits"]
    doclens = np.asarray(ix["doclens"], dtype=np.int64)
    Np, Ne = len(doclens), int(doclens.sum())
    cfg = {"use_gpu": False, "rank": 0, "nranks": 1, "query_token_id": "[unused0]", "doc_token_id": "[unused1]", "query_token": "[Q]",
           "doc_token": "[D]", "checkpoint": "colbert-ir/colbertv2.0", "collection": "synthetic", "dim": dim, "doc_maxlen": 300,
           "mask_punctuation": True, "query_maxlen": 32, "attend_to_mask_tokens": False, "index_path": path, "index_bsize": 64,
           "chunksize": None, "passages_batch_size": 300, "nbits": nbits, "kmeans_niters": 20, "nprobe": nprobe, "ncandidates": 8192}
    json.dump(cfg, open(os.path.join(path, "config.json"), "w"), indent=4)
    bounds = np.linspace(0, Np, n_chunks + 1).astype(np.int64)
    cs = np.concatenate([[0], np.cumsum(doclens)])
    plan = {"chunksize": int(bounds[1] - bounds[0]), "num_chunks": n_chunks, "avg_doclen_est": float(doclens.mean()) if Np else 0.0,
            "num_documents": Np, "num_embeddings_est": float(Ne), "num_embeddings": Ne, "num_partitions": int(ix["centroids"].shape[0]),
            "embeddings_offsets": [int(cs[b]) + 1 for b in bounds[:-1]]}
    json.dump(plan, open(os.path.join(path, "plan.json"), "w"), indent=4)
    so = lambda name, a: save_object(os.path.join(path, name), a, compact_below)
    so("centroids.jld2", np.asarray(ix["centroids"], np.float32))                  # Matrix{Float32}(dim, K) == C [K][dim]
    so("bucket_weights.jld2", np.asarray(ix["bucket_weights"], np.float32))
    so("bucket_cutoffs.jld2", np.zeros((1 << nbits) - 1, np.float32))
    so("avg_residual.jld2", np.float32(0.0123))                                    # a Float32 scalar
    so("ivf.jld2", np.asarray(ix["ivf"], np.int64))
    so("ivf_lengths.jld2", np.asarray(ix["ivf_lengths"], np.int64))
    for c in range(n_chunks):
        p0, p1 = int(bounds[c]), int(bounds[c + 1])
        e0, e1 = int(cs[p0]), int(cs[p1])
        so(f"{c + 1}.codes.jld2", np.asarray(ix["codes"][e0:e1], np.uint32))
        so(f"{c + 1}.residuals.jld2", np.asarray(ix["residuals"][e0:e1], np.uint8))   # Matrix{UInt8}(R, n) == C [n][R]
        so(f"doclens.{c + 1}.jld2", doclens[p0:p1])
        json.dump({"passage_offset": p0 + 1, "num_passages": p1 - p0, "num_embeddings": e1 - e0, "embedding_offset": e0 + 1},
                  open(os.path.join(path, f"{c + 1}.metadata.json"), "w"), indent=4)
