"""CPU: the native JLD2 reader (csrc/jld2.cpp, through cb_jld2_read) against files laid out as JLD2 0.4 lays them out
(tests/jld2_writer.py: 512-byte header, superblock v2 at 512 with base address 512, v2 object headers, Link messages,
compact layout below 8 KB, contiguous above, reversed dims).  Mirrors the reference's loader/saver round-trip tests
(test/loaders_and_savers.jl:6-120: save_codec -> load_codec, save_chunk -> load_doclens / load_compressed_embs).
Parity against bytes written by the real JLD2.jl is UNPINNED in this container (no Julia): see bench/julia_write_index.jl."""
import os
import struct

import numpy as np
import pytest

import colbert_jl_b200 as cb
from tests import jld2_writer as W

rng = np.random.default_rng(77)


@pytest.mark.parametrize("dtype", [np.float32, np.float64, np.int64, np.uint32, np.uint8, np.int32, np.uint16])
@pytest.mark.parametrize("shape", [(), (5,), (3000,), (7, 16), (300, 128)])
def test_save_object_round_trip(tmp_path, dtype, shape):
    # small arrays are compact (inside the object header), large ones contiguous: both paths, every element type
    a = (rng.random(shape) * 100).astype(dtype) if np.dtype(dtype).kind == "f" else rng.integers(0, 200, shape).astype(dtype)
    p = str(tmp_path / "x.jld2")
    W.save_object(p, a)
    b = cb.load_object(p)
    assert b.dtype == a.dtype and b.shape == a.shape and np.array_equal(a, b)


def test_codec_round_trip(tmp_path):   # test/loaders_and_savers.jl:6-38 (save_codec / load_codec)
    centroids = rng.random((128, 500), dtype=np.float32)         # Julia shape (dim, K)
    cutoffs, weights, avg = rng.random(3, dtype=np.float32), rng.random(4, dtype=np.float32), np.float32(rng.random())
    for name, arr in (("centroids", centroids.T), ("bucket_cutoffs", cutoffs), ("bucket_weights", weights), ("avg_residual", avg)):
        W.save_object(str(tmp_path / f"{name}.jld2"), arr)
    c = cb.load_object(str(tmp_path / "centroids.jld2"))
    assert c.shape == (500, 128) and np.array_equal(c.T, centroids)               # C [K][dim] == Julia (dim, K)
    assert np.array_equal(cb.load_object(str(tmp_path / "bucket_cutoffs.jld2")), cutoffs)
    assert np.array_equal(cb.load_object(str(tmp_path / "bucket_weights.jld2")), weights)
    a = cb.load_object(str(tmp_path / "avg_residual.jld2"))
    assert a.shape == () and a.dtype == np.float32 and a == avg


def test_layout_of_the_written_file_is_what_jld2_documents(tmp_path):
    p = str(tmp_path / "v.jld2")
    W.save_object(p, np.arange(5000, dtype=np.int64))
    raw = open(p, "rb").read()
    assert raw.startswith(b"HDF5-based Julia Data Format, version ")
    assert raw[512:520] == b"\x89HDF\r\n\x1a\n" and raw[520] == 2                  # superblock v2 at 512
    base, _, eof, root = struct.unpack("<QQQQ", raw[524:556])
    assert base == 512 and eof + 512 == len(raw) and raw[512 + root:512 + root + 4] == b"OHDR"
    assert struct.unpack("<I", raw[556:560])[0] == W.lookup3(raw[512:556])
    # lookup3 known answers (Jenkins' own test vectors for hashlittle)
    assert W.lookup3(b"") == 0xDEADBEEF
    assert W.lookup3(b"Four score and seven years ago") == 0x17770551
    assert W.lookup3(b"Four score and seven years ago", 1) == 0xCD628161


def test_reader_errors(tmp_path):
    p = str(tmp_path / "x.jld2")
    with pytest.raises(cb.DimensionMismatch):
        cb.load_object(p)                                                          # missing file
    open(p, "wb").write(b"not a jld2 file" * 100)
    with pytest.raises(cb.DimensionMismatch, match="superblock"):
        cb.load_object(p)
    W.save_object(p, np.arange(10, dtype=np.int64))
    with pytest.raises(cb.DimensionMismatch, match="no dataset named"):
        cb.load_object(p, "something_else")
    W.save_object(p, np.arange(5000, dtype=np.int64))
    raw = open(p, "rb").read()
    open(p, "wb").write(raw[:20000])                                               # truncated contiguous data
    with pytest.raises(cb.DimensionMismatch):
        cb.load_object(p)


def test_continuation_chunks_and_foreign_members(tmp_path):
    """A root group whose links live in an object-header continuation chunk (what JLD2 does when a group grows), with
    other members before ours."""
    a = np.arange(12, dtype=np.float32).reshape(3, 4)
    data = a.tobytes()
    space = struct.pack("<BBBB", 2, 2, 0, 1) + struct.pack("<QQ", 3, 4)
    layout = struct.pack("<BBH", 3, 0, len(data)) + data
    ds = W._object_header(W._msg(0x01, space) + W._msg(0x03, W._datatype(np.float32), flags=1) + W._msg(0x08, layout))
    ds_addr = 48
    name = b"single_stored_object"
    mk = lambda nm, addr: W._msg(0x06, struct.pack("<BB", 1, 0x10) + bytes([1, len(nm)]) + nm + struct.pack("<Q", addr))
    cont_body = b"OCHK" + mk(b"_types", 7) + mk(name, ds_addr)
    cont = cont_body + struct.pack("<I", W.lookup3(cont_body))
    cont_addr = ds_addr + len(ds)
    root = W._object_header(W._msg(0x02, struct.pack("<BBQQ", 0, 0, W.UNDEF, W.UNDEF)) + W._msg(0x0A, b"\0\0") +
                            W._msg(0x10, struct.pack("<QQ", cont_addr, len(cont))))
    root_addr = cont_addr + len(cont)
    sb = b"\x89HDF\r\n\x1a\n" + struct.pack("<BBBB", 2, 8, 8, 0) + struct.pack("<QQQQ", 512, W.UNDEF, root_addr + len(root), root_addr)
    sb += struct.pack("<I", W.lookup3(sb))
    p = str(tmp_path / "c.jld2")
    open(p, "wb").write(b"HDF5-based Julia Data Format, version 0.1.0".ljust(512, b"\0") + sb + ds + cont + root)
    assert np.array_equal(cb.load_object(p), a)
