"""Decode the reference's golden arrays (zarr v2, single chunk, blosc-lz4 + byte shuffle).

TEST INFRASTRUCTURE ONLY.  zarr / numcodecs / blosc are not installed, so the Blosc-1
frame is parsed by hand and the LZ4 *block* streams are decoded with pyarrow's
``lz4_raw`` codec (or system liblz4 through ctypes).

Blosc-1 frame: 16-byte header
    [0] version  [1] versionlz  [2] flags  [3] typesize
    u32 nbytes   u32 blocksize  u32 cbytes
followed by ``i32 bstarts[nblocks]``; each block is ``nsplits`` streams, each prefixed by an
``i32`` compressed size (== raw size => stored verbatim).  flags: bit0 byte-shuffle,
bit1 memcpy'd, bit2 bit-shuffle, bit4 "do not split".
"""
import ctypes
import ctypes.util
import json
import os
import struct

import numpy as np


def _lz4_block_decoder():
    try:
        import pyarrow as pa

        codec = pa.Codec("lz4_raw")

        def dec(src, raw_size):
            return codec.decompress(src, decompressed_size=raw_size).to_pybytes()

        dec(codec.compress(b"x" * 64).to_pybytes(), 64)
        return dec
    except Exception:
        pass
    name = ctypes.util.find_library("lz4") or "liblz4.so.1"
    lib = ctypes.CDLL(name)
    lib.LZ4_decompress_safe.restype = ctypes.c_int
    lib.LZ4_decompress_safe.argtypes = [ctypes.c_char_p, ctypes.c_char_p, ctypes.c_int, ctypes.c_int]

    def dec(src, raw_size):
        out = ctypes.create_string_buffer(raw_size)
        n = lib.LZ4_decompress_safe(bytes(src), out, len(src), raw_size)
        if n != raw_size:
            raise ValueError(f"lz4: got {n} bytes, expected {raw_size}")
        return out.raw

    return dec


def blosc1_decompress(buf):
    version, versionlz, flags, typesize = struct.unpack_from("<BBBB", buf, 0)
    nbytes, blocksize, cbytes = struct.unpack_from("<III", buf, 4)
    if cbytes != len(buf):
        raise ValueError("blosc: truncated frame")
    if flags & 0x2:  # memcpy'd
        return bytes(buf[16:16 + nbytes])
    if flags & 0x4:
        raise NotImplementedError("blosc bit-shuffle")
    codec = (flags >> 5) & 0x7
    if codec != 1:
        raise NotImplementedError(f"blosc codec {codec} (only LZ4 = 1)")
    byte_shuffle = bool(flags & 0x1)
    dont_split = bool(flags & 0x10)
    lz4 = _lz4_block_decoder()
    nblocks = (nbytes + blocksize - 1) // blocksize
    bstarts = struct.unpack_from(f"<{nblocks}i", buf, 16)
    out = bytearray()
    for b in range(nblocks):
        bsize = min(blocksize, nbytes - b * blocksize)
        leftover = bsize != blocksize
        split = (not dont_split) and (not leftover) and typesize <= 16 and bsize // typesize >= 128
        nsplits = typesize if split else 1
        neblock = bsize // nsplits
        pos = bstarts[b]
        block = bytearray()
        for _ in range(nsplits):
            (csize,) = struct.unpack_from("<i", buf, pos)
            pos += 4
            chunk = buf[pos:pos + csize]
            pos += csize
            block += bytes(chunk) if csize == neblock else lz4(bytes(chunk), neblock)
        if byte_shuffle and typesize > 1:
            n = bsize // typesize
            body = np.frombuffer(bytes(block[: n * typesize]), dtype=np.uint8).reshape(typesize, n).T
            block = bytearray(body.tobytes()) + block[n * typesize:]
        out += block
    return bytes(out)


def read_zarr_array(path):
    """Read a single-chunk zarr-v2 array directory written by the reference tests."""
    with open(os.path.join(path, ".zarray")) as fh:
        meta = json.load(fh)
    shape = tuple(meta["shape"])
    if tuple(meta["chunks"]) != shape:
        raise NotImplementedError("multi-chunk zarr")
    chunk = os.path.join(path, ".".join(["0"] * len(shape)))
    with open(chunk, "rb") as fh:
        raw = fh.read()
    comp = meta.get("compressor")
    if comp is not None:
        if comp["id"] != "blosc":
            raise NotImplementedError(comp["id"])
        raw = blosc1_decompress(raw)
    return np.frombuffer(raw, dtype=np.dtype(meta["dtype"])).reshape(shape, order=meta["order"]).copy()
