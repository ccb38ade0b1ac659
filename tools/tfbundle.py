"""Readers for the TensorFlow artefact formats the reference ships (no TensorFlow needed).

* ``saved_model.pb``  -- SavedModel proto; MetaGraphDefs are parsed with the protobuf classes
  that ship inside tensorboard (``tensorboard.compat.proto``).
* ``variables/variables.index`` -- tensor-bundle index = LevelDB table of BundleEntryProto.
* masked CRC32C as stored in BundleEntryProto.crc32c.

Formats are described in SURVEY.md Appendix A.  This module is a build-time tool: it is used by
``tools/extract_assets.py`` (run where /root/reference exists) and by the weight loader to verify a
user-supplied ``variables.data-00000-of-00001`` against the CRCs recorded in the committed table.
"""
from __future__ import annotations

import struct
from dataclasses import dataclass

import numpy as np

# ----------------------------------------------------------------------------- varint / proto wire

def _varint(buf: bytes, pos: int):
    out = 0
    shift = 0
    while True:
        b = buf[pos]
        pos += 1
        out |= (b & 0x7F) << shift
        if not b & 0x80:
            return out, pos
        shift += 7


def iter_fields(buf: bytes):
    """Yield (field_number, wire_type, value) over a serialized protobuf message."""
    pos = 0
    n = len(buf)
    while pos < n:
        key, pos = _varint(buf, pos)
        fno, wt = key >> 3, key & 7
        if wt == 0:
            v, pos = _varint(buf, pos)
        elif wt == 1:
            v = buf[pos:pos + 8]
            pos += 8
        elif wt == 2:
            ln, pos = _varint(buf, pos)
            v = buf[pos:pos + ln]
            pos += ln
        elif wt == 5:
            v = buf[pos:pos + 4]
            pos += 4
        else:
            raise ValueError(f"unsupported wire type {wt}")
        yield fno, wt, v


# ----------------------------------------------------------------------------- crc32c (Castagnoli)

_CRC_TABLE = None


def _crc_table():
    global _CRC_TABLE
    if _CRC_TABLE is None:
        poly = 0x82F63B78
        tbl = np.zeros(256, dtype=np.uint32)
        for i in range(256):
            c = i
            for _ in range(8):
                c = (c >> 1) ^ poly if c & 1 else c >> 1
            tbl[i] = c
        _CRC_TABLE = tbl
    return _CRC_TABLE


def crc32c(data: bytes) -> int:
    """Plain CRC32C. Slicing-by-1 over numpy-table; fast enough for 13 MB (a few seconds)."""
    tbl = _crc_table().tolist()
    c = 0xFFFFFFFF
    for b in data:
        c = tbl[(c ^ b) & 0xFF] ^ (c >> 8)
    return c ^ 0xFFFFFFFF


def masked_crc32c(data: bytes) -> int:
    c = crc32c(data)
    return (((c >> 15) | (c << 17)) + 0xA282EAD8) & 0xFFFFFFFF


# ----------------------------------------------------------------------------- tensor bundle index

@dataclass
class BundleEntry:
    name: str
    dtype: int          # 1 = float32, 9 = int64, 7 = string
    shape: tuple
    shard: int
    offset: int
    size: int
    crc32c: int


def _block_handle(buf, pos):
    off, pos = _varint(buf, pos)
    sz, pos = _varint(buf, pos)
    return off, sz, pos


def _iter_block(data: bytes):
    """Iterate (key, value) over one LevelDB table block (uncompressed)."""
    n_restarts = struct.unpack_from("<I", data, len(data) - 4)[0]
    limit = len(data) - 4 - 4 * n_restarts
    pos = 0
    key = b""
    while pos < limit:
        shared, pos = _varint(data, pos)
        non_shared, pos = _varint(data, pos)
        vlen, pos = _varint(data, pos)
        key = key[:shared] + data[pos:pos + non_shared]
        pos += non_shared
        val = data[pos:pos + vlen]
        pos += vlen
        yield key, val


def _parse_entry(name: str, val: bytes) -> BundleEntry:
    dtype = 0
    shape = []
    shard = offset = size = crc = 0
    for fno, wt, v in iter_fields(val):
        if fno == 1:
            dtype = v
        elif fno == 2:
            for f2, _, v2 in iter_fields(v):
                if f2 == 2:
                    dim = 0
                    for f3, _, v3 in iter_fields(v2):
                        if f3 == 1:
                            dim = v3
                    shape.append(dim)
        elif fno == 3:
            shard = v
        elif fno == 4:
            offset = v
        elif fno == 5:
            size = v
        elif fno == 6:
            crc = struct.unpack("<I", v)[0]
    return BundleEntry(name, dtype, tuple(shape), shard, offset, size, crc)


def read_bundle_index(path: str) -> list[BundleEntry]:
    buf = open(path, "rb").read()
    footer = buf[-48:]
    assert footer[-8:] == struct.pack("<Q", 0xDB4775248B80FB57), "not a leveldb table"
    pos = 0
    _, _, pos = _block_handle(footer, pos)          # metaindex
    ioff, isz, pos = _block_handle(footer, pos)     # index
    entries = []
    for _, handle in _iter_block(buf[ioff:ioff + isz]):
        doff, dsz, _ = _block_handle(handle, 0)
        assert buf[doff + dsz] == 0, "compressed block"
        for key, val in _iter_block(buf[doff:doff + dsz]):
            if key == b"":
                continue                             # BundleHeaderProto
            entries.append(_parse_entry(key.decode(), val))
    return entries


# ----------------------------------------------------------------------------- saved_model.pb

def read_meta_graphs(path: str):
    from tensorboard.compat.proto import meta_graph_pb2
    buf = open(path, "rb").read()
    out = []
    for fno, wt, v in iter_fields(buf):
        if fno == 2 and wt == 2:
            mg = meta_graph_pb2.MetaGraphDef()
            mg.ParseFromString(v)
            out.append(mg)
    return out


def const_nodes(mg) -> dict:
    """Top-level Const nodes of a MetaGraphDef as numpy arrays."""
    from tensorboard.compat.proto import types_pb2  # noqa: F401
    out = {}
    np_of = {1: np.float32, 3: np.int32, 9: np.int64, 2: np.float64}
    for node in mg.graph_def.node:
        if node.op != "Const":
            continue
        t = node.attr["value"].tensor
        if t.dtype not in np_of:
            continue
        shape = tuple(d.size for d in t.tensor_shape.dim)
        dt = np_of[t.dtype]
        if t.tensor_content:
            arr = np.frombuffer(t.tensor_content, dtype=dt).reshape(shape)
        else:
            vals = {1: t.float_val, 3: t.int_val, 9: t.int64_val, 2: t.double_val}[t.dtype]
            arr = np.array(list(vals), dtype=dt)
            if shape and arr.size == 1:
                arr = np.full(shape, arr[0], dtype=dt)
            else:
                arr = arr.reshape(shape)
        out[node.name] = arr
    return out
