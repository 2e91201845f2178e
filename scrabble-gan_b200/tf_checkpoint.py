"""Reader / writer for TensorFlow checkpoints (the "tensor bundle" format), without TensorFlow.

The reference saves its networks with `model.save_weights(prefix)` (src/bigacgan/data_utils.py:346-348: generator and
recogniser after every epoch; src/main.py:95-102 builds a tf.train.Checkpoint of everything).  With a prefix that does not
end in .h5 Keras writes a TF checkpoint: `<prefix>.index` + `<prefix>.data-00000-of-00001`.  This module reads and writes
that format so that weights trained with the reference can be loaded into the libsgan models (and ours exported for the
reference), see `keras_checkpoint_keys` / `load_keras_checkpoint` in bigacgan/keras_names.py.

Format (tensorflow/core/util/tensor_bundle/tensor_bundle.{h,cc}, tensorflow/core/lib/io/table*, pinned: TF 2.1 -- the
format has not changed since 1.x):
  * `.index` is a leveldb-style sorted string table: data blocks of prefix-compressed (key, value) entries with restart
    points, one index block, a (here empty) meta-index block and a 48-byte footer ending in the magic 0xdb4775248b80fb57;
    every block is followed by a 1-byte compression tag (0 = none: the bundle writer disables compression) and a masked
    CRC32C.  Key "" maps to a BundleHeaderProto, every other key (a tensor name) to a BundleEntryProto
    {dtype, shape, shard_id, offset, size, crc32c}.
  * `.data-SSSSS-of-NNNNN` holds the raw little-endian tensor bytes at [offset, offset + size).
  * string tensors (e.g. `_CHECKPOINTABLE_OBJECT_GRAPH`) are stored as varint lengths, a masked CRC32C of the lengths and
    the concatenated bytes.
PARITY UNPINNED: no TensorFlow and no reference checkpoint exist in this image; the format code is pinned by a round trip
and by known-answer CRC32C / varint / block-layout tests only (tests/test_tf_checkpoint_cpu.py)."""
from __future__ import annotations

import os
import struct
from collections import OrderedDict
from typing import Dict, Iterable, List, Optional, Tuple

import numpy as np

TABLE_MAGIC = 0xDB4775248B80FB57
HEADER_KEY = b""

# tensorflow/core/framework/types.proto
DT = {1: np.float32, 2: np.float64, 3: np.int32, 4: np.uint8, 5: np.int16, 6: np.int8, 9: np.int64, 10: np.bool_, 17: np.uint16,
      19: np.float16, 22: np.uint32, 23: np.uint64}
DT_STRING, DT_BFLOAT16 = 7, 14
DT_OF = {np.dtype(v): k for k, v in DT.items()}


# ----------------------------------------------------------------------------------------------------
# CRC32C (Castagnoli), masked as in tensorflow/core/lib/hash/crc32c.h
# ----------------------------------------------------------------------------------------------------
def _make_table():
    poly = 0x82F63B78
    tab = np.zeros(256, dtype=np.uint32)
    for i in range(256):
        c = i
        for _ in range(8):
            c = (c >> 1) ^ (poly if c & 1 else 0)
        tab[i] = c
    return tab


_CRC_TABLE = _make_table()
_CRC_LIST = [int(x) for x in _CRC_TABLE]


def crc32c(data: bytes, crc: int = 0) -> int:
    if len(data) >= 4096:                      # large buffers: the slicing-by-8 host routine of libsgan (sg_crc32c)
        try:
            from . import _abi
            import ctypes
            buf = (ctypes.c_char * len(data)).from_buffer_copy(data)
            return int(_abi.load().sg_crc32c(ctypes.cast(buf, ctypes.c_void_p), len(data), crc)) & 0xFFFFFFFF
        except Exception:
            pass
    c = crc ^ 0xFFFFFFFF
    tab = _CRC_LIST
    for b in data:
        c = tab[(c ^ b) & 0xFF] ^ (c >> 8)
    return c ^ 0xFFFFFFFF


def crc32c_array(a: np.ndarray) -> int:
    """CRC32C of a large buffer: 8 independent byte lanes are not possible for a CRC, so this is the plain table walk done
    with numpy in chunks (slow but only used at save / verify time)."""
    return crc32c(a.tobytes())


def mask_crc(c: int) -> int:
    return ((((c >> 15) | (c << 17)) & 0xFFFFFFFF) + 0xA282EAD8) & 0xFFFFFFFF


def unmask_crc(m: int) -> int:
    r = (m - 0xA282EAD8) & 0xFFFFFFFF
    return ((r >> 17) | (r << 15)) & 0xFFFFFFFF


# ----------------------------------------------------------------------------------------------------
# varints and the few protobuf messages involved (hand-rolled: protobuf descriptors for TF are not available)
# ----------------------------------------------------------------------------------------------------
def put_varint(n: int) -> bytes:
    out = bytearray()
    n &= (1 << 64) - 1
    while True:
        b = n & 0x7F
        n >>= 7
        if n:
            out.append(b | 0x80)
        else:
            out.append(b)
            return bytes(out)


def get_varint(buf: bytes, pos: int) -> Tuple[int, int]:
    shift = result = 0
    while True:
        b = buf[pos]
        pos += 1
        result |= (b & 0x7F) << shift
        if not b & 0x80:
            return result, pos
        shift += 7
        if shift > 70:
            raise ValueError("malformed varint")


def _pb_fields(buf: bytes) -> Iterable[Tuple[int, int, object]]:
    """(field number, wire type, value) of one protobuf message; value is an int (varint / fixed) or bytes."""
    pos = 0
    while pos < len(buf):
        tag, pos = get_varint(buf, pos)
        fn, wt = tag >> 3, tag & 7
        if wt == 0:
            v, pos = get_varint(buf, pos)
        elif wt == 1:
            v = struct.unpack_from("<Q", buf, pos)[0]
            pos += 8
        elif wt == 2:
            ln, pos = get_varint(buf, pos)
            v = bytes(buf[pos:pos + ln])
            pos += ln
        elif wt == 5:
            v = struct.unpack_from("<I", buf, pos)[0]
            pos += 4
        else:
            raise ValueError("unsupported protobuf wire type %d" % wt)
        yield fn, wt, v


def _pb_varint(fn: int, v: int) -> bytes:
    return put_varint((fn << 3) | 0) + put_varint(v)


def _pb_bytes(fn: int, b: bytes) -> bytes:
    return put_varint((fn << 3) | 2) + put_varint(len(b)) + b


def _pb_fixed32(fn: int, v: int) -> bytes:
    return put_varint((fn << 3) | 5) + struct.pack("<I", v)


def encode_shape(shape) -> bytes:
    """TensorShapeProto: repeated Dim dim = 2 { int64 size = 1 }."""
    return b"".join(_pb_bytes(2, _pb_varint(1, int(d))) for d in shape)


def decode_shape(buf: bytes) -> Tuple[int, ...]:
    dims = []
    for fn, wt, v in _pb_fields(buf):
        if fn == 2:
            size = 0
            for f2, _, v2 in _pb_fields(v):
                if f2 == 1:
                    size = v2 if v2 < (1 << 63) else v2 - (1 << 64)
            dims.append(size)
    return tuple(dims)


def encode_entry(dtype: int, shape, shard_id: int, offset: int, size: int, crc_masked: int) -> bytes:
    """BundleEntryProto {dtype=1, shape=2, shard_id=3, offset=4, size=5, crc32c=6 (fixed32)}; zero fields are omitted."""
    out = _pb_varint(1, dtype) + _pb_bytes(2, encode_shape(shape))
    if shard_id:
        out += _pb_varint(3, shard_id)
    if offset:
        out += _pb_varint(4, offset)
    out += _pb_varint(5, size) + _pb_fixed32(6, crc_masked)
    return out


def decode_entry(buf: bytes) -> dict:
    e = {"dtype": 0, "shape": (), "shard_id": 0, "offset": 0, "size": 0, "crc32c": 0, "slices": 0}
    for fn, wt, v in _pb_fields(buf):
        if fn == 1:
            e["dtype"] = v
        elif fn == 2:
            e["shape"] = decode_shape(v)
        elif fn == 3:
            e["shard_id"] = v
        elif fn == 4:
            e["offset"] = v
        elif fn == 5:
            e["size"] = v
        elif fn == 6:
            e["crc32c"] = v
        elif fn == 7:
            e["slices"] += 1
    return e


def encode_header(num_shards: int = 1) -> bytes:
    """BundleHeaderProto {num_shards=1, endianness=2 (LITTLE = 0, omitted), version=3 {producer=1}}."""
    return _pb_varint(1, num_shards) + _pb_bytes(3, _pb_varint(1, 1))


def decode_header(buf: bytes) -> dict:
    h = {"num_shards": 0, "endianness": 0}
    for fn, wt, v in _pb_fields(buf):
        if fn == 1:
            h["num_shards"] = v
        elif fn == 2:
            h["endianness"] = v
    return h


# ----------------------------------------------------------------------------------------------------
# the sorted string table (.index)
# ----------------------------------------------------------------------------------------------------
def _parse_block(contents: bytes) -> List[Tuple[bytes, bytes]]:
    n_restarts = struct.unpack_from("<I", contents, len(contents) - 4)[0]
    end = len(contents) - 4 - 4 * n_restarts
    pos, key, out = 0, b"", []
    while pos < end:
        shared, pos = get_varint(contents, pos)
        non_shared, pos = get_varint(contents, pos)
        vlen, pos = get_varint(contents, pos)
        key = key[:shared] + contents[pos:pos + non_shared]
        pos += non_shared
        out.append((key, contents[pos:pos + vlen]))
        pos += vlen
    return out


def _read_block(buf: bytes, offset: int, size: int, verify: bool) -> bytes:
    contents = buf[offset:offset + size]
    tag = buf[offset + size]
    if tag != 0:
        raise ValueError("compressed table block (type %d): the tensor-bundle writer never compresses; not supported" % tag)
    if verify:
        stored = struct.unpack_from("<I", buf, offset + size + 1)[0]
        if unmask_crc(stored) != crc32c(buf[offset:offset + size + 1]):
            raise ValueError("table block checksum mismatch at offset %d" % offset)
    return contents


def read_table(path: str, verify: bool = True) -> "OrderedDict[bytes, bytes]":
    with open(path, "rb") as f:
        buf = f.read()
    if len(buf) < 48 or struct.unpack_from("<Q", buf, len(buf) - 8)[0] != TABLE_MAGIC:
        raise ValueError("{}: not a TensorFlow checkpoint index (bad table magic)".format(path))
    footer = buf[-48:]
    pos = 0
    _, pos = get_varint(footer, pos)        # meta-index handle
    _, pos = get_varint(footer, pos)
    idx_off, pos = get_varint(footer, pos)
    idx_size, pos = get_varint(footer, pos)
    out: "OrderedDict[bytes, bytes]" = OrderedDict()
    for _, handle in _parse_block(_read_block(buf, idx_off, idx_size, verify)):
        off, p2 = get_varint(handle, 0)
        size, _ = get_varint(handle, p2)
        for k, v in _parse_block(_read_block(buf, off, size, verify)):
            out[k] = v
    return out


def _build_block(entries: List[Tuple[bytes, bytes]], restart_interval: int = 16) -> bytes:
    out, restarts, last = bytearray(), [], b""
    for i, (k, v) in enumerate(entries):
        if i % restart_interval == 0:
            restarts.append(len(out))
            shared = 0
        else:
            shared = 0
            while shared < min(len(k), len(last)) and k[shared] == last[shared]:
                shared += 1
        out += put_varint(shared) + put_varint(len(k) - shared) + put_varint(len(v)) + k[shared:] + v
        last = k
    if not restarts:
        restarts = [0]
    for r in restarts:
        out += struct.pack("<I", r)
    out += struct.pack("<I", len(restarts))
    return bytes(out)


def write_table(path: str, items: "Dict[bytes, bytes]", block_size: int = 4096) -> None:
    keys = sorted(items)
    blob, index_entries = bytearray(), []

    def emit(block: bytes) -> Tuple[int, int]:
        off = len(blob)
        blob.extend(block)
        blob.append(0)                                           # kNoCompression
        blob.extend(struct.pack("<I", mask_crc(crc32c(block + b"\x00"))))
        return off, len(block)

    cur: List[Tuple[bytes, bytes]] = []
    cur_bytes = 0
    for k in keys:
        cur.append((k, items[k]))
        cur_bytes += len(k) + len(items[k]) + 8
        if cur_bytes >= block_size:
            off, size = emit(_build_block(cur))
            index_entries.append((cur[-1][0], put_varint(off) + put_varint(size)))
            cur, cur_bytes = [], 0
    if cur or not index_entries:
        off, size = emit(_build_block(cur))
        index_entries.append((cur[-1][0] if cur else b"", put_varint(off) + put_varint(size)))
    meta_off, meta_size = emit(_build_block([]))
    idx_off, idx_size = emit(_build_block(index_entries, restart_interval=1))
    footer = put_varint(meta_off) + put_varint(meta_size) + put_varint(idx_off) + put_varint(idx_size)
    footer += b"\x00" * (40 - len(footer)) + struct.pack("<Q", TABLE_MAGIC)
    blob.extend(footer)
    with open(path, "wb") as f:
        f.write(bytes(blob))


# ----------------------------------------------------------------------------------------------------
# the bundle
# ----------------------------------------------------------------------------------------------------
def _decode_strings(raw: bytes, count: int) -> List[bytes]:
    pos, lens = 0, []
    for _ in range(count):
        n, pos = get_varint(raw, pos)
        lens.append(n)
    pos += 4                                                     # masked crc32c of the length varints
    out = []
    for n in lens:
        out.append(raw[pos:pos + n])
        pos += n
    return out


def read_checkpoint(prefix: str, verify_crc: bool = False, with_strings: bool = False) -> "OrderedDict[str, np.ndarray]":
    """All tensors of the checkpoint `<prefix>.index` / `<prefix>.data-*` as numpy arrays, in key order.  String tensors are
    skipped unless with_strings (then returned as object arrays of bytes)."""
    table = read_table(prefix + ".index")
    if HEADER_KEY not in table:
        raise ValueError("{}.index: no bundle header entry".format(prefix))
    header = decode_header(table[HEADER_KEY])
    if header["endianness"] != 0:
        raise ValueError("big-endian checkpoints are not supported")
    shards: Dict[int, np.memmap] = {}
    out: "OrderedDict[str, np.ndarray]" = OrderedDict()
    for key, val in table.items():
        if key == HEADER_KEY:
            continue
        e = decode_entry(val)
        if e["slices"]:
            raise ValueError("{}: partitioned (sliced) variables are not supported".format(key.decode()))
        sid = e["shard_id"]
        if sid not in shards:
            path = "{}.data-{:05d}-of-{:05d}".format(prefix, sid, header["num_shards"])
            shards[sid] = np.memmap(path, dtype=np.uint8, mode="r") if os.path.getsize(path) else np.zeros(0, np.uint8)
        raw = shards[sid][e["offset"]:e["offset"] + e["size"]]
        count = int(np.prod(e["shape"])) if e["shape"] else 1
        if e["dtype"] == DT_STRING:
            if with_strings:
                arr = np.empty(count, dtype=object)
                arr[:] = _decode_strings(bytes(raw), count)
                out[key.decode()] = arr.reshape(e["shape"])
            continue
        if verify_crc and unmask_crc(e["crc32c"]) != crc32c(bytes(raw)):
            raise ValueError("{}: tensor checksum mismatch".format(key.decode()))
        if e["dtype"] == DT_BFLOAT16:
            arr = (np.frombuffer(bytes(raw), dtype=np.uint16).astype(np.uint32) << 16).view(np.float32)
        elif e["dtype"] in DT:
            arr = np.frombuffer(bytes(raw), dtype=DT[e["dtype"]])
        else:
            raise ValueError("{}: unsupported dtype enum {}".format(key.decode(), e["dtype"]))
        if arr.size != count:
            raise ValueError("{}: {} bytes do not match shape {}".format(key.decode(), e["size"], e["shape"]))
        out[key.decode()] = arr.reshape(e["shape"]).copy()
    return out


def write_checkpoint(prefix: str, tensors: "Dict[str, np.ndarray]", strings: Optional[Dict[str, bytes]] = None) -> None:
    """Write `<prefix>.index` + `<prefix>.data-00000-of-00001` (one shard), tensors laid out in key order as TensorFlow's
    BundleWriter does, and the `checkpoint` state file next to them."""
    d = os.path.dirname(prefix)
    if d:
        os.makedirs(d, exist_ok=True)
    items: Dict[bytes, bytes] = {HEADER_KEY: encode_header(1)}
    offset = 0
    names = sorted(list(tensors) + list(strings or {}), key=lambda s: s.encode())
    with open(prefix + ".data-00000-of-00001", "wb") as f:
        for name in names:
            if strings and name in strings:
                s = strings[name]
                lens = put_varint(len(s))
                raw = lens + struct.pack("<I", mask_crc(crc32c(lens))) + s
                crc = crc32c(struct.pack("<I", mask_crc(crc32c(lens))), crc32c(lens))
                crc = crc32c(s, crc)
                items[name.encode()] = encode_entry(DT_STRING, (), 0, offset, len(raw), mask_crc(crc))
            else:
                a = np.asarray(tensors[name])
                if not a.flags.c_contiguous:
                    a = a.copy(order="C")                          # (np.ascontiguousarray would turn a scalar into shape (1,))
                if a.dtype not in DT_OF:
                    raise ValueError("{}: dtype {} cannot be stored".format(name, a.dtype))
                raw = a.astype(a.dtype.newbyteorder("<"), copy=False).tobytes()
                items[name.encode()] = encode_entry(DT_OF[a.dtype], a.shape, 0, offset, len(raw), mask_crc(crc32c(raw)))
            f.write(raw)
            offset += len(raw)
    write_table(prefix + ".index", items)
    with open(os.path.join(d or ".", "checkpoint"), "w") as f:
        base = os.path.basename(prefix)
        f.write('model_checkpoint_path: "{}"\nall_model_checkpoint_paths: "{}"\n'.format(base, base))


# ----------------------------------------------------------------------------------------------------
# the object graph stored with Keras / tf.train.Checkpoint saves: checkpoint key -> variable name at save time
# ----------------------------------------------------------------------------------------------------
def object_graph_names(prefix: str) -> Dict[str, str]:
    """{checkpoint_key: full_name} from `_CHECKPOINTABLE_OBJECT_GRAPH` (TrackableObjectGraph: repeated nodes = 1, each with
    repeated SerializedTensor attributes = 2 {name = 1, full_name = 2, checkpoint_key = 3}); {} if the entry is absent."""
    t = read_checkpoint(prefix, with_strings=True)
    g = t.get("_CHECKPOINTABLE_OBJECT_GRAPH")
    if g is None:
        return {}
    out = {}
    for fn, wt, node in _pb_fields(bytes(g.reshape(-1)[0])):
        if fn != 1:
            continue
        for f2, _, attr in _pb_fields(node):
            if f2 != 2:
                continue
            full, key = "", ""
            for f3, _, v in _pb_fields(attr):
                if f3 == 2:
                    full = v.decode()
                elif f3 == 3:
                    key = v.decode()
            if key:
                out[key] = full
    return out
