"""TensorFlow V2 checkpoint ("tensor bundle") reader and writer, without TensorFlow.

The reference restores its weights with ``tf.train.Saver().restore`` and discovers
``id_num`` with ``pywrap_tensorflow.NewCheckpointReader(path).get_variable_to_shape_map()``
(reference ``synthesizer.py:23-25,33-34``); ``eval.py:45-48`` finds the newest
checkpoint through the ``checkpoint`` state file.  TensorFlow 1.x cannot be installed
in this image, so this module restates the on-disk format:

``<prefix>.index``
    a LevelDB-style sorted string table (tensorflow/core/lib/io/table*.cc):
    data blocks of prefix-compressed (shared, non_shared, value_len) entries with a
    restart array, each block followed by a 1-byte compression type and a masked
    CRC32C; an index block mapping separator keys to block handles; a 48-byte
    footer (metaindex handle, index handle, padding, magic 0xdb4775248b80fb57).
    Key ``""`` holds a ``BundleHeaderProto`` (num_shards, endianness, version), every
    other key is a tensor name holding a ``BundleEntryProto`` (dtype, shape,
    shard_id, offset, size, crc32c) -- tensorflow/core/protobuf/tensor_bundle.proto.
``<prefix>.data-SSSSS-of-NNNNN``
    the raw little-endian tensor bytes at (offset, size) of shard ``shard_id``.

:class:`CheckpointReader` mirrors the ``NewCheckpointReader`` surface
(``get_variable_to_shape_map``, ``get_variable_to_dtype_map``, ``has_tensor``,
``get_tensor``).  :func:`write_checkpoint` writes a bundle in the same format (one
shard, uncompressed blocks like ``BundleWriter``) so that weights can be exported for
the reference and so that the reader is testable here.

Parity note: no TensorFlow-written checkpoint exists in this container or in the
reference tree, so the reader is pinned only against this module's own writer and
against the format constants (magic, CRC32C known answer, masked-CRC formula).
"""
from __future__ import annotations

import os
import re
import struct
from typing import Dict, Iterator, List, Optional, Tuple

import numpy as np

TABLE_MAGIC = 0xDB4775248B80FB57
FOOTER_LEN = 48
BLOCK_TRAILER_LEN = 5            # compression type + masked crc32c
NO_COMPRESSION, SNAPPY_COMPRESSION = 0, 1

# tensorflow/core/framework/types.proto
_DTYPES = {1: np.float32, 2: np.float64, 3: np.int32, 4: np.uint8, 5: np.int16, 6: np.int8, 9: np.int64,
           10: np.bool_, 17: np.uint16, 19: np.float16, 22: np.uint32, 23: np.uint64}
_DTYPE_ENUM = {np.dtype(v): k for k, v in _DTYPES.items()}
DT_BFLOAT16 = 14


# ---- crc32c (Castagnoli), masked as in tensorflow/core/lib/hash/crc32c.h -------------------------------------------
def _make_crc_table():
    tab = []
    for i in range(256):
        c = i
        for _ in range(8):
            c = (c >> 1) ^ 0x82F63B78 if c & 1 else c >> 1
        tab.append(c)
    return tab


_CRC_TABLE = _make_crc_table()


def crc32c(data: bytes, crc: int = 0) -> int:
    c = crc ^ 0xFFFFFFFF
    tab = _CRC_TABLE
    for b in data:
        c = tab[(c ^ b) & 0xFF] ^ (c >> 8)
    return c ^ 0xFFFFFFFF


def mask_crc(crc: int) -> int:
    return (((crc >> 15) | (crc << 17)) + 0xA282EAD8) & 0xFFFFFFFF


def unmask_crc(masked: int) -> int:
    rot = (masked - 0xA282EAD8) & 0xFFFFFFFF
    return ((rot >> 17) | (rot << 15)) & 0xFFFFFFFF


# ---- varints and the few protobuf fields the bundle uses ---------------------------------------------------------
def _get_varint(buf: bytes, pos: int) -> Tuple[int, int]:
    shift = result = 0
    while True:
        if pos >= len(buf):
            raise ValueError("truncated varint")
        b = buf[pos]
        pos += 1
        result |= (b & 0x7F) << shift
        if not b & 0x80:
            return result, pos
        shift += 7
        if shift > 63:
            raise ValueError("varint too long")


def _put_varint(v: int) -> bytes:
    if v < 0:
        v += 1 << 64
    out = bytearray()
    while True:
        b = v & 0x7F
        v >>= 7
        if v:
            out.append(b | 0x80)
        else:
            out.append(b)
            return bytes(out)


def _pb_fields(buf: bytes) -> Iterator[Tuple[int, int, object]]:
    """Yield (field number, wire type, value) of one protobuf message."""
    pos = 0
    while pos < len(buf):
        key, pos = _get_varint(buf, pos)
        fno, wt = key >> 3, key & 7
        if wt == 0:
            v, pos = _get_varint(buf, pos)
        elif wt == 1:
            v = buf[pos:pos + 8]
            pos += 8
        elif wt == 2:
            n, pos = _get_varint(buf, pos)
            v = buf[pos:pos + n]
            pos += n
        elif wt == 5:
            v = buf[pos:pos + 4]
            pos += 4
        else:
            raise ValueError("unsupported protobuf wire type %d" % wt)
        yield fno, wt, v


def _signed64(v: int) -> int:
    return v - (1 << 64) if v >= 1 << 63 else v


def _pb_field(fno: int, wt: int, payload: bytes) -> bytes:
    return _put_varint((fno << 3) | wt) + payload


class BundleEntry:
    """BundleEntryProto: dtype=1, shape=2, shard_id=3, offset=4, size=5, crc32c=6 (fixed32), slices=7."""
    __slots__ = ("dtype", "shape", "shard_id", "offset", "size", "crc32c", "sliced")

    def __init__(self):
        self.dtype, self.shape, self.shard_id, self.offset, self.size, self.crc32c, self.sliced = 0, (), 0, 0, 0, 0, False

    @classmethod
    def parse(cls, buf: bytes) -> "BundleEntry":
        e = cls()
        for fno, _, v in _pb_fields(buf):
            if fno == 1:
                e.dtype = v
            elif fno == 2:                      # TensorShapeProto: repeated Dim dim = 2 {int64 size = 1}
                dims = []
                for f2, _, v2 in _pb_fields(v):
                    if f2 == 2:
                        size = 0
                        for f3, _, v3 in _pb_fields(v2):
                            if f3 == 1:
                                size = _signed64(v3)
                        dims.append(size)
                e.shape = tuple(dims)
            elif fno == 3:
                e.shard_id = v
            elif fno == 4:
                e.offset = v
            elif fno == 5:
                e.size = v
            elif fno == 6:
                e.crc32c = struct.unpack("<I", v)[0]
            elif fno == 7:
                e.sliced = True
        return e

    def serialize(self) -> bytes:
        shape = b"".join(_pb_field(2, 2, _put_varint(len(d)) + d)
                         for d in (_pb_field(1, 0, _put_varint(s)) for s in self.shape))
        out = _pb_field(1, 0, _put_varint(self.dtype)) + _pb_field(2, 2, _put_varint(len(shape)) + shape)
        if self.shard_id:
            out += _pb_field(3, 0, _put_varint(self.shard_id))
        if self.offset:
            out += _pb_field(4, 0, _put_varint(self.offset))
        out += _pb_field(5, 0, _put_varint(self.size))
        out += _pb_field(6, 5, struct.pack("<I", self.crc32c))
        return out


# ---- snappy (only needed if a table was written with compressed blocks; BundleWriter does not) --------------------
def _snappy_uncompress(buf: bytes) -> bytes:
    n, pos = _get_varint(buf, 0)
    out = bytearray()
    while pos < len(buf):
        tag = buf[pos]
        pos += 1
        kind = tag & 3
        if kind == 0:
            ln = tag >> 2
            if ln >= 60:
                nb = ln - 59
                ln = int.from_bytes(buf[pos:pos + nb], "little")
                pos += nb
            ln += 1
            out += buf[pos:pos + ln]
            pos += ln
            continue
        if kind == 1:
            ln = ((tag >> 2) & 7) + 4
            off = ((tag >> 5) << 8) | buf[pos]
            pos += 1
        elif kind == 2:
            ln = (tag >> 2) + 1
            off = int.from_bytes(buf[pos:pos + 2], "little")
            pos += 2
        else:
            ln = (tag >> 2) + 1
            off = int.from_bytes(buf[pos:pos + 4], "little")
            pos += 4
        if off == 0 or off > len(out):
            raise ValueError("corrupt snappy block")
        for _ in range(ln):                      # copies may overlap their own output
            out.append(out[-off])
    if len(out) != n:
        raise ValueError("corrupt snappy block: length mismatch")
    return bytes(out)


# ---- sorted string table -----------------------------------------------------------------------------------------
def _read_block(f: bytes, offset: int, size: int, verify: bool) -> bytes:
    if offset + size + BLOCK_TRAILER_LEN > len(f):
        raise ValueError("block handle out of range")
    contents, ctype = f[offset:offset + size], f[offset + size]
    if verify:
        stored = struct.unpack("<I", f[offset + size + 1:offset + size + 5])[0]
        actual = crc32c(f[offset:offset + size + 1])
        if unmask_crc(stored) != actual:
            raise ValueError("block checksum mismatch at offset %d" % offset)
    if ctype == NO_COMPRESSION:
        return contents
    if ctype == SNAPPY_COMPRESSION:
        return _snappy_uncompress(contents)
    raise ValueError("unknown block compression type %d" % ctype)


def _block_entries(block: bytes) -> Iterator[Tuple[bytes, bytes]]:
    if len(block) < 4:
        raise ValueError("block too small")
    num_restarts = struct.unpack("<I", block[-4:])[0]
    limit = len(block) - 4 - 4 * num_restarts
    if limit < 0:
        raise ValueError("bad restart array")
    pos, key = 0, b""
    while pos < limit:
        shared, pos = _get_varint(block, pos)
        non_shared, pos = _get_varint(block, pos)
        vlen, pos = _get_varint(block, pos)
        if shared > len(key):
            raise ValueError("corrupt entry: shared prefix longer than the previous key")
        key = key[:shared] + block[pos:pos + non_shared]
        pos += non_shared
        yield key, block[pos:pos + vlen]
        pos += vlen


def read_table(data: bytes, verify: bool = True) -> List[Tuple[bytes, bytes]]:
    """All (key, value) pairs of a table file, in key order."""
    if len(data) < FOOTER_LEN:
        raise ValueError("not a table: file shorter than the footer")
    footer = data[-FOOTER_LEN:]
    if struct.unpack("<Q", footer[-8:])[0] != TABLE_MAGIC:
        raise ValueError("not a table: bad magic number")
    _, p = _get_varint(footer, 0)          # metaindex handle (unused)
    _, p = _get_varint(footer, p)
    ioff, p = _get_varint(footer, p)
    isize, p = _get_varint(footer, p)
    out = []
    for _, handle in _block_entries(_read_block(data, ioff, isize, verify)):
        boff, q = _get_varint(handle, 0)
        bsize, q = _get_varint(handle, q)
        out.extend(_block_entries(_read_block(data, boff, bsize, verify)))
    return out


class _BlockBuilder:
    def __init__(self, restart_interval: int = 16):
        self.buf = bytearray()
        self.restarts = [0]
        self.counter = 0
        self.last_key = b""
        self.restart_interval = restart_interval

    def add(self, key: bytes, value: bytes):
        shared = 0
        if self.counter < self.restart_interval:
            m = min(len(key), len(self.last_key))
            while shared < m and key[shared] == self.last_key[shared]:
                shared += 1
        else:
            self.restarts.append(len(self.buf))
            self.counter = 0
        self.buf += _put_varint(shared) + _put_varint(len(key) - shared) + _put_varint(len(value))
        self.buf += key[shared:] + value
        self.last_key = key
        self.counter += 1

    def empty(self) -> bool:
        return not self.buf

    def size(self) -> int:
        return len(self.buf) + 4 * len(self.restarts) + 4

    def finish(self) -> bytes:
        return bytes(self.buf) + b"".join(struct.pack("<I", r) for r in self.restarts) + struct.pack("<I", len(self.restarts))


def build_table(items: List[Tuple[bytes, bytes]], block_size: int = 4096) -> bytes:
    """Serialize sorted (key, value) pairs as a table file (uncompressed blocks)."""
    out = bytearray()
    index = _BlockBuilder(restart_interval=1)

    def write_block(contents: bytes) -> bytes:
        off = len(out)
        out.extend(contents)
        out.append(NO_COMPRESSION)
        out.extend(struct.pack("<I", mask_crc(crc32c(contents + bytes([NO_COMPRESSION])))))
        return _put_varint(off) + _put_varint(len(contents))

    blk = _BlockBuilder()
    prev = None
    for key, value in items:
        if prev is not None and key <= prev:
            raise ValueError("table keys must be strictly increasing")
        blk.add(key, value)
        prev = key
        if blk.size() >= block_size:
            index.add(key, write_block(blk.finish()))       # separator = the block's last key
            blk = _BlockBuilder()
    if not blk.empty():
        index.add(prev, write_block(blk.finish()))
    meta_handle = write_block(_BlockBuilder().finish())
    index_handle = write_block(index.finish())
    footer = meta_handle + index_handle
    footer += b"\0" * (FOOTER_LEN - 8 - len(footer)) + struct.pack("<Q", TABLE_MAGIC)
    out.extend(footer)
    return bytes(out)


# ---- bundle -------------------------------------------------------------------------------------------------------
def _data_path(prefix: str, shard: int, num_shards: int) -> str:
    return "%s.data-%05d-of-%05d" % (prefix, shard, num_shards)


class CheckpointReader:
    """``pywrap_tensorflow.NewCheckpointReader`` without TensorFlow (reference ``synthesizer.py:23``)."""

    def __init__(self, prefix: str, verify_index: bool = True):
        self.prefix = prefix
        index_path = prefix + ".index"
        if not os.path.exists(index_path):
            raise FileNotFoundError("no V2 checkpoint at %r (%s is missing)" % (prefix, index_path))
        with open(index_path, "rb") as f:
            items = read_table(f.read(), verify=verify_index)
        self.num_shards = 1
        self._entries: Dict[str, BundleEntry] = {}
        for key, value in items:
            if key == b"":
                for fno, _, v in _pb_fields(value):
                    if fno == 1:
                        self.num_shards = v
                    elif fno == 2 and v != 0:
                        raise ValueError("big-endian bundles are not supported")
                continue
            self._entries[key.decode("utf-8")] = BundleEntry.parse(value)

    def get_variable_to_shape_map(self) -> Dict[str, List[int]]:
        return {k: list(e.shape) for k, e in self._entries.items()}

    def get_variable_to_dtype_map(self) -> Dict[str, object]:
        return {k: ("bfloat16" if e.dtype == DT_BFLOAT16 else np.dtype(_DTYPES[e.dtype]).name if e.dtype in _DTYPES else e.dtype)
                for k, e in self._entries.items()}

    def has_tensor(self, name: str) -> bool:
        return name in self._entries

    def get_tensor(self, name: str, verify: bool = False) -> np.ndarray:
        if name not in self._entries:
            raise KeyError("tensor %r not found in checkpoint %s" % (name, self.prefix))
        e = self._entries[name]
        if e.sliced:
            raise NotImplementedError("partitioned variable %r (tensor slices) is not supported" % name)
        with open(_data_path(self.prefix, e.shard_id, self.num_shards), "rb") as f:
            f.seek(e.offset)
            raw = f.read(e.size)
        if len(raw) != e.size:
            raise ValueError("tensor %r: data shard is truncated" % name)
        if verify and unmask_crc(e.crc32c) != crc32c(raw):
            raise ValueError("tensor %r: checksum mismatch" % name)
        if e.dtype == DT_BFLOAT16:
            a = (np.frombuffer(raw, dtype="<u2").astype(np.uint32) << 16).view(np.float32)
        elif e.dtype in _DTYPES:
            a = np.frombuffer(raw, dtype=np.dtype(_DTYPES[e.dtype]).newbyteorder("<"))
        else:
            raise NotImplementedError("tensor %r: dtype enum %d is not supported" % (name, e.dtype))
        n = int(np.prod(e.shape, dtype=np.int64)) if e.shape else 1
        if a.size != n:
            raise ValueError("tensor %r: %d elements on disk, shape %s" % (name, a.size, e.shape))
        return a.reshape(e.shape).copy()

    def tensors(self, skip_slots: bool = True) -> Dict[str, np.ndarray]:
        """Every (non-optimizer) tensor of the checkpoint as a dict keyed by variable name."""
        out = {}
        for name, e in self._entries.items():
            if skip_slots and (name.endswith("/Adam") or name.endswith("/Adam_1") or e.sliced):
                continue
            if e.dtype not in _DTYPES and e.dtype != DT_BFLOAT16:
                continue
            out[name] = self.get_tensor(name)
        return out


def write_checkpoint(prefix: str, tensors: Dict[str, np.ndarray], block_size: int = 4096) -> None:
    """Write ``tensors`` as a one-shard V2 checkpoint at ``prefix`` (what ``tf.train.Saver().save`` produces for
    unpartitioned variables); also usable to hand this package's weights to the reference."""
    names = sorted(tensors, key=lambda s: s.encode("utf-8"))
    if "" in tensors:
        raise ValueError("the empty name is reserved for the bundle header")
    os.makedirs(os.path.dirname(os.path.abspath(prefix)), exist_ok=True)
    items = []
    header = _pb_field(1, 0, _put_varint(1)) + _pb_field(3, 2, _put_varint(2) + _pb_field(1, 0, _put_varint(1)))
    items.append((b"", header))                 # num_shards = 1, endianness LITTLE (default), version.producer = 1
    offset = 0
    with open(_data_path(prefix, 0, 1), "wb") as f:
        for name in names:
            a = np.asarray(tensors[name])          # (ascontiguousarray would turn a scalar into shape (1,))
            dt = a.dtype.newbyteorder("=") if a.dtype.byteorder not in ("=", "|", "<") else a.dtype
            if np.dtype(dt) not in _DTYPE_ENUM:
                raise TypeError("tensor %r: dtype %s cannot be stored" % (name, a.dtype))
            raw = a.astype(np.dtype(dt).newbyteorder("<"), copy=False).tobytes()
            e = BundleEntry()
            e.dtype, e.shape, e.offset, e.size = _DTYPE_ENUM[np.dtype(dt)], tuple(int(s) for s in a.shape), offset, len(raw)
            e.crc32c = mask_crc(crc32c(raw))       # byte-serial in pure Python: a few seconds for the 35 MB of this model
            f.write(raw)
            offset += len(raw)
            items.append((name.encode("utf-8"), e.serialize()))
    with open(prefix + ".index", "wb") as f:
        f.write(build_table(items, block_size))


# ---- checkpoint state file (tf.train.get_checkpoint_state / latest_checkpoint) --------------------------------------
def latest_checkpoint(checkpoint_dir: str) -> Optional[str]:
    """``model_checkpoint_path`` of ``<dir>/checkpoint`` (reference ``eval.py:45-48``), or None."""
    state = os.path.join(checkpoint_dir, "checkpoint")
    if not os.path.exists(state):
        return None
    with open(state, "r") as f:
        for line in f:
            m = re.match(r'\s*model_checkpoint_path:\s*"(.*)"\s*$', line)
            if m:
                p = m.group(1)
                return p if os.path.isabs(p) else os.path.join(checkpoint_dir, p)
    return None


def load_weights(path: str) -> Dict[str, np.ndarray]:
    """Weights by TF variable name from ``path``: a V2 checkpoint prefix (``model.ckpt-1000``), a log directory with
    a ``checkpoint`` state file, or an ``.npz`` archive keyed the same way."""
    if os.path.isdir(path):
        latest = latest_checkpoint(path)
        if latest is None:
            raise FileNotFoundError("no checkpoint state file in %r" % path)
        path = latest
    if path.endswith(".npz"):
        with np.load(path) as z:
            return {k: z[k] for k in z.files}
    if path.endswith(".index"):
        path = path[:-len(".index")]
    return CheckpointReader(path).tensors()
