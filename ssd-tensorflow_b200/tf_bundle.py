"""Reader / writer of TensorFlow's checkpoint-V2 "tensor bundle" (``<prefix>.index`` + ``<prefix>.data-00000-of-00001``)
without TensorFlow -- SURVEY.md 8f row 2.

Why it is here: the reference starts from a pretrained VGG-16 *saved-model* (``<vgg_dir>/vgg/variables/variables.*``,
ssdvgg.py:157-161,190-207) and saves / restores its own checkpoints with ``tf.train.Saver`` (train.py:208,336-343,
infer.py:111-126).  Both are tensor bundles.  This module lets the B200 engine load those weights and write checkpoints
under the reference's variable names with no TensorFlow installed.

Format (restated from the published TensorFlow / LevelDB sources; TensorFlow is not installable here, so the module is
verified by round trips and by hand-built byte-level cases only -- "parity unpinned" in DESIGN.md):
  * ``.index`` is a LevelDB-format sorted table (tensorflow/core/lib/io/table*): data blocks of prefix-compressed
    entries ``varint shared | varint non_shared | varint value_len | key suffix | value`` with a restart array every 16
    entries, each block followed by a 5-byte trailer (compression type 0 + masked CRC32C), an (empty) metaindex block, an
    index block whose values are BlockHandles (varint offset, varint size) and a 48-byte footer ending in the magic
    0xdb4775248b80fb57.  BundleWriter disables compression, so only type 0 is handled.
  * key "" holds a BundleHeaderProto (num_shards = 1, little endian, version); every other key is a tensor name whose
    value is a BundleEntryProto: dtype (1), shape (2), shard_id (3), offset (4), size (5), masked crc32c (6).
  * ``.data-00000-of-00001`` is the raw little-endian tensor bytes at those offsets.
"""
import os
import struct

import numpy as np

MAGIC = 0xdb4775248b80fb57
RESTART_INTERVAL = 16
BLOCK_SIZE = 4096

# tensorflow/core/framework/types.proto
DTYPES = {1: np.float32, 2: np.float64, 3: np.int32, 4: np.uint8, 5: np.int16, 6: np.int8, 9: np.int64, 10: np.bool_,
          17: np.uint16, 19: np.float16, 22: np.uint32, 23: np.uint64}
DTYPE_IDS = {np.dtype(v): k for k, v in DTYPES.items()}


# ---------------------------------------------------------------- CRC32C (Castagnoli), masked as in tensorflow/core/lib/hash/crc32c.h
def _make_table():
    t = np.zeros(256, np.uint32)
    for i in range(256):
        c = i
        for _ in range(8):
            c = (c >> 1) ^ (0x82f63b78 if c & 1 else 0)
        t[i] = c
    return t


_TABLE = _make_table()
_native_crc = None


def _native():
    """ssdb_crc32c of libssd_b200.so when the library is built (tensor payloads are hundreds of MB); else pure Python."""
    global _native_crc
    if _native_crc is None:
        try:
            import ctypes
            import ssdb
            f = ssdb.lib().ssdb_crc32c
            f.restype = ctypes.c_uint32
            f.argtypes = [ctypes.c_uint32, ctypes.c_void_p, ctypes.c_size_t]
            _native_crc = f
        except Exception:
            _native_crc = False
    return _native_crc


def crc32c(data, crc=0):
    data = bytes(data) if not isinstance(data, (bytes, bytearray, memoryview)) else data
    f = _native()
    if f and len(data) > 4096:
        import ctypes
        buf = (ctypes.c_char * len(data)).from_buffer_copy(data)
        return int(f(crc, buf, len(data)))
    c = crc ^ 0xffffffff
    tab = _TABLE
    for b in data:
        c = int(tab[(c ^ b) & 0xff]) ^ (c >> 8)
    return c ^ 0xffffffff


def mask_crc(c):
    return (((c >> 15) | (c << 17)) + 0xa282ead8) & 0xffffffff


# ---------------------------------------------------------------- varints / protobuf subset
def _put_varint(out, v):
    v &= (1 << 64) - 1
    while v >= 0x80:
        out.append((v & 0x7f) | 0x80)
        v >>= 7
    out.append(v)


def _get_varint(buf, pos):
    shift = v = 0
    while True:
        b = buf[pos]; pos += 1
        v |= (b & 0x7f) << shift
        if b < 0x80:
            return v, pos
        shift += 7


def _fields(buf):
    """Iterate (field number, wire type, value) of one protobuf message; value is an int or a bytes slice."""
    pos, n = 0, len(buf)
    while pos < n:
        tag, pos = _get_varint(buf, pos)
        num, wt = tag >> 3, tag & 7
        if wt == 0:
            v, pos = _get_varint(buf, pos)
        elif wt == 1:
            v = struct.unpack_from('<Q', buf, pos)[0]; pos += 8
        elif wt == 2:
            ln, pos = _get_varint(buf, pos)
            v = bytes(buf[pos:pos + ln]); pos += ln
        elif wt == 5:
            v = struct.unpack_from('<I', buf, pos)[0]; pos += 4
        else:
            raise ValueError('unsupported protobuf wire type %d' % wt)
        yield num, wt, v


def _parse_entry(buf):
    e = dict(dtype=0, shape=[], shard=0, offset=0, size=0, crc=0, sliced=False)
    for num, wt, v in _fields(buf):
        if num == 1:
            e['dtype'] = v
        elif num == 2:
            for n2, _, v2 in _fields(v):
                if n2 == 2:
                    size = 0
                    for n3, _, v3 in _fields(v2):
                        if n3 == 1:
                            size = v3 if v3 < (1 << 63) else v3 - (1 << 64)
                    e['shape'].append(size)
        elif num == 3:
            e['shard'] = v
        elif num == 4:
            e['offset'] = v
        elif num == 5:
            e['size'] = v
        elif num == 6:
            e['crc'] = v
        elif num == 7:
            e['sliced'] = True
    return e


def _encode_entry(dtype_id, shape, offset, size, crc):
    out = bytearray()
    out.append(0x08); _put_varint(out, dtype_id)
    shp = bytearray()
    for d in shape:
        dim = bytearray([0x08]); _put_varint(dim, d)
        shp.append(0x12); _put_varint(shp, len(dim)); shp += dim
    out.append(0x12); _put_varint(out, len(shp)); out += shp
    if offset:
        out.append(0x20); _put_varint(out, offset)
    out.append(0x28); _put_varint(out, size)
    out.append(0x35); out += struct.pack('<I', crc)
    return bytes(out)


def _encode_header():
    version = bytes([0x08, 0x01])                                 # VersionDef.producer = 1
    return bytes([0x08, 0x01]) + bytes([0x1a, len(version)]) + version   # num_shards = 1, endianness LITTLE (default), version


# ---------------------------------------------------------------- table blocks
def _block_entries(block):
    """(key, value) pairs of one table block (restart array ignored: entries are walked sequentially)."""
    nrestarts = struct.unpack_from('<I', block, len(block) - 4)[0]
    end = len(block) - 4 - 4 * nrestarts
    pos, key = 0, b''
    while pos < end:
        shared, pos = _get_varint(block, pos)
        non_shared, pos = _get_varint(block, pos)
        vlen, pos = _get_varint(block, pos)
        key = key[:shared] + bytes(block[pos:pos + non_shared]); pos += non_shared
        yield key, bytes(block[pos:pos + vlen]); pos += vlen


def _read_block(buf, offset, size, verify=True):
    block = buf[offset:offset + size]
    ctype = buf[offset + size]
    if ctype != 0:
        raise ValueError('compressed table blocks are not supported (tensor bundles are written uncompressed)')
    if verify:
        want = struct.unpack_from('<I', buf, offset + size + 1)[0]
        if mask_crc(crc32c(bytes(block) + bytes([ctype]))) != want:
            raise ValueError('table block checksum mismatch at offset %d' % offset)
    return block


class _BlockBuilder:
    def __init__(self):
        self.buf = bytearray(); self.restarts = [0]; self.count = 0; self.last = b''

    def add(self, key, value):
        shared = 0
        if self.count % RESTART_INTERVAL == 0 and self.count:
            self.restarts.append(len(self.buf))
        elif self.count:
            m = min(len(key), len(self.last))
            while shared < m and key[shared] == self.last[shared]:
                shared += 1
        _put_varint(self.buf, shared); _put_varint(self.buf, len(key) - shared); _put_varint(self.buf, len(value))
        self.buf += key[shared:]; self.buf += value
        self.last = key; self.count += 1

    def finish(self):
        out = bytes(self.buf) + b''.join(struct.pack('<I', r) for r in self.restarts) + struct.pack('<I', len(self.restarts))
        return out

    def size(self):
        return len(self.buf) + 4 * len(self.restarts) + 4


def _emit_block(f, content):
    off = f.tell()
    f.write(content); f.write(b'\x00'); f.write(struct.pack('<I', mask_crc(crc32c(content + b'\x00'))))
    return off, len(content)


def _handle(off, size):
    out = bytearray(); _put_varint(out, off); _put_varint(out, size)
    return bytes(out)


# ---------------------------------------------------------------- public API
def list_entries(prefix, verify=True):
    """{tensor name: entry dict (dtype id, shape, offset, size, crc)} of ``<prefix>.index``."""
    buf = memoryview(open(prefix + '.index', 'rb').read())
    if len(buf) < 48 or struct.unpack_from('<Q', buf, len(buf) - 8)[0] != MAGIC:
        raise ValueError('%s.index is not a TensorFlow table file (bad magic)' % prefix)
    foot = buf[len(buf) - 48:]
    _, p = _get_varint(foot, 0); _, p = _get_varint(foot, p)            # metaindex handle
    ioff, p = _get_varint(foot, p); isize, p = _get_varint(foot, p)
    entries = {}
    for _, hv in _block_entries(_read_block(buf, ioff, isize, verify)):
        off, q = _get_varint(hv, 0); size, q = _get_varint(hv, q)
        for key, value in _block_entries(_read_block(buf, off, size, verify)):
            if key == b'':
                for num, _, v in _fields(value):
                    if num == 1 and v != 1:
                        raise ValueError('sharded bundles (%d shards) are not supported' % v)
                    if num == 2 and v != 0:
                        raise ValueError('big-endian bundles are not supported')
                continue
            entries[key.decode()] = _parse_entry(value)
    return entries


def read_bundle(prefix, names=None, verify_tensors=False):
    """{name: ndarray} of the bundle at ``prefix`` (e.g. ``vgg_graph/vgg/variables/variables``); `names` restricts the
    tensors that are materialised.  Table blocks are always checksummed; tensor payloads only on request."""
    entries = list_entries(prefix)
    data = np.memmap(prefix + '.data-00000-of-00001', dtype=np.uint8, mode='r')
    out = {}
    for name, e in entries.items():
        if names is not None and name not in names:
            continue
        if e['sliced'] or e['dtype'] not in DTYPES:
            continue                                               # partitioned variables / strings: not on this path
        raw = data[e['offset']:e['offset'] + e['size']]
        if verify_tensors and mask_crc(crc32c(raw.tobytes())) != e['crc']:
            raise ValueError('tensor checksum mismatch: ' + name)
        out[name] = np.frombuffer(raw.tobytes(), dtype=np.dtype(DTYPES[e['dtype']]).newbyteorder('<')).reshape(e['shape'])
    return out


def write_bundle(prefix, tensors):
    """Write {name: ndarray} as ``<prefix>.index`` + ``<prefix>.data-00000-of-00001`` (one shard, sorted by name)."""
    os.makedirs(os.path.dirname(os.path.abspath(prefix)), exist_ok=True)
    items = sorted(((k.encode(), np.asarray(v)) for k, v in tensors.items()), key=lambda kv: kv[0])     # 0-d stays 0-d
    metas = []
    with open(prefix + '.data-00000-of-00001', 'wb') as f:
        for key, a in items:
            if a.dtype not in DTYPE_IDS:
                raise TypeError('unsupported dtype %s for %s' % (a.dtype, key.decode()))
            raw = a.astype(a.dtype.newbyteorder('<'), copy=False).tobytes(order='C')
            metas.append((key, DTYPE_IDS[a.dtype], a.shape, f.tell(), len(raw), mask_crc(crc32c(raw))))
            f.write(raw)
    with open(prefix + '.index', 'wb') as f:
        index = _BlockBuilder()
        blk = _BlockBuilder()
        last_key = b''

        def flush():
            nonlocal blk
            if blk.count:
                off, size = _emit_block(f, blk.finish())
                index.add(last_key, _handle(off, size))            # separator = the block's last key (valid, not shortened)
                blk = _BlockBuilder()

        blk.add(b'', _encode_header())
        for key, dt, shape, off, size, crc in metas:
            if blk.size() >= BLOCK_SIZE:
                flush()
            blk.add(key, _encode_entry(dt, shape, off, size, crc)); last_key = key
        flush()
        moff, msize = _emit_block(f, _BlockBuilder().finish())
        ioff, isize = _emit_block(f, index.finish())
        foot = _handle(moff, msize) + _handle(ioff, isize)
        f.write(foot + b'\x00' * (40 - len(foot)) + struct.pack('<Q', MAGIC))
