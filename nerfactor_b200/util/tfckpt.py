"""Pure-Python reader / writer for TensorFlow 2 checkpoints (the "tensor bundle" format) so
weights trained with the reference -- the released NeRF / shape / BRDF / NeRFactor models
(reference README.md:40-42) -- can drive these kernels, and weights trained here can be written
back under the reference's names.  No TensorFlow needed.

What the reference saves (nerfactor/trainvali.py:134-141, nerfactor/util/io.py:36-45):
`tf.train.Checkpoint(step=..., optimizer=..., net=model)` with every trainable Dense aliased
directly under the model as `net_<name>_layer<i>` (nerfactor/models/base.py:81-104).  TF2's
object-based saver names each variable by its attribute path from the root:

    net/net_<name>_layer<i>/kernel/.ATTRIBUTES/VARIABLE_VALUE        float32 [in, out]
    net/net_<name>_layer<i>/bias/.ATTRIBUTES/VARIABLE_VALUE          float32 [out]
    net/_light/.ATTRIBUTES/VARIABLE_VALUE                            float32 [h, 2h, 3]
    net/latent_code/_z/.ATTRIBUTES/VARIABLE_VALUE                    float32 [n_brdfs, z_dim]
    step/.ATTRIBUTES/VARIABLE_VALUE                                  int32 []
    optimizer/iter/.ATTRIBUTES/VARIABLE_VALUE                        int64 []
    <variable path>/.OPTIMIZER_SLOT/optimizer/{m,v,vhat}/.ATTRIBUTES/VARIABLE_VALUE

On disk a checkpoint `prefix` is two files:

  prefix.index                an SSTable (LevelDB table format): sorted key -> protobuf.
                              key ""  -> BundleHeaderProto {num_shards, endianness, version}
                              key k   -> BundleEntryProto {dtype, shape, shard_id, offset, size, crc32c}
  prefix.data-00000-of-00001  the raw little-endian tensor bytes at [offset, offset + size)

Table format (leveldb/doc/table_format.md): data blocks, a meta-index block, an index block and a
48-byte footer (two BlockHandles, padding, magic 0xdb4775248b80fb57).  A block is a run of
prefix-compressed entries (varint shared, varint non_shared, varint value_len, key suffix, value)
followed by uint32 restart offsets and their count; on disk it is followed by a 1-byte
compression type (0 none, 1 snappy) and a masked CRC32C.  TensorFlow writes bundles uncompressed;
snappy blocks are decoded anyway.

`_CHECKPOINTABLE_OBJECT_GRAPH` (the serialized object graph `tf.train.Checkpoint.restore` matches
against) is skipped on read and not produced on write: files written here are name-based bundles,
readable with `tf.train.load_checkpoint` / `tf.train.list_variables` and by this module.
"""
import os
import re
import struct

import numpy as np

TABLE_MAGIC = 0xdb4775248b80fb57
_SUFFIX = '/.ATTRIBUTES/VARIABLE_VALUE'

# tensorflow/core/framework/types.proto
_DTYPES = {1: np.dtype('<f4'), 2: np.dtype('<f8'), 3: np.dtype('<i4'), 4: np.dtype('u1'),
           5: np.dtype('<i2'), 6: np.dtype('i1'), 9: np.dtype('<i8'), 10: np.dtype('?'),
           17: np.dtype('<u2'), 19: np.dtype('<f2'), 22: np.dtype('<u4'), 23: np.dtype('<u8')}
_DT_STRING = 7
_DTYPE_CODES = {v: k for k, v in _DTYPES.items()}


# ------------------------------------------------------------------------ crc32c (Castagnoli)
def _make_crc_table():
    tbl = []
    for i in range(256):
        c = i
        for _ in range(8):
            c = (c >> 1) ^ 0x82F63B78 if c & 1 else c >> 1
        tbl.append(c)
    return tbl


_CRC_TABLE = _make_crc_table()
_CRC_NP = np.array(_CRC_TABLE, dtype=np.uint32)


def crc32c(data, crc=0):
    crc ^= 0xFFFFFFFF
    tbl = _CRC_TABLE
    for b in bytes(data):
        crc = tbl[(crc ^ b) & 0xFF] ^ (crc >> 8)
    return crc ^ 0xFFFFFFFF


def mask_crc(crc):
    """leveldb / TF store crcs rotated and offset (crc32c.h: Mask)."""
    return ((((crc >> 15) | (crc << 17)) & 0xFFFFFFFF) + 0xa282ead8) & 0xFFFFFFFF


def unmask_crc(m):
    rot = (m - 0xa282ead8) & 0xFFFFFFFF
    return ((rot >> 17) | (rot << 15)) & 0xFFFFFFFF


# ------------------------------------------------------------------------------- varints
def _get_varint(buf, pos):
    result, shift = 0, 0
    while True:
        b = buf[pos]
        pos += 1
        result |= (b & 0x7F) << shift
        if not b & 0x80:
            return result, pos
        shift += 7


def _put_varint(v):
    out = bytearray()
    while True:
        b = v & 0x7F
        v >>= 7
        if v:
            out.append(b | 0x80)
        else:
            out.append(b)
            return bytes(out)


# ---------------------------------------------------------------------------- snappy (decode)
def _snappy_decompress(buf):
    n, pos = _get_varint(buf, 0)
    out = bytearray()
    while pos < len(buf):
        tag = buf[pos]
        pos += 1
        kind = tag & 3
        if kind == 0:                                   # literal
            ln = tag >> 2
            if ln >= 60:
                nb = ln - 59
                ln = int.from_bytes(buf[pos:pos + nb], 'little')
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
            off = buf[pos] | (buf[pos + 1] << 8)
            pos += 2
        else:
            ln = (tag >> 2) + 1
            off = int.from_bytes(buf[pos:pos + 4], 'little')
            pos += 4
        if off == 0 or off > len(out):
            raise ValueError("corrupt snappy block")
        for _ in range(ln):                             # may overlap: byte by byte
            out.append(out[-off])
    if len(out) != n:
        raise ValueError("snappy length mismatch")
    return bytes(out)


# ------------------------------------------------------------------------ table (SSTable) read
def _read_block(f_bytes, offset, size, verify=True):
    raw = f_bytes[offset:offset + size]
    ctype = f_bytes[offset + size]
    if verify:
        stored = struct.unpack('<I', f_bytes[offset + size + 1:offset + size + 5])[0]
        if unmask_crc(stored) != crc32c(f_bytes[offset:offset + size + 1]):
            raise ValueError("table block checksum mismatch at offset %d" % offset)
    if ctype == 0:
        return raw
    if ctype == 1:
        return _snappy_decompress(raw)
    raise ValueError("unknown block compression type %d" % ctype)


def _block_entries(block):
    """Yields (key, value) of one block (prefix-compressed entries, restart array at the end)."""
    n_restarts = struct.unpack('<I', block[-4:])[0]
    limit = len(block) - 4 - 4 * n_restarts
    pos, key = 0, b''
    while pos < limit:
        shared, pos = _get_varint(block, pos)
        non_shared, pos = _get_varint(block, pos)
        vlen, pos = _get_varint(block, pos)
        key = key[:shared] + block[pos:pos + non_shared]
        pos += non_shared
        yield key, block[pos:pos + vlen]
        pos += vlen


def read_table(path, verify=True):
    """-> list of (key bytes, value bytes) of an SSTable file, in key order."""
    data = open(path, 'rb').read()
    if len(data) < 48:
        raise ValueError("%s: too short for a table file" % path)
    footer = data[-48:]
    if struct.unpack('<Q', footer[40:])[0] != TABLE_MAGIC:
        raise ValueError("%s: not an SSTable (bad magic)" % path)
    pos = 0
    _, pos = _get_varint(footer, pos)          # metaindex handle (unused)
    _, pos = _get_varint(footer, pos)
    idx_off, pos = _get_varint(footer, pos)
    idx_size, pos = _get_varint(footer, pos)
    out = []
    for _, handle in _block_entries(_read_block(data, idx_off, idx_size, verify)):
        off, p = _get_varint(handle, 0)
        size, p = _get_varint(handle, p)
        out.extend(_block_entries(_read_block(data, off, size, verify)))
    return out


# ----------------------------------------------------------------------- table (SSTable) write
def _build_block(entries, restart_interval=16):
    buf, restarts, last = bytearray(), [], b''
    for i, (k, v) in enumerate(entries):
        if i % restart_interval == 0:
            restarts.append(len(buf))
            shared = 0
        else:
            shared = 0
            for a, b in zip(last, k):
                if a != b:
                    break
                shared += 1
        buf += _put_varint(shared) + _put_varint(len(k) - shared) + _put_varint(len(v))
        buf += k[shared:] + v
        last = k
    if not restarts:
        restarts = [0]
    for r in restarts:
        buf += struct.pack('<I', r)
    buf += struct.pack('<I', len(restarts))
    return bytes(buf)


def write_table(path, entries, block_size=4096):
    """entries: iterable of (key bytes, value bytes); written sorted, uncompressed."""
    entries = sorted(entries)
    out = bytearray()
    index = []

    def emit(block):
        off = len(out)
        out.extend(block)
        out.append(0)                                              # no compression
        out.extend(struct.pack('<I', mask_crc(crc32c(block + b'\x00'))))
        return off, len(block)

    cur, cur_size = [], 0
    for kv in entries:
        cur.append(kv)
        cur_size += len(kv[0]) + len(kv[1]) + 3
        if cur_size >= block_size:
            off, size = emit(_build_block(cur))
            index.append((cur[-1][0], _put_varint(off) + _put_varint(size)))
            cur, cur_size = [], 0
    if cur or not index:
        off, size = emit(_build_block(cur))
        index.append((cur[-1][0] if cur else b'', _put_varint(off) + _put_varint(size)))
    meta_off, meta_size = emit(_build_block([]))
    idx_off, idx_size = emit(_build_block(index, restart_interval=1))
    footer = _put_varint(meta_off) + _put_varint(meta_size) + _put_varint(idx_off) + _put_varint(idx_size)
    footer += b'\x00' * (40 - len(footer)) + struct.pack('<Q', TABLE_MAGIC)
    out.extend(footer)
    with open(path, 'wb') as f:
        f.write(bytes(out))


# --------------------------------------------------------------------------- protobuf (minimal)
def _parse_fields(buf):
    """-> list of (field number, wire type, value) of one protobuf message."""
    pos, out = 0, []
    while pos < len(buf):
        tag, pos = _get_varint(buf, pos)
        fn, wt = tag >> 3, tag & 7
        if wt == 0:
            v, pos = _get_varint(buf, pos)
        elif wt == 1:
            v = buf[pos:pos + 8]
            pos += 8
        elif wt == 2:
            ln, pos = _get_varint(buf, pos)
            v = buf[pos:pos + ln]
            pos += ln
        elif wt == 5:
            v = buf[pos:pos + 4]
            pos += 4
        else:
            raise ValueError("unsupported protobuf wire type %d" % wt)
        out.append((fn, wt, v))
    return out


def _parse_entry(buf):
    """BundleEntryProto (tensorflow/core/protobuf/tensor_bundle.proto)."""
    e = {'dtype': 0, 'shape': [], 'shard_id': 0, 'offset': 0, 'size': 0, 'crc32c': None,
         'slices': 0}
    for fn, wt, v in _parse_fields(buf):
        if fn == 1:
            e['dtype'] = v
        elif fn == 2:                                    # TensorShapeProto
            for f2, _, v2 in _parse_fields(v):
                if f2 == 2:                              # repeated Dim
                    size = 0
                    for f3, _, v3 in _parse_fields(v2):
                        if f3 == 1:
                            size = v3 - (1 << 64) if v3 >= (1 << 63) else v3
                    e['shape'].append(size)
        elif fn == 3:
            e['shard_id'] = v
        elif fn == 4:
            e['offset'] = v
        elif fn == 5:
            e['size'] = v
        elif fn == 6:
            e['crc32c'] = struct.unpack('<I', v)[0]
        elif fn == 7:
            e['slices'] += 1
    return e


def _put_field(fn, wt, payload):
    return _put_varint((fn << 3) | wt) + payload


def _encode_entry(dtype_code, shape, offset, size, crc):
    shp = b''.join(_put_field(2, 2, (lambda d: _put_varint(len(d)) + d)(_put_field(1, 0, _put_varint(int(s)))))
                   for s in shape)
    msg = _put_field(1, 0, _put_varint(dtype_code))
    msg += _put_field(2, 2, _put_varint(len(shp)) + shp)
    # shard_id 0 and offset 0 are proto3 defaults (omitted)
    if offset:
        msg += _put_field(4, 0, _put_varint(offset))
    msg += _put_field(5, 0, _put_varint(size))
    msg += _put_field(6, 5, struct.pack('<I', crc))
    return msg


def _encode_header(num_shards=1):
    version = _put_field(1, 0, _put_varint(1))           # VersionDef.producer = 1
    return (_put_field(1, 0, _put_varint(num_shards)) +   # endianness LITTLE = 0 (default)
            _put_field(3, 2, _put_varint(len(version)) + version))


# ------------------------------------------------------------------------------ bundle read
def _shard_path(prefix, shard, num_shards):
    return '%s.data-%05d-of-%05d' % (prefix, shard, num_shards)


def read_checkpoint(prefix, verify=True, verify_limit=1 << 22):
    """-> {tensor name: np.ndarray} of every numeric tensor in checkpoint `prefix`
    (`.../ckpt-100`: the path without `.index`).  String tensors (the object graph, save
    counters' names) are skipped.  Data checksums are verified for tensors up to `verify_limit`
    bytes (pure-Python crc32c is slow); table checksums always when `verify`."""
    entries = read_table(prefix + '.index', verify)
    if not entries or entries[0][0] != b'':
        raise ValueError("%s.index: missing bundle header" % prefix)
    num_shards, endian = 1, 0
    for fn, _, v in _parse_fields(entries[0][1]):
        if fn == 1:
            num_shards = v
        elif fn == 2:
            endian = v
    if endian != 0:
        raise ValueError("big-endian bundles are not supported")
    shards = {}
    out = {}
    for key, val in entries[1:]:
        e = _parse_entry(val)
        if e['dtype'] == _DT_STRING or e['slices']:
            continue
        if e['dtype'] not in _DTYPES:
            raise ValueError("%s: unsupported dtype enum %d" % (key.decode(), e['dtype']))
        sid = e['shard_id']
        if sid not in shards:
            shards[sid] = np.memmap(_shard_path(prefix, sid, num_shards), dtype=np.uint8, mode='r')
        raw = shards[sid][e['offset']:e['offset'] + e['size']]
        dt = _DTYPES[e['dtype']]
        n = int(np.prod(e['shape'])) if e['shape'] else 1
        if n * dt.itemsize != e['size']:
            raise ValueError("%s: size %d does not match shape %s" % (key.decode(), e['size'], e['shape']))
        if verify and e['crc32c'] is not None and e['size'] <= verify_limit:
            if unmask_crc(e['crc32c']) != crc32c(raw.tobytes()):
                raise ValueError("%s: data checksum mismatch" % key.decode())
        out[key.decode()] = np.frombuffer(raw.tobytes(), dtype=dt).reshape(e['shape']).copy()
    return out


def write_checkpoint(prefix, tensors):
    """Writes {name: array} as a single-shard bundle `prefix.index` + `prefix.data-00000-of-00001`."""
    os.makedirs(os.path.dirname(os.path.abspath(prefix)), exist_ok=True)
    entries = [(b'', _encode_header(1))]
    offset = 0
    with open(_shard_path(prefix, 0, 1), 'wb') as f:
        for name in sorted(tensors):
            shape = np.asarray(tensors[name]).shape          # ascontiguousarray promotes 0-d to 1-d
            a = np.ascontiguousarray(np.asarray(tensors[name]))
            if a.dtype.byteorder == '>':
                a = a.astype(a.dtype.newbyteorder('<'))
            code = _DTYPE_CODES.get(a.dtype)
            if code is None:
                raise ValueError("%s: dtype %s not supported" % (name, a.dtype))
            raw = a.tobytes()
            f.write(raw)
            entries.append((name.encode(), _encode_entry(code, shape, offset, len(raw),
                                                         mask_crc(crc32c(raw)))))
            offset += len(raw)
    write_table(prefix + '.index', entries)


def latest_checkpoint(ckpt_dir):
    """tf.train.latest_checkpoint: reads the text-proto `checkpoint` state file."""
    state = os.path.join(ckpt_dir, 'checkpoint')
    if not os.path.exists(state):
        return None
    m = re.search(r'^model_checkpoint_path:\s*"([^"]+)"', open(state).read(), re.M)
    if not m:
        return None
    p = m.group(1)
    return p if os.path.isabs(p) else os.path.join(ckpt_dir, p)


# ------------------------------------------------------------------- reference naming <-> params
_LAYER_RE = re.compile(r'^(?P<root>[^/]+)/(?P<sub>(?:[A-Za-z_0-9]+/)*)net_(?P<net>[A-Za-z_0-9]+?)_layer(?P<i>\d+)/'
                       r'(?P<kind>kernel|bias)' + re.escape(_SUFFIX) + '$')


def params_from_tensors(tensors, submodel=''):
    """Tensor dict of a reference checkpoint -> the params dict the models here take
    (`{'<net>': {'layers': [(W, b), ...]}, 'light': ..., 'z': ...}`; synth.make_stage_b_params
    layout).  `submodel`: attribute path of a nested model, e.g. 'brdf_model' for the BRDF prior
    inside a NeRFactor checkpoint; '' = the checkpointed model itself.  Optimizer slots and
    counters are ignored (tf.train.Checkpoint.restore(...).expect_partial(), util/io.py:45)."""
    want_sub = submodel.strip('/') + '/' if submodel else ''
    nets = {}
    params = {}
    for name, arr in tensors.items():
        if '.OPTIMIZER_SLOT' in name:
            continue
        m = _LAYER_RE.match(name)
        if m and m.group('sub') == want_sub:
            nets.setdefault(m.group('net'), {}).setdefault(int(m.group('i')), {})[m.group('kind')] = arr
            continue
        root_sub = re.match(r'^[^/]+/(?P<rest>.*)' + re.escape(_SUFFIX) + '$', name)
        if not root_sub:
            continue
        rest = root_sub.group('rest')
        if rest == want_sub + '_light':
            params['light'] = arr
        elif rest == want_sub + 'latent_code/_z':
            params['z'] = arr
    for net, layers in nets.items():
        idx = sorted(layers)
        if idx != list(range(len(idx))):
            raise ValueError("network %s: layers %s are not contiguous from 0" % (net, idx))
        params[net] = {'layers': [(layers[i]['kernel'].astype(np.float32),
                                   layers[i]['bias'].astype(np.float32)) for i in idx]}
    return params


def params_from_checkpoint(prefix, submodel='', verify=True):
    return params_from_tensors(read_checkpoint(prefix, verify), submodel)


def tensors_from_params(params, root='net', step=None, adam=None):
    """Inverse of params_from_tensors: reference key names for a params dict (plus optional
    `step` and AMSGrad state `adam = {'iter': int, 'slots': {key: (m, v, vhat)}}`, keys being the
    variable paths returned here without the suffix)."""
    out = {}
    for net, v in params.items():
        if net == 'light':
            out['%s/_light%s' % (root, _SUFFIX)] = np.asarray(v, np.float32)
        elif net == 'z':
            out['%s/latent_code/_z%s' % (root, _SUFFIX)] = np.asarray(v, np.float32)
        else:
            for i, (w, b) in enumerate(v['layers']):
                out['%s/net_%s_layer%d/kernel%s' % (root, net, i, _SUFFIX)] = np.asarray(w, np.float32)
                out['%s/net_%s_layer%d/bias%s' % (root, net, i, _SUFFIX)] = np.asarray(b, np.float32)
    if step is not None:
        out['step' + _SUFFIX] = np.asarray(step, np.int32)
    if adam is not None:
        out['optimizer/iter' + _SUFFIX] = np.asarray(adam['iter'], np.int64)
        for key, (m, v, vhat) in adam.get('slots', {}).items():
            for slot, arr in (('m', m), ('v', v), ('vhat', vhat)):
                out['%s/.OPTIMIZER_SLOT/optimizer/%s%s' % (key, slot, _SUFFIX)] = np.asarray(arr, np.float32)
    return out
