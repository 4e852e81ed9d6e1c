"""TensorFlow-free checkpoint reader / writer (nerfactor_b200/util/tfckpt.py): the checksum and
table primitives against published known answers, hand-built tables exercising every reader
path, and the reference's variable naming (models/base.py:81-104, util/io.py:36-45)."""
import os
import struct

import numpy as np
import pytest

from nerfactor_b200 import synth
from nerfactor_b200.util import tfckpt


def test_crc32c_known_answers():
    # RFC 3720 B.4 / leveldb util/crc32c_test.cc
    assert tfckpt.crc32c(b'123456789') == 0xE3069283
    assert tfckpt.crc32c(b'\x00' * 32) == 0x8A9136AA
    assert tfckpt.crc32c(b'\xff' * 32) == 0x62A8AB43
    assert tfckpt.crc32c(bytes(range(32))) == 0x46DD794E
    assert tfckpt.crc32c(bytes(range(31, -1, -1))) == 0x113FDB5C
    c = tfckpt.crc32c(b'foo')
    assert tfckpt.mask_crc(c) != c and tfckpt.unmask_crc(tfckpt.mask_crc(c)) == c
    assert tfckpt.crc32c(b'world', tfckpt.crc32c(b'hello ')) == tfckpt.crc32c(b'hello world')


def test_varint_and_snappy():
    for v in (0, 1, 127, 128, 300, 2 ** 32 - 1, 2 ** 63 + 5):
        enc = tfckpt._put_varint(v)
        assert tfckpt._get_varint(enc, 0) == (v, len(enc))
    # snappy stream: literal "abcd", copy(offset 4, len 8) -> "abcdabcdabcd", literal "xyz"
    stream = bytes([15]) + bytes([3 << 2]) + b'abcd' + bytes([((8 - 4) << 2) | 1, 4]) + \
        bytes([2 << 2]) + b'xyz'
    assert tfckpt._snappy_decompress(stream) == b'abcdabcdabcdxyz'


def test_table_round_trip_multi_block(tmp_path):
    rng = np.random.default_rng(0)
    entries = [(b'', b'header')]
    for i in range(500):        # long shared prefixes -> prefix compression; > 4 KB -> several blocks
        k = ('net/net_albedo_mlp_layer%03d/kernel/.ATTRIBUTES/VARIABLE_VALUE' % i).encode()
        entries.append((k, rng.bytes(int(rng.integers(0, 40)))))
    path = str(tmp_path / 't.index')
    tfckpt.write_table(path, entries, block_size=1024)
    got = tfckpt.read_table(path)
    assert got == sorted(entries)
    raw = bytearray(open(path, 'rb').read())
    assert struct.unpack('<Q', raw[-8:])[0] == tfckpt.TABLE_MAGIC
    raw[10] ^= 0x40                                       # corrupt a data block
    open(path, 'wb').write(bytes(raw))
    with pytest.raises(ValueError, match='checksum'):
        tfckpt.read_table(path)


def test_reader_on_hand_built_snappy_table(tmp_path):
    """A table assembled byte by byte (not with write_table): one snappy-compressed data block
    with restart interval 2, so the reader's decompression / restart handling is exercised
    independently of the writer."""
    kv = [(b'', b'H'), (b'aa/x', b'1'), (b'aa/y', b'22'), (b'ab', b'333')]
    blk, last, restarts = bytearray(), b'', []
    for i, (k, v) in enumerate(kv):
        shared = 0
        if i % 2 == 0:
            restarts.append(len(blk))
        else:
            while shared < min(len(last), len(k)) and last[shared] == k[shared]:
                shared += 1
        blk += tfckpt._put_varint(shared) + tfckpt._put_varint(len(k) - shared) + \
            tfckpt._put_varint(len(v)) + k[shared:] + v
        last = k
    for r in restarts:
        blk += struct.pack('<I', r)
    blk += struct.pack('<I', len(restarts))
    # snappy: one literal covering the whole block
    n = len(blk)
    lit = bytes([(n - 1) << 2]) if n <= 60 else bytes([60 << 2, n - 1])
    comp = tfckpt._put_varint(n) + lit + bytes(blk)
    out = bytearray()

    def emit(body, ctype):
        off = len(out)
        out.extend(body)
        out.append(ctype)
        out.extend(struct.pack('<I', tfckpt.mask_crc(tfckpt.crc32c(bytes(body) + bytes([ctype])))))
        return off, len(body)

    d_off, d_size = emit(comp, 1)
    m_off, m_size = emit(tfckpt._build_block([]), 0)
    i_off, i_size = emit(tfckpt._build_block(
        [(b'ab', tfckpt._put_varint(d_off) + tfckpt._put_varint(d_size))]), 0)
    footer = b''.join(tfckpt._put_varint(x) for x in (m_off, m_size, i_off, i_size))
    out.extend(footer + b'\x00' * (40 - len(footer)) + struct.pack('<Q', tfckpt.TABLE_MAGIC))
    path = str(tmp_path / 'h.index')
    open(path, 'wb').write(bytes(out))
    assert tfckpt.read_table(path) == kv


def test_bundle_entry_proto_layout():
    """BundleEntryProto bytes for a float32 [63, 128] tensor at offset 1024, field by field
    (tensor_bundle.proto: 1 dtype, 2 shape{2 dim{1 size}}, 4 offset, 5 size, 6 crc32c fixed32)."""
    e = tfckpt._encode_entry(1, (63, 128), 1024, 63 * 128 * 4, 0xDEADBEEF)
    assert e == bytes([0x08, 0x01,                       # dtype = DT_FLOAT
                       0x12, 0x09,                       # shape, 9 bytes
                       0x12, 0x02, 0x08, 63,             #   dim {size: 63}
                       0x12, 0x03, 0x08, 0x80, 0x01,     #   dim {size: 128}
                       0x20, 0x80, 0x08,                 # offset = 1024
                       0x28, 0x80, 0xFC, 0x01,           # size = 32256
                       0x35, 0xEF, 0xBE, 0xAD, 0xDE])    # crc32c
    p = tfckpt._parse_entry(e)
    assert (p['dtype'], p['shape'], p['offset'], p['size'], p['crc32c']) == \
        (1, [63, 128], 1024, 32256, 0xDEADBEEF)


def test_checkpoint_round_trip_and_reference_names(tmp_path):
    params = synth.make_stage_b_params(3, 'learned', light_hw=(4, 8))
    params['z'] = np.random.default_rng(1).standard_normal((5, 3)).astype(np.float32)
    tensors = tfckpt.tensors_from_params(
        params, step=17, adam={'iter': 1700, 'slots': {
            'net/net_albedo_mlp_layer0/kernel': tuple(
                np.full((63, 128), v, np.float32) for v in (1., 2., 3.))}})
    assert 'net/net_lvis_mlp_layer3/kernel/.ATTRIBUTES/VARIABLE_VALUE' in tensors
    assert 'net/_light/.ATTRIBUTES/VARIABLE_VALUE' in tensors
    assert 'net/latent_code/_z/.ATTRIBUTES/VARIABLE_VALUE' in tensors
    # sub-model of a NeRFactor checkpoint (nerfactor.py:58-60: self.brdf_model)
    tensors['net/brdf_model/net_brdf_mlp_layer0/kernel/.ATTRIBUTES/VARIABLE_VALUE'] = \
        np.ones((18, 128), np.float32)
    tensors['net/brdf_model/net_brdf_mlp_layer0/bias/.ATTRIBUTES/VARIABLE_VALUE'] = \
        np.zeros((128,), np.float32)
    prefix = str(tmp_path / 'checkpoints' / 'ckpt-17')
    tfckpt.write_checkpoint(prefix, tensors)
    assert os.path.exists(prefix + '.index') and os.path.exists(prefix + '.data-00000-of-00001')
    back = tfckpt.read_checkpoint(prefix)
    assert set(back) == set(tensors)
    for k in tensors:
        assert back[k].dtype == np.asarray(tensors[k]).dtype and np.array_equal(back[k], tensors[k])
    assert back['step/.ATTRIBUTES/VARIABLE_VALUE'].shape == () and int(back['optimizer/iter/.ATTRIBUTES/VARIABLE_VALUE']) == 1700
    p2 = tfckpt.params_from_checkpoint(prefix)
    assert set(p2) == set(params)
    for k, v in params.items():
        if k in ('light', 'z'):
            assert np.array_equal(p2[k], v)
        else:
            assert len(p2[k]['layers']) == len(v['layers'])
            for (w0, b0), (w1, b1) in zip(v['layers'], p2[k]['layers']):
                assert np.array_equal(w0, w1) and np.array_equal(b0, b1)
    sub = tfckpt.params_from_checkpoint(prefix, submodel='brdf_model')
    assert list(sub) == ['brdf_mlp'] and sub['brdf_mlp']['layers'][0][0].shape == (18, 128)
    # data corruption is detected
    raw = bytearray(open(prefix + '.data-00000-of-00001', 'rb').read())
    raw[100] ^= 1
    open(prefix + '.data-00000-of-00001', 'wb').write(bytes(raw))
    with pytest.raises(ValueError, match='checksum'):
        tfckpt.read_checkpoint(prefix)


def test_latest_checkpoint(tmp_path):
    d = tmp_path / 'checkpoints'
    d.mkdir()
    assert tfckpt.latest_checkpoint(str(d)) is None
    (d / 'checkpoint').write_text('model_checkpoint_path: "ckpt-40"\n'
                                  'all_model_checkpoint_paths: "ckpt-20"\n'
                                  'all_model_checkpoint_paths: "ckpt-40"\n')
    assert tfckpt.latest_checkpoint(str(d)) == str(d / 'ckpt-40')
