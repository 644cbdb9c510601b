"""TensorFlow checkpoint-V2 bundle reader / writer and the VGG import -- CPU only.

TensorFlow is not installable here, so the format is checked by (1) byte-level known answers that follow from the
published format (CRC32C test vector, masked CRC, footer magic, a hand-assembled one-entry table), (2) round trips
through many-block tables, and (3) the fc6 / fc7 decimation against the reference's own loops (ssdvgg.py:245-280)."""
import os
import struct

import numpy as np
import pytest

import tf_bundle as tb
import vgg_import


def test_crc32c_known_answers():
    assert tb.crc32c(b'123456789') == 0xe3069283                      # the standard CRC-32C check value
    assert tb.crc32c(b'') == 0
    assert tb.crc32c(bytes(32)) == 0x8a9136aa                         # RFC 3720 B.4: 32 bytes of zeros
    assert tb.crc32c(bytes([0xff] * 32)) == 0x62a8ab43                # RFC 3720 B.4: 32 bytes of ones
    assert tb.crc32c(bytes(range(32))) == 0x46dd794e                  # RFC 3720 B.4: 0x00..0x1f
    # incremental == one shot; native (libssd_b200.so) == pure Python
    data = np.random.default_rng(0).integers(0, 256, 70001, dtype=np.uint8).tobytes()
    assert tb.crc32c(data[30000:], tb.crc32c(data[:30000])) == tb.crc32c(data)
    saved = tb._native_crc
    try:
        tb._native_crc = False
        py = tb.crc32c(data)
    finally:
        tb._native_crc = saved
    assert py == tb.crc32c(data)
    # leveldb's mask: rotate right by 15 and add a constant (crc32c.h)
    assert tb.mask_crc(0) == 0xa282ead8 and tb.mask_crc(0xe3069283) == ((0xe3069283 >> 15 | 0xe3069283 << 17) + 0xa282ead8) & 0xffffffff


def test_round_trip_many_blocks(tmp_path):
    rng = np.random.default_rng(1)
    tensors = {}
    for i in range(600):                                               # ~35 bytes of key + entry each: several 4 KB blocks
        shape = tuple(int(v) for v in rng.integers(1, 5, int(rng.integers(0, 4))))
        name = 'scope%d/block_%03d/%s' % (i % 7, i, 'filter' if i % 2 else 'biases')
        dt = [np.float32, np.float64, np.int32, np.int64, np.uint8][i % 5]
        tensors[name] = (rng.standard_normal(shape) * 100).astype(dt)
    tensors['global_step'] = np.array(123456, np.int64)                # a scalar
    prefix = str(tmp_path / 'sub' / 'model.ckpt')
    tb.write_bundle(prefix, tensors)
    assert os.path.getsize(prefix + ".index") > 4 * tb.BLOCK_SIZE
    back = tb.read_bundle(prefix, verify_tensors=True)
    assert sorted(back) == sorted(tensors)
    for k, v in tensors.items():
        assert back[k].dtype == v.dtype and back[k].shape == v.shape and np.array_equal(back[k], v), k
    some = tb.read_bundle(prefix, names={'global_step'})
    assert list(some) == ['global_step'] and int(some['global_step']) == 123456
    ent = tb.list_entries(prefix)
    offs = sorted((e['offset'], e['size']) for e in ent.values())
    assert offs[0][0] == 0 and all(a[0] + a[1] == b[0] for a, b in zip(offs, offs[1:]))       # packed back to back, name order


def test_file_layout_follows_the_table_format(tmp_path):
    prefix = str(tmp_path / 'one')
    tb.write_bundle(prefix, {'w': np.arange(6, dtype=np.float32).reshape(2, 3)})
    raw = open(prefix + '.index', 'rb').read()
    assert struct.unpack('<Q', raw[-8:])[0] == 0xdb4775248b80fb57 and len(raw[-48:]) == 48
    # first data block: header entry (empty key) then 'w'; entries are (shared, non_shared, value_len, key, value)
    assert raw[0] == 0 and raw[1] == 0                                 # shared = 0, key length 0: the header key ""
    hlen = raw[2]
    header = raw[3:3 + hlen]
    assert header[:2] == b'\x08\x01'                                   # BundleHeaderProto.num_shards = 1
    pos = 3 + hlen
    assert raw[pos] == 0 and raw[pos + 1] == 1 and raw[pos + 3:pos + 4] == b'w'
    entry = raw[pos + 4:pos + 4 + raw[pos + 2]]
    fields = {n: v for n, _, v in tb._fields(entry)}
    assert fields[1] == 1 and fields[5] == 24 and fields[6] == tb.mask_crc(tb.crc32c(np.arange(6, dtype=np.float32).tobytes()))
    assert open(prefix + '.data-00000-of-00001', 'rb').read() == np.arange(6, dtype='<f4').tobytes()
    # a flipped bit in a table block is detected
    bad = bytearray(raw); bad[5] ^= 1
    open(prefix + '.index', 'wb').write(bytes(bad))
    with pytest.raises(ValueError):
        tb.read_bundle(prefix)


def test_decimation_equals_the_reference_loops():
    rng = np.random.default_rng(2)
    cin, n = 8, 16                                                     # reduced channel counts, same index pattern
    fc6_w, fc6_b = rng.standard_normal((7, 7, cin, 4 * n)), rng.standard_normal(4 * n)
    fc7_w, fc7_b = rng.standard_normal((1, 1, 4 * n, 4 * n)), rng.standard_normal(4 * n)
    # the reference's loops (ssdvgg.py:245-253, 273-280), restated
    w6 = np.zeros((3, 3, cin, n)); b6 = np.zeros(n)
    for i in range(n):
        b6[i] = fc6_b[4 * i]
        for h in range(3):
            for w in range(3):
                w6[h, w, :, i] = fc6_w[3 * h, 3 * w, :, 4 * i]
    w7 = np.zeros((1, 1, n, n)); b7 = np.zeros(n)
    for i in range(n):
        b7[i] = fc7_b[4 * i]
        for j in range(n):
            w7[:, :, j, i] = fc7_w[:, :, 4 * j, 4 * i]
    g6, gb6 = vgg_import.decimate_fc6(fc6_w, fc6_b)
    g7, gb7 = vgg_import.decimate_fc7(fc7_w, fc7_b)
    assert np.array_equal(g6, w6) and np.array_equal(gb6, b6) and np.array_equal(g7, w7) and np.array_equal(gb7, b7)


def test_vgg_saved_model_directory_to_engine_tensors(tmp_path):
    """A saved-model-shaped directory (reduced channel counts) -> engine tensor names with conv6 / conv7 decimated."""
    rng = np.random.default_rng(3)
    variables = {}
    cin = 3
    for l in vgg_import.VGG_CONVS:
        variables[l + '/filter'] = rng.standard_normal((3, 3, cin, 4)).astype(np.float32)
        variables[l + '/biases'] = rng.standard_normal(4).astype(np.float32)
        cin = 4
    variables['fc6/weights'] = rng.standard_normal((7, 7, 4, 32)).astype(np.float32)
    variables['fc6/biases'] = rng.standard_normal(32).astype(np.float32)
    variables['fc7/weights'] = rng.standard_normal((1, 1, 32, 32)).astype(np.float32)
    variables['fc7/biases'] = rng.standard_normal(32).astype(np.float32)
    variables['fc8/weights'] = rng.standard_normal((1, 1, 32, 10)).astype(np.float32)      # present in the file, unused
    tb.write_bundle(str(tmp_path / 'vgg' / 'variables' / 'variables'), variables)
    assert vgg_import.find_bundle(str(tmp_path)) is not None
    P = vgg_import.load_vgg_dir(str(tmp_path))
    assert len(P) == 2 * 13 + 4 and 'fc8/weights' not in P
    assert np.array_equal(P['conv3_2/filter'], variables['conv3_2/filter'])
    assert P['mod_conv6/filter'].shape == (3, 3, 4, 8) and np.array_equal(P['mod_conv6/filter'][1, 2, :, 3], variables['fc6/weights'][3, 6, :, 12])
    assert P['mod_conv7/filter'].shape == (1, 1, 8, 8) and P['mod_conv7/filter'][0, 0, 2, 5] == variables['fc7/weights'][0, 0, 8, 20]
    assert np.array_equal(P['mod_conv7/biases'], variables['fc7/biases'][::4])
    with pytest.raises(FileNotFoundError):
        vgg_import.load_vgg_dir(str(tmp_path / 'nowhere'))


def test_ssdvgg_checkpoint_round_trip_as_tf_bundle(tmp_path):
    """SSDVGG.save(tf_checkpoint=True) -> build_from_metagraph: every tensor back under the reference's variable names
    (host side only: no engine is created, so this runs without a GPU)."""
    from ssdvgg import SSDVGG, Session
    a = SSDVGG(Session(), 'vgg300'); a.build_from_vgg(None, 20)
    prefix = str(tmp_path / 'e5.ckpt')
    a.save(prefix, tf_checkpoint=True)
    names = tb.list_entries(prefix)
    assert 'conv4_3/filter' in names and 'l2_norm_conv4_3/scale' in names and 'classifiers/classifier5_3/biases' in names
    assert names['mod_conv6/filter']['shape'] == [3, 3, 512, 1024]
    b = SSDVGG(Session(), 'vgg300'); b.build_from_metagraph(None, prefix)
    pa, pb = a.get_params(), b.get_params()
    assert sorted(pa) == sorted(pb) and b.num_vars == 25 and b.num_classes == 21
    assert all(np.array_equal(pa[k], pb[k]) for k in pa)
