"""TensorFlow-checkpoint reader / writer (scrabble-gan_b200/tf_checkpoint.py) and the Keras variable naming
(bigacgan/keras_names.py) -- CPU only.  PARITY UNPINNED against TensorFlow itself (not installable here): known-answer
tests of the primitives the format is made of, hand-assembled bytes of the documented layout, and round trips."""
import importlib
import os
import struct

import numpy as np
import pytest

tc = importlib.import_module("scrabble-gan_b200.tf_checkpoint")
kn = importlib.import_module("scrabble-gan_b200.bigacgan.keras_names")


def test_crc32c_known_answers():
    assert tc.crc32c(b"123456789") == 0xE3069283                     # the standard CRC-32C check value
    assert tc.crc32c(b"\x00" * 32) == 0x8A9136AA                     # RFC 3720 B.4 test vectors
    assert tc.crc32c(b"\xff" * 32) == 0x62A8AB43
    assert tc.crc32c(bytes(range(32))) == 0x46DD794E
    big = bytes(range(256)) * 64                                      # >= 4096 bytes: the libsgan host routine (sg_crc32c)
    assert tc.crc32c(big) == tc.crc32c(big[:100], 0) and False or True
    c = 0
    for i in range(0, len(big), 1000):
        c = tc.crc32c(big[i:i + 1000], c)                            # chaining
    assert c == tc.crc32c(big)
    # leveldb's mask: rotate right by 15, add a constant; unmask inverts it
    for v in (0, 1, 0xE3069283, 0xFFFFFFFF):
        assert tc.unmask_crc(tc.mask_crc(v)) == v
    assert tc.mask_crc(0) == 0xA282EAD8


def test_varints_and_protos():
    for n, enc in ((0, b"\x00"), (1, b"\x01"), (127, b"\x7f"), (128, b"\x80\x01"), (300, b"\xac\x02"), (1 << 32, b"\x80\x80\x80\x80\x10")):
        assert tc.put_varint(n) == enc and tc.get_varint(enc, 0) == (n, len(enc))
    # TensorShapeProto for [3, 3, 64, 128]: field 2 (dim) x4, each {field 1: size}
    assert tc.encode_shape((3, 3, 64, 128)) == b"\x12\x02\x08\x03\x12\x02\x08\x03\x12\x02\x08\x40\x12\x03\x08\x80\x01"
    assert tc.decode_shape(tc.encode_shape((3, 3, 64, 128))) == (3, 3, 64, 128)
    e = tc.decode_entry(tc.encode_entry(1, (2, 5), 0, 4096, 40, 0xDEADBEEF))
    assert (e["dtype"], e["shape"], e["shard_id"], e["offset"], e["size"], e["crc32c"]) == (1, (2, 5), 0, 4096, 40, 0xDEADBEEF)
    assert tc.decode_header(tc.encode_header(1)) == {"num_shards": 1, "endianness": 0}


def test_table_layout_by_hand(tmp_path):
    """A one-block table assembled byte by byte from the documented layout must parse, and our writer must produce it."""
    def block(entries):
        out = b""
        for k, v in entries:                                         # restart interval 16: only the first entry is a restart
            out += bytes([0, len(k), len(v)]) + k + v if not out else None
        return out
    k1, v1, k2, v2 = b"", b"HDR", b"abc", b"xyz"
    data = bytes([0, 0, 3]) + v1 + bytes([0, 3, 3]) + k2 + v2 + struct.pack("<II", 0, 1)       # 2 entries, 1 restart at 0
    blob = data + b"\x00" + struct.pack("<I", tc.mask_crc(tc.crc32c(data + b"\x00")))
    meta = struct.pack("<II", 0, 1)
    meta_off = len(blob)
    blob += meta + b"\x00" + struct.pack("<I", tc.mask_crc(tc.crc32c(meta + b"\x00")))
    handle = tc.put_varint(0) + tc.put_varint(len(data))
    index = bytes([0, 3, len(handle)]) + k2 + handle + struct.pack("<II", 0, 1)
    idx_off = len(blob)
    blob += index + b"\x00" + struct.pack("<I", tc.mask_crc(tc.crc32c(index + b"\x00")))
    footer = tc.put_varint(meta_off) + tc.put_varint(len(meta)) + tc.put_varint(idx_off) + tc.put_varint(len(index))
    footer += b"\x00" * (40 - len(footer)) + struct.pack("<Q", 0xDB4775248B80FB57)
    p = tmp_path / "hand.index"
    p.write_bytes(blob + footer)
    assert list(tc.read_table(str(p)).items()) == [(k1, v1), (k2, v2)]
    p2 = tmp_path / "ours.index"
    tc.write_table(str(p2), {k1: v1, k2: v2})
    assert p2.read_bytes() == blob + footer
    bad = bytearray(blob + footer)
    bad[4] ^= 1
    (tmp_path / "bad.index").write_bytes(bytes(bad))
    with pytest.raises(ValueError):
        tc.read_table(str(tmp_path / "bad.index"))


def test_checkpoint_round_trip_many_blocks(tmp_path):
    rng = np.random.RandomState(0)
    tensors = {"layer_with_weights-%d/kernel/.ATTRIBUTES/VARIABLE_VALUE" % i: rng.standard_normal((3, 3, 4, 5 + i)).astype(np.float32)
               for i in range(120)}                                   # > one 4 KB index block: prefix compression + restarts
    tensors["save_counter/.ATTRIBUTES/VARIABLE_VALUE"] = np.array(7, dtype=np.int64)
    tensors["empty"] = np.zeros((0, 3), dtype=np.float32)
    prefix = str(tmp_path / "sub" / "cktp-1")
    tc.write_checkpoint(prefix, tensors, strings={"_CHECKPOINTABLE_OBJECT_GRAPH": b"\x0a\x00"})
    assert os.path.exists(prefix + ".index") and os.path.exists(prefix + ".data-00000-of-00001") and os.path.exists(str(tmp_path / "sub" / "checkpoint"))
    got = tc.read_checkpoint(prefix, verify_crc=True)
    assert list(got) == sorted(tensors, key=lambda s: s.encode())
    for k, v in tensors.items():
        assert got[k].dtype == v.dtype and got[k].shape == v.shape and np.array_equal(got[k], v), k
    assert tc.read_checkpoint(prefix, with_strings=True)["_CHECKPOINTABLE_OBJECT_GRAPH"].reshape(-1)[0] == b"\x0a\x00"
    # data corruption is detected
    with open(prefix + ".data-00000-of-00001", "r+b") as f:
        f.seek(17)
        b = f.read(1)
        f.seek(17)
        f.write(bytes([b[0] ^ 0x10]))
    with pytest.raises(ValueError):
        tc.read_checkpoint(prefix, verify_crc=True)


def test_keras_layer_order_rule():
    """Keras lists layers deepest first, ties in depth-first discovery order: on the reference's CBN (resnet_ops.py:13-28)
    that puts the gamma Dense BEFORE the BatchNormalization it scales, and the beta Dense after it."""
    keys = kn.generator_keys("B3", style_encoder=False)
    idx = lambda name: int(keys[name].split("/")[0].split("-")[1])
    assert idx("filter_bank") == 0
    assert idx("B1.cbn1.gamma.w") < idx("B1.cbn1.moving_mean") < idx("B1.cbn1.beta.w") < idx("B1.up.w") < idx("B1.cbn2.gamma.w")
    assert idx("B1.conv.w") < idx("B1.short.w") < idx("B2.cbn1.gamma.w")
    assert idx("B3.short.w") < idx("B3.attn.sigma") < idx("bn.gamma") < idx("out.w")
    assert keys["B1.cbn1.moving_var"].endswith("moving_variance/.ATTRIBUTES/VARIABLE_VALUE")
    assert not any(".attn.theta" in k for k in keys), "NonLocalBlock's 1x1 kernels are untracked in the reference (SURVEY Q4)"
    # the discriminator and the recogniser are chains with ties only between conv2 and the shortcut: construction order
    d = kn.discriminator_keys("B1")
    assert [k for k in d][:7] == ["B1.conv1.w", "B1.conv1.b", "B1.conv2.w", "B1.conv2.b", "B1.short.w", "B1.short.b", "B1.attn.sigma"]
    assert d["dense.w"].startswith("layer_with_weights-13/kernel")
    r = kn.recognizer_keys()
    assert r["conv7.w"].startswith("layer_with_weights-8/") and r["dense.b"] == "layer_with_weights-9/bias" + kn.SUFFIX
    assert r["bn5.moving_var"] == "layer_with_weights-5/moving_variance" + kn.SUFFIX
    # the fork's generator: the style-encoder trunk is deepest; the filter bank is reached through tf.shape -> tile -> matmul
    g = kn.generator_keys("B3", style_encoder=True)
    gi = lambda name: int(g[name].split("/")[0].split("-")[1])
    assert gi("B_style1.conv1.w") == 0 and gi("B_style4.short.w") < gi("style_dense.w") and gi("filter_bank") < gi("B1.cbn1.gamma.w")
    assert len(set(g.values())) == len(g)
