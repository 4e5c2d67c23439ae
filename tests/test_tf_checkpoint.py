"""CPU tests of the TF-format checkpoint reader (scope row f2): LevelDB-format table, tensor bundle entries, object
graph walk, and `load_weights` / `save_weights` of the layer classes.  Parity unpinned against TensorFlow-written
files (TensorFlow cannot be installed here); these tests pin the reader to the format description and to the writer."""
import struct

import numpy as np
import pytest
import torch

from strajnet_b200 import tf_checkpoint as T
from strajnet_b200 import weights as W


def test_table_roundtrip_multiblock_and_prefix_compression(tmp_path):
    p = str(tmp_path / "t.index")
    items = {f"layer/{i:04d}/kernel/.ATTRIBUTES/VARIABLE_VALUE".encode(): bytes([i % 251]) * (i % 37) for i in range(500)}
    items[b""] = b"header"
    T.write_table(p, items)
    raw = open(p, "rb").read()
    assert struct.unpack("<Q", raw[-8:])[0] == 0xDB4775248B80FB57 and len(raw[-48:]) == 48
    assert T.read_table(p) == items
    # prefix compression really happened (keys share 'layer/0'): the file is much smaller than the raw keys
    assert len(raw) < sum(len(k) + len(v) for k, v in items.items())
    bad = bytearray(raw)
    bad[10] ^= 0x40
    open(p, "wb").write(bytes(bad))
    with pytest.raises(ValueError):
        T.read_table(p)
    open(p, "wb").write(raw[:-1])
    with pytest.raises(ValueError):
        T.read_table(p)


def test_bundle_roundtrip_dtypes_and_object_graph(tmp_path):
    rng = np.random.Generator(np.random.PCG64(0))
    w = {
        "encoder.patch_embed_map.proj.kernel": rng.standard_normal((4, 4, 3, 96)).astype(np.float32),
        "encoder.patch_embed_map.proj.bias": rng.standard_normal(96).astype(np.float32),
        "encoder.basic_layers.0.blocks.1.attn.relative_position_bias_table": rng.standard_normal((225, 3)).astype(np.float32),
        "encoder.basic_layers.1.blocks.0.norm1.gamma": np.ones(192, np.float32),
        "step": np.array(7, np.int64),
        "flags": np.array([True, False]),
    }
    prefix = str(tmp_path / "ckpt")
    T.save_keras_checkpoint(prefix, w)
    assert T.is_tf_checkpoint(prefix)
    b = T.TensorBundle(prefix)
    assert T.OBJECT_GRAPH_KEY in b.keys() and b.num_shards == 1
    key = "encoder/patch_embed_map/proj/kernel/.ATTRIBUTES/VARIABLE_VALUE"
    assert b.entries[key]["shape"] == [4, 4, 3, 96] and b.entries[key]["dtype"] == 1
    assert np.array_equal(b.read(key), w["encoder.patch_embed_map.proj.kernel"])
    paths = b.variables_by_attribute_path()
    assert set(paths) == set(w)
    got = T.load_keras_checkpoint(prefix)
    for k in w:
        assert got[k].dtype == w[k].dtype and np.array_equal(got[k], w[k]), k
    assert got["step"].shape == ()
    with pytest.raises(KeyError):
        T.load_keras_checkpoint(prefix, ["encoder.missing.kernel"])
    # a flipped byte in the data shard is caught by the per-tensor checksum
    d = bytearray(open(prefix + ".data-00000-of-00001", "rb").read())
    d[5] ^= 1
    open(prefix + ".data-00000-of-00001", "wb").write(bytes(d))
    with pytest.raises(ValueError):
        T.load_keras_checkpoint(prefix)


def test_bfloat16_and_string_entries(tmp_path):
    # hand-built bundle: one bf16 tensor and one string vector, to exercise the non-numpy dtypes of the reader
    from strajnet_b200.records import _enc_ld, _enc_varint, crc32c
    prefix = str(tmp_path / "b")
    vals = np.array([1.0, -2.5, 3.140625], np.float32)
    bf = (vals.view(np.uint32) >> 16).astype("<u2").tobytes()
    strs = [b"ab", b"", b"xyz"]
    lens = b"".join(_enc_varint(len(s)) for s in strs)
    # TensorFlow's WriteStringTensor rule, spelled out (NOT through the module's helper): the CRC runs over each length
    # as a fixed-width uint32, then over the 4 bytes of the masked length checksum, then over the string bytes
    c = 0
    for s_ in strs:
        c = crc32c(struct.pack("<I", len(s_)), c)
    len_ck = T._mask(c)
    c = crc32c(struct.pack("<I", len_ck), c)
    c = crc32c(b"".join(strs), c)
    str_crc = T._mask(c)
    sp = lens + struct.pack("<I", len_ck) + b"".join(strs)

    def entry(dtype, shape, off, payload, crc=None):
        sm = b"".join(_enc_ld(2, _enc_varint(8) + _enc_varint(d)) for d in shape)
        return (_enc_varint(8) + _enc_varint(dtype) + _enc_ld(2, sm) + (_enc_varint(32) + _enc_varint(off) if off else b"") +
                _enc_varint(40) + _enc_varint(len(payload)) + _enc_varint(53) +
                struct.pack("<I", T._mask(crc32c(payload)) if crc is None else crc))

    def write(string_entry_crc, string_payload):
        T.write_table(prefix + ".index", {b"": _enc_varint(8) + _enc_varint(1), b"h": entry(14, [3], 0, bf),
                                          b"s": entry(7, [3], len(bf), string_payload, string_entry_crc)})
        open(prefix + ".data-00000-of-00001", "wb").write(bf + string_payload)

    write(str_crc, sp)
    b = T.TensorBundle(prefix)
    assert np.array_equal(b.read("h"), vals) and b.read("s") == strs
    assert T.string_tensor_crc(strs) == (len_ck, str_crc)
    # the checksum over the raw on-disk bytes (varint lengths included) is NOT what TensorFlow stores: must be rejected
    write(T._mask(crc32c(sp)), sp)
    with pytest.raises(ValueError):
        T.TensorBundle(prefix).read("s")
    # ... and so must a wrong length checksum
    write(str_crc, lens + struct.pack("<I", T._mask(crc32c(lens))) + b"".join(strs))
    with pytest.raises(ValueError):
        T.TensorBundle(prefix).read("s")


def test_object_graph_skips_keras_aliases(tmp_path):
    from strajnet_b200.records import _enc_ld, _enc_varint
    prefix = str(tmp_path / "g")
    T.save_keras_checkpoint(prefix, {"dense.kernel": np.zeros((2, 2), np.float32)})
    b = T.TensorBundle(prefix)
    nodes = b.object_graph()
    assert nodes[0]["children"] == {"dense": 1} and nodes[1]["children"] == {"kernel": 2}
    assert nodes[2]["attributes"] == {"VARIABLE_VALUE": "dense/kernel/.ATTRIBUTES/VARIABLE_VALUE"}
    # re-write the graph with the alias edges Keras adds; the attribute path must still be the only result
    def ref(name, nid):
        return _enc_ld(1, _enc_varint(8) + _enc_varint(nid) + _enc_ld(2, name.encode()))
    g = (_enc_ld(1, ref("layer_with_weights-0", 1) + ref("dense", 1) + ref("layer-0", 1) + ref("variables", 3)) +
         _enc_ld(1, ref("kernel", 2)) +
         _enc_ld(1, _enc_ld(2, _enc_ld(1, b"VARIABLE_VALUE") + _enc_ld(3, b"dense/kernel/.ATTRIBUTES/VARIABLE_VALUE"))) +
         _enc_ld(1, ref("0", 2)))
    b.read = lambda key, _r=b.read: g if key == T.OBJECT_GRAPH_KEY else _r(key)
    assert b.variables_by_attribute_path() == {"dense.kernel": "dense/kernel/.ATTRIBUTES/VARIABLE_VALUE"}


def test_layer_save_and_load_weights_tf_format(tmp_path):
    import strajnet_b200 as sj
    mk = lambda: sj.Mlp(96, 384)  # a layer whose build() needs no GPU
    shapes = mk().weight_shapes()
    w = W.default_init(shapes, seed=3)
    m = mk()
    m.set_weights(w)
    prefix = str(tmp_path / "final_model")  # train.py:366 naming
    m.save_weights(prefix)
    assert T.is_tf_checkpoint(prefix)
    m2 = mk()
    m2.load_weights(prefix)
    for k in shapes:
        assert torch.equal(m2.get_weights()[k], w[k].float()), k
    # a checkpoint lacking a parameter is an error, not a silent default
    T.save_keras_checkpoint(prefix, {k: v.numpy() for k, v in w.items() if k != "fc2.bias"})
    with pytest.raises(KeyError):
        mk().load_weights(prefix)
    # the .npz route is unchanged
    m.save_weights(str(tmp_path / "w.npz"))
    m3 = mk()
    m3.load_weights(str(tmp_path / "w.npz"))
    assert torch.equal(m3.get_weights()["fc1.kernel"], w["fc1.kernel"].float())
