"""CPU tests of the record reader (scope row f3): CRC-32C known answers, TFRecord framing, tf.train.Example wire
format and the reference's record decode (inference.py:84-96)."""
import os
import struct

import numpy as np
import pytest

from strajnet_b200 import records as R


def test_crc32c_known_answers():
    # RFC 3720 appendix B.4 test vectors + the classic check value
    assert R.crc32c(bytes(32)) == 0x8A9136AA
    assert R.crc32c(b"\xff" * 32) == 0x62A8AB43
    assert R.crc32c(bytes(range(32))) == 0x46DD794E
    assert R.crc32c(bytes(reversed(range(32)))) == 0x113FDB5C
    assert R.crc32c(b"123456789") == 0xE3069283
    assert R.crc32c(b"") == 0


def test_crc32c_incremental_and_unaligned():
    rng = np.random.Generator(np.random.PCG64(0))
    data = rng.integers(0, 256, size=100003, dtype=np.uint8).tobytes()
    whole = R.crc32c(data)
    for cut in (0, 1, 7, 8, 9, 4096, 99999):
        assert R.crc32c(data[cut:], R.crc32c(data[:cut])) == whole
    # bitwise reference implementation on a short prefix
    def slow(b):
        c = 0xFFFFFFFF
        for x in b:
            c ^= x
            for _ in range(8):
                c = (c >> 1) ^ 0x82F63B78 if c & 1 else c >> 1
        return c ^ 0xFFFFFFFF
    for n in (1, 3, 8, 15, 64, 1000):
        assert R.crc32c(data[5:5 + n]) == slow(data[5:5 + n])


def test_tfrecord_framing_layout_and_roundtrip(tmp_path):
    p = str(tmp_path / "a.tfrecords")
    payloads = [b"", b"x", bytes(range(256)) * 40, b"last"]
    R.write_tfrecords(p, payloads)
    raw = open(p, "rb").read()
    # first record: empty payload = 8-byte length 0, crc of the length, no data, crc of b""
    assert raw[:8] == struct.pack("<Q", 0)
    c = R.crc32c(raw[:8])
    assert struct.unpack("<I", raw[8:12])[0] == ((((c >> 15) | (c << 17)) & 0xFFFFFFFF) + 0xA282EAD8) & 0xFFFFFFFF
    assert struct.unpack("<I", raw[12:16])[0] == 0xA282EAD8  # masked crc of the empty string (crc 0)
    assert list(R.read_tfrecords(p)) == payloads
    # corruption of a data byte and truncation are detected
    bad = bytearray(raw)
    bad[100] ^= 1  # inside the payload of the third record
    open(p, "wb").write(bytes(bad))
    with pytest.raises(ValueError):
        list(R.read_tfrecords(p))
    assert len(list(R.read_tfrecords(p, verify=False))) == 4
    open(p, "wb").write(raw[:-3])
    with pytest.raises(ValueError):
        list(R.read_tfrecords(p))


def test_example_wire_format_hand_encoded():
    # Example{features{feature{key:"a" value{bytes_list{value:"xy"}}}}} assembled by hand from the protobuf spec
    msg = bytes([0x0A, 0x0D, 0x0A, 0x0B, 0x0A, 0x01, ord("a"), 0x12, 0x06, 0x0A, 0x04, 0x0A, 0x02, ord("x"), ord("y")])
    d = R.parse_example(msg)
    assert list(d) == ["a"] and [bytes(v) for v in d["a"]] == [b"xy"]
    assert R.serialize_example({"a": b"xy"}) == msg
    # float_list (packed) and int64_list (packed varints, incl. a negative value)
    msg2 = R.serialize_example({"f": np.array([1.5, -2.0], np.float32), "i": np.array([3, -1], np.int64), "e": []})
    d2 = R.parse_example(msg2)
    assert np.array_equal(d2["f"], np.array([1.5, -2.0], np.float32))
    assert np.array_equal(d2["i"], np.array([3, -1], np.int64))
    assert d2["e"] == []


def _synthetic_example(seed):
    rng = np.random.Generator(np.random.PCG64(seed))
    arrs = {
        "centerlines": rng.standard_normal((256, 10, 7)),
        "actors": rng.standard_normal((48, 11, 8)),
        "occl_actors": rng.standard_normal((16, 11, 8)),
        "ogm": rng.random((512, 512, 11, 2)) < 0.03,
        "map_image": rng.integers(-128, 128, size=(256, 256, 3)).astype(np.int8),
        "vec_flow": rng.standard_normal((512, 512, 2)).astype(np.float32),
    }
    feats = {k: v.tobytes() for k, v in arrs.items()}  # data_preprocessing.py:405-426
    feats["scenario/id"] = f"scene{seed}".encode()
    feats["byc_flow"] = b"ignored"
    feats["gt_flow"] = []
    return arrs, R.serialize_example(feats)


def test_decode_example_matches_reference_decode(tmp_path):
    arrs, payload = _synthetic_example(1)
    p = str(tmp_path / "s.tfrecords")
    R.write_tfrecords(p, [payload])
    (rec,) = list(R.read_tfrecords(p))
    ref = R.decode_example(rec, raw=False)  # the reference's float32 tensors (inference.py:87-94)
    assert ref["ogm"].dtype == np.float32 and np.array_equal(ref["ogm"], arrs["ogm"].astype(np.float32))
    assert np.array_equal(ref["map_image"], arrs["map_image"].astype(np.float32) / 256)
    assert np.array_equal(ref["actors"], arrs["actors"].astype(np.float32))
    assert np.array_equal(ref["occl_actors"], arrs["occl_actors"].astype(np.float32))
    assert np.array_equal(ref["centerlines"], arrs["centerlines"].astype(np.float32))
    assert np.array_equal(ref["vec_flow"], arrs["vec_flow"]) and ref["scenario/id"] == b"scene1"
    raw = R.decode_example(rec, raw=True)  # record dtypes kept for the device path
    assert raw["ogm"].dtype == np.uint8 and np.array_equal(raw["ogm"].astype(np.float32), ref["ogm"])
    assert raw["map_image"].dtype == np.int8 and np.array_equal(raw["map_image"].astype(np.float32) / 256, ref["map_image"])
    b = R.batch_examples([raw, raw])
    assert b["ogm"].shape == (2, 512, 512, 11, 2) and b["obs"].shape == (2, 48, 11, 8) and b["map_img"].shape == (2, 256, 256, 3)
    assert b["occ"].shape == (2, 16, 11, 8) and b["mapt"].shape == (2, 256, 10, 7) and b["flow"].shape == (2, 512, 512, 2)
    veh = R.decode_example(rec, raw=True, vehicle_plane_only=True)  # the one plane the model reads (modules.py:572)
    assert veh["ogm"].shape == (512, 512, 11) and np.array_equal(veh["ogm"], raw["ogm"][..., 0])


def test_decode_example_rejects_wrong_sizes():
    _, payload = _synthetic_example(2)
    d = {k: (bytes(v[0]) if isinstance(v, list) and v else b"") for k, v in R.parse_example(payload).items()}
    d["actors"] = d["actors"][:-8]
    with pytest.raises(ValueError):
        R.decode_example(R.serialize_example(d))
    del d["actors"]
    with pytest.raises(ValueError):
        R.decode_example(R.serialize_example(d))
