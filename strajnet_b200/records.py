"""Input records without TensorFlow (scope row f3): TFRecord framing, `tf.train.Example` wire format and the
reference's record decode.

The reference reads its pre-processed scenes with `tf.data.TFRecordDataset(...).map(_parse_image_function_test)`
(`inference.py:256-258`; schema written at `data_preprocessing.py:417-440`, decoded at `inference.py:84-96`).  This
module restates the three layers it goes through, from their published formats:

* TFRecord framing: `uint64 length | uint32 masked_crc32c(length) | data | uint32 masked_crc32c(data)`, little endian,
  mask(c) = ((c >> 15) | (c << 17)) + 0xa282ead8.  CRC-32C comes from the C ABI (`sj_crc32c`, host code).
* `tf.train.Example` protobuf: `Example{1: Features{1: map<string, Feature>}}`,
  `Feature{1: BytesList | 2: FloatList | 3: Int64List}`, parsed straight from the wire (no generated classes).
* `decode_example`: the `tf.io.decode_raw` + reshape + cast of `_parse_image_function_test`, except that the rasters
  KEEP their record dtypes (bool/uint8 occupancy, int8 map) -- the CUDA patch embedding consumes them directly
  (`SjIoSpec`), which is 4x less host->device traffic than the reference's float32 casts and bit-identical.
"""
from __future__ import annotations

import ctypes as C
import struct
from typing import Dict, Iterator, List, Union

import numpy as np

_MASK_DELTA = 0xA282EAD8


def crc32c(data, crc: int = 0) -> int:
    """CRC-32C (Castagnoli) through the library's host helper."""
    from . import _lib
    mv = memoryview(data).cast("B")
    if len(mv) == 0:
        return crc
    arr = np.frombuffer(mv, dtype=np.uint8)
    return int(_lib.lib().sj_crc32c(C.c_void_p(arr.ctypes.data), len(mv), crc))


def masked_crc32c(data) -> int:
    c = crc32c(data)
    return ((((c >> 15) | (c << 17)) & 0xFFFFFFFF) + _MASK_DELTA) & 0xFFFFFFFF


# ------------------------------------------------------------------------------------------- TFRecord framing
def read_tfrecords(path: str, verify: bool = True) -> Iterator[bytes]:
    """Yield the payload of every record of an uncompressed TFRecord file (`compression_type=''`, train.py:380)."""
    import os
    size = os.path.getsize(path)
    with open(path, "rb") as f:
        while True:
            head = f.read(12)
            if not head:
                return
            if len(head) < 12:
                raise ValueError(f"{path}: truncated record header")
            (length,), (lcrc,) = struct.unpack("<Q", head[:8]), struct.unpack("<I", head[8:])
            if verify and masked_crc32c(head[:8]) != lcrc:
                raise ValueError(f"{path}: corrupted record length")
            if length + 4 > size - f.tell():
                raise ValueError(f"{path}: truncated record (or corrupted length)")
            data = f.read(length)
            tail = f.read(4)
            if len(data) < length or len(tail) < 4:
                raise ValueError(f"{path}: truncated record")
            if verify and masked_crc32c(data) != struct.unpack("<I", tail)[0]:
                raise ValueError(f"{path}: corrupted record data")
            yield data


def write_tfrecords(path: str, payloads) -> None:
    with open(path, "wb") as f:
        for data in payloads:
            head = struct.pack("<Q", len(data))
            f.write(head + struct.pack("<I", masked_crc32c(head)) + data + struct.pack("<I", masked_crc32c(data)))


# ------------------------------------------------------------------------------------------- protobuf wire format
def _varint(buf, pos: int):
    out = shift = 0
    while True:
        b = buf[pos]
        pos += 1
        out |= (b & 0x7F) << shift
        if not b & 0x80:
            return out, pos
        shift += 7


def _fields(buf):
    """Iterate (field number, wire type, value) over one message; length-delimited values are memoryviews."""
    mv = memoryview(buf)
    pos, n = 0, len(mv)
    while pos < n:
        key, pos = _varint(mv, pos)
        num, wt = key >> 3, key & 7
        if wt == 0:
            v, pos = _varint(mv, pos)
        elif wt == 1:
            v, pos = mv[pos:pos + 8], pos + 8
        elif wt == 2:
            ln, pos = _varint(mv, pos)
            v, pos = mv[pos:pos + ln], pos + ln
        elif wt == 5:
            v, pos = mv[pos:pos + 4], pos + 4
        else:
            raise ValueError(f"unsupported protobuf wire type {wt}")
        yield num, wt, v


def _enc_varint(v: int) -> bytes:
    out = bytearray()
    while True:
        b = v & 0x7F
        v >>= 7
        out.append(b | (0x80 if v else 0))
        if not v:
            return bytes(out)


def _enc_ld(num: int, payload: bytes) -> bytes:
    return _enc_varint((num << 3) | 2) + _enc_varint(len(payload)) + payload


FeatureValue = Union[List[bytes], np.ndarray]


def parse_example(payload) -> Dict[str, FeatureValue]:
    """`tf.io.parse_single_example` for the three feature kinds: bytes -> list of memoryviews, float -> float32 array,
    int64 -> int64 array."""
    out: Dict[str, FeatureValue] = {}
    for num, _, features in _fields(payload):
        if num != 1:
            continue
        for fnum, _, entry in _fields(features):  # map<string, Feature> entries
            if fnum != 1:
                continue
            key, feat = None, None
            for enum_, _, v in _fields(entry):
                if enum_ == 1:
                    key = bytes(v).decode("utf-8")
                elif enum_ == 2:
                    feat = v
            if key is None:
                continue
            value: FeatureValue = []
            for knum, _, lst in _fields(feat if feat is not None else b""):
                if knum == 1:  # BytesList
                    value = [v for n2, _, v in _fields(lst) if n2 == 1]
                elif knum == 2:  # FloatList (packed or not)
                    vals = []
                    for n2, wt, v in _fields(lst):
                        if n2 == 1:
                            vals.append(np.frombuffer(v, dtype="<f4"))
                    value = np.concatenate(vals) if vals else np.zeros(0, np.float32)
                elif knum == 3:  # Int64List
                    vals = []
                    for n2, wt, v in _fields(lst):
                        if n2 != 1:
                            continue
                        if wt == 0:
                            vals.append(v)
                        else:
                            pos, mv = 0, memoryview(v)
                            while pos < len(mv):
                                x, pos = _varint(mv, pos)
                                vals.append(x)
                    value = np.array([x - (1 << 64) if x >> 63 else x for x in vals], dtype=np.int64)
            out[key] = value
    return out


def serialize_example(features: Dict[str, Union[bytes, List[bytes], np.ndarray]]) -> bytes:
    """`tf.train.Example(...).SerializeToString()` for bytes / float32 / int64 features (used by tests and tools)."""
    entries = b""
    for key in sorted(features):
        v = features[key]
        if isinstance(v, (bytes, bytearray, memoryview)):
            v = [bytes(v)]
        if isinstance(v, list):
            feat = _enc_ld(1, b"".join(_enc_ld(1, bytes(x)) for x in v))
        elif np.issubdtype(np.asarray(v).dtype, np.floating):
            feat = _enc_ld(2, _enc_ld(1, np.asarray(v, "<f4").tobytes()))
        else:
            feat = _enc_ld(3, _enc_ld(1, b"".join(_enc_varint(int(x) & ((1 << 64) - 1)) for x in np.asarray(v).ravel())))
        entries += _enc_ld(1, _enc_ld(1, key.encode("utf-8")) + _enc_ld(2, feat))
    return _enc_ld(1, entries)


# ------------------------------------------------------------------------------------------- the reference's schema
# name -> (raw dtype, shape); data_preprocessing.py:417-440 writes ndarray.tobytes() of exactly these
SCHEMA = {
    "centerlines": (np.float64, (256, 10, 7)),
    "actors": (np.float64, (48, 11, 8)),
    "occl_actors": (np.float64, (16, 11, 8)),
    "ogm": (np.bool_, (512, 512, 11, 2)),
    "map_image": (np.int8, (256, 256, 3)),
    "vec_flow": (np.float32, (512, 512, 2)),
}


def decode_example(payload, raw: bool = True, vehicle_plane_only: bool = False) -> Dict[str, object]:
    """`_parse_image_function_test` (inference.py:84-96).

    raw=True keeps the raster dtypes of the record (ogm uint8 in {0,1}, map_image int8 -- the model divides by 256 on
    the device); raw=False reproduces the reference's float32 tensors exactly (ogm 0/1, map_image int8/256).
    vehicle_plane_only=True returns `ogm[..., 0]` ([S,S,11]), the only plane the model reads (modules.py:572): half the
    bytes of the largest input on its way to the GPU (STrajNet and InferencePipeline accept either shape)."""
    d = parse_example(payload)
    out: Dict[str, object] = {}
    for name, (dt, shape) in SCHEMA.items():
        if name not in d or not isinstance(d[name], list) or len(d[name]) != 1:
            raise ValueError(f"record has no single bytes feature '{name}'")
        buf = d[name][0]
        n = int(np.prod(shape)) * np.dtype(dt).itemsize
        if len(buf) != n:
            raise ValueError(f"feature '{name}': {len(buf)} bytes, expected {n} for {np.dtype(dt).name}{list(shape)}")
        a = np.frombuffer(buf, dtype=np.uint8 if dt is np.bool_ else dt).reshape(shape)
        if name in ("centerlines", "actors", "occl_actors"):
            a = a.astype(np.float32)  # tf.cast(float64 -> float32)
        elif name == "ogm":
            if vehicle_plane_only:
                a = a[..., 0]
            a = (a != 0).astype(np.uint8) if raw else (a != 0).astype(np.float32)
        elif name == "map_image" and not raw:
            a = a.astype(np.float32) / 256
        out[name] = a
    sid = d.get("scenario/id")
    out["scenario/id"] = bytes(sid[0]) if isinstance(sid, list) and sid else b""
    return out


def batch_examples(examples: List[Dict[str, object]]) -> Dict[str, np.ndarray]:
    """`dataset.batch(n)`: stack the decoded examples and rename to the model's argument names
    (`test_step`, inference.py:144-152: ogm, map_img=map_image, obs=actors, occ=occl_actors, mapt=centerlines,
    flow=vec_flow)."""
    names = {"ogm": "ogm", "map_image": "map_img", "actors": "obs", "occl_actors": "occ", "centerlines": "mapt",
             "vec_flow": "flow"}
    out = {dst: np.stack([e[src] for e in examples]) for src, dst in names.items()}
    out["scenario/id"] = [e["scenario/id"] for e in examples]
    return out
