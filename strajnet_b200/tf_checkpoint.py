"""TF-format (tensor bundle) checkpoints without TensorFlow (scope row f2).

The reference saves and restores its weights with Keras `model.save_weights(path)` / `model.load_weights(path)` in the
TF checkpoint format (`train.py:358,366`, `inference.py:283`): `<path>.index` + `<path>.data-00000-of-00001`.  This
module restates the published formats those files are made of, so trained weights drop onto `STrajNet.load_weights`:

* `<path>.index` is a LevelDB-format table (SSTable): 48-byte footer (metaindex + index `BlockHandle`s as varint64
  pairs, padding, magic 0xdb4775248b80fb57), blocks = entries with prefix-compressed keys + restart array, each block
  followed by a 1-byte compression type and a masked CRC-32C.  TensorFlow's `BundleWriter` writes it uncompressed.
* values: key "" -> `BundleHeaderProto{num_shards, endianness, version}`; every other key ->
  `BundleEntryProto{dtype=1, shape=2, shard_id=3, offset=4, size=5, crc32c=6}` pointing into a data shard.
* key `_CHECKPOINTABLE_OBJECT_GRAPH` is a scalar DT_STRING tensor (`varint64 length | uint32 masked crc of the lengths |
  bytes`) holding a `TrackableObjectGraph`: nodes with `children{node_id=1, local_name=2}` and
  `attributes{name=1, full_name=2, checkpoint_key=3}`.  Keras names children after the Python attributes, so the
  attribute paths of SURVEY App. B (`encoder.basic_layers.0.blocks.1.attn.qkv.kernel`) are walked from the root.

**Parity unpinned**: TensorFlow is not installable in this environment, so no TF-written checkpoint was available to
read; the reader is tested against files produced by the writer below (same format description), against the format's
known constants (magic, CRC-32C vectors) and against byte vectors assembled by hand from TensorFlow's published rules
(the DT_STRING checksum of `WriteStringTensor`: lengths as fixed-width integers, then the length checksum, then bytes).
"""
from __future__ import annotations

import os
import struct
from typing import Dict, List, Optional, Tuple

import numpy as np

from .records import _enc_ld, _enc_varint, _fields, _varint, crc32c

TABLE_MAGIC = 0xDB4775248B80FB57
OBJECT_GRAPH_KEY = "_CHECKPOINTABLE_OBJECT_GRAPH"
_MASK_DELTA = 0xA282EAD8

# tensorflow/core/framework/types.proto
DTYPES = {1: np.float32, 2: np.float64, 3: np.int32, 4: np.uint8, 5: np.int16, 6: np.int8, 9: np.int64, 10: np.bool_,
          17: np.uint16, 19: np.float16, 22: np.uint32, 23: np.uint64}
DT_STRING, DT_BFLOAT16 = 7, 14
_DTYPE_CODE = {np.dtype(v): k for k, v in DTYPES.items()}


def _mask(c: int) -> int:
    return ((((c >> 15) | (c << 17)) & 0xFFFFFFFF) + _MASK_DELTA) & 0xFFFFFFFF


def string_tensor_crc(values: List[bytes]) -> Tuple[int, int]:
    """Checksums of a DT_STRING tensor as TensorFlow's `WriteStringTensor` / `ReadStringTensor` compute them
    (tensorflow/core/util/tensor_bundle/tensor_bundle.cc): the running CRC-32C covers every element LENGTH as a
    fixed-width little-endian integer (uint32, or uint64 when it exceeds UINT32_MAX) -- not the varint bytes on disk --
    then the 4 bytes of the masked length checksum, then all string bytes.  Returns (masked length checksum stored
    after the varints, masked total stored in BundleEntryProto.crc32c)."""
    c = 0
    for v in values:
        c = crc32c(struct.pack("<I", len(v)) if len(v) <= 0xFFFFFFFF else struct.pack("<Q", len(v)), c)
    len_cksum = _mask(c)
    c = crc32c(struct.pack("<I", len_cksum), c)
    for v in values:
        c = crc32c(v, c)
    return len_cksum, _mask(c)


# ------------------------------------------------------------------------------------------------ SSTable
def _read_block(buf: bytes, offset: int, size: int, verify: bool) -> bytes:
    body, trailer = buf[offset:offset + size], buf[offset + size:offset + size + 5]
    if len(body) != size or len(trailer) != 5:
        raise ValueError("index file: block outside the file")
    if trailer[0] != 0:
        raise ValueError("index file: compressed block (snappy) -- TensorFlow writes the bundle index uncompressed")
    if verify and _mask(crc32c(trailer[:1], crc32c(body))) != struct.unpack("<I", trailer[1:])[0]:
        raise ValueError("index file: block checksum mismatch")
    return body


def _block_entries(block: bytes) -> List[Tuple[bytes, bytes]]:
    (n_restarts,) = struct.unpack("<I", block[-4:])
    end = len(block) - 4 - 4 * n_restarts
    out, pos, key = [], 0, b""
    while pos < end:
        shared, pos = _varint(block, pos)
        non_shared, pos = _varint(block, pos)
        vlen, pos = _varint(block, pos)
        key = key[:shared] + block[pos:pos + non_shared]
        pos += non_shared
        out.append((key, block[pos:pos + vlen]))
        pos += vlen
    return out


def read_table(path: str, verify: bool = True) -> Dict[bytes, bytes]:
    """All key/value pairs of a LevelDB-format table file."""
    buf = open(path, "rb").read()
    if len(buf) < 48 or struct.unpack("<Q", buf[-8:])[0] != TABLE_MAGIC:
        raise ValueError(f"{path}: not a table file (bad magic)")
    footer = buf[-48:]
    pos = 0
    _, pos = _varint(footer, pos)  # metaindex handle (unused)
    _, pos = _varint(footer, pos)
    ioff, pos = _varint(footer, pos)
    isize, pos = _varint(footer, pos)
    out: Dict[bytes, bytes] = {}
    for _, handle in _block_entries(_read_block(buf, ioff, isize, verify)):
        boff, p2 = _varint(handle, 0)
        bsize, _ = _varint(handle, p2)
        for k, v in _block_entries(_read_block(buf, boff, bsize, verify)):
            out[k] = v
    return out


def _write_block(entries: List[Tuple[bytes, bytes]], restart_interval: int = 16) -> bytes:
    body, restarts, prev = bytearray(), [], b""
    for i, (k, v) in enumerate(entries):
        shared = 0
        if i % restart_interval == 0:
            restarts.append(len(body))
        else:
            while shared < min(len(prev), len(k)) and prev[shared] == k[shared]:
                shared += 1
        body += _enc_varint(shared) + _enc_varint(len(k) - shared) + _enc_varint(len(v)) + k[shared:] + v
        prev = k
    if not restarts:
        restarts = [0]
    body += b"".join(struct.pack("<I", r) for r in restarts) + struct.pack("<I", len(restarts))
    return bytes(body)


def write_table(path: str, items: Dict[bytes, bytes], block_entries: int = 64) -> None:
    """Uncompressed LevelDB-format table with sorted keys (what BundleWriter::Finish produces)."""
    keys = sorted(items)
    out, index = bytearray(), []

    def emit(block: bytes) -> Tuple[int, int]:
        off = len(out)
        out.extend(block + b"\x00" + struct.pack("<I", _mask(crc32c(b"\x00", crc32c(block)))))
        return off, len(block)

    for i in range(0, len(keys), block_entries):
        chunk = keys[i:i + block_entries]
        off, size = emit(_write_block([(k, items[k]) for k in chunk]))
        index.append((chunk[-1], _enc_varint(off) + _enc_varint(size)))
    moff, msize = emit(_write_block([]))
    ioff, isize = emit(_write_block(index, restart_interval=1))
    footer = _enc_varint(moff) + _enc_varint(msize) + _enc_varint(ioff) + _enc_varint(isize)
    out.extend(footer + b"\x00" * (40 - len(footer)) + struct.pack("<Q", TABLE_MAGIC))
    open(path, "wb").write(bytes(out))


# ------------------------------------------------------------------------------------------------ tensor bundle
class TensorBundle:
    """Read access to `<prefix>.index` + `<prefix>.data-XXXXX-of-YYYYY`."""

    def __init__(self, prefix: str, verify: bool = True):
        self.prefix, self.verify = prefix, verify
        table = read_table(prefix + ".index", verify)
        if b"" not in table:
            raise ValueError(f"{prefix}.index: no bundle header")
        self.num_shards, self.endianness = 1, 0
        for num, wt, v in _fields(table[b""]):
            if num == 1:
                self.num_shards = v
            elif num == 2:
                self.endianness = v
        if self.endianness != 0:
            raise ValueError("big-endian bundles are not supported")
        self.entries: Dict[str, dict] = {}
        for k, v in table.items():
            if k == b"":
                continue
            e = {"dtype": 0, "shape": [], "shard": 0, "offset": 0, "size": 0, "crc": None}
            for num, wt, val in _fields(v):
                if num == 1:
                    e["dtype"] = val
                elif num == 2:
                    e["shape"] = [next((x for n3, _, x in _fields(dim) if n3 == 1), 0)
                                  for n2, _, dim in _fields(val) if n2 == 2]
                elif num == 3:
                    e["shard"] = val
                elif num == 4:
                    e["offset"] = val
                elif num == 5:
                    e["size"] = val
                elif num == 6:
                    e["crc"] = struct.unpack("<I", val)[0]
                elif num == 7:
                    raise ValueError(f"{k!r}: sliced (partitioned) variables are not supported")
            self.entries[k.decode("utf-8")] = e
        self._shards: Dict[int, np.memmap] = {}

    def keys(self) -> List[str]:
        return sorted(self.entries)

    def _bytes(self, e: dict) -> memoryview:
        s = e["shard"]
        if s not in self._shards:
            self._shards[s] = np.memmap(f"{self.prefix}.data-{s:05d}-of-{self.num_shards:05d}", dtype=np.uint8, mode="r")
        raw = self._shards[s][e["offset"]:e["offset"] + e["size"]]
        if len(raw) != e["size"]:
            raise ValueError("data shard shorter than the index says")
        # DT_STRING entries carry a different checksum (string_tensor_crc): verified in read()
        if (self.verify and e["crc"] is not None and e["dtype"] != DT_STRING
                and _mask(crc32c(np.ascontiguousarray(raw))) != e["crc"]):
            raise ValueError("tensor checksum mismatch")
        return memoryview(np.ascontiguousarray(raw))

    def read(self, key: str):
        """The tensor stored under a checkpoint key: ndarray, or bytes / list of bytes for DT_STRING."""
        e = self.entries[key]
        raw = self._bytes(e)
        n = int(np.prod(e["shape"])) if e["shape"] else 1
        if e["dtype"] == DT_STRING:
            pos, lens = 0, []
            for _ in range(n):
                ln, pos = _varint(raw, pos)
                lens.append(ln)
            (len_cksum,) = struct.unpack("<I", raw[pos:pos + 4])
            pos += 4
            vals = []
            for ln in lens:
                vals.append(bytes(raw[pos:pos + ln]))
                pos += ln
            if self.verify:
                want_len, want_all = string_tensor_crc(vals)
                if len_cksum != want_len or (e["crc"] is not None and e["crc"] != want_all):
                    raise ValueError("string tensor checksum mismatch")
            return vals[0] if not e["shape"] else vals
        if e["dtype"] == DT_BFLOAT16:
            a = np.frombuffer(raw, dtype="<u2").astype(np.uint32) << 16
            return a.view(np.float32).reshape(e["shape"])
        if e["dtype"] not in DTYPES:
            raise ValueError(f"{key}: unsupported dtype code {e['dtype']}")
        return np.frombuffer(raw, dtype=DTYPES[e["dtype"]]).reshape(e["shape"]).copy()

    # ---- object graph -------------------------------------------------------------------------------
    def object_graph(self) -> List[dict]:
        nodes = []
        for num, _, node in _fields(self.read(OBJECT_GRAPH_KEY)):
            if num != 1:
                continue
            children, attrs = {}, {}
            for n2, _, v in _fields(node):
                if n2 == 1:
                    nid, name = 0, ""
                    for n3, _, x in _fields(v):
                        if n3 == 1:
                            nid = x
                        elif n3 == 2:
                            name = bytes(x).decode("utf-8")
                    children[name] = nid
                elif n2 == 2:
                    name, ckey = "", ""
                    for n3, _, x in _fields(v):
                        if n3 == 1:
                            name = bytes(x).decode("utf-8")
                        elif n3 == 3:
                            ckey = bytes(x).decode("utf-8")
                    attrs[name] = ckey
            nodes.append({"children": children, "attributes": attrs})
        return nodes

    def variables_by_attribute_path(self) -> Dict[str, str]:
        """{'encoder.patch_embed_map.proj.kernel': checkpoint key, ...} for every variable reachable from the root
        object by Python attribute names (list elements by index); the shortest path wins (Keras also tracks
        `layer-N` / `layer_with_weights-N` aliases, which are skipped)."""
        nodes = self.object_graph()
        out: Dict[str, str] = {}
        seen = set()
        frontier = [(0, "")]
        while frontier:
            nxt = []
            for nid, path in frontier:
                if nid in seen:
                    continue
                seen.add(nid)
                node = nodes[nid]
                if "VARIABLE_VALUE" in node["attributes"] and path:
                    out[path] = node["attributes"]["VARIABLE_VALUE"]
                for name, cid in sorted(node["children"].items()):
                    if name.startswith(("layer-", "layer_with_weights-", "_")) or name in ("variables", "trainable_variables",
                            "non_trainable_variables", "layers", "keras_api", "optimizer", "metrics", "regularization_losses"):
                        continue
                    nxt.append((cid, f"{path}.{name}" if path else name))
            frontier = nxt
        return out


def load_keras_checkpoint(prefix: str, names: Optional[List[str]] = None, verify: bool = True) -> Dict[str, np.ndarray]:
    """Variables of a Keras TF-format checkpoint keyed by attribute path (SURVEY App. B names).  `names`: the
    parameters the caller needs; a missing one raises KeyError naming it (Keras would silently keep its init)."""
    b = TensorBundle(prefix, verify)
    by_path = b.variables_by_attribute_path()
    wanted = names if names is not None else sorted(by_path)
    out = {}
    for n in wanted:
        if n not in by_path:
            raise KeyError(f"{prefix}: no variable at attribute path '{n}'")
        out[n] = b.read(by_path[n])
    return out


def is_tf_checkpoint(path: str) -> bool:
    return os.path.exists(path + ".index")


# ------------------------------------------------------------------------------------------------ writer (tests / export)
def save_keras_checkpoint(prefix: str, weights: Dict[str, np.ndarray]) -> None:
    """Write {attribute path: array} as a one-shard tensor bundle with the object graph Keras would record."""
    nodes: List[dict] = [{"children": {}, "key": None}]

    def node_for(path: List[str]) -> int:
        cur = 0
        for name in path:
            if name not in nodes[cur]["children"]:
                nodes.append({"children": {}, "key": None})
                nodes[cur]["children"][name] = len(nodes) - 1
            cur = nodes[cur]["children"][name]
        return cur

    items: Dict[bytes, bytes] = {}
    data = bytearray()

    def add_entry(key: str, dtype_code: int, shape, payload: bytes, crc: Optional[int] = None):
        shape_msg = b"".join(_enc_ld(2, _enc_varint((1 << 3) | 0) + _enc_varint(int(d))) for d in shape)
        e = (_enc_varint((1 << 3) | 0) + _enc_varint(dtype_code) + _enc_ld(2, shape_msg) +
             (_enc_varint((4 << 3) | 0) + _enc_varint(len(data)) if len(data) else b"") +
             _enc_varint((5 << 3) | 0) + _enc_varint(len(payload)) +
             _enc_varint((6 << 3) | 5) + struct.pack("<I", _mask(crc32c(payload)) if crc is None else crc))
        items[key.encode("utf-8")] = e
        data.extend(payload)

    for name in sorted(weights):
        a = np.asarray(weights[name])  # (ascontiguousarray would promote scalars to 1-D)
        ckey = name.replace(".", "/") + "/.ATTRIBUTES/VARIABLE_VALUE"
        nodes[node_for(name.split("."))]["key"] = ckey
        add_entry(ckey, _DTYPE_CODE[a.dtype], a.shape, a.tobytes())
    graph = b""
    for nd in nodes:
        msg = b"".join(_enc_ld(1, _enc_varint((1 << 3) | 0) + _enc_varint(cid) + _enc_ld(2, nm.encode("utf-8")))
                       for nm, cid in nd["children"].items())
        if nd["key"]:
            msg += _enc_ld(2, _enc_ld(1, b"VARIABLE_VALUE") + _enc_ld(3, nd["key"].encode("utf-8")))
        graph += _enc_ld(1, msg)
    len_cksum, total = string_tensor_crc([graph])
    add_entry(OBJECT_GRAPH_KEY, DT_STRING, (), _enc_varint(len(graph)) + struct.pack("<I", len_cksum) + graph, crc=total)
    items[b""] = _enc_varint((1 << 3) | 0) + _enc_varint(1) + _enc_ld(3, _enc_varint((1 << 3) | 0) + _enc_varint(1))
    write_table(prefix + ".index", items)
    open(prefix + ".data-00000-of-00001", "wb").write(bytes(data))
