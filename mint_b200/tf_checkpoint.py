"""TensorFlow checkpoint (TensorBundle V2) import / export for the FACT weights -- SURVEY.md §8(f) row N3.

The reference saves `tf.train.Checkpoint(optimizer=, model=)` through a `CheckpointManager` (trainer.py:168-173) and
the released weights are such a checkpoint (README.md:39-40): `ckpt-N.index` (an SSTable: key -> BundleEntryProto)
plus `ckpt-N.data-00000-of-00001` (raw little-endian tensor bytes).  TensorFlow is not installable in this image, so
this module reads and writes the format itself (host-side, NumPy):

  * SSTable (LevelDB table format): prefix-compressed key blocks with restart arrays, 5-byte block trailers (type +
    masked crc32c), index block, 48-byte footer with magic 0xdb4775248b80fb57;
  * BundleHeaderProto under the empty key, one BundleEntryProto (dtype, shape, shard, offset, size, masked crc32c)
    per tensor;
  * the TF2 object graph (`_CHECKPOINTABLE_OBJECT_GRAPH`, a serialised TrackableObjectGraph): variables are found by
    walking attribute names from the root (`model` -> `motion_transformer` -> `net` -> `layer_with_weights-0` -> ...),
    which is how Keras resolves them, instead of trusting key strings.

Pinning: the table / entry / tensor layer is checked against a REAL TensorFlow-written bundle that ships inside the
reference tree (third_party/tf_models/research/lfads/synth_data/trained_itb/model-65000.*: every block and tensor
checksum verifies and the writer reproduces its `.index` and `.data` files byte for byte -- the committed golden is
the decoded tensors plus TensorFlow's file hashes, tests/golden/make_tf_bundle_golden.py).  That file is a
TF1 Saver bundle without an object graph.  The object-graph MESSAGE layer is pinned to TensorFlow's own generated
protobuf module for tensorflow/core/protobuf/trackable_object_graph.proto, which ships in this image inside
tensorboard (`tensorboard.compat.proto.trackable_object_graph_pb2`): tests/test_tf_checkpoint.py parses the bytes this
module writes with that class, and feeds this module's reader graphs serialised by it (slot variables, registered
savers, children in any order, fields this reader does not know).  Field numbers, from that descriptor:
    TrackableObjectGraph.nodes = 1 (repeated TrackableObject)
    TrackableObject: children = 1 (ObjectReference), attributes = 2 (SerializedTensor), slot_variables = 3,
                     registered_saver = 4, has_checkpoint_values = 5
    ObjectReference: node_id = 1 (int32), local_name = 2
    SerializedTensor: name = 1, full_name = 2, checkpoint_key = 3
    SlotVariableReference: original_variable_node_id = 1, slot_name = 2, slot_variable_node_id = 3
DataType numbers: tensorflow/core/framework/types.proto (same check against tensorboard.compat.proto.types_pb2).
What remains from the published TensorFlow sources only (tensor_bundle.cc WriteStringTensor) is the string-tensor
framing that carries the graph: varint64 lengths | masked crc32c of the lengths taken as uint32 each (uint64 only
above 4 GiB) | bytes, entry checksum = masked running crc over lengths, length checksum and bytes.  No
TensorFlow-written TF2 checkpoint exists in the reference tree or the image: **loading the released FACT checkpoint
is untested**; INTEGRATION.md says so.
"""
from __future__ import annotations

import os
import struct

import numpy as np

from .inputs import crc32c, masked_crc

MAGIC = 0xDB4775248B80FB57
OBJECT_GRAPH_KEY = "_CHECKPOINTABLE_OBJECT_GRAPH"
VAR_SUFFIX = "/.ATTRIBUTES/VARIABLE_VALUE"

# tensorflow/core/framework/types.proto
DT_FLOAT, DT_DOUBLE, DT_INT32, DT_STRING, DT_INT64, DT_BOOL, DT_BFLOAT16, DT_HALF = 1, 2, 3, 7, 9, 10, 14, 19
_NP = {DT_FLOAT: np.dtype("<f4"), DT_DOUBLE: np.dtype("<f8"), DT_INT32: np.dtype("<i4"), DT_INT64: np.dtype("<i8"),
       DT_BOOL: np.dtype("?"), DT_HALF: np.dtype("<f2")}
_DT = {np.dtype("float32"): DT_FLOAT, np.dtype("float64"): DT_DOUBLE, np.dtype("int32"): DT_INT32,
       np.dtype("int64"): DT_INT64, np.dtype("bool"): DT_BOOL, np.dtype("float16"): DT_HALF}


# ------------------------------------------------------------------------------------------- varints / tiny protobuf
def _varint(buf: bytes, pos: int):
    shift = result = 0
    while True:
        b = buf[pos]
        pos += 1
        result |= (b & 0x7F) << shift
        if not b & 0x80:
            return result, pos
        shift += 7


def _put_varint(v: int) -> bytes:
    out = bytearray()
    while True:
        b = v & 0x7F
        v >>= 7
        if v:
            out.append(b | 0x80)
        else:
            out.append(b)
            return bytes(out)


def _fields(buf: bytes):
    """Yield (field number, wire type, value) of one protobuf message (varint, fixed32/64 and length-delimited)."""
    pos = 0
    while pos < len(buf):
        tag, pos = _varint(buf, pos)
        num, wt = tag >> 3, tag & 7
        if wt == 0:
            val, pos = _varint(buf, pos)
        elif wt == 1:
            val, pos = struct.unpack_from("<Q", buf, pos)[0], pos + 8
        elif wt == 2:
            n, pos = _varint(buf, pos)
            val, pos = buf[pos:pos + n], pos + n
        elif wt == 5:
            val, pos = struct.unpack_from("<I", buf, pos)[0], pos + 4
        else:
            raise ValueError(f"unsupported protobuf wire type {wt}")
        yield num, wt, val


def _tag(num: int, wt: int) -> bytes:
    return _put_varint((num << 3) | wt)


def _ld(num: int, payload: bytes) -> bytes:
    return _tag(num, 2) + _put_varint(len(payload)) + payload


# ------------------------------------------------------------------------------------------- SSTable
def _mask_ok(stored: int, data: bytes) -> bool:
    return stored == masked_crc(data)


def _read_block(buf: bytes, offset: int, size: int, verify: bool) -> bytes:
    body, trailer = buf[offset:offset + size], buf[offset + size:offset + size + 5]
    if len(trailer) != 5:
        raise ValueError("truncated table block")
    if trailer[0] != 0:
        raise ValueError("compressed table block (TensorBundle writes uncompressed blocks)")
    if verify and not _mask_ok(struct.unpack("<I", trailer[1:])[0], body + trailer[:1]):
        raise ValueError(f"table block at {offset}: checksum mismatch")
    return body


def _block_entries(block: bytes):
    n_restarts = struct.unpack_from("<I", block, len(block) - 4)[0]
    end = len(block) - 4 - 4 * n_restarts
    pos, key = 0, b""
    while pos < end:
        shared, pos = _varint(block, pos)
        non_shared, pos = _varint(block, pos)
        vlen, pos = _varint(block, pos)
        key = key[:shared] + block[pos:pos + non_shared]
        pos += non_shared
        yield key, block[pos:pos + vlen]
        pos += vlen


def read_table(path: str, verify: bool = True) -> dict:
    """All (key, value) pairs of an SSTable file, in key order."""
    buf = open(path, "rb").read()
    if len(buf) < 48 or struct.unpack("<Q", buf[-8:])[0] != MAGIC:
        raise ValueError(f"{path}: not a table file (bad magic)")
    footer = buf[-48:]
    _, p = _varint(footer, 0)          # metaindex handle (unused)
    _, p = _varint(footer, p)
    idx_off, p = _varint(footer, p)
    idx_size, p = _varint(footer, p)
    out = {}
    for _, handle in _block_entries(_read_block(buf, idx_off, idx_size, verify)):
        off, q = _varint(handle, 0)
        size, _ = _varint(handle, q)
        for k, v in _block_entries(_read_block(buf, off, size, verify)):
            out[k] = v
    return out


def _build_block(items, restart_interval: int = 16) -> bytes:
    out, restarts, prev = bytearray(), [], b""
    for i, (k, v) in enumerate(items):
        shared = 0
        if i % restart_interval == 0:
            restarts.append(len(out))
        else:
            while shared < min(len(prev), len(k)) and prev[shared] == k[shared]:
                shared += 1
        out += _put_varint(shared) + _put_varint(len(k) - shared) + _put_varint(len(v)) + k[shared:] + v
        prev = k
    if not restarts:
        restarts = [0]
    for r in restarts:
        out += struct.pack("<I", r)
    out += struct.pack("<I", len(restarts))
    return bytes(out)


def _separator(last: bytes, nxt: bytes | None) -> bytes:
    """Index key of a data block, as LevelDB's bytewise comparator shortens it: a short key k with
    last <= k < next (FindShortestSeparator), or a short successor of the table's final key (FindShortSuccessor)."""
    if nxt is None:
        for i, b in enumerate(last):
            if b != 0xFF:
                return last[:i] + bytes([b + 1])
        return last
    n = min(len(last), len(nxt))
    d = 0
    while d < n and last[d] == nxt[d]:
        d += 1
    if d < n and last[d] < 0xFF and last[d] + 1 < nxt[d]:
        return last[:d] + bytes([last[d] + 1])
    return last


def write_table(path: str, items: dict, entries_per_block: int = 64) -> None:
    keys = sorted(items)
    out, index = bytearray(), []

    def emit(block: bytes):
        off = len(out)
        out.extend(block + b"\x00" + struct.pack("<I", masked_crc(block + b"\x00")))
        return _put_varint(off) + _put_varint(len(block))

    for i in range(0, len(keys), entries_per_block):
        chunk = keys[i:i + entries_per_block]
        nxt = keys[i + entries_per_block] if i + entries_per_block < len(keys) else None
        index.append((_separator(chunk[-1], nxt), emit(_build_block([(k, items[k]) for k in chunk]))))
    meta = emit(_build_block([]))
    idx = emit(_build_block(index, restart_interval=1))
    footer = (meta + idx).ljust(40, b"\x00") + struct.pack("<Q", MAGIC)
    with open(path, "wb") as f:
        f.write(bytes(out) + footer)


# ------------------------------------------------------------------------------------------- bundle entries / tensors
class Entry:
    __slots__ = ("dtype", "shape", "shard", "offset", "size", "crc")

    def __init__(self, dtype, shape, shard, offset, size, crc):
        self.dtype, self.shape, self.shard, self.offset, self.size, self.crc = dtype, shape, shard, offset, size, crc

    def __repr__(self):
        return f"Entry(dtype={self.dtype}, shape={self.shape}, shard={self.shard}, offset={self.offset}, size={self.size})"


def _parse_entry(buf: bytes) -> Entry:
    dtype = shard = offset = size = crc = 0
    shape = []
    for num, _, val in _fields(buf):
        if num == 1:
            dtype = val
        elif num == 2:                                   # TensorShapeProto: repeated Dim dim = 2 { int64 size = 1 }
            for n2, _, dim in _fields(val):
                if n2 == 2:
                    shape.append(next((v for n3, _, v in _fields(dim) if n3 == 1), 0))
        elif num == 3:
            shard = val
        elif num == 4:
            offset = val
        elif num == 5:
            size = val
        elif num == 6:
            crc = val
        elif num == 7:
            raise ValueError("sliced (partitioned) variables are not supported")
    return Entry(dtype, tuple(shape), shard, offset, size, crc)


def _entry_bytes(e: Entry) -> bytes:
    shape = b"".join(_ld(2, _tag(1, 0) + _put_varint(d)) for d in e.shape)
    out = _tag(1, 0) + _put_varint(e.dtype) + _ld(2, shape)
    if e.shard:
        out += _tag(3, 0) + _put_varint(e.shard)
    if e.offset:
        out += _tag(4, 0) + _put_varint(e.offset)
    out += _tag(5, 0) + _put_varint(e.size) + _tag(6, 5) + struct.pack("<I", e.crc)
    return out


def _length_word(n: int) -> bytes:
    """How WriteStringTensor feeds a string length to the checksum."""
    return struct.pack("<I", n) if n <= 0xFFFFFFFF else struct.pack("<Q", n)


class Bundle:
    """Reader for `<prefix>.index` + `<prefix>.data-XXXXX-of-YYYYY`."""

    def __init__(self, prefix: str, verify_table: bool = True):
        self.prefix = prefix
        raw = read_table(prefix + ".index", verify=verify_table)
        header = raw.pop(b"", b"")
        self.num_shards = next((v for n, _, v in _fields(header) if n == 1), 1) or 1
        if next((v for n, _, v in _fields(header) if n == 2), 0) != 0:
            raise ValueError("big-endian bundles are not supported")
        self.entries = {k.decode(): _parse_entry(v) for k, v in raw.items()}

    def keys(self):
        return sorted(self.entries)

    def _raw(self, e: Entry) -> bytes:
        path = f"{self.prefix}.data-{e.shard:05d}-of-{self.num_shards:05d}"
        with open(path, "rb") as f:
            f.seek(e.offset)
            data = f.read(e.size)
        if len(data) != e.size:
            raise ValueError(f"{path}: truncated tensor")
        return data

    def tensor(self, key: str, verify: bool | None = None) -> np.ndarray:
        """The tensor stored under `key`.  verify=None checks the checksum of tensors up to 8 MB (the pure-Python
        crc32c runs at a few MB/s); True / False force it."""
        e = self.entries[key]
        data = self._raw(e)
        if e.dtype == DT_STRING:
            return self._strings(e, data)
        if e.dtype == DT_BFLOAT16:
            raise ValueError("bfloat16 tensors are not expected in a FACT checkpoint")
        if e.dtype not in _NP:
            raise ValueError(f"{key}: unsupported dtype {e.dtype}")
        if verify or (verify is None and e.size <= (8 << 20)):
            if not _mask_ok(e.crc, data):
                raise ValueError(f"{key}: tensor checksum mismatch")
        arr = np.frombuffer(data, dtype=_NP[e.dtype])
        n = int(np.prod(e.shape)) if e.shape else 1
        if arr.size != n:
            raise ValueError(f"{key}: {arr.size} elements for shape {e.shape}")
        return arr.reshape(e.shape).copy()

    @staticmethod
    def _strings(e: Entry, data: bytes, verify: bool = True):
        """tensor_bundle.cc ReadStringTensor: lengths, their checksum, bytes; both checksums verified."""
        n = int(np.prod(e.shape)) if e.shape else 1
        lens, pos = [], 0
        for _ in range(n):
            v, pos = _varint(data, pos)
            lens.append(v)
        len_bytes = b"".join(_length_word(v) for v in lens)
        stored = data[pos:pos + 4]
        if verify and (len(stored) != 4 or struct.unpack("<I", stored)[0] != masked_crc(len_bytes)):
            raise ValueError("string tensor: length checksum mismatch")
        pos += 4
        out = []
        for ln in lens:
            if pos + ln > len(data):
                raise ValueError("string tensor: truncated")
            out.append(data[pos:pos + ln])
            pos += ln
        if verify and not _mask_ok(e.crc, len_bytes + stored + b"".join(out)):
            raise ValueError("string tensor: checksum mismatch")
        return np.array(out, dtype=object).reshape(e.shape)

    # ---- TF2 object graph
    def object_graph(self):
        """[(children {local_name: node_id}, attributes {name: checkpoint_key})] per node; node 0 is the root."""
        raw = self.tensor(OBJECT_GRAPH_KEY).reshape(-1)[0]
        nodes = []
        for num, _, node in _fields(raw):
            if num != 1:
                continue
            children, attrs = {}, {}
            for n2, _, val in _fields(node):
                if n2 == 1:
                    f = {n3: v for n3, _, v in _fields(val)}
                    children[f.get(2, b"").decode()] = f.get(1, 0)
                elif n2 == 2:
                    f = {n3: v for n3, _, v in _fields(val)}
                    attrs[f.get(1, b"").decode()] = f.get(3, b"").decode()
            nodes.append((children, attrs))
        return nodes

    def variable_key(self, path: str, nodes=None) -> str:
        """Checkpoint key of the variable reached from the root through the attribute path `a/b/c`."""
        nodes = nodes if nodes is not None else self.object_graph()
        node = 0
        for part in path.split("/"):
            children = nodes[node][0]
            if part not in children:
                raise KeyError(f"object graph has no edge {part!r} on the way to {path!r} (has {sorted(children)[:8]})")
            node = children[part]
        attrs = nodes[node][1]
        if "VARIABLE_VALUE" not in attrs:
            raise KeyError(f"{path}: not a variable")
        return attrs["VARIABLE_VALUE"]


class BundleWriter:
    """Single-shard writer; `object_paths` (attribute path -> tensor name) also emits a TF2 object graph."""

    def __init__(self, prefix: str):
        self.prefix = prefix
        self._items = {}
        self._data = bytearray()

    def add(self, key: str, array) -> None:
        arr = np.asarray(array)
        if not arr.flags.c_contiguous:                   # (np.ascontiguousarray would turn a scalar into shape (1,))
            arr = arr.copy(order="C")
        if arr.dtype not in _DT:
            raise ValueError(f"{key}: dtype {arr.dtype} not supported")
        data = arr.astype(arr.dtype.newbyteorder("<")).tobytes()
        self._items[key.encode()] = _entry_bytes(Entry(_DT[arr.dtype], arr.shape, 0, len(self._data), len(data),
                                                       masked_crc(data)))
        self._data += data

    def add_string_scalar(self, key: str, value: bytes) -> None:
        """tensor_bundle.cc WriteStringTensor for a scalar: the running crc takes each length as a uint32 (a uint64
        only above UINT32_MAX), then the 4 bytes of the masked length checksum, then the string bytes."""
        lens = _put_varint(len(value))
        word = _length_word(len(value))
        len_ck = struct.pack("<I", masked_crc(word))
        data = lens + len_ck + value
        self._items[key.encode()] = _entry_bytes(Entry(DT_STRING, (), 0, len(self._data), len(data),
                                                       masked_crc(word + len_ck + value)))
        self._data += data

    def add_object_graph(self, path_to_key: dict) -> None:
        """Trie of attribute names -> TrackableObjectGraph (children = 1 {node_id = 1, local_name = 2},
        attributes = 2 {name = 1, full_name = 2, checkpoint_key = 3})."""
        nodes = [({}, None)]                              # (children name -> id, checkpoint key)
        for path in sorted(path_to_key):
            node = 0
            for part in path.split("/"):
                kids = nodes[node][0]
                if part not in kids:
                    kids[part] = len(nodes)
                    nodes.append(({}, None))
                node = kids[part]
            nodes[node] = (nodes[node][0], path_to_key[path])
        graph = b""
        for kids, key in nodes:
            body = b"".join(_ld(1, _tag(1, 0) + _put_varint(i) + _ld(2, name.encode())) for name, i in kids.items())
            if key is not None:
                body += _ld(2, _ld(1, b"VARIABLE_VALUE") + _ld(2, key.split("/.ATTRIBUTES")[0].encode()) +
                            _ld(3, key.encode()))
            graph += _ld(1, body)
        self.add_string_scalar(OBJECT_GRAPH_KEY, graph)

    def close(self) -> None:
        header = _tag(1, 0) + _put_varint(1) + _ld(3, _tag(1, 0) + _put_varint(1))   # num_shards 1, version.producer 1
        items = dict(self._items)
        items[b""] = header
        os.makedirs(os.path.dirname(os.path.abspath(self.prefix)), exist_ok=True)
        with open(self.prefix + ".data-00000-of-00001", "wb") as f:
            f.write(bytes(self._data))
        write_table(self.prefix + ".index", items)


# ------------------------------------------------------------------------------------------- FACT variable mapping
def fact_object_paths(dims) -> dict:
    """mint_b200 weight name -> attribute path below the checkpoint root (`model/...`), following how Keras tracks the
    reference's layers: Transformer.net = Sequential([Residual(Norm(Attention)), Residual(Norm(MLP)), ...])
    (base_models.py:20-109) gives layer_with_weights-(2l) / (2l+1); Residual.fn -> Norm; Norm.norm / Norm.fn;
    Attention.to_qkv / to_out; MLP.net = Sequential([Dense, Dense]); LinearEmbedding.net; PositionEmbedding
    .pos_embedding (base_models.py:130-156); CrossModalLayer.transformer_layer / cross_output_layer (:159-181)."""
    from . import weights as W
    out = {}
    stacks = {"motion_transformer": "model/motion_transformer", "audio_transformer": "model/audio_transformer",
              "cross_modal_layer/transformer": "model/cross_modal_layer/transformer_layer"}
    for name in W.variable_shapes(dims):
        if name in ("motion_pos_embedding", "audio_pos_embedding"):
            out[name] = f"model/{name}/pos_embedding"
            continue
        if name.startswith(("motion_linear_embedding/", "audio_linear_embedding/")):
            mod, leaf = name.split("/")
            out[name] = f"model/{mod}/net/{leaf}"
            continue
        if name.startswith("cross_modal_layer/output/"):
            out[name] = "model/cross_modal_layer/cross_output_layer/" + name.rsplit("/", 1)[1]
            continue
        for prefix, root in stacks.items():
            if name.startswith(prefix + "/layer_"):
                rest = name[len(prefix) + 1:]             # layer_3/attn/to_qkv/kernel
                layer, block, *tail = rest.split("/")
                li = int(layer.split("_")[1])
                k = 2 * li + (0 if block == "attn" else 1)
                base = f"{root}/net/layer_with_weights-{k}/fn"
                if tail[0] == "norm":
                    out[name] = f"{base}/norm/{tail[1]}"
                elif block == "attn":
                    out[name] = f"{base}/fn/{tail[0]}/{tail[1]}"
                else:                                     # mlp/dense_0|dense_1/kernel|bias
                    out[name] = f"{base}/fn/net/layer_with_weights-{int(tail[0].split('_')[1])}/{tail[1]}"
                break
        else:
            raise KeyError(f"no checkpoint path rule for {name}")
    return out


def load_fact_weights(prefix: str, dims, verify: bool | None = None) -> dict:
    """name -> float32 array (Keras layout) for `FACTModel.set_weights`, resolved through the object graph."""
    from . import weights as W
    bundle = Bundle(prefix)
    nodes = bundle.object_graph()
    shapes = W.variable_shapes(dims)
    out = {}
    for name, path in fact_object_paths(dims).items():
        arr = bundle.tensor(bundle.variable_key(path, nodes), verify=verify)
        if tuple(arr.shape) != tuple(shapes[name]):
            raise ValueError(f"{name}: checkpoint shape {arr.shape} != model shape {tuple(shapes[name])}")
        out[name] = arr.astype(np.float32)
    return out


def save_fact_weights(prefix: str, weights: dict, dims, step: int | None = None) -> None:
    """Write `weights` (name -> array, Keras layout) as a TF2-style checkpoint the reference's
    `tf.train.Checkpoint(model=...)` layout describes (`model/.../.ATTRIBUTES/VARIABLE_VALUE` keys + object graph)."""
    paths = fact_object_paths(dims)
    missing = set(paths) - set(weights)
    if missing:
        raise KeyError(f"missing weights: {sorted(missing)[:3]}")
    w = BundleWriter(prefix)
    path_to_key = {}
    for name, path in paths.items():
        key = path + VAR_SUFFIX
        w.add(key, np.asarray(weights[name], dtype=np.float32))
        path_to_key[path] = key
    if step is not None:
        key = "optimizer/iter" + VAR_SUFFIX
        w.add(key, np.asarray(step, dtype=np.int64))
        path_to_key["optimizer/iter"] = key
    w.add_object_graph(path_to_key)
    w.close()


def latest_checkpoint(model_dir: str):
    """Prefix of the newest TF checkpoint in `model_dir`, or None: the `checkpoint` state file CheckpointManager
    maintains (`model_checkpoint_path: "ckpt-12"`, what tf.train.latest_checkpoint reads), else the highest-numbered
    `ckpt-N.index`."""
    import re
    state = os.path.join(model_dir, "checkpoint")
    if os.path.isfile(state):
        m = re.search(r'^model_checkpoint_path:\s*"([^"]+)"', open(state).read(), flags=re.M)
        if m:
            p = m.group(1)
            p = p if os.path.isabs(p) else os.path.join(model_dir, p)
            if os.path.isfile(p + ".index"):
                return p
    best = None
    if os.path.isdir(model_dir):
        for f in os.listdir(model_dir):
            m = re.fullmatch(r"(.+-(\d+))\.index", f)
            if m and (best is None or int(m.group(2)) > best[0]):
                best = (int(m.group(2)), os.path.join(model_dir, m.group(1)))
    return best[1] if best else None


def write_checkpoint_state(model_dir: str, prefixes: list) -> None:
    """The `checkpoint` text file next to the bundles (newest last), as CheckpointManager writes it."""
    names = [os.path.basename(p) for p in prefixes]
    with open(os.path.join(model_dir, "checkpoint"), "w") as f:
        f.write(f'model_checkpoint_path: "{names[-1]}"\n')
        for n in names:
            f.write(f'all_model_checkpoint_paths: "{n}"\n')


def _main(argv):
    if len(argv) != 2:
        print("usage: python -m mint_b200.tf_checkpoint <checkpoint prefix>")
        return 2
    b = Bundle(argv[1])
    for k in b.keys():
        print(k, b.entries[k])
    return 0


if __name__ == "__main__":
    import sys
    raise SystemExit(_main(sys.argv))
