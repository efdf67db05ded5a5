"""mint_b200/tf_checkpoint.py: the TensorBundle layer against a checkpoint written by TensorFlow
itself (its decoded tensors + file hashes: tests/golden/tf_bundle_*), and the FACT variable mapping / object graph
through a write-read round trip."""
import hashlib
import json
import os
import shutil
import struct

import numpy as np
import pytest

from mint_b200 import config_util, tf_checkpoint as T, weights as W
from tests.helpers import make_config


def _dims(**kw):
    return config_util.resolve_fact_dims(make_config(**kw))


def _small_dims():
    return _dims(d=64, heads=4, ff=128, layers=(1, 2, 3), motion_seq=12, audio_seq=20)

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
MANIFEST = json.load(open(os.path.join(GOLD, "tf_bundle_manifest.json")))


@pytest.fixture()
def tf_written(tmp_path):
    """A bundle written by THIS repo's writer from the decoded tensors of a TensorFlow-written checkpoint; both files
    must hash to what TensorFlow wrote (tests/golden/make_tf_bundle_golden.py), so everything read back from it is read
    from TensorFlow's own bytes."""
    tensors = np.load(os.path.join(GOLD, "tf_bundle_tensors.npz"))
    prefix = str(tmp_path / "model-65000")
    w = T.BundleWriter(prefix)
    for k in tensors.files:
        w.add(k, tensors[k])
    w.close()
    for ext, meta in MANIFEST["files"].items():
        raw = open(prefix + ext, "rb").read()
        assert len(raw) == meta["bytes"], ext
        assert hashlib.sha256(raw).hexdigest() == meta["sha256"], f"{ext}: not byte-identical to TensorFlow's file"
    return prefix


def test_writer_reproduces_tensorflows_bytes(tf_written):
    assert os.path.getsize(tf_written + ".index") == 266


def test_reads_tensorflows_bytes_with_every_checksum_verified(tf_written):
    b = T.Bundle(tf_written, verify_table=True)
    assert b.num_shards == MANIFEST["num_shards"] == 1
    assert b.keys() == sorted(MANIFEST["entries"])
    for k, want in MANIFEST["entries"].items():       # the entry table TensorFlow wrote (decoded in the build container)
        e = b.entries[k]
        assert (e.dtype, list(e.shape), e.offset, e.size, e.crc) == \
            (want["dtype"], want["shape"], want["offset"], want["size"], want["masked_crc32c"]), k
    step = b.tensor("global_step", verify=True)
    assert step.dtype == np.int32 and step.shape == () and int(step) == 65000      # the checkpoint is model-65000
    w = b.tensor("Variable", verify=True)
    assert w.dtype == np.float32 and w.shape == (50, 50) and np.isfinite(w).all()
    end = 0                                                                        # offsets tile the data file exactly
    for k in sorted(b.keys(), key=lambda k: b.entries[k].offset):
        assert b.entries[k].offset == end
        end += b.entries[k].size
    assert end == os.path.getsize(tf_written + ".data-00000-of-00001")


def test_corruption_is_detected(tf_written, tmp_path):
    for ext in (".index", ".data-00000-of-00001"):
        shutil.copyfile(tf_written + ext, str(tmp_path / "m") + ext)
    data = bytearray(open(str(tmp_path / "m.data-00000-of-00001"), "rb").read())
    data[123] ^= 0x40
    open(str(tmp_path / "m.data-00000-of-00001"), "wb").write(bytes(data))
    b = T.Bundle(str(tmp_path / "m"))
    with pytest.raises(ValueError, match="checksum"):
        b.tensor("Variable", verify=True)
    assert b.tensor("global_step", verify=True) == 65000                          # other tensors are intact
    idx = bytearray(open(str(tmp_path / "m.index"), "rb").read())
    idx[20] ^= 0x01
    open(str(tmp_path / "m.index"), "wb").write(bytes(idx))
    with pytest.raises(ValueError, match="checksum"):
        T.Bundle(str(tmp_path / "m"))
    open(str(tmp_path / "m.index"), "wb").write(bytes(idx[:-3]))
    with pytest.raises(ValueError):
        T.Bundle(str(tmp_path / "m"))


def test_table_with_many_blocks_and_shared_prefixes(tmp_path):
    rng = np.random.default_rng(0)
    items = {f"model/layer_with_weights-{i}/fn/{'kernel' if i % 3 else 'bias'}{j}".encode():
             bytes(rng.integers(0, 256, int(rng.integers(0, 40)), dtype=np.uint8)) for i in range(40) for j in range(7)}
    items[b""] = b"\x08\x01"
    items[b"\xff\xff"] = b"last"
    path = str(tmp_path / "t.index")
    T.write_table(path, items, entries_per_block=16)
    assert T.read_table(path) == items
    assert list(T.read_table(path)) == sorted(items)


def test_fact_variable_paths_follow_the_keras_object_graph():
    dims = _dims()                                                                # fact_v5
    paths = T.fact_object_paths(dims)
    assert set(paths) == set(W.variable_shapes(dims))
    assert paths["cross_modal_layer/transformer/layer_3/attn/to_qkv/kernel"] == \
        "model/cross_modal_layer/transformer_layer/net/layer_with_weights-6/fn/fn/to_qkv/kernel"
    assert paths["cross_modal_layer/transformer/layer_11/mlp/dense_1/bias"] == \
        "model/cross_modal_layer/transformer_layer/net/layer_with_weights-23/fn/fn/net/layer_with_weights-1/bias"
    assert paths["motion_transformer/layer_1/mlp/norm/gamma"] == "model/motion_transformer/net/layer_with_weights-3/fn/norm/gamma"
    assert paths["audio_transformer/layer_0/attn/to_out/bias"] == "model/audio_transformer/net/layer_with_weights-0/fn/fn/to_out/bias"
    assert paths["motion_pos_embedding"] == "model/motion_pos_embedding/pos_embedding"
    assert paths["audio_linear_embedding/kernel"] == "model/audio_linear_embedding/net/kernel"
    assert paths["cross_modal_layer/output/kernel"] == "model/cross_modal_layer/cross_output_layer/kernel"
    assert len(set(paths.values())) == len(paths)


def test_fact_checkpoint_round_trip(tmp_path):
    dims = _dims(d=64, heads=4, ff=128, layers=(1, 2, 3), motion_seq=12, audio_seq=20)
    rng = np.random.default_rng(1)
    weights = {n: rng.standard_normal(s).astype(np.float32) for n, s in W.variable_shapes(dims).items()}
    prefix = str(tmp_path / "ckpt-7")
    T.save_fact_weights(prefix, weights, dims, step=7000)
    b = T.Bundle(prefix)
    assert T.OBJECT_GRAPH_KEY in b.entries
    keys = [k for k in b.keys() if k.endswith(T.VAR_SUFFIX)]
    assert len(keys) == len(weights) + 1
    assert "model/cross_modal_layer/transformer_layer/net/layer_with_weights-4/fn/fn/to_qkv/kernel" + T.VAR_SUFFIX in keys
    assert int(b.tensor("optimizer/iter" + T.VAR_SUFFIX)) == 7000
    # variables are found by walking attribute names from the root, not by composing key strings
    nodes = b.object_graph()
    assert set(nodes[0][0]) == {"model", "optimizer"}
    key = b.variable_key("model/motion_transformer/net/layer_with_weights-1/fn/fn/net/layer_with_weights-0/kernel", nodes)
    np.testing.assert_array_equal(b.tensor(key), weights["motion_transformer/layer_0/mlp/dense_0/kernel"])
    with pytest.raises(KeyError):
        b.variable_key("model/motion_transformer/net/layer_with_weights-9/fn/norm/gamma", nodes)
    loaded = T.load_fact_weights(prefix, dims, verify=True)
    assert set(loaded) == set(weights)
    for n in weights:
        np.testing.assert_array_equal(loaded[n], weights[n])
    wrong = _dims(d=64, heads=4, ff=256, layers=(1, 2, 3), motion_seq=12, audio_seq=20)
    with pytest.raises(ValueError, match="shape"):
        T.load_fact_weights(prefix, wrong)


def test_latest_checkpoint_follows_the_state_file_then_the_numbers(tmp_path):
    d = str(tmp_path)
    assert T.latest_checkpoint(d) is None
    for n in (1000, 3000, 2000):
        open(os.path.join(d, f"ckpt-{n}.index"), "wb").close()
    assert T.latest_checkpoint(d) == os.path.join(d, "ckpt-3000")
    T.write_checkpoint_state(d, [os.path.join(d, "ckpt-1000"), os.path.join(d, "ckpt-2000")])
    assert open(os.path.join(d, "checkpoint")).read().splitlines()[0] == 'model_checkpoint_path: "ckpt-2000"'
    assert T.latest_checkpoint(d) == os.path.join(d, "ckpt-2000")          # the state file wins, as in TF
    os.remove(os.path.join(d, "ckpt-2000.index"))
    assert T.latest_checkpoint(d) == os.path.join(d, "ckpt-3000")          # stale state file: fall back


# ------------------------------------------------------- object-graph layer vs TensorFlow's own generated proto module
def _tf_graph_pb2():
    """tensorflow/core/protobuf/trackable_object_graph.proto as generated by TensorFlow's build, shipped in tensorboard."""
    return pytest.importorskip("tensorboard.compat.proto.trackable_object_graph_pb2")


def test_dtype_numbers_match_tensorflows_types_proto():
    types_pb2 = pytest.importorskip("tensorboard.compat.proto.types_pb2")
    for name in ("DT_FLOAT", "DT_DOUBLE", "DT_INT32", "DT_STRING", "DT_INT64", "DT_BOOL", "DT_BFLOAT16", "DT_HALF"):
        assert getattr(T, name) == getattr(types_pb2, name), name


def test_written_object_graph_parses_with_tensorflows_proto_class(tmp_path):
    pb = _tf_graph_pb2()
    dims = _small_dims()
    w = {n: np.random.default_rng(0).standard_normal(s).astype(np.float32) for n, s in W.variable_shapes(dims).items()}
    prefix = str(tmp_path / "ckpt-7")
    T.save_fact_weights(prefix, w, dims, step=7)
    b = T.Bundle(prefix)
    raw = b.tensor(T.OBJECT_GRAPH_KEY).reshape(-1)[0]          # both string checksums verified on the way
    g = pb.TrackableObjectGraph.FromString(raw)
    # walk it with TensorFlow's message classes exactly as the checkpoint loader would: by local_name from the root
    paths = T.fact_object_paths(dims)
    for name, path in list(paths.items()) + [("iter", "optimizer/iter")]:
        node = g.nodes[0]
        for part in path.split("/"):
            nxt = [c for c in node.children if c.local_name == part]
            assert len(nxt) == 1, (path, part)
            node = g.nodes[nxt[0].node_id]
        attr = [a for a in node.attributes if a.name == "VARIABLE_VALUE"]
        assert len(attr) == 1 and attr[0].checkpoint_key == path + T.VAR_SUFFIX
        assert attr[0].full_name == path
        if name != "iter":
            np.testing.assert_array_equal(b.tensor(attr[0].checkpoint_key), w[name])
    # and the bytes are canonical: TensorFlow's class re-serialises them identically
    assert g.SerializeToString() == raw


def test_reader_on_a_hostile_graph_serialised_by_tensorflows_proto_class(tmp_path):
    """A graph as tf.train.Checkpoint(optimizer=, model=) lays it out, built with TensorFlow's message classes: children
    in shuffled order, optimizer slot variables (m / v per weight), a registered_saver, has_checkpoint_values wrappers,
    node ids that are not in path order, and unknown fields at every level.  The reader must still resolve every model
    variable through the attribute path and ignore the rest."""
    pb = _tf_graph_pb2()
    dims = _small_dims()
    rng = np.random.default_rng(3)
    w = {n: rng.standard_normal(s).astype(np.float32) for n, s in W.variable_shapes(dims).items()}
    paths = T.fact_object_paths(dims)
    # build the trie with node ids assigned in a scrambled order
    trie = {"": {}}
    for path in list(paths.values()) + ["optimizer/iter", "optimizer/beta_1", "save_counter"]:
        parts = path.split("/")
        for i in range(len(parts)):
            trie.setdefault("/".join(parts[:i + 1]), {})
            trie["/".join(parts[:i])][parts[i]] = "/".join(parts[:i + 1])
    ids = [p for p in trie if p != ""]
    rng.shuffle(ids)
    ids = [""] + ids                                                     # the root stays node 0
    slot_base = len(ids)
    index = {p: i for i, p in enumerate(ids)}
    g = pb.TrackableObjectGraph()
    for _ in range(slot_base):
        g.nodes.add()
    key_of = {}
    for p, i in index.items():
        node = g.nodes[i]
        kids = list(trie[p].items())
        rng.shuffle(kids)
        for local, child in kids:
            c = node.children.add()
            c.node_id, c.local_name = index[child], local
        if not trie[p] and p:                                            # a leaf: a variable
            a = node.attributes.add()
            a.name, a.full_name, a.checkpoint_key = "VARIABLE_VALUE", "some/graph/name:0", p + T.VAR_SUFFIX
            key_of[p] = a.checkpoint_key
            node.has_checkpoint_values.value = True
    # optimizer slot variables hang off the optimizer node and point at extra nodes past the trie
    opt = g.nodes[index["optimizer"]]
    opt.registered_saver.name = "Custom>Saver"
    opt.registered_saver.object_name = "optimizer"
    slots = {}
    for name, path in list(paths.items())[:6]:
        for slot in ("m", "v"):
            n = g.nodes.add()
            a = n.attributes.add()
            a.name = "VARIABLE_VALUE"
            a.checkpoint_key = f"{path}/.OPTIMIZER_SLOT/optimizer/{slot}{T.VAR_SUFFIX}"
            s = opt.slot_variables.add()
            s.original_variable_node_id, s.slot_name, s.slot_variable_node_id = index[path], slot, len(g.nodes) - 1
            slots[a.checkpoint_key] = np.zeros(w[name].shape, np.float32)
    raw = g.SerializeToString()
    # unknown fields: one appended to the graph, one spliced into the root node (field 15 varint, field 14 bytes)
    root = g.nodes[0].SerializeToString() + T._tag(15, 0) + T._put_varint(99) + T._ld(14, b"future")
    rest = b"".join(T._ld(1, n.SerializeToString()) for n in g.nodes[1:])
    raw_hostile = T._ld(1, root) + rest + T._tag(9, 5) + (1234).to_bytes(4, "little")
    assert pb.TrackableObjectGraph.FromString(raw_hostile).nodes[3] == g.nodes[3]      # still valid protobuf for TF
    prefix = str(tmp_path / "ckpt-3")
    bw = T.BundleWriter(prefix)
    for name, path in paths.items():
        bw.add(key_of[path], w[name])
    for k, v in slots.items():
        bw.add(k, v)
    bw.add(key_of["optimizer/iter"], np.asarray(3, np.int64))
    bw.add(key_of["optimizer/beta_1"], np.asarray(0.9, np.float32))
    bw.add(key_of["save_counter"], np.asarray(1, np.int64))
    bw.add_string_scalar(T.OBJECT_GRAPH_KEY, raw_hostile)
    bw.close()
    back = T.load_fact_weights(prefix, dims, verify=True)
    assert set(back) == set(w)
    for name in w:
        np.testing.assert_array_equal(back[name], w[name])
    b = T.Bundle(prefix)
    nodes = b.object_graph()
    assert len(nodes) == len(g.nodes)
    with pytest.raises(KeyError):
        b.variable_key("model/no_such_layer/kernel", nodes)
    with pytest.raises(KeyError):
        b.variable_key("model", nodes)                                   # an inner node is not a variable


def test_string_tensor_checksums_are_verified(tmp_path):
    prefix = str(tmp_path / "s")
    bw = T.BundleWriter(prefix)
    bw.add_string_scalar("k", b"hello object graph")
    bw.close()
    b = T.Bundle(prefix)
    assert b.tensor("k").reshape(-1)[0] == b"hello object graph"
    e = b.entries["k"]
    # WriteStringTensor: the length enters the checksums as a uint32
    assert e.size == 1 + 4 + 18
    data = b._raw(e)
    assert data[1:5] == struct.pack("<I", T.masked_crc(struct.pack("<I", 18)))
    assert e.crc == T.masked_crc(struct.pack("<I", 18) + data[1:5] + b"hello object graph")
    path = prefix + ".data-00000-of-00001"
    raw = bytearray(open(path, "rb").read())
    raw[-1] ^= 1
    open(path, "wb").write(bytes(raw))
    with pytest.raises(ValueError):
        T.Bundle(prefix).tensor("k")


def test_tfrecord_framing_agrees_with_tensorboards_reader_and_writer(tmp_path):
    """tensorboard (TensorFlow team's code, in the image) carries its own TFRecord writer and reader: frames written by
    either side are read by the other, byte-identical for the same payloads."""
    rw = pytest.importorskip("tensorboard.summary.writer.record_writer")
    pw = pytest.importorskip("tensorboard.compat.tensorflow_stub.pywrap_tensorflow")
    from mint_b200 import inputs
    payloads = [b"", b"x", bytes(range(256)) * 9, np.random.default_rng(1).bytes(70001)]
    ours, theirs = tmp_path / "ours.record", tmp_path / "theirs.record"
    with inputs.TFRecordWriter(str(ours)) as wtr:
        for p in payloads:
            wtr.write(p)
    f = open(theirs, "wb")
    tw = rw.RecordWriter(f)
    for p in payloads:
        tw.write(p)
    tw.close()
    assert ours.read_bytes() == theirs.read_bytes()
    assert list(inputs.read_tfrecords(str(theirs), verify_payload_crc=True)) == payloads
    rd = pw.PyRecordReader_New(str(ours))
    got = []
    while True:
        try:
            rd.GetNext()
        except Exception:
            break
        got.append(rd.record())
    assert got == payloads
    for p in payloads[1:]:
        assert inputs.masked_crc(p) == pw.masked_crc32c(p)
