"""mint_b200/tf_checkpoint.py: the TensorBundle layer against a checkpoint written by TensorFlow
itself (its decoded tensors + file hashes: tests/golden/tf_bundle_*), and the FACT variable mapping / object graph
through a write-read round trip."""
import hashlib
import json
import os
import shutil

import numpy as np
import pytest

from mint_b200 import config_util, tf_checkpoint as T, weights as W
from tests.helpers import make_config


def _dims(**kw):
    return config_util.resolve_fact_dims(make_config(**kw))

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
