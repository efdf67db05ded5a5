"""Config boundary: the fact_v5 proto config parses to the same message the reference's own protobuf classes produce."""
import os

import pytest
from google.protobuf import text_format

from mint_b200 import config_util, protos

GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "fact_v5_pipeline_config.binpb")


def test_wire_identical_to_reference_pb2():
    """Golden bytes were serialised by the reference's mint/protos/pipeline_pb2.py from its own config file
    (tests/golden/make_config_golden.py); our runtime-built schema must produce and parse the same bytes."""
    cfg = config_util.get_configs_from_pipeline_file(config_util.DEFAULT_CONFIG)
    ours = config_util.create_pipeline_proto_from_configs(cfg)
    golden = open(GOLDEN, "rb").read()
    assert ours.SerializeToString(deterministic=True) == golden
    back = protos.TrainEvalPipelineConfig()
    back.ParseFromString(golden)
    assert back == ours


def test_fact_v5_dims():
    cfg = config_util.get_configs_from_pipeline_file(config_util.DEFAULT_CONFIG)
    d = config_util.resolve_fact_dims(cfg["model"].fact_model)
    assert (d.motion.seq_len, d.motion.feature_dim, d.motion.layers) == (120, 225, 2)
    assert (d.audio.seq_len, d.audio.feature_dim, d.audio.layers) == (240, 35, 2)
    assert (d.cross_hidden, d.cross_heads, d.cross_ff, d.cross_layers, d.out_dim) == (800, 10, 3072, 12, 225)
    assert d.cross_seq == 360
    assert cfg["train_config"].batch_size == 32 and cfg["eval_config"].batch_size == 1
    lr = cfg["train_config"].learning_rate.manual_step_learning_rate
    assert abs(lr.initial_learning_rate - 1e-4) < 1e-9 and [s.step for s in lr.schedule] == [100000, 150000]


def test_proto_defaults_match_reference_schema():
    t = protos.Transformer()
    assert (t.hidden_size, t.num_hidden_layers, t.num_attention_heads, t.intermediate_size) == (768, 12, 12, 3072)
    assert abs(protos.MLP().initializer_range - 0.02) < 1e-9
    assert protos.CrossModalModel().cross_modal_concat_dim == protos.CrossModalModel.SEQUENCE_WISE
    assert protos.TrainConfig().batch_size == 4 and protos.EvalConfig().batch_size == 4


def test_override_string_and_unknown_field():
    cfg = config_util.get_configs_from_pipeline_file(config_util.DEFAULT_CONFIG, "train_config { batch_size: 128 }")
    assert cfg["train_config"].batch_size == 128
    with pytest.raises(text_format.ParseError):
        text_format.Merge("no_such_field: 1", protos.TrainEvalPipelineConfig())


def test_channel_wise_concat_rejected():
    cfg = config_util.get_configs_from_pipeline_file(config_util.DEFAULT_CONFIG)
    m = cfg["model"].fact_model
    m.cross_modal_model.cross_modal_concat_dim = protos.CrossModalModel.CHANNEL_WISE
    with pytest.raises(NotImplementedError):  # mint/core/base_models.py:194-196
        config_util.resolve_fact_dims(m)
