"""Load a fact_v5 pipeline config (text proto) and resolve it to the dims the kernels need.

Mirrors the observable behaviour of the reference loader
(mint/utils/config_util.py:22-50: text_format.Merge of a TrainEvalPipelineConfig, optional
override string merged on top, result returned as a dict of sub-configs) and of
`build_modalities_model` (mint/core/multi_modal_model_util.py:24-56).
"""
from __future__ import annotations

import dataclasses
import os

from google.protobuf import text_format

from . import protos

DEFAULT_CONFIG = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                              "configs", "fact_v5_deeper_t10_cm12.config")


def create_pipeline_proto_from_configs(configs: dict):
    """Inverse of `get_configs_from_pipeline_file` (reference config_util.py:53-70)."""
    p = protos.TrainEvalPipelineConfig()
    p.multi_modal_model.CopyFrom(configs["model"])
    p.train_config.CopyFrom(configs["train_config"])
    p.train_dataset.CopyFrom(configs["train_dataset"])
    p.eval_config.CopyFrom(configs["eval_config"])
    p.eval_dataset.CopyFrom(configs["eval_dataset"])
    return p


def get_configs_from_pipeline_file(pipeline_config_path: str, config_override: str | None = None) -> dict:
    """Reads a TrainEvalPipelineConfig text proto -> dict of its five sub-configs."""
    pipeline = protos.TrainEvalPipelineConfig()
    with open(pipeline_config_path, "r") as f:
        text_format.Merge(f.read(), pipeline)
    if config_override:
        text_format.Merge(config_override, pipeline)
    return {
        "model": pipeline.multi_modal_model,
        "train_config": pipeline.train_config,
        "train_dataset": pipeline.train_dataset,
        "eval_config": pipeline.eval_config,
        "eval_dataset": pipeline.eval_dataset,
    }


@dataclasses.dataclass(frozen=True)
class EncoderDims:
    seq_len: int
    feature_dim: int   # 0 = unset in the config (inferred from data, as Keras Dense does)
    hidden: int
    layers: int
    heads: int
    ff: int


@dataclasses.dataclass(frozen=True)
class FactDims:
    """Frozen shape summary of a `mint.protos.FACTModel` config."""
    motion: EncoderDims
    audio: EncoderDims
    cross_hidden: int
    cross_layers: int
    cross_heads: int
    cross_ff: int
    out_dim: int
    out_init_range: float

    @property
    def cross_seq(self) -> int:
        return self.motion.seq_len + self.audio.seq_len


def resolve_fact_dims(fact_config, audio_feature_dim: int = 35) -> FactDims:
    """`fact_config`: a mint.protos.FACTModel message (model.proto:27-31)."""
    enc = {}
    for m in fact_config.modality:
        # build_modalities_model keeps the LAST transformer listed per modality
        tr = None
        for mm in m.model:
            if mm.WhichOneof("model") == "transformer":
                tr = mm.transformer
        if tr is None:
            raise ValueError(f"modality {m.feature_name!r} has no transformer model")
        fdim = m.feature_dim if m.HasField("feature_dim") else 0
        enc[m.feature_name] = EncoderDims(m.sequence_length, fdim, tr.hidden_size, tr.num_hidden_layers,
                                          tr.num_attention_heads, tr.intermediate_size)
    if "motion" not in enc or "audio" not in enc:
        raise ValueError("FACT needs a 'motion' and an 'audio' modality")
    if enc["audio"].feature_dim == 0:
        enc["audio"] = dataclasses.replace(enc["audio"], feature_dim=audio_feature_dim)
    cm = fact_config.cross_modal_model
    if cm.WhichOneof("model") != "transformer":
        raise NotImplementedError("cross_modal_model must be a transformer")
    if cm.cross_modal_concat_dim != protos.CrossModalModel.SEQUENCE_WISE:
        raise NotImplementedError(
            "cross_modal_concat_dim %s is not supported." % cm.cross_modal_concat_dim)
    t = cm.transformer
    return FactDims(enc["motion"], enc["audio"], t.hidden_size, t.num_hidden_layers,
                    t.num_attention_heads, t.intermediate_size, cm.output_layer.out_dim,
                    cm.output_layer.initializer_range)
