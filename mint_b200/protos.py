"""Runtime construction of the `mint.protos` config schema (fact_v5 pipeline config).

The reference keeps its configuration as proto2 text files parsed into
`mint.protos.TrainEvalPipelineConfig` (reference: mint/protos/pipeline.proto:26-33,
model.proto:20-124, train.proto:20-82, eval.proto:20-46, dataset.proto:22-90,
preprocessor.proto:20-27; loader mint/utils/config_util.py:22-50).  The committed
`*_pb2.py` of the reference are pre-3.20 generated code and do not import under
protobuf >= 4, and there is no `protoc` in this image, so the schema is rebuilt here
from a compact table with `descriptor_pb2` -- same package, message names, field names,
field numbers, types and defaults, so any text config the reference accepts parses
here and `msg.SerializeToString()` is wire-compatible with the reference's messages.

Only the schema (an interface) is restated; no reference code is used.
"""
from __future__ import annotations

from google.protobuf import descriptor_pb2 as _dpb
from google.protobuf import descriptor_pool as _pool_mod
from google.protobuf import message_factory as _mf

_F = _dpb.FieldDescriptorProto
_T = {
    "int32": _F.TYPE_INT32, "uint32": _F.TYPE_UINT32, "float": _F.TYPE_FLOAT,
    "bool": _F.TYPE_BOOL, "string": _F.TYPE_STRING,
}
PACKAGE = "mint.protos"

# message -> list of (name, number, type, label, default, oneof)
#   type: scalar name, or ".Msg" for a message / "#Enum" for an enum of this package
#   label: "opt" | "rep"
_SCHEMA = {
    # ---- preprocessor.proto
    "FACTPreprocessor": [],
    "Preprocessor": [("fact_preprocessor", 1, ".FACTPreprocessor", "opt", None, "preprocessor")],
    # ---- model.proto
    "MultiModalModel": [("fact_model", 1, ".FACTModel", "opt", None, "model")],
    "FACTModel": [
        ("modality", 1, ".Modality", "rep", None, None),
        ("cross_modal_model", 2, ".CrossModalModel", "opt", None, None),
        ("fk_path", 3, "string", "opt", None, None),
    ],
    "ModalityInputConfig": [("use_look_ahead_mask", 2, "bool", "opt", "false", None)],
    "Modality": [
        ("feature_name", 1, "string", "opt", None, None),
        ("feature_dim", 2, "int32", "opt", None, None),
        ("sequence_length", 3, "int32", "opt", None, None),
        ("input_config", 4, ".ModalityInputConfig", "opt", None, None),
        ("preprocessor", 5, ".ModalityPreprocessor", "rep", None, None),
        ("model", 6, ".ModalityModel", "rep", None, None),
    ],
    "ModalityModel": [
        ("transformer", 1, ".Transformer", "opt", None, "model"),
        ("mlp", 2, ".MLP", "opt", None, "model"),
    ],
    "MLP": [
        ("initializer_range", 1, "float", "opt", "0.02", None),
        ("hidden_act", 2, "string", "opt", "gelu", None),
        ("out_dim", 3, "int32", "opt", None, None),
    ],
    "Conv2D": [
        ("hidden_size", 1, "int32", "opt", "512", None),
        ("kernel_size", 2, "int32", "opt", "16", None),
    ],
    "ModalityPreprocessor": [("fact_preprocessor", 1, ".FACTPreprocessor", "opt", None, "preprocessor")],
    "CrossModalModel": [
        ("modality_a", 1, "string", "opt", None, None),
        ("modality_b", 2, "string", "opt", None, None),
        ("transformer", 3, ".Transformer", "opt", None, "model"),
        ("mlp", 4, ".MLP", "opt", None, "model"),
        ("cross_modal_concat_dim", 5, "#CrossModalModel.CrossModalConcatDim", "opt", "SEQUENCE_WISE", None),
        ("output_layer", 6, ".MLP", "opt", None, None),
        ("preprocess", 7, "#CrossModalModel.Preprocess", "opt", None, None),
    ],
    "Transformer": [
        ("hidden_size", 1, "int32", "opt", "768", None),
        ("num_hidden_layers", 2, "int32", "opt", "12", None),
        ("num_attention_heads", 3, "int32", "opt", "12", None),
        ("max_position_embeddings", 4, "int32", "opt", "512", None),
        ("intermediate_size", 5, "int32", "opt", "3072", None),
        ("hidden_act", 6, "string", "opt", "gelu", None),
        ("hidden_dropout_prob", 7, "float", "opt", "0.1", None),
        ("attention_probs_dropout_prob", 8, "float", "opt", "0.1", None),
        ("initializer_range", 9, "float", "opt", "0.02", None),
        ("masked_loss_type", 10, "string", "opt", "nce", None),
        ("add_spatial_attention", 11, "bool", "opt", "false", None),
        ("sp_hidden_size", 12, "int32", "opt", "768", None),
        ("sp_num_attention_heads", 13, "int32", "opt", "12", None),
        ("sp_num_hidden_layers", 14, "int32", "opt", "12", None),
        ("add_cls_token", 15, "bool", "opt", "false", None),
        ("weight_decay", 16, "float", "opt", "0", None),
    ],
    # ---- train.proto
    "TrainConfig": [
        ("num_steps", 1, "int32", "opt", "10000", None),
        ("batch_size", 2, "int32", "opt", "4", None),
        ("use_bfloat16", 3, "bool", "opt", "false", None),
        ("learning_rate", 4, ".LearningRate", "opt", None, None),
        ("grad_clip_norm", 5, "float", "opt", "1", None),
        ("fine_tune_checkpoint", 6, "string", "opt", "", None),
        ("fine_tune_checkpoint_type", 7, "#TrainConfig.CheckpointType", "opt", None, None),
    ],
    "LearningRate": [
        ("constant_learning_rate", 1, ".ConstantLearningRate", "opt", None, "learning_rate"),
        ("exponential_decay_learning_rate", 2, ".ExponentialDecayLearningRate", "opt", None, "learning_rate"),
        ("manual_step_learning_rate", 3, ".ManualStepLearningRate", "opt", None, "learning_rate"),
        ("cosine_decay_learning_rate", 4, ".CosineDecayLearningRate", "opt", None, "learning_rate"),
    ],
    "ConstantLearningRate": [("learning_rate", 1, "float", "opt", "0.002", None)],
    "ExponentialDecayLearningRate": [
        ("initial_learning_rate", 1, "float", "opt", "0.002", None),
        ("decay_steps", 2, "uint32", "opt", "4000000", None),
        ("decay_factor", 3, "float", "opt", "0.95", None),
        ("staircase", 4, "bool", "opt", "true", None),
        ("burnin_learning_rate", 5, "float", "opt", "0", None),
        ("burnin_steps", 6, "uint32", "opt", "0", None),
        ("min_learning_rate", 7, "float", "opt", "0", None),
    ],
    "ManualStepLearningRate": [
        ("initial_learning_rate", 1, "float", "opt", "0.002", None),
        ("schedule", 2, ".ManualStepLearningRate.LearningRateSchedule", "rep", None, None),
        ("warmup", 3, "bool", "opt", "false", None),
    ],
    "CosineDecayLearningRate": [
        ("learning_rate_base", 1, "float", "opt", "0.002", None),
        ("total_steps", 2, "uint32", "opt", "4000000", None),
        ("warmup_learning_rate", 3, "float", "opt", "0.0002", None),
        ("warmup_steps", 4, "uint32", "opt", "10000", None),
        ("hold_base_rate_steps", 5, "uint32", "opt", "0", None),
    ],
    # ---- eval.proto
    "EvalConfig": [
        ("batch_size", 1, "int32", "opt", "4", None),
        ("eval_metric", 2, ".EvalMetric", "opt", None, None),
    ],
    "EvalMetric": [
        ("motion_prediction_metrics", 1, ".MotionPredictionMetrics", "opt", None, "metric_oneof"),
        ("motion_generation_metrics", 2, ".MotionGenerationMetrics", "opt", None, "metric_oneof"),
    ],
    "MotionPredictionMetrics": [
        ("add_positional_metrics", 1, "bool", "opt", "false", None),
        ("pck_thresholds", 3, "float", "rep", None, None),
    ],
    "MotionGenerationMetrics": [
        ("pck_thresholds", 1, "float", "rep", None, None),
        ("num_joints", 2, "int32", "opt", "24", None),
    ],
    # ---- dataset.proto
    "Dataset": [
        ("name", 1, "string", "opt", None, None),
        ("data_files", 2, "string", "opt", None, None),
        ("window_type", 3, "#Dataset.WindowType", "opt", "DEFAULT_WINDOW", None),
        ("data_target_field", 4, "string", "opt", None, None),
        ("create_bert_masks", 5, "bool", "opt", "false", None),
        ("bert_mask_type", 6, "#Dataset.BERTMaskType", "opt", "DEFAULT_MASK", None),
        ("data_augmentation_options", 7, ".Preprocessor", "rep", None, None),
        ("sample_window", 8, "bool", "opt", "true", None),
        ("target_num_categories", 9, "int32", "opt", None, None),
        ("modality", 10, ".DataModality", "rep", None, None),
        ("input_length_sec", 11, "float", "opt", None, None),
        ("target_length_sec", 12, "float", "opt", None, None),
        ("target_shift_sec", 13, "float", "opt", None, None),
        ("length_threshold_sec", 14, "float", "opt", "0", None),
    ],
    "DataModality": [("general_modality", 2, ".GeneralModality", "opt", None, "modality")],
    "GeneralModality": [
        ("feature_name", 1, "string", "opt", None, None),
        ("dimension", 2, "int32", "opt", None, None),
        ("sample_rate", 3, "int32", "opt", None, None),
        ("resize", 4, "int32", "opt", None, None),
        ("crop_size", 5, "int32", "opt", None, None),
    ],
    # ---- pipeline.proto
    "TrainEvalPipelineConfig": [
        ("multi_modal_model", 1, ".MultiModalModel", "opt", None, None),
        ("train_config", 2, ".TrainConfig", "opt", None, None),
        ("train_dataset", 3, ".Dataset", "opt", None, None),
        ("eval_config", 4, ".EvalConfig", "opt", None, None),
        ("eval_dataset", 5, ".Dataset", "opt", None, None),
    ],
}

# nested messages: parent -> {name: fields}
_NESTED = {
    "ManualStepLearningRate": {
        "LearningRateSchedule": [
            ("step", 1, "uint32", "opt", None, None),
            ("learning_rate", 2, "float", "opt", "0.002", None),
        ]
    }
}

# enums nested in messages: parent -> {enum: [(name, value)]}
_ENUMS = {
    "CrossModalModel": {
        "CrossModalConcatDim": [("DEFAULT_CONCAT", 0), ("SEQUENCE_WISE", 1), ("CHANNEL_WISE", 2)],
        "Preprocess": [("DEFAULT_NONE", 0), ("CONTRASTIVE", 1)],
    },
    "TrainConfig": {"CheckpointType": [("DEFAULT", 0)]},
    "Dataset": {
        "BERTMaskType": [("DEFAULT_MASK", 0), ("CONTIGUOUS", 1)],
        "WindowType": [("DEFAULT_WINDOW", 0), ("BEGINNING", 1), ("CENTER", 2), ("RANDOM", 3)],
    },
}


def _fill_message(msg: _dpb.DescriptorProto, name: str, fields) -> None:
    msg.name = name
    oneofs: list[str] = []
    for fname, num, ftype, label, default, oneof in fields:
        f = msg.field.add()
        f.name, f.number = fname, num
        f.label = _F.LABEL_REPEATED if label == "rep" else _F.LABEL_OPTIONAL
        if ftype.startswith("."):
            f.type = _F.TYPE_MESSAGE
            f.type_name = f".{PACKAGE}{ftype}"
        elif ftype.startswith("#"):
            f.type = _F.TYPE_ENUM
            f.type_name = f".{PACKAGE}.{ftype[1:]}"
        else:
            f.type = _T[ftype]
        if default is not None:
            f.default_value = default
        if oneof is not None:
            if oneof not in oneofs:
                oneofs.append(oneof)
                msg.oneof_decl.add().name = oneof
            f.oneof_index = oneofs.index(oneof)
    for ename, values in _ENUMS.get(name, {}).items():
        e = msg.enum_type.add()
        e.name = ename
        for vname, vnum in values:
            v = e.value.add()
            v.name, v.number = vname, vnum
    for nname, nfields in _NESTED.get(name, {}).items():
        _fill_message(msg.nested_type.add(), nname, nfields)


def _build():
    fd = _dpb.FileDescriptorProto()
    fd.name = "mint_b200/fact_v5_schema.proto"
    fd.package = PACKAGE
    fd.syntax = "proto2"
    for name, fields in _SCHEMA.items():
        _fill_message(fd.message_type.add(), name, fields)
    # the pipeline message reserves an extension range in the reference (pipeline.proto:32)
    for m in fd.message_type:
        if m.name == "TrainEvalPipelineConfig":
            r = m.extension_range.add()
            r.start, r.end = 1000, 536870912
    pool = _pool_mod.DescriptorPool()
    pool.Add(fd)
    classes = {}
    for name in _SCHEMA:
        classes[name] = _mf.GetMessageClass(pool.FindMessageTypeByName(f"{PACKAGE}.{name}"))
    return pool, classes


_POOL, _CLASSES = _build()

MultiModalModel = _CLASSES["MultiModalModel"]
FACTModel = _CLASSES["FACTModel"]
Modality = _CLASSES["Modality"]
ModalityModel = _CLASSES["ModalityModel"]
CrossModalModel = _CLASSES["CrossModalModel"]
Transformer = _CLASSES["Transformer"]
MLP = _CLASSES["MLP"]
TrainConfig = _CLASSES["TrainConfig"]
LearningRate = _CLASSES["LearningRate"]
EvalConfig = _CLASSES["EvalConfig"]
Dataset = _CLASSES["Dataset"]
TrainEvalPipelineConfig = _CLASSES["TrainEvalPipelineConfig"]


def message_class(name: str):
    """Message class of `mint.protos.<name>`."""
    return _CLASSES[name]
