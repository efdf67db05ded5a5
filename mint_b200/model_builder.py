"""model_builder.build -- same dispatch as the reference (mint/core/model_builder.py:19-33)."""
from .fact_model import FACTModel

MODEL_BUILDER_MAP = {"fact_model": FACTModel}


def build(model_config, is_training, **kwargs):
    """model_config: mint.protos.MultiModalModel; dispatches on its `model` oneof."""
    model_type = model_config.WhichOneof("model")
    if model_type not in MODEL_BUILDER_MAP:
        raise ValueError("Unknown model type: {}".format(model_type))
    return MODEL_BUILDER_MAP[model_type](getattr(model_config, model_type), is_training, **kwargs)
