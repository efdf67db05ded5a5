"""Golden vectors produced by the REFERENCE'S OWN MODEL CODE (run from the repo root, build container only):

    PROTOCOL_BUFFERS_PYTHON_IMPLEMENTATION=python python tests/golden/make_reference_golden.py

Imports /root/reference/mint/core/fact_model.py (+ base_models.py, base_model_util.py, model_builder.py, the committed
*_pb2.py) with `tensorflow` resolving to oracle/tf_shim (NumPy semantics of the few TF primitives they call: TensorFlow
itself is not installable here) and `einops.layers.tensorflow.Rearrange` resolving to einops' own NumPy rearrange.
Runs FACTModel.call, .infer_auto_regressive and .loss in float64, walks the reference object tree to collect every
weight under this repo's variable names, and writes tests/golden/fact_reference_code_small.npz.
"""
import json
import os
import sys
import types

os.environ.setdefault("PROTOCOL_BUFFERS_PYTHON_IMPLEMENTATION", "python")
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [os.path.join(ROOT, "oracle", "tf_shim"), "/root/reference"]

import einops  # noqa: E402
import numpy as np  # noqa: E402
import tensorflow as tf  # noqa: E402  (the shim)

# einops.layers.tensorflow imports the real TF: give the reference an equivalent layer built on einops.rearrange
_m = types.ModuleType("einops.layers.tensorflow")


class Rearrange(tf.keras.layers.Layer):
    def __init__(self, pattern, **axes_lengths):
        super().__init__()
        self.pattern, self.axes_lengths = pattern, axes_lengths

    def call(self, x):
        return tf.convert_to_tensor(einops.rearrange(np.asarray(x).view(np.ndarray), self.pattern, **self.axes_lengths))


_m.Rearrange = Rearrange
sys.modules["einops.layers.tensorflow"] = _m

from google.protobuf import text_format  # noqa: E402
from mint.core import model_builder  # noqa: E402  (the reference)
from mint.protos import model_pb2  # noqa: E402

META = {"d": 64, "heads": 4, "ff": 96, "layers": [1, 2, 2], "motion_seq": 8, "audio_seq": 12, "motion_dim": 225,
        "audio_dim": 35, "out_dim": 225, "batch": 2, "audio_len": 15, "steps": 1200, "target_len": 5, "seed": 21}

CONFIG = """
fact_model {
  modality { feature_name: "audio" sequence_length: %(audio_seq)d
             model { transformer { hidden_size: %(d)d num_attention_heads: %(heads)d num_hidden_layers: %(la)d intermediate_size: %(ff)d } } }
  modality { feature_name: "motion" sequence_length: %(motion_seq)d feature_dim: %(motion_dim)d
             model { transformer { hidden_size: %(d)d num_attention_heads: %(heads)d num_hidden_layers: %(lm)d intermediate_size: %(ff)d } } }
  cross_modal_model { modality_a: "motion" modality_b: "audio"
    transformer { hidden_size: %(d)d num_attention_heads: %(heads)d num_hidden_layers: %(lc)d intermediate_size: %(ff)d }
    output_layer { out_dim: %(out_dim)d } }
}
"""


def collect_weights(model, randomize, rng):
    """Reference object tree -> this repo's flat names."""
    out = {}

    def stack(prefix, transformer):
        blocks = transformer.net.layers                      # [Residual(Norm(Attention)), Residual(Norm(MLP))] * L
        for i in range(len(blocks) // 2):
            attn_norm, mlp_norm = blocks[2 * i].fn, blocks[2 * i + 1].fn
            attn, mlp = attn_norm.fn, mlp_norm.fn
            p = f"{prefix}/layer_{i}"
            if randomize:                                    # biases / LN affine are trivial at init: perturb in place
                for t in (attn_norm.norm.gamma, mlp_norm.norm.gamma):
                    t += 0.1 * rng.standard_normal(t.shape)
                for t in (attn_norm.norm.beta, mlp_norm.norm.beta, attn.to_out.bias, mlp.net.layers[0].bias,
                          mlp.net.layers[1].bias):
                    t += 0.05 * rng.standard_normal(t.shape)
            out[f"{p}/attn/norm/gamma"] = attn_norm.norm.gamma
            out[f"{p}/attn/norm/beta"] = attn_norm.norm.beta
            out[f"{p}/attn/to_qkv/kernel"] = attn.to_qkv.kernel
            out[f"{p}/attn/to_out/kernel"] = attn.to_out.kernel
            out[f"{p}/attn/to_out/bias"] = attn.to_out.bias
            out[f"{p}/mlp/norm/gamma"] = mlp_norm.norm.gamma
            out[f"{p}/mlp/norm/beta"] = mlp_norm.norm.beta
            out[f"{p}/mlp/dense_0/kernel"] = mlp.net.layers[0].kernel
            out[f"{p}/mlp/dense_0/bias"] = mlp.net.layers[0].bias
            out[f"{p}/mlp/dense_1/kernel"] = mlp.net.layers[1].kernel
            out[f"{p}/mlp/dense_1/bias"] = mlp.net.layers[1].bias

    stack("cross_modal_layer/transformer", model.cross_modal_layer.transformer_layer)
    head = model.cross_modal_layer.cross_output_layer
    if randomize:
        head.bias += 0.05 * rng.standard_normal(head.bias.shape)
    out["cross_modal_layer/output/kernel"], out["cross_modal_layer/output/bias"] = head.kernel, head.bias
    for name in ("motion", "audio"):
        stack(f"{name}_transformer", getattr(model, f"{name}_transformer"))
        out[f"{name}_pos_embedding"] = getattr(model, f"{name}_pos_embedding").pos_embedding
        lin = getattr(model, f"{name}_linear_embedding").net
        if randomize:
            lin.bias += 0.05 * rng.standard_normal(lin.bias.shape)
        out[f"{name}_linear_embedding/kernel"], out[f"{name}_linear_embedding/bias"] = lin.kernel, lin.bias
    for v in out.values():                                   # fp32-representable values: the CUDA path stores fp32 weights
        v[...] = np.asarray(v, dtype=np.float32)
    return {k: np.array(v, dtype=np.float32) for k, v in out.items()}


def assign_weights(model, weights):
    """Inverse of collect_weights: write this repo's flat dict into the (already built) reference object tree."""
    live = collect_weights(model, False, None)               # arrays are copies: fetch the live tensors again below

    def stack(prefix, transformer):
        blocks = transformer.net.layers
        for i in range(len(blocks) // 2):
            attn_norm, mlp_norm = blocks[2 * i].fn, blocks[2 * i + 1].fn
            attn, mlp = attn_norm.fn, mlp_norm.fn
            p = f"{prefix}/layer_{i}"
            pairs = [(attn_norm.norm.gamma, "attn/norm/gamma"), (attn_norm.norm.beta, "attn/norm/beta"),
                     (attn.to_qkv.kernel, "attn/to_qkv/kernel"), (attn.to_out.kernel, "attn/to_out/kernel"),
                     (attn.to_out.bias, "attn/to_out/bias"), (mlp_norm.norm.gamma, "mlp/norm/gamma"),
                     (mlp_norm.norm.beta, "mlp/norm/beta"), (mlp.net.layers[0].kernel, "mlp/dense_0/kernel"),
                     (mlp.net.layers[0].bias, "mlp/dense_0/bias"), (mlp.net.layers[1].kernel, "mlp/dense_1/kernel"),
                     (mlp.net.layers[1].bias, "mlp/dense_1/bias")]
            for t, suf in pairs:
                t[...] = weights[f"{p}/{suf}"]

    stack("cross_modal_layer/transformer", model.cross_modal_layer.transformer_layer)
    head = model.cross_modal_layer.cross_output_layer
    head.kernel[...] = weights["cross_modal_layer/output/kernel"]
    head.bias[...] = weights["cross_modal_layer/output/bias"]
    for name in ("motion", "audio"):
        stack(f"{name}_transformer", getattr(model, f"{name}_transformer"))
        getattr(model, f"{name}_pos_embedding").pos_embedding[...] = weights[f"{name}_pos_embedding"]
        lin = getattr(model, f"{name}_linear_embedding").net
        lin.kernel[...] = weights[f"{name}_linear_embedding/kernel"]
        lin.bias[...] = weights[f"{name}_linear_embedding/bias"]
    assert set(live) == set(weights)


def main_v5():
    """Full fact_v5 dims: the reference code on THIS repo's seeded Keras-default weights (oracle.init_weights(seed 0),
    too large to commit but reproducible) and synthetic inputs (seed 0, batch 1); only inputs' seed and the outputs are
    stored."""
    sys.path.insert(0, ROOT)
    from oracle import fact_oracle as O
    cfg = model_pb2.MultiModalModel()
    with open("/root/reference/configs/fact_v5_deeper_t10_cm12.config") as f:
        from mint.protos import pipeline_pb2
        pipe = pipeline_pb2.TrainEvalPipelineConfig()
        text_format.Merge(f.read(), pipe)
    model = model_builder.build(pipe.multi_modal_model, is_training=True)
    dims = O.FACT_V5
    w = {k: v.astype(np.float32).astype(np.float64) for k, v in O.init_weights(dims, seed=0).items()}
    inp = O.synthetic_inputs(dims, batch=1, audio_len=dims.audio_seq + 1, seed=0)
    motion = inp["motion_input"].astype(np.float32).astype(np.float64)
    audio = inp["audio_input"].astype(np.float32).astype(np.float64)
    window = {"motion_input": tf.constant(motion), "audio_input": tf.constant(audio[:, :dims.audio_seq])}
    model(window)                                             # build
    assign_weights(model, w)
    call = np.asarray(model(window))
    ar = np.asarray(model.infer_auto_regressive({"motion_input": tf.constant(motion), "audio_input": tf.constant(audio)},
                                                steps=1200))
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "fact_reference_code_v5.npz")
    np.savez_compressed(path, call=call, ar=ar, meta=json.dumps({"weights_seed": 0, "inputs_seed": 0, "batch": 1,
                                                                 "audio_len": dims.audio_seq + 1}))
    print(path, call.shape, ar.shape, float(np.abs(call).mean()))


def main():
    m = META
    tf.set_seed(m["seed"])
    cfg = model_pb2.MultiModalModel()
    text_format.Merge(CONFIG % dict(m, lm=m["layers"][0], la=m["layers"][1], lc=m["layers"][2]), cfg)
    model = model_builder.build(cfg, is_training=True)           # mint/core/model_builder.py:29-33
    rng = np.random.default_rng(m["seed"] + 1)
    f32 = lambda a: a.astype(np.float32).astype(np.float64)
    motion = f32(0.5 * rng.standard_normal((m["batch"], m["motion_seq"], m["motion_dim"])))
    motion[..., :6] = 0
    audio = f32(rng.standard_normal((m["batch"], m["audio_len"], m["audio_dim"])))
    target = f32(0.5 * rng.standard_normal((m["batch"], m["target_len"], m["out_dim"])))
    window = {"motion_input": tf.constant(motion), "audio_input": tf.constant(audio[:, :m["audio_seq"]])}
    model(window)                                                # first call builds the lazily-created weights
    weights = collect_weights(model, True, rng)                  # ... then make biases / LN affine non-trivial
    call = np.asarray(model(window))                             # FACTModel.call
    ar = np.asarray(model.infer_auto_regressive(
        {"motion_input": tf.constant(motion), "audio_input": tf.constant(audio)}, steps=m["steps"]))
    loss = float(np.asarray(model.loss(tf.constant(target), tf.constant(call))))
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "fact_reference_code_small.npz")
    np.savez_compressed(path, meta=json.dumps(m), motion=motion.astype(np.float32), audio=audio.astype(np.float32),
                        target=target.astype(np.float32), call=call, ar=ar, loss=np.float64(loss),
                        **{"w:" + k: v for k, v in weights.items()})
    print(path, call.shape, ar.shape, loss, len(weights), "weights",
          sum(v.size for v in weights.values()), "params")


if __name__ == "__main__":
    main()
    main_v5()
