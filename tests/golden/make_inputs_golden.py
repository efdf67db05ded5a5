#!/usr/bin/env python
"""Golden vectors for mint_b200/inputs.py from THE REFERENCE'S OWN mint/utils/inputs_util.py (build container only):

    PROTOCOL_BUFFERS_PYTHON_IMPLEMENTATION=python python tests/golden/make_inputs_golden.py

`get_modality_to_param_dict` runs on the reference's pipeline_pb2 parse of the fact_v5 config; `fact_preprocessing`
runs with `tensorflow` resolving to oracle/tf_shim (pad / maximum / random.uniform / checked set_shape on NumPy).  The
random window start is forced to known values so that both implementations slice the same frames; the bounds of the
draw (minval, maxval) are recorded as the reference computed them.
"""
import json
import os
import sys

os.environ.setdefault("PROTOCOL_BUFFERS_PYTHON_IMPLEMENTATION", "python")
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [os.path.join(ROOT, "oracle", "tf_shim"), "/root/reference"]

import numpy as np  # noqa: E402
import tensorflow as tf  # noqa: E402  (the shim)
from google.protobuf import text_format  # noqa: E402
from mint.protos import pipeline_pb2  # noqa: E402  (the reference)
from mint.utils import inputs_util  # noqa: E402  (the reference)

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "inputs_reference.npz")
CONFIG = "/root/reference/configs/fact_v5_deeper_t10_cm12.config"


def main():
    cfg = pipeline_pb2.TrainEvalPipelineConfig()
    text_format.Merge(open(CONFIG).read(), cfg)
    params = inputs_util.get_modality_to_param_dict(cfg.train_dataset)
    T = 401
    # value = 1000 * frame + channel (+ 0.5 for audio): exact in float32, self-describing, and it compresses to a few KB
    motion = (1000.0 * np.arange(T)[:, None] + np.arange(219)[None, :]).astype(np.float32)
    audio = (1000.0 * np.arange(T)[:, None] + np.arange(35)[None, :] + 0.5).astype(np.float32)
    out = {"motion_sequence": motion, "audio_sequence": audio,
           "params_json": np.frombuffer(json.dumps({k: {kk: (float(vv) if isinstance(vv, float) else int(vv))
                                                        for kk, vv in v.items()} for k, v in params.items()},
                                                   sort_keys=True).encode(), dtype=np.uint8)}
    bounds = []
    real_uniform = tf.random.uniform

    def spy(shape, minval=0, maxval=None, dtype=None):
        bounds.append((int(minval), int(maxval)))
        return real_uniform(shape, minval, maxval, dtype)

    tf.random.uniform = spy
    starts = [0, 137, T - 240]
    for i, start in enumerate(starts):
        tf.FORCED_START = start
        ex = inputs_util.fact_preprocessing({"motion_sequence": tf.convert_to_tensor(motion).astype(np.float32).view(tf.Tensor),
                                             "audio_sequence": tf.convert_to_tensor(audio).astype(np.float32).view(tf.Tensor),
                                             "motion_name": b"m", "audio_name": b"a"}, params, True)
        for k in ("motion_input", "target", "audio_input"):
            out[f"train{i}_{k}"] = np.asarray(ex[k])
        assert "motion_sequence" not in ex and "audio_sequence" not in ex
    tf.FORCED_START = None
    ex = inputs_util.fact_preprocessing({"motion_sequence": tf.convert_to_tensor(motion).astype(np.float32).view(tf.Tensor),
                                         "audio_sequence": tf.convert_to_tensor(audio).astype(np.float32).view(tf.Tensor)},
                                        params, False)
    out["eval_motion_input"], out["eval_audio_input"] = np.asarray(ex["motion_input"]), np.asarray(ex["audio_input"])
    assert "target" not in ex
    out["starts"] = np.array(starts)
    out["uniform_bounds"] = np.array(bounds)
    np.savez_compressed(OUT, **out)
    print("wrote", OUT, "params", params, "bounds", bounds)


if __name__ == "__main__":
    main()
