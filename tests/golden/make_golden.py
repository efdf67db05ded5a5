"""Regenerate tests/golden/fact_small_fp64.npz from the fp64 oracle (run from the repo root).

The reference itself cannot produce vectors here (TensorFlow is not installable in this image), so these pin the
oracle, not the reference: PARITY UNPINNED (see oracle/fact_oracle.py)."""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from oracle import fact_oracle as O  # noqa: E402

META = {"dims": dict(d=32, heads=4, ff=48, layers=(1, 1, 2), motion_seq=6, audio_seq=10, motion_dim=225, out_dim=225),
        "seed": 11, "batch": 2, "audio_len": 13, "steps": 1200, "target_len": 5}


def main():
    dm = META["dims"]
    dims = O.Dims(dm["d"], dm["heads"], dm["ff"], *dm["layers"], dm["motion_seq"], dm["audio_seq"], dm["motion_dim"],
                  35, dm["out_dim"])
    w = O.init_weights(dims, META["seed"], randomize_affine=True)
    inp = O.synthetic_inputs(dims, META["batch"], audio_len=META["audio_len"], seed=META["seed"],
                             target_len=META["target_len"])
    call = O.call(w, dims, {"motion_input": inp["motion_input"], "audio_input": inp["audio_input"][:, :dims.audio_seq]})
    ar = O.infer_auto_regressive(w, dims, inp, steps=META["steps"])
    out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "fact_small_fp64.npz")
    np.savez_compressed(out, call=call, ar=ar, loss=np.float64(O.loss(inp["target"], call)), meta=json.dumps(META))
    print(out, call.shape, ar.shape)


if __name__ == "__main__":
    main()
