"""SingleTaskEvaluator -- counterpart of mint/ctl/single_task_evaluator.py:26-97.

eval_step: `model.infer_auto_regressive(inputs, steps=1200)`, prepend the seed motion, save one
`{motion_name}_{audio_name}.npy` of shape [motion_seq + n, 225] per clip (single_task_evaluator.py:64-89) -- the
layout tools/calculate_scores.py:197-215 of the reference reads.  Clips are independent, so N GPUs shard them
(mint_b200.parallel.shard_clips) with no collective.
"""
from __future__ import annotations

import os

import numpy as np
import torch


def _name(v) -> str:
    if isinstance(v, bytes):
        return v.decode("utf-8")
    if isinstance(v, (np.ndarray, torch.Tensor)):
        v = v.item() if v.ndim == 0 else v.tolist()
        return _name(v)
    if isinstance(v, (list, tuple)):
        return _name(v[0])
    return str(v)


class SingleTaskEvaluator:
    def __init__(self, eval_dataset, model, metrics=None, output_dir=None, evaluator_options=None, steps: int = 1200):
        self.eval_dataset = eval_dataset
        self.model = model
        self.metrics = metrics if isinstance(metrics, list) else ([metrics] if metrics else [])
        self.output_dir = output_dir
        self.steps = steps

    def eval_step(self, inputs: dict):
        outputs = self.model.infer_auto_regressive(inputs, steps=self.steps)            # :69
        seed = torch.as_tensor(inputs["motion_input"]).to(outputs.device, torch.float32)
        outputs = torch.cat([seed, outputs], dim=1)                                       # :71
        batch_size = outputs.shape[0]
        if self.output_dir is not None:
            os.makedirs(self.output_dir, exist_ok=True)
            host = outputs.cpu().numpy()
            for i in range(batch_size):                                                    # :73-83
                m = _name(inputs["motion_name"][i]) if "motion_name" in inputs else f"motion{i}"
                a = _name(inputs["audio_name"][i]) if "audio_name" in inputs else f"audio{i}"
                np.save(os.path.join(self.output_dir, "%s_%s.npy" % (m, a)), host[i])
        return outputs

    def evaluate(self, num_steps: int = -1):
        """orbit StandardEvaluator.evaluate: run eval_step until the dataset is exhausted (num_steps < 0)."""
        n = 0
        for batch in self.eval_dataset:
            if 0 <= num_steps <= n:
                break
            self.eval_step(batch)
            n += 1
        return {"clips_batches": n}
