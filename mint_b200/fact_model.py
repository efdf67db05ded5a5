"""FACTModel -- B200-native drop-in for mint/core/fact_model.py:26-148.

Same constructor and methods as the reference class (`FACTModel(config, is_training)`, `__call__(inputs)`,
`infer_auto_regressive(inputs, steps=1200)`, `loss(target, pred)`, `get_metrics(eval_config)`,
`trainable_variables`, `losses`, settable `global_step`), but tensors are torch CUDA tensors and all math runs in
libfact_sm100.so (hand-written sm_100a kernels) through the C ABI of include/fact_sm100.h.  No CPU fallback.
"""
from __future__ import annotations

import copy
import ctypes as C

import torch

from . import config_util, lib, weights as W

_TORCH_BF16 = torch.bfloat16


class FACTModel:
    """Audio Motion Multi-Modal model (reference docstring: fact_model.py:26-27)."""

    def __init__(self, config, is_training: bool, device: str | torch.device = "cuda", mode: str = "precise",
                 seed: int = 0, use_graph: bool = True):
        """config: mint.protos.FACTModel message (model.proto:27-31); is_training kept for signature parity
        (the reference only uses it to gate dropout, which FACT never applies)."""
        self._lib = lib.load()
        if not torch.cuda.is_available():
            raise RuntimeError("FACTModel needs a CUDA device (sm_100a); there is no CPU path")
        self.config = copy.deepcopy(config)
        self.is_training = is_training
        self.dims = config_util.resolve_fact_dims(self.config)
        d = self.dims
        for enc in (d.motion, d.audio):
            if (enc.hidden, enc.heads, enc.ff) != (d.cross_hidden, d.cross_heads, d.cross_ff):
                # the reference raises on a width mismatch at call time (base_models.py:184-189); heads/ff could
                # differ there, but the C ABI carries one (d, heads, ff) triple
                raise ValueError("The modal_a hidden size (%d) should be the same with the modal_b hidden size (%d)"
                                 % (d.motion.hidden, d.audio.hidden))
        self.device = torch.device(device)
        if mode not in lib.MODES:
            raise ValueError(f"mode must be one of {sorted(lib.MODES)}")
        self.mode = mode
        self.use_graph = use_graph
        self.global_step = None        # set by trainer / evaluator (trainer.py:151, evaluator.py:57)
        self.losses = []               # Keras regularisation losses: none in FACT
        self._params: dict[str, torch.Tensor] = {}     # views into self._flat (one bucket: one NCCL all-reduce)
        self._grads: dict[str, torch.Tensor] | None = None
        self._flat: torch.Tensor | None = None
        self._grad_flat: torch.Tensor | None = None
        self._packed: dict[str, tuple[torch.Tensor, torch.Tensor]] = {}
        self._keras_bf16: dict[str, torch.Tensor] = {}  # Keras-layout bf16 copies (backward dX GEMMs), training only
        self._kl_flat: torch.Tensor | None = None       # ... which are views into this bf16 mirror of the bucket
        self._kl_end = 0
        self._train_ws: torch.Tensor | None = None
        self._ws: dict[tuple, torch.Tensor] = {}
        self._side_stream = None
        self._step_counter = torch.zeros(1, dtype=torch.int32, device=self.device)
        # owner of the captured AR frame graphs: dies with the model, so no graph outlives the buffers it points into
        self._session = self._lib.fact_ar_session_create()
        if not self._session:
            raise MemoryError("fact_ar_session_create failed")
        self._ar_staging: dict[tuple, tuple[torch.Tensor, torch.Tensor]] = {}
        self._cdims = lib.Dims(d.cross_hidden, d.cross_heads, d.cross_ff, d.motion.layers, d.audio.layers,
                               d.cross_layers, d.motion.seq_len, d.audio.seq_len, d.motion.feature_dim,
                               d.audio.feature_dim, d.out_dim)
        self.set_weights(W.keras_default_init(d, seed))

    def __del__(self):
        sess, self._session = getattr(self, "_session", None), None
        if sess:
            try:
                self._lib.fact_ar_session_destroy(sess)
            except Exception:        # interpreter shutdown
                pass

    # ------------------------------------------------------------------ variables
    @property
    def trainable_variables(self):
        return list(self._params.values())

    def variable_names(self):
        return list(self._params.keys())

    def get_weights(self) -> dict:
        return {k: v.detach().clone() for k, v in self._params.items()}

    def set_weights(self, weights: dict) -> None:
        """weights: name -> array-like in Keras layout (see mint_b200/weights.py for the names)."""
        shapes = W.variable_shapes(self.dims)
        missing = set(shapes) - set(weights)
        extra = set(weights) - set(shapes)
        if missing or extra:
            raise KeyError(f"weight names mismatch: missing {sorted(missing)[:3]} extra {sorted(extra)[:3]}")
        if self._flat is None:
            self._offsets, off = {}, 0
            for name, shp in shapes.items():
                n = 1
                for v in shp:
                    n *= v
                self._offsets[name] = (off, n)
                off += (n + 7) // 8 * 8                      # every tensor 16-byte aligned, also in the bf16 copy
            self._flat = torch.zeros(off, dtype=torch.float32, device=self.device)
            for name, shp in shapes.items():
                o, n = self._offsets[name]
                self._params[name] = self._flat[o:o + n].view(shp)
        for name, shp in shapes.items():
            t = torch.as_tensor(weights[name]).to(torch.float32)
            if tuple(t.shape) != tuple(shp):
                raise ValueError(f"{name}: shape {tuple(t.shape)} != {shp}")
            self._params[name].copy_(t.to(self.device))
        self.repack()

    def repack(self, bf16_copy_done: bool = False) -> None:
        """(Re)build the bf16 operand copies the tensor-core GEMMs read (after any weight update): the K-major hi/lo
        (transposed) copies in ONE launch, and, in training, the Keras-layout bf16 copy of the whole bucket (skipped
        when the optimizer kernel has just written it)."""
        st = torch.cuda.current_stream(self.device).cuda_stream
        names = [n for n in self._params if n.endswith("/kernel") and "linear_embedding" not in n]
        precise = self.mode != "bf16"
        for name in names:
            if name not in self._packed:
                k_in, n_out = self._params[name].shape
                hi = torch.empty((n_out, k_in), dtype=_TORCH_BF16, device=self.device)
                self._packed[name] = (hi, torch.empty_like(hi) if precise else None)
        cnt = len(names)
        vp = C.c_void_p * cnt
        w_arr = vp(*[self._params[n].data_ptr() for n in names])
        hi_arr = vp(*[self._packed[n][0].data_ptr() for n in names])
        lo_arr = vp(*[self._packed[n][1].data_ptr() if precise else None for n in names])
        k_arr = (C.c_int * cnt)(*[self._params[n].shape[0] for n in names])
        n_arr = (C.c_int * cnt)(*[self._params[n].shape[1] for n in names])
        with torch.cuda.device(self.device):
            lib.check(self._lib.fact_pack_weights(w_arr, hi_arr, lo_arr, k_arr, n_arr, cnt, st), "fact_pack_weights")
            if self._kl_flat is not None:
                if not bf16_copy_done:
                    self._kl_flat.copy_(self._flat)              # round-to-nearest, as the optimizer kernel does
                for name, kl in self._keras_bf16.items():        # padded copies (the head: 225 -> 256 columns)
                    if kl.data_ptr() < self._kl_flat.data_ptr() or kl.data_ptr() >= self._kl_end:
                        p = self._params[name]
                        lib.check(self._lib.fact_cast_weight(p.data_ptr(), kl.data_ptr(), p.shape[0], p.shape[1],
                                                             kl.shape[1], st), "fact_cast_weight")
        self._build_tables()

    @property
    def flat_bf16_parameters(self):
        """bf16 mirror of flat_parameters (training only; None before the first training call)."""
        return self._kl_flat

    def _build_tables(self) -> None:
        P, K = self._params, self._packed
        precise = self.mode != "bf16"

        def layer(prefix: str) -> lib.LayerWeights:
            g = lambda s: P[f"{prefix}/{s}"].data_ptr()
            hi = lambda s: K[f"{prefix}/{s}"][0].data_ptr()
            lo = lambda s: K[f"{prefix}/{s}"][1].data_ptr() if precise else None
            kl = lambda s: (self._keras_bf16[f"{prefix}/{s}"].data_ptr()
                            if f"{prefix}/{s}" in self._keras_bf16 else None)
            return lib.LayerWeights(
                g("attn/norm/gamma"), g("attn/norm/beta"), hi("attn/to_qkv/kernel"), lo("attn/to_qkv/kernel"),
                hi("attn/to_out/kernel"), lo("attn/to_out/kernel"), g("attn/to_out/bias"),
                g("mlp/norm/gamma"), g("mlp/norm/beta"), hi("mlp/dense_0/kernel"), lo("mlp/dense_0/kernel"),
                g("mlp/dense_0/bias"), hi("mlp/dense_1/kernel"), lo("mlp/dense_1/kernel"), g("mlp/dense_1/bias"),
                g("attn/to_qkv/kernel"), g("attn/to_out/kernel"), g("mlp/dense_0/kernel"), g("mlp/dense_1/kernel"),
                kl("attn/to_qkv/kernel"), kl("attn/to_out/kernel"), kl("mlp/dense_0/kernel"), kl("mlp/dense_1/kernel"))

        d = self.dims
        self._lw_motion = (lib.LayerWeights * d.motion.layers)(
            *[layer(f"motion_transformer/layer_{i}") for i in range(d.motion.layers)])
        self._lw_audio = (lib.LayerWeights * d.audio.layers)(
            *[layer(f"audio_transformer/layer_{i}") for i in range(d.audio.layers)])
        self._lw_cross = (lib.LayerWeights * d.cross_layers)(
            *[layer(f"cross_modal_layer/transformer/layer_{i}") for i in range(d.cross_layers)])
        ok, ob = "cross_modal_layer/output/kernel", "cross_modal_layer/output/bias"
        self._cw = lib.Weights(
            self._lw_motion, self._lw_audio, self._lw_cross,
            P["motion_linear_embedding/kernel"].data_ptr(), P["motion_linear_embedding/bias"].data_ptr(),
            P["motion_pos_embedding"].data_ptr(),
            P["audio_linear_embedding/kernel"].data_ptr(), P["audio_linear_embedding/bias"].data_ptr(),
            P["audio_pos_embedding"].data_ptr(),
            P[ok].data_ptr(), P[ob].data_ptr(), K[ok][0].data_ptr(), K[ok][1].data_ptr() if precise else None,
            self._keras_bf16[ok].data_ptr() if ok in self._keras_bf16 else None)

    # ------------------------------------------------------------------ plumbing
    def _workspace(self, batch: int):
        mode = lib.MODES[self.mode]
        key = (batch, mode)
        if key not in self._ws:
            need = self._lib.fact_workspace_bytes(C.byref(self._cdims), batch, mode)
            self._ws.clear()  # one live workspace: shapes change rarely and the buffers are large
            self._ws[key] = torch.empty(need + 1024, dtype=torch.uint8, device=self.device)
        buf = self._ws[key]
        base = (buf.data_ptr() + 1023) & ~1023
        return base, buf.numel() - (base - buf.data_ptr())

    @staticmethod
    def _check(t, last_dim: int, what: str) -> torch.Tensor:
        t = torch.as_tensor(t)
        if t.dim() != 3 or t.shape[-1] != last_dim:
            raise ValueError(f"{what} must be [batch, seq, {last_dim}], got {tuple(t.shape)}")
        return t

    def _to_dev(self, t, last_dim: int, what: str) -> torch.Tensor:
        t = self._check(t, last_dim, what)
        return t.to(device=self.device, dtype=torch.float32, non_blocking=True).contiguous()

    # ------------------------------------------------------------------ the reference interface
    def __call__(self, inputs: dict, training: bool = False) -> torch.Tensor:
        """Single forward pass (fact_model.py:72-101).  inputs["motion_input"]: [B, 120, 225],
        inputs["audio_input"]: [B, 240, 35]; other keys are ignored.  Returns [B, 360, 225]."""
        d = self.dims
        motion = self._to_dev(inputs["motion_input"], d.motion.feature_dim, "motion_input")
        audio = self._to_dev(inputs["audio_input"], d.audio.feature_dim, "audio_input")
        if motion.shape[1] != d.motion.seq_len or audio.shape[1] != d.audio.seq_len:
            # PositionEmbedding adds a [seq, dim] table (base_models.py:154-156): other lengths fail to broadcast
            raise ValueError(f"sequence lengths must be {d.motion.seq_len}/{d.audio.seq_len}, got "
                             f"{motion.shape[1]}/{audio.shape[1]}")
        if motion.shape[0] != audio.shape[0]:
            raise ValueError("motion_input and audio_input batch sizes differ")
        batch = motion.shape[0]
        out = torch.empty((batch, d.cross_seq, d.out_dim), dtype=torch.float32, device=self.device)
        ws, ws_bytes = self._workspace(batch)
        with torch.cuda.device(self.device):
            st = torch.cuda.current_stream(self.device).cuda_stream
            lib.check(self._lib.fact_forward(C.byref(self._cdims), C.byref(self._cw), motion.data_ptr(),
                                             audio.data_ptr(), out.data_ptr(), batch, ws, ws_bytes,
                                             lib.MODES[self.mode], st), "fact_forward")
        return out

    call = __call__

    def infer_auto_regressive(self, inputs: dict, steps: int = 1200) -> torch.Tensor:
        """Auto-regressive generation (fact_model.py:103-132): returns [B, n, 225], n = min(steps, T_audio - 239).
        Inputs are staged in buffers the model keeps per (batch, n, T_audio), so repeated calls of one shape replay the
        same captured frame graph by construction (its key holds those buffers' addresses)."""
        d = self.dims
        motion = self._check(inputs["motion_input"], d.motion.feature_dim, "motion_input")
        audio = self._check(inputs["audio_input"], d.audio.feature_dim, "audio_input")
        if motion.shape[1] != d.motion.seq_len:
            raise ValueError(f"motion seed must have {d.motion.seq_len} frames")
        if motion.shape[0] != audio.shape[0]:
            raise ValueError("motion_input and audio_input batch sizes differ")
        batch, audio_len = audio.shape[0], audio.shape[1]
        n = min(int(steps), audio_len - d.audio.seq_len + 1)   # the loop breaks on the first short window (:125-126)
        if n <= 0:
            raise ValueError("no full audio window: tf.concat of an empty list fails in the reference too")
        key = (batch, n, audio_len)
        if key not in self._ar_staging:
            while len(self._ar_staging) >= 4:                   # a few shapes stay warm; oldest out
                self._ar_staging.pop(next(iter(self._ar_staging)))
            self._ar_staging[key] = (
                torch.empty((batch, d.motion.seq_len + n, d.motion.feature_dim), dtype=torch.float32, device=self.device),
                torch.empty((batch, audio_len, d.audio.feature_dim), dtype=torch.float32, device=self.device))
        hist, audio_buf = self._ar_staging[key]
        hist[:, :d.motion.seq_len].copy_(motion, non_blocking=True)     # host -> device directly into the staging
        audio_buf.copy_(audio, non_blocking=True)
        self.generate_into(hist, audio_buf, 0, n)
        return hist[:, d.motion.seq_len:].clone()

    def new_history(self, motion_seed: torch.Tensor, capacity: int) -> torch.Tensor:
        """[B, motion_seq + capacity, motion_dim] device buffer whose head is the seed; frames are appended in place."""
        d = self.dims
        motion_seed = self._to_dev(motion_seed, d.motion.feature_dim, "motion_input")
        hist = torch.empty((motion_seed.shape[0], d.motion.seq_len + capacity, d.motion.feature_dim),
                           dtype=torch.float32, device=self.device)
        hist[:, :d.motion.seq_len].copy_(motion_seed)
        return hist

    def generate_into(self, hist: torch.Tensor, audio: torch.Tensor, start: int, n: int) -> None:
        """Generate frames [start, start + n) into `hist` (device tensors; continues a previous call when start > 0)."""
        d = self.dims
        batch, audio_len = audio.shape[0], audio.shape[1]
        capacity = hist.shape[1] - d.motion.seq_len
        ws, ws_bytes = self._workspace(batch)
        with torch.cuda.device(self.device):
            cur = torch.cuda.current_stream(self.device)
            run = cur
            if self.use_graph and cur.cuda_stream == 0:       # stream capture is illegal on the legacy stream
                if self._side_stream is None:
                    self._side_stream = torch.cuda.Stream(self.device)
                run = self._side_stream
                run.wait_stream(cur)
            lib.check(self._lib.fact_infer_auto_regressive(
                C.byref(self._cdims), C.byref(self._cw), hist.data_ptr(), capacity, audio.data_ptr(), audio_len,
                batch, start, n, self._step_counter.data_ptr(), ws, ws_bytes, lib.MODES[self.mode],
                1 if self.use_graph else 0, self._session, run.cuda_stream), "fact_infer_auto_regressive")
            if run is not cur:
                cur.wait_stream(run)
                for t in (hist, audio):
                    t.record_stream(run)

    def loss(self, target, pred) -> torch.Tensor:
        """L2 motion generation loss on the first target_seq_len frames (fact_model.py:134-148)."""
        pred = torch.as_tensor(pred).to(self.device, torch.float32).contiguous()
        target = torch.as_tensor(target).to(self.device, torch.float32).contiguous()
        b, n, od = pred.shape
        t_len = target.shape[1]
        out = torch.empty((), dtype=torch.float32, device=self.device)
        partial = torch.empty(1024, dtype=torch.float32, device=self.device)
        with torch.cuda.device(self.device):
            st = torch.cuda.current_stream(self.device).cuda_stream
            lib.check(self._lib.fact_mse(target.data_ptr(), pred.data_ptr(), out.data_ptr(), None,
                                         partial.data_ptr(), b, t_len, n, od, 1.0, st), "fact_mse")
        return out

    def compute_motion_generation_loss(self, pred_tensors, target_tensors):
        return self.loss(target_tensors, pred_tensors)

    # ------------------------------------------------------------------ training (single_task_trainer.py:138-187)
    @property
    def flat_parameters(self) -> torch.Tensor:
        """All trainable variables as ONE fp32 bucket (the named variables are views into it)."""
        return self._flat

    @property
    def flat_gradients(self) -> torch.Tensor:
        self._ensure_training_state()
        return self._grad_flat

    def gradients(self) -> dict:
        self._ensure_training_state()
        return self._grads

    def _ensure_training_state(self) -> None:
        if self._grad_flat is not None:
            return
        self._grad_flat = torch.zeros_like(self._flat)
        self._grads = {}
        for name, p in self._params.items():
            o, n = self._offsets[name]
            self._grads[name] = self._grad_flat[o:o + n].view(p.shape)
        pad64 = lambda v: (v + 63) // 64 * 64
        self._kl_flat = torch.zeros(self._flat.numel(), dtype=_TORCH_BF16, device=self.device)
        self._kl_end = self._kl_flat.data_ptr() + 2 * self._kl_flat.numel()
        for name, p in self._params.items():
            if name.endswith("/kernel") and "linear_embedding" not in name:
                cols = p.shape[1]
                if name == "cross_modal_layer/output/kernel":    # 16-byte row pitch for TMA: own padded buffer
                    self._keras_bf16[name] = torch.zeros((p.shape[0], pad64(cols)), dtype=_TORCH_BF16,
                                                         device=self.device)
                else:
                    o, n = self._offsets[name]
                    self._keras_bf16[name] = self._kl_flat[o:o + n].view(p.shape)
        self.repack()
        self._build_grad_tables()

    def adopt_symmetric(self, arena) -> None:
        """Move the flat parameter bucket, the gradient bucket and the bf16 mirror into `arena`
        (mint_b200.parallel.SymmetricArena) so that peers can read / write them over NVLink; the named variables,
        gradients and operand copies become views of the arena.  Values are preserved."""
        self._ensure_training_state()
        n = self._flat.numel()
        if arena.numel != n:
            raise ValueError("arena size does not match the parameter bucket")
        arena.w.copy_(self._flat)
        arena.grad.copy_(self._grad_flat)
        arena.wb.copy_(self._kl_flat)
        self._flat, self._grad_flat, self._kl_flat = arena.w, arena.grad, arena.wb
        self._kl_end = self._kl_flat.data_ptr() + 2 * n
        for name, p in list(self._params.items()):
            o, cnt = self._offsets[name]
            shp = tuple(p.shape)
            self._params[name] = self._flat[o:o + cnt].view(shp)
            self._grads[name] = self._grad_flat[o:o + cnt].view(shp)
            if name in self._keras_bf16 and name != "cross_modal_layer/output/kernel":
                self._keras_bf16[name] = self._kl_flat[o:o + cnt].view(shp)
        self._arena = arena               # keeps the symmetric allocation alive as long as the model
        self.repack()
        self._build_grad_tables()

    def _build_grad_tables(self) -> None:
        G, d = self._grads, self.dims

        def layer(prefix: str) -> lib.LayerGrads:
            g = lambda s: G[f"{prefix}/{s}"].data_ptr()
            return lib.LayerGrads(g("attn/norm/gamma"), g("attn/norm/beta"), g("attn/to_qkv/kernel"),
                                  g("attn/to_out/kernel"), g("attn/to_out/bias"), g("mlp/norm/gamma"),
                                  g("mlp/norm/beta"), g("mlp/dense_0/kernel"), g("mlp/dense_0/bias"),
                                  g("mlp/dense_1/kernel"), g("mlp/dense_1/bias"))

        self._lg_motion = (lib.LayerGrads * d.motion.layers)(
            *[layer(f"motion_transformer/layer_{i}") for i in range(d.motion.layers)])
        self._lg_audio = (lib.LayerGrads * d.audio.layers)(
            *[layer(f"audio_transformer/layer_{i}") for i in range(d.audio.layers)])
        self._lg_cross = (lib.LayerGrads * d.cross_layers)(
            *[layer(f"cross_modal_layer/transformer/layer_{i}") for i in range(d.cross_layers)])
        self._cg = lib.Grads(
            self._lg_motion, self._lg_audio, self._lg_cross,
            G["motion_linear_embedding/kernel"].data_ptr(), G["motion_linear_embedding/bias"].data_ptr(),
            G["motion_pos_embedding"].data_ptr(),
            G["audio_linear_embedding/kernel"].data_ptr(), G["audio_linear_embedding/bias"].data_ptr(),
            G["audio_pos_embedding"].data_ptr(),
            G["cross_modal_layer/output/kernel"].data_ptr(), G["cross_modal_layer/output/bias"].data_ptr())

    def gradient_stages(self) -> list[tuple[int, int, int]]:
        """(offset, count, event) slices of `flat_gradients` in the order fact_train_step finishes them; `event` is the
        index of the stage event (include/fact_sm100.h) after which the slice is final: the head, cross layers top to
        bottom, motion layers top to bottom, audio layers likewise.  A modality's last event covers two ranges (its
        layer 0, and its position table + LinearEmbedding, which sit behind the other layers in the bucket).  Slices
        are contiguous, 16-byte aligned and together cover the bucket."""
        d = self.dims
        first = lambda prefix: self._offsets[next(n for n in self._offsets if n.startswith(prefix))][0]
        total = self._flat.numel()
        cuts = {("head", 0): first("cross_modal_layer/output/"), ("motion_tail", 0): first("motion_pos_embedding"),
                ("audio_tail", 0): first("audio_pos_embedding")}
        for i in range(d.cross_layers):
            cuts[("cross", i)] = first(f"cross_modal_layer/transformer/layer_{i}/")
        for i in range(d.motion.layers):
            cuts[("motion", i)] = first(f"motion_transformer/layer_{i}/")
        for i in range(d.audio.layers):
            cuts[("audio", i)] = first(f"audio_transformer/layer_{i}/")
        order = sorted(cuts.values()) + [total]
        end = {off: order[i + 1] for i, off in enumerate(order[:-1])}
        span = lambda key: (cuts[key], end[cuts[key]] - cuts[key])
        stages = [(*span(("head", 0)), 0)]
        ev = 1
        for i in reversed(range(d.cross_layers)):
            stages.append((*span(("cross", i)), ev))
            ev += 1
        for name, layers in (("motion", d.motion.layers), ("audio", d.audio.layers)):
            for i in reversed(range(layers)):
                stages.append((*span((name, i)), ev))
                if i == 0:
                    stages.append((*span((name + "_tail", 0)), ev))
                ev += 1
        assert sum(c for _, c, _ in stages) == total and ev == self.num_gradient_stage_events
        return stages

    @property
    def num_gradient_stage_events(self) -> int:
        d = self.dims
        return 1 + d.cross_layers + d.motion.layers + d.audio.layers

    def forward_backward(self, inputs: dict, target, loss_scale: float = 1.0, stage_events=None) -> torch.Tensor:
        """One replica's forward + backward (single_task_trainer.py:145-178): returns FACTModel.loss(target, pred)
        and leaves d(loss * loss_scale)/d(variable) in `flat_gradients` (zeroed first).  bf16 products, fp32 stats.
        stage_events: optional list of torch.cuda.Event (entries may be None), recorded as the stages of the backward
        finish (order: include/fact_sm100.h, fact_train_step; see gradient_stages / gradient_stage_events)."""
        self._ensure_training_state()
        d = self.dims
        motion = self._to_dev(inputs["motion_input"], d.motion.feature_dim, "motion_input")
        audio = self._to_dev(inputs["audio_input"], d.audio.feature_dim, "audio_input")
        target = self._to_dev(target, d.out_dim, "target")
        if motion.shape[1] != d.motion.seq_len or audio.shape[1] != d.audio.seq_len:
            raise ValueError("sequence lengths must match the position tables")
        batch = motion.shape[0]
        if audio.shape[0] != batch or target.shape[0] != batch:
            raise ValueError("batch sizes differ")
        need = self._lib.fact_train_workspace_bytes(C.byref(self._cdims), batch)
        if self._train_ws is None or self._train_ws.numel() < need + 1024:
            self._train_ws = None
            self._train_ws = torch.empty(need + 1024, dtype=torch.uint8, device=self.device)
        base = (self._train_ws.data_ptr() + 1023) & ~1023
        loss = torch.zeros((), dtype=torch.float32, device=self.device)
        self._grad_flat.zero_()
        with torch.cuda.device(self.device):
            st = torch.cuda.current_stream(self.device).cuda_stream
            ev, n_ev = None, 0
            if stage_events is not None:
                n_ev = len(stage_events)
                ev = (C.c_void_p * n_ev)(*[e.cuda_event if e is not None else None for e in stage_events])
            lib.check(self._lib.fact_train_step(
                C.byref(self._cdims), C.byref(self._cw), C.byref(self._cg), motion.data_ptr(), audio.data_ptr(),
                target.data_ptr(), target.shape[1], batch, float(loss_scale), loss.data_ptr(), base,
                self._train_ws.numel() - (base - self._train_ws.data_ptr()), ev, n_ev, st), "fact_train_step")
        return loss

    def get_metrics(self, eval_config):
        """Metrics are computed offline in the reference (fact_model.py:138-141)."""
        return []
