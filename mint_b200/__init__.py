"""mint_b200: the FACT cross-modal transformer hot path of google-research/mint, rebuilt for B200 (sm_100a).

    from mint_b200 import config_util, model_builder
    cfg = config_util.get_configs_from_pipeline_file(config_util.DEFAULT_CONFIG)
    model = model_builder.build(cfg["model"], is_training=False)
    frames = model.infer_auto_regressive({"motion_input": seed, "audio_input": audio}, steps=1200)
"""
__version__ = "0.1.0"
