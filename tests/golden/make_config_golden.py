"""Serialise the reference's fact_v5 pipeline config with the REFERENCE's own generated protobuf classes.

Runs only in the build container (needs /root/reference); the output is committed so the GPU box never reads the
reference tree:
    PROTOCOL_BUFFERS_PYTHON_IMPLEMENTATION=python python tests/golden/make_config_golden.py
(the committed *_pb2.py predate protobuf 3.20 and only import under the pure-python implementation).
"""
import os
import sys

os.environ.setdefault("PROTOCOL_BUFFERS_PYTHON_IMPLEMENTATION", "python")
sys.path.insert(0, "/root/reference")
from google.protobuf import text_format  # noqa: E402
from mint.protos import pipeline_pb2  # noqa: E402

cfg = pipeline_pb2.TrainEvalPipelineConfig()
with open("/root/reference/configs/fact_v5_deeper_t10_cm12.config") as f:
    text_format.Merge(f.read(), cfg)
out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "fact_v5_pipeline_config.binpb")
with open(out, "wb") as f:
    f.write(cfg.SerializeToString(deterministic=True))
print(out, len(cfg.SerializeToString()))
