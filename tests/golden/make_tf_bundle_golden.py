#!/usr/bin/env python
"""Golden TensorBundle for mint_b200/tf_checkpoint.py: a checkpoint WRITTEN BY TENSORFLOW that ships inside the
reference tree (a tiny LFADS synthetic-data model: 5 float32 variables + an int32 global_step, 10.6 KB).  It is data,
not source; it pins the SSTable / BundleEntryProto / tensor layout this repo's reader and writer implement.

Run in the build container (the GPU box has no /root/reference):   python tests/golden/make_tf_bundle_golden.py
"""
import hashlib
import json
import os
import shutil

SRC = "/root/reference/third_party/tf_models/research/lfads/synth_data/trained_itb/model-65000"
DST = os.path.join(os.path.dirname(os.path.abspath(__file__)), "tf_bundle")


def main():
    os.makedirs(DST, exist_ok=True)
    meta = {"source": SRC, "files": {}}
    for ext in (".index", ".data-00000-of-00001"):
        shutil.copyfile(SRC + ext, os.path.join(DST, "model-65000" + ext))
        os.chmod(os.path.join(DST, "model-65000" + ext), 0o644)
        meta["files"]["model-65000" + ext] = hashlib.sha256(open(SRC + ext, "rb").read()).hexdigest()
    with open(os.path.join(DST, "MANIFEST.json"), "w") as f:
        json.dump(meta, f, indent=1)
    print(meta)


if __name__ == "__main__":
    main()
