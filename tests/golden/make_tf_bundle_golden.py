#!/usr/bin/env python
"""Golden for mint_b200/tf_checkpoint.py from a checkpoint WRITTEN BY TENSORFLOW that ships inside the reference tree
(a tiny LFADS synthetic-data model: 5 float32 variables + an int32 global_step, 10.6 KB).

The TensorFlow files themselves are not copied.  This script (build container only) reads them with this repo's
reader -- verifying every block and tensor checksum TensorFlow stored -- and commits
  * tf_bundle_tensors.npz : the decoded tensors (name -> array), and
  * tf_bundle_manifest.json : sha256 / size of TensorFlow's .index and .data files, entry table (dtype, shape, offset,
    size, masked crc32c per tensor).
tests/test_tf_checkpoint.py then WRITES a bundle from the npz and requires both files to hash to TensorFlow's: the
writer is byte-exact, and the reader is exercised on bytes identical to what TensorFlow produced.

    python tests/golden/make_tf_bundle_golden.py
"""
import hashlib
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from mint_b200 import tf_checkpoint as T  # noqa: E402

SRC = "/root/reference/third_party/tf_models/research/lfads/synth_data/trained_itb/model-65000"
HERE = os.path.dirname(os.path.abspath(__file__))


def main():
    b = T.Bundle(SRC, verify_table=True)
    tensors = {k: b.tensor(k, verify=True) for k in b.keys()}
    np.savez_compressed(os.path.join(HERE, "tf_bundle_tensors.npz"), **tensors)
    manifest = {"source": SRC, "num_shards": b.num_shards, "files": {}, "entries": {}}
    for ext in (".index", ".data-00000-of-00001"):
        raw = open(SRC + ext, "rb").read()
        manifest["files"][ext] = {"sha256": hashlib.sha256(raw).hexdigest(), "bytes": len(raw)}
    for k in b.keys():
        e = b.entries[k]
        manifest["entries"][k] = {"dtype": e.dtype, "shape": list(e.shape), "offset": e.offset, "size": e.size,
                                  "masked_crc32c": e.crc}
    with open(os.path.join(HERE, "tf_bundle_manifest.json"), "w") as f:
        json.dump(manifest, f, indent=1)
    print(json.dumps(manifest, indent=1)[:600])


if __name__ == "__main__":
    main()
