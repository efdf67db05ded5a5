#!/usr/bin/env python
"""Golden vectors for mint_b200/scores.py from THE REFERENCE'S OWN tools/calculate_scores.py.

The tool imports vedo, torch and aist_plusplus at module scope; only torch exists here, so the other two are stubbed
(empty modules: the functions exercised below never touch them) and the file is executed from /root/reference.
Run in the build container (the GPU box has no /root/reference):   python tests/golden/make_scores_golden.py
"""
import importlib.util
import os
import sys
import types

import numpy as np

REF = "/root/reference/tools/calculate_scores.py"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "scores_reference.npz")


def load_reference():
    for name in ("vedo", "aist_plusplus", "aist_plusplus.features", "aist_plusplus.features.kinetic",
                 "aist_plusplus.features.manual"):
        sys.modules.setdefault(name, types.ModuleType(name))
    sys.modules["aist_plusplus.features.kinetic"].extract_kinetic_features = None
    sys.modules["aist_plusplus.features.manual"].extract_manual_features = None
    spec = importlib.util.spec_from_file_location("ref_calculate_scores", REF)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    # The tool was written for a SciPy whose sqrtm took `disp`: with disp=False it returned (sqrtm, error estimate).
    # The SciPy in this image dropped the argument, so the module's `linalg` name gets a proxy with the old signature
    # (the numerical routine itself is SciPy's, untouched).
    import scipy.linalg as sl

    class _Linalg:
        def __getattr__(self, name):
            return getattr(sl, name)

        @staticmethod
        def sqrtm(a, disp=True, blocksize=64):
            r = sl.sqrtm(a)
            return r if disp else (r, 0.0)

    mod.linalg = _Linalg()
    return mod


def main():
    ref = load_reference()
    rng = np.random.default_rng(20240917)
    out = {}
    # noisy rotations (what the network emits), plus a few reflections / near-singular blocks
    from scipy.spatial.transform import Rotation
    rots = Rotation.random(2 * 5 * 24, random_state=3).as_matrix().reshape(2, 5, 24, 3, 3)
    noisy = rots + 0.2 * rng.standard_normal(rots.shape)
    noisy[0, 0, 0] = np.diag([1.0, 1.0, -1.0])            # improper: determinant -1
    noisy[0, 0, 1] = 1e-3 * rng.standard_normal((3, 3))   # almost zero
    out["rot_in"] = noisy
    out["rot_out"] = ref.get_closest_rotmat(noisy)
    motion = np.zeros((2, 5, 225))
    motion[..., 6:9] = rng.standard_normal((2, 5, 3))
    motion[..., 9:] = noisy.reshape(2, 5, 216)
    poses, transl = ref.recover_to_axis_angles(motion)
    out["motion"], out["poses"], out["transl"] = motion, poses, transl
    # Frechet distance: generic, identical, and rank-deficient covariances (more dims than samples)
    cases = []
    for k, (dim, n1, n2) in enumerate([(6, 40, 30), (12, 200, 180), (20, 12, 15)]):
        a = rng.standard_normal((n1, dim)) @ rng.standard_normal((dim, dim))
        b = 0.7 * rng.standard_normal((n2, dim)) @ rng.standard_normal((dim, dim)) + 0.3
        out[f"fa{k}"], out[f"fb{k}"] = a, b
        out[f"fd{k}"] = ref.calculate_frechet_distance(a.mean(0), np.cov(a, rowvar=False), b.mean(0),
                                                       np.cov(b, rowvar=False))
        out[f"ffd{k}"] = ref.calculate_frechet_feature_distance(list(a), list(b))
        cases.append(k)
    out["cases"] = np.array(cases)
    out["fd_same"] = ref.calculate_frechet_distance(out["fa0"].mean(0), np.cov(out["fa0"], rowvar=False),
                                                    out["fa0"].mean(0), np.cov(out["fa0"], rowvar=False))
    np.savez_compressed(OUT, **out)
    print("wrote", OUT, {k: np.asarray(v).shape for k, v in out.items()})


if __name__ == "__main__":
    main()
