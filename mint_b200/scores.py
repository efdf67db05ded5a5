"""Offline scoring of generated motion: the numeric core of the reference's
tools/calculate_scores.py -- projection of the predicted 3x3 blocks onto SO(3) (:21-38), axis-angle recovery of the
225-dim motion vector (:41-52), Frechet distance between feature sets (:82-150, :172-190).

Host-side NumPy / SciPy, like the reference's tool (it runs once per evaluation over ~40 clips).  The SMPL forward pass
and the aist_plusplus kinetic / manual feature extractors the reference calls (:153-169) are external packages that are
not in this image; `frechet_feature_distance` therefore takes the extracted feature vectors, and `motion_to_smpl`
produces exactly the (axis-angle poses, translation) pair those extractors' SMPL front end consumes.
"""
from __future__ import annotations

import numpy as np

MOTION_DIM = 225          # 6 padding zeros + 3 translation + 24 joints x 9 rotation-matrix entries (inputs_util.py:70-73)
N_JOINTS = 24


def closest_rotation(mats: np.ndarray) -> np.ndarray:
    """Nearest proper rotation (Frobenius norm) to each 3x3 matrix of `mats` [..., 3, 3]: with M = U S V^T the
    minimiser is U diag(1, 1, det(U V^T)) V^T (calculate_scores.py:21-38)."""
    mats = np.asarray(mats, dtype=np.float64)
    if mats.shape[-2:] != (3, 3):
        raise ValueError(f"expected [..., 3, 3], got {mats.shape}")
    u, _, vt = np.linalg.svd(mats)
    sign = np.sign(np.linalg.det(u @ vt))
    fix = np.ones(mats.shape[:-1], dtype=np.float64)          # [..., 3]: scales the columns of U
    fix[..., 2] = sign
    return (u * fix[..., None, :]) @ vt


def motion_to_smpl(motion: np.ndarray):
    """[B, T, 225] network output -> (axis-angle poses [B, T, 24, 3], root translation [B, T, 3])
    (calculate_scores.py:41-52): columns 6:9 are the translation, 9: the 24 row-major rotation matrices."""
    from scipy.spatial.transform import Rotation
    motion = np.asarray(motion, dtype=np.float64)
    if motion.ndim != 3 or motion.shape[-1] != MOTION_DIM:
        raise ValueError(f"expected [B, T, {MOTION_DIM}], got {motion.shape}")
    b, t, _ = motion.shape
    rot = closest_rotation(motion[..., 9:].reshape(b, t, N_JOINTS, 3, 3))
    poses = Rotation.from_matrix(rot.reshape(-1, 3, 3)).as_rotvec().reshape(b, t, N_JOINTS, 3)
    return poses, motion[..., 6:9].copy()


def frechet_distance(mu1, sigma1, mu2, sigma2, eps: float = 1e-6) -> float:
    """d^2 = |mu1 - mu2|^2 + tr(S1) + tr(S2) - 2 tr((S1 S2)^(1/2))  between N(mu1, S1) and N(mu2, S2)
    (calculate_scores.py:82-150).  tr((S1 S2)^(1/2)) is the sum of the square roots of the eigenvalues of S1 S2, which
    are real and non-negative for covariance matrices; tiny negative / imaginary parts are numerical noise.  A singular
    product (non-finite spectrum) is regularised with eps on both diagonals, as the reference does."""
    mu1, mu2 = np.atleast_1d(np.asarray(mu1, np.float64)), np.atleast_1d(np.asarray(mu2, np.float64))
    s1, s2 = np.atleast_2d(np.asarray(sigma1, np.float64)), np.atleast_2d(np.asarray(sigma2, np.float64))
    if mu1.shape != mu2.shape or s1.shape != s2.shape:
        raise ValueError("the two Gaussians have different dimensions")

    def trace_sqrt(a, b):
        ev = np.linalg.eigvals(a @ b)
        if not np.isfinite(ev).all():
            return None
        if np.abs(ev.imag).max(initial=0.0) > 1e-3 * max(1.0, np.abs(ev.real).max(initial=0.0)):
            raise ValueError(f"imaginary component {np.abs(ev.imag).max()}")
        return float(np.sqrt(np.clip(ev.real, 0.0, None)).sum())

    tr = trace_sqrt(s1, s2)
    if tr is None:
        off = np.eye(s1.shape[0]) * eps
        tr = trace_sqrt(s1 + off, s2 + off)
        if tr is None:
            raise ValueError("covariance product is not finite")
    d = mu1 - mu2
    return float(d @ d + np.trace(s1) + np.trace(s2) - 2.0 * tr)


def frechet_feature_distance(real_features, generated_features) -> float:
    """FID between two lists of feature vectors (calculate_scores.py:172-190): both sets are standardised with the
    mean / std (+1e-10) of the FIRST one, then compared through their sample mean and covariance."""
    a = np.stack([np.asarray(f, np.float64) for f in real_features])
    b = np.stack([np.asarray(f, np.float64) for f in generated_features])
    mean, std = a.mean(axis=0), a.std(axis=0) + 1e-10
    a, b = (a - mean) / std, (b - mean) / std
    return frechet_distance(a.mean(axis=0), np.cov(a, rowvar=False), b.mean(axis=0), np.cov(b, rowvar=False))


def score_feature_files(real_glob: str, generated_glob: str) -> float:
    """FID from cached per-clip feature vectors (`*.npy`), the layout of ./data/aist_features in the reference's
    __main__ (calculate_scores.py:197-200)."""
    import glob
    real = [np.load(f) for f in sorted(glob.glob(real_glob))]
    gen = [np.load(f) for f in sorted(glob.glob(generated_glob))]
    if not real or not gen:
        raise FileNotFoundError("no feature files matched")
    return frechet_feature_distance(real, gen)
