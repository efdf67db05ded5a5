"""mint_b200/scores.py against vectors produced by the reference's own tools/calculate_scores.py
(tests/golden/make_scores_golden.py), plus the properties the quantities must have."""
import os

import numpy as np
import pytest

from mint_b200 import scores

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "scores_reference.npz")


@pytest.fixture(scope="module")
def gold():
    return np.load(GOLD)


def test_closest_rotation_matches_the_reference(gold):
    got = scores.closest_rotation(gold["rot_in"])
    np.testing.assert_allclose(got, gold["rot_out"], rtol=0, atol=1e-12)
    # proper rotations: R R^T = I, det = +1 (also for the reflected and the near-zero block)
    eye = np.einsum("...ij,...kj->...ik", got, got)
    np.testing.assert_allclose(eye, np.broadcast_to(np.eye(3), eye.shape), atol=1e-12)
    np.testing.assert_allclose(np.linalg.det(got), 1.0, atol=1e-12)
    # idempotent, and a rotation is its own projection
    np.testing.assert_allclose(scores.closest_rotation(got), got, atol=1e-12)


def test_motion_to_smpl_matches_the_reference(gold):
    poses, transl = scores.motion_to_smpl(gold["motion"])
    np.testing.assert_allclose(transl, gold["transl"], rtol=0, atol=0)
    np.testing.assert_allclose(poses, gold["poses"], rtol=0, atol=1e-10)
    assert poses.shape == (2, 5, 24, 3)
    with pytest.raises(ValueError):
        scores.motion_to_smpl(np.zeros((2, 5, 219)))


def test_frechet_distance_matches_the_reference(gold):
    for k in gold["cases"]:
        a, b = gold[f"fa{k}"], gold[f"fb{k}"]
        fd = scores.frechet_distance(a.mean(0), np.cov(a, rowvar=False), b.mean(0), np.cov(b, rowvar=False))
        assert fd == pytest.approx(float(gold[f"fd{k}"]), rel=1e-6, abs=1e-6), k
        ffd = scores.frechet_feature_distance(list(a), list(b))
        assert ffd == pytest.approx(float(gold[f"ffd{k}"]), rel=1e-6, abs=1e-6), k
    a = gold["fa0"]
    same = scores.frechet_distance(a.mean(0), np.cov(a, rowvar=False), a.mean(0), np.cov(a, rowvar=False))
    assert abs(same) < 1e-6 and abs(float(gold["fd_same"])) < 1e-6


def test_frechet_distance_properties():
    rng = np.random.default_rng(0)
    mu1, mu2 = rng.standard_normal(5), rng.standard_normal(5)
    a, b = rng.standard_normal((5, 5)), rng.standard_normal((5, 5))
    s1, s2 = a @ a.T + np.eye(5), b @ b.T + np.eye(5)
    d12 = scores.frechet_distance(mu1, s1, mu2, s2)
    assert d12 == pytest.approx(scores.frechet_distance(mu2, s2, mu1, s1), rel=1e-9)     # symmetric
    assert d12 > 0
    # commuting (diagonal) covariances: closed form  |dmu|^2 + sum (sqrt(a) - sqrt(b))^2
    da, db = rng.uniform(0.5, 2, 5), rng.uniform(0.5, 2, 5)
    want = float((mu1 - mu2) @ (mu1 - mu2) + ((np.sqrt(da) - np.sqrt(db)) ** 2).sum())
    assert scores.frechet_distance(mu1, np.diag(da), mu2, np.diag(db)) == pytest.approx(want, rel=1e-9)
    with pytest.raises(ValueError):
        scores.frechet_distance(mu1, s1, mu2[:4], s2[:4, :4])


def test_score_feature_files(tmp_path, gold):
    a, b = gold["fa0"], gold["fb0"]
    for i, v in enumerate(a):
        np.save(tmp_path / f"real_{i:03d}_kinetic.npy", v)
    for i, v in enumerate(b):
        np.save(tmp_path / f"gen_{i:03d}_kinetic.npy", v)
    got = scores.score_feature_files(str(tmp_path / "real_*_kinetic.npy"), str(tmp_path / "gen_*_kinetic.npy"))
    assert got == pytest.approx(float(gold["ffd0"]), rel=1e-6)
    with pytest.raises(FileNotFoundError):
        scores.score_feature_files(str(tmp_path / "none_*.npy"), str(tmp_path / "gen_*.npy"))
