"""Host-side training logic that needs no GPU: the reference's ManualStepping schedule (learning_schedules.py:19-67) and
the config -> schedule factory (trainer.py:49-96)."""
import pytest

from mint_b200 import config_util
from mint_b200.optim import ManualStepping, learning_rate_from_config


def test_manual_stepping_fact_v5_schedule():
    cfg = config_util.get_configs_from_pipeline_file(config_util.DEFAULT_CONFIG)
    lr = learning_rate_from_config(cfg["train_config"])
    assert lr(0) == pytest.approx(1e-4) and lr(99999) == pytest.approx(1e-4)
    assert lr(100000) == pytest.approx(1e-5) and lr(149999) == pytest.approx(1e-5)      # boundary is inclusive (>=)
    assert lr(150000) == pytest.approx(1e-6) and lr(2400000) == pytest.approx(1e-6)


def test_manual_stepping_warmup_and_validation():
    s = ManualStepping([10, 20], [0.0, 1.0, 0.1], warmup=True)
    assert [s(i) for i in (0, 5, 9, 10, 19, 20, 99)] == pytest.approx([0.0, 0.5, 0.9, 1.0, 1.0, 0.1, 0.1])
    with pytest.raises(ValueError):
        ManualStepping([10, 10], [1.0, 0.5, 0.1])
    with pytest.raises(ValueError):
        ManualStepping([10], [1.0])
    with pytest.raises(ValueError):
        ManualStepping([0, 5], [1.0, 0.5, 0.1])


def test_unsupported_learning_rate_kind():
    cfg = config_util.get_configs_from_pipeline_file(
        config_util.DEFAULT_CONFIG, "train_config { learning_rate { cosine_decay_learning_rate {} } }")
    with pytest.raises(ValueError):
        learning_rate_from_config(cfg["train_config"])


def test_shards_of_a_bucket_slice_tile_it_exactly():
    """fact_dp_adam_range / Adam.dp_fused_range: every rank updates its piece of a slice of the gradient bucket; the
    pieces of all ranks must tile the slice exactly (no element updated twice or never), start on 4-element boundaries
    (16-byte fp32 loads) and be multiples of 8 except the last (16-byte bf16 stores) -- for every world size the kernel
    accepts and for slice sizes like the ones FACTModel.gradient_stages produces (multiples of 4, from one bias vector to
    a whole transformer layer)."""
    from mint_b200.optim import shard_of_range
    for world in range(1, 17):
        for offset, count in ((0, 4), (8, 800), (1024, 180228), (4096, 7483072), (12, 120406980 // 4 * 4)):
            pieces = [shard_of_range(offset, count, r, world) for r in range(world)]
            pos = offset
            for lo, cnt in pieces:
                assert cnt >= 0 and lo % 4 == 0
                if cnt:
                    assert lo == pos
                    pos += cnt
            assert pos == offset + count
            assert all(cnt % 8 == 0 for _, cnt in pieces[:-1] if cnt and _ + cnt != offset + count)
