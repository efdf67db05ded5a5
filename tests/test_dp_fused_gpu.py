"""fact_dp_adam_step on two GPUs: the fused cross-replica sum + Adam + weight broadcast over NVLink peer memory must
reproduce `NCCL all-reduce -> fact_adam_step` (mint/ctl/single_task_trainer.py:180-187, trainer.py:150), keep the
replicas bit-identical, and reassemble the sharded optimizer moments.  Needs >= 2 GPUs and torch symmetric memory;
skipped otherwise (the single-GPU driver run skips it)."""
import os

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu

SMALL = dict(d=64, heads=4, ff=128, layers=(1, 1, 2), motion_seq=12, audio_seq=20, motion_dim=225, out_dim=225)


def _worker(rank, world, port, q):
    try:
        os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
        torch.cuda.set_device(rank)
        dev = torch.device("cuda", rank)
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
        from mint_b200.fact_model import FACTModel
        from mint_b200.optim import Adam
        from mint_b200.trainer import SingleTaskTrainer
        from oracle import fact_oracle as O
        from tests.helpers import make_config, oracle_dims
        dims = oracle_dims(audio_dim=35, **SMALL)
        out = {"rank": rank}
        # "fused": per slice of the bucket, on a side stream, under the backward; "fused_tail": once, behind it
        for force_p2p, mode in ((False, "fused"), (True, "fused"), (False, "fused_tail")):
            fused_m = FACTModel(make_config(**SMALL), is_training=True, mode="bf16", seed=11, device=dev)
            ref_m = FACTModel(make_config(**SMALL), is_training=True, mode="bf16", seed=11, device=dev)
            fused_o, ref_o = Adam(fused_m, learning_rate=2e-3), Adam(ref_m, learning_rate=2e-3)
            fused_t = SingleTaskTrainer([], "target", fused_m, optimizer=fused_o, overlap=mode)
            ref_t = SingleTaskTrainer([], "target", ref_m, optimizer=ref_o, overlap="none")
            if fused_t.overlap != mode:
                out["skip"] = f"symmetric memory unavailable: {fused_t.fused_error}"
                break
            out["multicast"] = fused_t.arena.mc_ptr is not None
            if force_p2p:
                fused_t.arena.mc_ptr = None
            worst = 0.0
            for step in range(3):
                inp = O.synthetic_inputs(dims, batch=2, seed=100 * rank + step, target_len=5)
                tin = {k: torch.from_numpy(v).float() for k, v in inp.items()}
                fused_t.train_step(dict(tin))
                # the reference path gets exactly the local gradients the fused path reduced
                ref_t.model.forward_backward({k: v for k, v in tin.items() if k != "target"}, tin["target"],
                                             loss_scale=1.0 / world)
                ref_m.flat_gradients.copy_(fused_m.flat_gradients)
                dist.all_reduce(ref_m.flat_gradients)
                ref_o.apply_gradients()
                torch.cuda.synchronize(dev)
                a, b = fused_m.flat_parameters.double(), ref_m.flat_parameters.double()
                worst = max(worst, float((a - b).abs().max() / b.abs().max()))
                ref_m.flat_parameters.copy_(fused_m.flat_parameters)       # keep both trajectories on the same weights
                ref_m.repack()
                # replicas bit-identical: every rank holds the same bytes (MAX == MIN over ranks)
                for t in (fused_m.flat_parameters, fused_m.flat_bf16_parameters.float()):
                    hi, lo = t.clone(), t.clone()
                    dist.all_reduce(hi, op=dist.ReduceOp.MAX)
                    dist.all_reduce(lo, op=dist.ReduceOp.MIN)
                    assert torch.equal(hi, lo), "replicas diverged"
            key = "p2p" if force_p2p else ("default" if mode == "fused" else "tail")
            out[key + "_slices"] = len(fused_t._plan) if fused_t._plan else 0
            out[key + "_worst_rel"] = worst
            sd = fused_o.state_dict()                                       # collective: reassembles the shards
            out[key + "_m_rel"] = float((sd["m"].double() - ref_o.m.double()).abs().max() /
                                        ref_o.m.double().abs().max())
            out[key + "_iter"] = sd["iterations"]
            # the named variables are views of the arena: the model still works (forward through the C ABI)
            pred = fused_m({k: v for k, v in tin.items() if k != "target"})
            assert torch.isfinite(pred).all()
            del fused_t, ref_t
        q.put(out)
        dist.barrier()
        dist.destroy_process_group()
    except Exception as exc:  # surface the failure in the parent
        import traceback
        q.put({"rank": rank, "error": traceback.format_exc()[-1500:]})


def test_fused_dp_step_matches_allreduce_then_adam(fact_lib):
    if not torch.cuda.is_available() or torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29700 + os.getpid() % 200
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    results = [q.get(timeout=600) for _ in procs]
    for p in procs:
        p.join(60)
    for r in results:
        assert "error" not in r, r["error"]
    if any("skip" in r for r in results):
        pytest.skip(results[0].get("skip") or results[1].get("skip"))
    print(results)
    for r in results:
        for key in ("default", "p2p", "tail"):
            assert r[key + "_worst_rel"] < 2e-6, r           # same sums up to the order of two addends
            assert r[key + "_m_rel"] < 1e-5, r
            assert r[key + "_iter"] == 3
        assert r["default_slices"] > 1 and r["tail_slices"] == 0, r
