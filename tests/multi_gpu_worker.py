"""Worker for tests/test_gpu_multi.py, launched under torchrun on N GPUs: a sharded hierarchical fit
(subjects split over ranks, phi replicated, one NCCL all-reduce per phi step) must reproduce the
single-GPU trajectories."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    import torch
    import torch.distributed as dist
    from ggdmc_b200 import _lib as B
    from ggdmc_b200 import engine as E
    from ggdmc_b200 import workloads as W

    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group(backend="cpu:gloo,cuda:nccl", rank=rank, world_size=world)
    S, ntr, n_iter = 10, 96, 12
    # (schedule, exchange path, kernel path): peer-memory window vs ncclAllReduce; launch sequence vs persistent sampler kernel
    variants = [(B.SCHEDULE_PARALLEL, "p2p", "launches"), (B.SCHEDULE_REFERENCE, "p2p", "launches"), (B.SCHEDULE_PARALLEL, "nccl", "launches"),
                (B.SCHEDULE_PARALLEL, "p2p", "persistent")]
    for schedule, exchange, kernel in variants:
        for k in ("GGDMC_B200_NO_P2P", "GGDMC_B200_PERSIST"):
            os.environ.pop(k, None)
        if exchange == "nccl":
            os.environ["GGDMC_B200_NO_P2P"] = "1"
        if kernel == "persistent":
            os.environ["GGDMC_B200_PERSIST"] = "1"
        ref = None
        if rank == 0:  # the whole problem on one GPU, before any communicator exists
            w = W.hierarchical("t", 2, S, ntr, n_replicate=2)
            tun = W.tuning_for(w, nmc=2, thin=1 << 30, seeds=[11, 12], schedule=schedule, pop_migration_prob=0.3,
                               sub_migration_prob=0.3, device=local)
            eng = E.Engine(w.spec.ct, w.trials, w.spec.p_prior, w.spec.h_prior, tun, w.phi_start, w.subj_start)
            eng.iterate(n_iter)
            ref = eng.state()
            eng.close()
        dist.barrier()
        uid = [E.comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(uid, src=0)
        E.comm_init(world, rank, uid[0], local)
        b, e = W.shard_bounds(S, rank, world)
        w = W.hierarchical("t", 2, S, ntr, n_replicate=2, subject_begin=b, subject_end=e)
        tun = W.tuning_for(w, nmc=2, thin=1 << 30, seeds=[11, 12], schedule=schedule, pop_migration_prob=0.3, sub_migration_prob=0.3,
                           subject_begin=b, n_subject_total=S, device=local)
        eng = E.Engine(w.spec.ct, w.trials, w.spec.p_prior, w.spec.h_prior, tun, w.phi_start, w.subj_start)
        eng.iterate(n_iter)
        st = eng.state()
        eng.close()
        E.comm_finalize()
        out = [None] * world
        dist.all_gather_object(out, st)
        if rank == 0:
            theta = np.concatenate([o["theta"] for o in out])
            ll = np.concatenate([o["ll"] for o in out])
            assert np.array_equal(theta, ref["theta"]), f"schedule {schedule}: sharded subject thetas differ from the single-GPU run"
            assert np.allclose(ll, ref["ll"], rtol=1e-9, atol=0)
            for o in out:  # phi replicated: identical on every rank and equal to the single-GPU phi
                assert np.array_equal(o["phi_theta"], ref["phi_theta"])
                assert np.array_equal(o["phi_theta"], out[0]["phi_theta"]) and np.array_equal(o["phi_ll"], out[0]["phi_ll"])
                assert np.allclose(o["phi_ll"], ref["phi_ll"], rtol=1e-9, atol=0)
            assert not np.array_equal(ref["phi_theta"], w.phi_start.theta)
            print(f"schedule {schedule}, exchange {exchange}, {kernel}: {world}-GPU sharded run == single-GPU run ({S} subjects, {n_iter} iterations, "
                  f"theta bit-identical, phi identical on every rank)", flush=True)
        dist.barrier()
    dist.destroy_process_group()
    if rank == 0:
        print("MULTI_GPU_OK", flush=True)


if __name__ == "__main__":
    main()
