"""Worker for tests/test_zz_multigpu.py (launched by torch.distributed.run, one process per GPU, NCCL):
N TrainSteps of a PHM model on per-rank batches; every rank must hold bit-identical parameters, optimizer moments and
BN-independent state afterwards, and the averaged gradient must equal the mean of the per-rank gradients."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    from phc_gnn_b200.nn import PHMSkipConnectAdd
    from phc_gnn_b200.parallel import DataParallelPHC
    from phc_gnn_b200.synthetic import make_batch, tiny, workloads
    from phc_gnn_b200.train import TrainStep
    steps = int(sys.argv[1]) if len(sys.argv) > 1 else 5
    for wname in ("hiv", "ppa"):
        wl = tiny(workloads(4)[wname], 64, 3, 24, 8, 20, und_edges=40 if wname == "ppa" else None, head=[32, 16])
        wl.model["dropout_mpnn"] = [0.0] * 3
        wl.model["dropout_dn"] = [0.0] * 2
        finals = {}
        for mode in ("overlap", "single", "graph"):
            torch.manual_seed(100 + rank)                   # different init per rank: the wrapper's broadcast must fix it
            import numpy as np
            np.random.seed(100 + rank)
            model = PHMSkipConnectAdd(**wl.model).to(dev)
            dp = DataParallelPHC(model)
            step = TrainStep(model, wl, None, dp)
            bucket = step.opt.bucket
            if mode == "overlap":
                assert bucket.enable_overlap(model, dp.group)   # opt-in (PHC_OVERLAP_ALLREDUCE=1 does the same in TrainStep)
                bucket.min_chunk_bytes = 0                  # tiny model: force one all-reduce per completed stage
            else:
                assert not bucket.overlap                   # default: one all-reduce after backward
            run = step
            if mode == "graph":
                from phc_gnn_b200.graphed import GraphedTrainStep
                run = GraphedTrainStep(step, capture_after=1)
            model.train()
            losses = []
            for i in range(steps):
                losses.append(float(run(make_batch(wl, seed=1000 * rank + (i % 2)).to(dev))))
            if mode == "overlap":
                assert bucket.launched_last == len(bucket.stage_ends), (bucket.launched_last, bucket.stage_ends)
            if mode == "single":
                assert bucket.launched_last == 1
            if mode == "graph":
                assert run.stats()["replays"] >= steps - 2, run.stats()
            finals[mode] = torch.cat([p.detach().reshape(-1) for p in model.parameters()]).clone()
        # two ranks: (a + b) / 2 is exact and commutative, so slicing the all-reduce cannot change a bit; the replayed graph
        # runs the same kernels and collectives as the eager overlapped step (dropout is off in these tiny configurations)
        assert torch.equal(finals["overlap"], finals["single"]) or world != 2, f"{wname}: overlapped != single all-reduce"
        assert torch.equal(finals["overlap"], finals["graph"]), f"{wname}: graph replay != eager"
        flat = finals["graph"]
        moments = torch.cat([step.opt.exp_avg, step.opt.exp_avg_sq])
        for name, t in (("parameters", flat), ("adam moments", moments)):
            gathered = [torch.empty_like(t) for _ in range(world)]
            dist.all_gather(gathered, t)
            for r in range(1, world):
                assert torch.equal(gathered[0], gathered[r]), f"{wname}: {name} of rank {r} differ from rank 0 after {steps} steps"
        assert torch.isfinite(flat).all()
        # per-rank losses differ (different batches) -> the ranks really trained on different data
        ls = [None] * world
        dist.all_gather_object(ls, losses)
        if world > 1:
            assert ls[0] != ls[1], "ranks saw the same batches"
    dist.barrier()
    if rank == 0:
        print("MULTIGPU_OK", world, flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
