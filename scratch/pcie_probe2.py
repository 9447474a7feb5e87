"""Copy-only ceiling of the host link through gcs_b200_pcie_probe (library streams), per rank.
Usage: [torchrun ...] python scratch/pcie_probe2.py   (all ranks probe at the same time after a barrier)"""
import importlib, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
rank, local, world = (int(os.environ.get(k, d)) for k, d in (("RANK", 0), ("LOCAL_RANK", 0), ("WORLD_SIZE", 1)))
torch.cuda.set_device(local)
if world > 1:
    import torch.distributed as dist
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
gcs = importlib.import_module("2d_geometry_constraint_solver_b200")
capi = gcs.capi
capi.init([local])
rows = []
quick = os.environ.get("PROBE_QUICK")
if quick:
    plan = [(66060288, 26214400, "bench step (compact wire format)", 1, False), (66060288, 26214400, "bench step (compact wire format)", 16, False),
            (66060288, 26214400, "bench step (compact wire format)", 1, True), (256 << 20, 256 << 20, "256 MiB", 1, False)]
else:
    plan = [(up, down, what, pieces, wc)
            for up, down, what in ((80740352, 32505856, "bench step r1"), (62 << 20, 25 << 20, "bench step compact"), (256 << 20, 256 << 20, "256 MiB"))
            for pieces in (1, 4, 10) for wc in (False, True)]
for up, down, what, pieces, wc in plan:
    if True:
        if True:
            if world > 1:
                dist.barrier()
            r = capi.pcie_probe(local, up, down, pieces, wc, 6)
            t = torch.tensor([r["h2d_gbs"], r["d2h_gbs"], r["both_ms_median"]], device="cuda", dtype=torch.float64)
            agg = t.clone()
            if world > 1:
                dist.all_reduce(agg, op=dist.ReduceOp.SUM)
                tmax = t.clone(); dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
            else:
                tmax = t
            if rank == 0:
                rows.append({"what": what, "up_mb": up / 1e6, "down_mb": down / 1e6, "pieces": pieces, "write_combined": wc, "ranks": world,
                             "rank0": r, "sum_h2d_gbs": float(agg[0]), "sum_d2h_gbs": float(agg[1]), "both_ms_median_max_over_ranks": float(tmax[2])})
                print(json.dumps(rows[-1]), flush=True)
if world > 1:
    dist.destroy_process_group()
