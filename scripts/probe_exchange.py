"""torchrun script: batch-1 (and a few more) step times of the sharded index with the peer exchange (K5x) and with
NCCL all_gather + K5, same shards, same process.  usage: torchrun ... probe_exchange.py rows [nq,nq,...]"""
import os, sys, statistics
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch
import torch.distributed as dist

rank, world, local_rank = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local_rank)
dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
from sgic_b200.sharded import ShardedIndexFlatIP, shard_range
from sgic_b200.synth import fill_index_random, random_unit_queries

R = int(sys.argv[1]); nqs = [int(x) for x in sys.argv[2].split(",")] if len(sys.argv) > 2 else [1]
d, k = 512, 10
chunk = 1 << 16
lo, hi = shard_range(R, world, rank)
idxs = {}
for mode in ("peer", "nccl"):
    os.environ["SGIC_EXCHANGE"] = mode
    idx = ShardedIndexFlatIP(d)
    idx.add_local(lambda local: fill_index_random(local, hi - lo, row0=lo, chunk_rows=chunk), lo, hi - lo, R)
    idxs[mode] = idx
for nq in nqs:
    q = torch.from_numpy(random_unit_queries(nq, d)).cuda()
    res = {}
    for mode, idx in idxs.items():
        for _ in range(20):
            idx.search_torch(q, k)
        torch.cuda.synchronize(); dist.barrier()
        samples = []
        for _ in range(5):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(50):
                idx.search_torch(q, k)
            e1.record(); torch.cuda.synchronize()
            t = torch.tensor([e0.elapsed_time(e1) / 50], device="cuda", dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            samples.append(float(t.item()))
            dist.barrier()
        res[mode] = (statistics.median(samples), min(samples))
    if rank == 0:
        print(f"rows={R} world={world} nq={nq} k={k}: exchange={idxs['peer'].exchange}: peer {res['peer'][0]*1e3:.1f} us/step (best {res['peer'][1]*1e3:.1f}) | "
              f"nccl {res['nccl'][0]*1e3:.1f} us/step (best {res['nccl'][1]*1e3:.1f})", flush=True)
for idx in idxs.values():
    idx.close()
dist.destroy_process_group()
