"""Multi-rank check of the device-side pooled adaptor (run under torchrun, one rank per GPU):
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 scripts/nccl_exchange_check.py
Every rank feeds its own (theta, alpha) per iteration to ahmc_adapt_exchange_f64 over an ahmc_comm (NCCL all-gather inside the
C ABI); the result must equal the host-side pooled adaptors fed with the rank-ordered Chan merge of all ranks' K5 records
(exchanged here through torch.distributed as the independent path), and be bit-identical on every rank."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist

import ahmc_b200 as A
from ahmc_b200 import adaptation as ad

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
D, N, n_adapts = 33, 200 + 17 * rank, 40  # ragged: every rank owns a different number of chains
windows = (5, 4, 6)
comm = ad.Comm.from_torch_distributed(local)
assert comm is not None and comm.nranks == world
dev_ad = ad.PooledDeviceAdaptor(local, D, N, n_adapts, eps0=0.11, adapt_metric=True, init_buffer=windows[0], term_buffer=windows[1],
                                window_size=windows[2], n_min=3)
host = ad.StanHMCAdaptor(ad.WelfordVar(D, n_min=3), ad.NesterovDualAveraging(0.8, 0.11), *windows)
host.initialize(n_adapts)
rng = np.random.default_rng(100 + rank)
scale = np.exp(np.random.default_rng(7).uniform(-1, 1, D))
worst = 0.0
for i in range(1, n_adapts + 1):
    th = torch.as_tensor(rng.normal(size=(N, D)) * scale + 0.1 * rank, device=dev)
    al = torch.as_tensor(rng.uniform(0.2, 1.3, N), device=dev)
    dev_ad.exchange(th, al, comm, None, flags=0)
    rec = A.adapt_summary(th, al)
    recs = [torch.empty_like(rec) for _ in range(world)]
    dist.all_gather(recs, rec)
    merged = ad.merge_records([r.cpu().numpy() for r in recs])
    host.adapt(merged)
    if i == n_adapts:
        host.finalize()
    s = dev_ad.state()
    assert s["iteration"] == i
    assert np.allclose(s["merged_record"], merged, rtol=1e-12, atol=0), (rank, i)
    assert abs(s["eps"] - host.eps) <= 1e-12 * host.eps, (rank, i, s["eps"], host.eps)
    assert np.allclose(s["Minv"], host.Minv, rtol=1e-12, atol=0), (rank, i)
    worst = max(worst, abs(s["eps"] - host.eps) / host.eps)
    # bit-identical on every rank
    mine = torch.as_tensor(np.concatenate([[s["eps"]], s["Minv"]]), device=dev)
    allv = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(allv, mine)
    assert all(torch.equal(allv[0], v) for v in allv), (rank, i)
# the raw all-gather entry point
g = comm.allgather(torch.full((5,), float(rank), dtype=torch.float64, device=dev))
torch.cuda.synchronize()
assert g.shape == (world, 5) and all(float(g[r, 0]) == r for r in range(world))
dev_ad.destroy()
comm.destroy()
dist.barrier()
if rank == 0:
    print(f"nccl exchange ok: {world} ranks, {n_adapts} iterations, max rel eps error vs host adaptors {worst:.2e}, identical on all ranks")
dist.destroy_process_group()
