"""Debug driver for PeerBandedFilter on N GPUs (torchrun)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch, torch.distributed as dist

def log(*a):
    print(f"[rank {os.environ.get('RANK')}] {time.time() % 1000:.2f}", *a, flush=True)

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
log("pg ready")
import torch.distributed._symmetric_memory as symm_mem
t = symm_mem.empty(1 << 20, dtype=torch.uint8, device=f"cuda:{local}")
log("symm empty ok")
hdl = symm_mem.rendezvous(t, dist.group.WORLD)
log("rendezvous ok", [hex(p) for p in hdl.buffer_ptrs])
t.fill_(rank + 1); torch.cuda.synchronize(); dist.barrier()
peer = hdl.get_buffer((rank + 1) % world, (16,), torch.uint8, 0)
log("peer read", peer[:4].tolist())
from gcm_filters_b200 import Filter, GridType
from gcm_filters_b200.scheduler import PeerBandedFilter
from oracle import fixtures
(f,), gv = fixtures.fixture("IRREGULAR_WITH_LAND", (90, 160))
fb = np.stack([f, f * f])
flt = Filter(grid_type=GridType.IRREGULAR_WITH_LAND, grid_vars=gv, filter_scale=6.0, dx_min=1.0)
single = flt.apply(fb, None)
log("single ok")
pbf = PeerBandedFilter(flt, rank, world)
st = pbf.stage(fb)
log("staged", st["north"], st["south"], st["nyl"], st["nyl_north"], st["nyl_south"])
bar = pbf.run(st)
log("launched")
torch.cuda.synchronize()
log("synced")
out = bar[0].cpu().numpy()
log("equal:", np.array_equal(out, single[..., st["j0"]:st["j1"], :], equal_nan=True), float(np.nanmax(np.abs(out - single[..., st["j0"]:st["j1"], :]))))
dist.barrier(); dist.destroy_process_group()
