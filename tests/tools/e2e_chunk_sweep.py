import sys, time
sys.path.insert(0, '.')
import torch
import bench_inputs
from gcm_filters_b200 import Filter, FilterShape, GridType, engine
cfg = bench_inputs.cfg3(nb=62)
fa = dict(cfg["filter_args"]); fa["filter_shape"] = FilterShape[fa["filter_shape"]]
flt = Filter(grid_type=GridType[cfg["grid_type"]], grid_vars=cfg["grid_vars"], **fa)
hin = torch.from_numpy(cfg["fields"][0]).pin_memory(); hout = torch.empty_like(hin).pin_memory()
units = hin.numel() * int(flt.n_steps)
for chunks, nbuf in ((8, 2), (8, 3), (8, 4), (12, 3), (16, 4)):
    engine.PIPELINE_TARGET_CHUNKS = chunks
    engine.PIPELINE_NBUF = nbuf
    for _ in range(2): flt.apply(hin, None, out=hout)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(3): flt.apply(hin, None, out=hout)
    torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 3
    print("nbuf", nbuf, "chunks", chunks, "chunk_nb", engine._pipeline_chunk(62, 69120000), round(dt * 1e3, 1), "ms", round(units / dt / 1e9, 1), "G/s", flush=True)
