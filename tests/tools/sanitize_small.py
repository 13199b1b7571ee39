"""compute-sanitizer driver: a small case of every kernel family, checked against the oracle.
    compute-sanitizer --tool memcheck python tests/tools/sanitize_small.py"""
import sys, numpy as np
sys.path.insert(0, '.')
from gcm_filters_b200 import Filter, GridType
from oracle import fixtures, np_oracle
for g, shape in (("IRREGULAR_WITH_LAND", (70, 250)), ("REGULAR_WITH_LAND", (64, 256)), ("TRIPOLAR_POP_WITH_LAND", (66, 140)), ("VECTOR_C_GRID", (40, 70))):
    fields, gv = fixtures.fixture(g, shape)
    fields = tuple(np.stack([f, f * f]) for f in fields)
    fa = dict(filter_scale=6.0, dx_min=1.0)
    if g.startswith("VECTOR"):
        dxm = float(min(gv["dxT"].min(), gv["dyT"].min())); fa = dict(filter_scale=6.0 * dxm, dx_min=dxm)
    flt = Filter(grid_type=GridType[g], grid_vars=gv, **fa)
    out = flt.apply_to_vector(*fields) if len(fields) == 2 else (flt.apply(fields[0]),)
    ref = np_oracle.apply_filter(g, gv, fields, **fa)
    ref = ref if isinstance(ref, tuple) else (ref,)
    err = max(np.nanmax(np.abs(a - b)) for a, b in zip(out, ref))
    print(g, "max abs err", err)
