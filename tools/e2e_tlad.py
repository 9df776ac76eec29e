#!/usr/bin/env python
"""Time the host-pointer TL / AD / Taylor-test / adjoint-test entries (chunked 3-stream pipeline, used slabs
only) on page-locked caller arrays: ms per call and the PCIe volume they move.  usage: tools/e2e_tlad.py [NGPTOT]"""
import importlib, json, sys, time
from pathlib import Path
import numpy as np
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
pkg = importlib.import_module("dwarf-p-cloudsc2-tl-ad_b200")
ngptot = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
nproma, klev = 128, 137
src = pkg.synth_source(seed=0, klon=100, klev=klev)
gpu = pkg.Cloudsc2(pkg.default_params(lregcl=False), klev, src.ceta)
st = pkg.ArrayState(src, nproma, ngptot)
din, dout = pkg.driver.alloc_increments(st.nblocks, klev, nproma)
for n in ("paph", "pap", "pq", "pt", "plude", "plu", "pmfu", "pmfd"):
    din[n][...] = 0.01 * st.a[n]
for a in list(st.a.values()) + list(din.values()) + list(dout.values()):
    gpu.pin(a)
n2, n2h = nproma * klev * st.nblocks, nproma * (klev + 1) * st.nblocks
traj_up, traj_dn = 8 * (14 * n2 + n2h), 8 * (6 * n2 + 4 * n2h)
inc_in, inc_out = 8 * (15 * n2 + n2h), 8 * (6 * n2 + 4 * n2h)
def t(fn, n=3):
    fn(); t0 = time.perf_counter()
    for _ in range(n): fn()
    return (time.perf_counter() - t0) / n
res = {"ngptot": ngptot}
res["nl"] = {"ms": 1e3 * t(lambda: gpu.nl(st)), "h2d": traj_up, "d2h": 8 * (5 * n2 + 2 * n2h)}
res["tl"] = {"ms": 1e3 * t(lambda: gpu.tl(st, din, dout)), "h2d": traj_up + inc_in, "d2h": traj_dn + inc_out}
res["ad"] = {"ms": 1e3 * t(lambda: gpu.ad(st, din, dout)), "h2d": traj_up + inc_in + inc_out, "d2h": traj_dn + inc_in + inc_out}
res["tl_taylor"] = {"ms": 1e3 * t(lambda: gpu.tl_taylor(st)), "h2d": traj_up, "d2h": traj_dn + 8 * n2}
gpu.set_option("lregcl", 1)
res["ad_test"] = {"ms": 1e3 * t(lambda: gpu.ad_test(st)), "h2d": traj_up, "d2h": traj_dn + 8 * n2}
for k, v in res.items():
    if isinstance(v, dict):
        v["columns_per_s"] = ngptot / (v["ms"] * 1e-3)
        v["pcie_gbs_h2d"] = v["h2d"] / (v["ms"] * 1e-3) / 1e9
print(json.dumps(res))
