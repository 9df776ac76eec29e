import importlib, sys, numpy as np
sys.path.insert(0, "/root/repo")
pkg = importlib.import_module("dwarf-p-cloudsc2-tl-ad_b200")
src = pkg.synth_source(seed=0, klon=100, klev=137)
gpu = pkg.Cloudsc2(pkg.default_params(lregcl=False), 137, src.ceta)
small = pkg.ArrayState(src, 100, 100); gpu.nl(small)
ref = small.a["pa"][0]            # (137, 100)
for nproma, ngptot in [(128, 163840), (128, 655360), (128, 1310720), (32, 1310720)]:
    ds = pkg.DeviceState.from_source(gpu, src, nproma, ngptot)
    for rep in range(2):
        gpu.nl_dev(ds, src.ptsphy); gpu.sync()
        nb = ds.nblocks
        pa = np.empty((nb, 137, nproma)); gpu.d2h(pa, ds.ptr["pa"])
        flat = np.ascontiguousarray(np.transpose(pa, (1, 0, 2))).reshape(137, -1)[:, :ngptot]
        want = np.tile(ref, ngptot // 100 + 1)[:, :ngptot]
        bad = np.argwhere(flat != want)
        print(nproma, ngptot, "rep", rep, "mismatches", len(bad), "sum|err|", np.abs(flat - want).sum())
        if len(bad):
            lev = np.bincount(bad[:, 0], minlength=137); cols = bad[:, 1]
            print("  levels with errors:", np.nonzero(lev)[0][:20], "first cols:", np.unique(cols)[:12], "col%100:", np.unique(cols % 100)[:12],
                  "blocks:", np.unique(cols // nproma)[:12], "... last", np.unique(cols // nproma)[-3:])
        # also check an input survived expansion
        pt = np.empty((nb, 137, nproma)); gpu.d2h(pt, ds.ptr["pt"])
        fl = np.ascontiguousarray(np.transpose(pt, (1, 0, 2))).reshape(137, -1)[:, :ngptot]
        print("  input pt intact:", np.array_equal(fl, np.tile(src.f["pt"], ngptot // 100 + 1)[:, :ngptot]))
    ds.free()
