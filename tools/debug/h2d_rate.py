import importlib, sys, time, numpy as np
sys.path.insert(0, "/root/repo")
pkg = importlib.import_module("dwarf-p-cloudsc2-tl-ad_b200")
src = pkg.synth_source(seed=0, klon=100, klev=137)
gpu = pkg.Cloudsc2(pkg.default_params(), 137, src.ceta)
st = pkg.ArrayState(src, 128, 163840)
for mode in ("pageable", "registered", "hostalloc"):
    arrs = {}
    ptrs = []
    for n in ("pt", "pq", "pap", "paph", "plu", "plude", "pmfu", "pmfd", "psupsat", "pclv", "b_cml"):
        a = st.a[n]
        if mode == "registered": gpu.pin(a)
        if mode == "hostalloc":
            a, p = gpu.host_alloc_like(a); ptrs.append(p)
        arrs[n] = a
    d = {n: gpu.malloc(a.nbytes) for n, a in arrs.items()}
    tot = sum(a.nbytes for a in arrs.values())
    for rep in range(2):
        t0 = time.perf_counter()
        for n, a in arrs.items(): gpu.h2d(d[n], a)
        t = time.perf_counter() - t0
    print(f"{mode:10s} H2D of {tot/1e9:.2f} GB (whole arrays, sequential): {t*1e3:.1f} ms = {tot/t/1e9:.1f} GB/s")
    for n in d: gpu.free(d[n])
    if mode == "registered":
        for a in arrs.values(): gpu.unpin(a)
    for p in ptrs: gpu.host_free(p)
