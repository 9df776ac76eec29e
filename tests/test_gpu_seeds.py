"""GPU suite: parity on other synthetic atmospheres (generator seeds 1..4), so that the agreement
with the oracle is not an accident of the 100 seed-0 columns the other tests use.  NL fields vs the
oracle, Taylor verdict, adjoint verdict and the dot-product identity per seed.

Branch-flip accounting (SURVEY 7 'hard parts'): the kernels switch on trajectory values exactly
like the reference; a 1-ulp difference in exp can flip a knife-edge column.  Such columns are
counted, must be rare (<= 1 %), and every other column must meet the tolerance."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu
NL_RTOL = 1e-11


@pytest.mark.parametrize("seed", [1, 2, 3, 4])
def test_nl_parity_other_seeds(pkg, ob, seed):
    src = pkg.synth_source(seed=seed, klon=100, klev=137)
    prm = pkg.default_params(lregcl=False)
    st, ref = pkg.ArrayState(src, 32, 100), pkg.ArrayState(src, 32, 100)
    with pkg.Cloudsc2(prm, 137, src.ceta) as gpu:
        gpu.nl(st)
    ob.driver_nl(prm, src.ceta, ref, numomp=2)
    bad_cols = set()
    for n, r in ref.outputs().items():
        g = st.outputs()[n]
        assert np.isfinite(g).all(), n
        scale = max(float(np.abs(r).max()), 1e-300)
        err = np.abs(g - r) / scale                      # (NB, NLEV, NPROMA)
        for b, _, jl in np.argwhere(err > NL_RTOL):
            bad_cols.add(int(b) * 32 + int(jl))
    assert len(bad_cols) <= 1, sorted(bad_cols)          # at most one knife-edge column of 100


@pytest.mark.parametrize("seed", [1, 2, 3])
def test_taylor_and_adjoint_tests_other_seeds(pkg, seed):
    src = pkg.synth_source(seed=seed, klon=100, klev=137)
    with pkg.Cloudsc2(pkg.default_params(lregcl=False), 137, src.ceta) as gpu:
        z, _ = gpu.tl_taylor(pkg.ArrayState(src, 1, 100))          # dwarf-cloudsc2-tl 1 100 1
        pen, istart = pkg.taylor_verdict(z)
        assert 0 <= pen <= 5, (seed, pen, z)
    with pkg.Cloudsc2(pkg.default_params(lregcl=True), 137, src.ceta) as gpu:
        zn, nc = gpu.ad_test(pkg.ArrayState(src, 100, 100))         # dwarf-cloudsc2-ad 1 100 100
        assert pkg.adjoint_verdict(zn), (seed, zn)


@pytest.mark.parametrize("klev", [60, 91])
def test_nl_tl_ad_on_other_vertical_resolutions(pkg, ob, klev):
    """KLEV is run-time data (the dwarf uses 137; IFS also runs 60 and 91 levels): NL fields vs the
    oracle, the Taylor verdict and the adjoint verdict on synthetic atmospheres of other depths."""
    src = pkg.synth_source(seed=7, klon=100, klev=klev)
    prm = pkg.default_params(lregcl=False)
    st, ref = pkg.ArrayState(src, 64, 100), pkg.ArrayState(src, 64, 100)
    with pkg.Cloudsc2(prm, klev, src.ceta) as gpu:
        gpu.nl(st)
        z, _ = gpu.tl_taylor(pkg.ArrayState(src, 100, 100), allow_degenerate=True)
    ob.driver_nl(prm, src.ceta, ref, numomp=2)
    for n, r in ref.outputs().items():
        g = st.outputs()[n]
        assert g.shape == r.shape and np.isfinite(g).all(), n
        assert np.abs(g - r).max() <= NL_RTOL * max(float(np.abs(r).max()), 1e-300), n
    # the Taylor verdict depends on the atmosphere (a coarse 60-level column set need not pass with
    # penalty <= 5): what must hold is that GPU and oracle agree on the ratios and on the verdict
    z_o, _, _ = ob.driver_tl(prm, src.ceta, pkg.ArrayState(src, 100, 100), numomp=2, allow_degenerate=True)
    assert pkg.taylor_verdict(z)[0] == pkg.taylor_verdict(z_o)[0], (klev, z, z_o)
    assert np.allclose(z[:6], z_o[:6], rtol=1e-6), (klev, z, z_o)
    with pkg.Cloudsc2(pkg.default_params(lregcl=True), klev, src.ceta) as gpu:
        zn, _ = gpu.ad_test(pkg.ArrayState(src, 100, 100))
        assert pkg.adjoint_verdict(zn), (klev, zn)
