"""Parameter-file interface and passband loader (host side, CPU)."""
import os
import struct

import numpy as np
import pytest

from mcmc_spec_b200 import paramfile, passbands

REF = '/root/reference'

PARAM_TEXT = """### header
models btsettl\t\t\t\t#Which set of model spectra to use
dirname koi2298\t\t\t\t#dir
fname 2298 \t\t\t\t\t#the emcee filename
res 1700 \t\t\t\t\t#spectral resolution
tmin 3000 \t\t\t\t\t#min
tmax 5500 \t\t\t\t\t#max
lgmin 4\t\t\t\t\t#ignored
lgmax 5\t\t\t\t\t\t#ignored
specmin 3000 \t\t\t\t#
specmax 30000\t\t\t\t#
mask True
rad_prior True

filename Data/koi2298_HET.txt   #path
spmin 0.645 \t\t\t\t\t#minimum spectrum in um
spmax 0.84\t\t\t\t\t\t#max

cmag [2.08,1.3]\t\t\t#contrast magnitudes
cerr [0.14,0.02]\t\t\t#contrast error
cfilt ['lp600','Kp']\t\t\t#contrast filters
pmag [13.815,13.505,13.355,12.323,11.826,11.735] \t#unresolved magnitudes
perr [0.1,0.1,0.1,0.026,0.022,0.019]\t#unresolved mag error
pfilt ['sdss,r','sdss,i','sdss,z','j','h','k']\t\t\t#unresolved mag filters
plx 2.0732e-3\t\t\t\t#parallax
plx_err 0.0277e-3 \t\t\t#err
dist_fit True
av 0.106\t\t\t\t\t#extinction
av_err 0.01 \t\t\t\t#err
ra 288.456118
dec 45.802226
nwalk 150 \t\t\t\t\t#walkers
nstep 400\t\t\t\t\t#steps
nspec 2\t\t\t\t\t\t#stars
ndust 0\t\t\t\t\t\t#disks
nburn 301\t\t\t\t\t#burn
nsteps 15001\t\t\t\t\t#sample
real_values [0,0,0,0,0,0,0,0]
"""


def _check_koi2298(fp):
    assert fp.models == 'btsettl' and fp.res == 1700 and fp.mask and fp.rad_prior and fp.dist_fit
    assert (fp.tmin, fp.tmax, fp.specmin, fp.specmax) == (3000, 5500, 3000, 30000)
    assert fp.cfilt == ['lp600', 'Kp'] and fp.pfilt == ['sdss,r', 'sdss,i', 'sdss,z', 'j', 'h', 'k']
    assert fp.cmag == [2.08, 1.3] and fp.cerr == [0.14, 0.02] and len(fp.pmag) == 6
    assert fp.plx == pytest.approx(2.0732e-3) and fp.plx_err == pytest.approx(0.0277e-3)
    assert (fp.nwalk, fp.nstep, fp.nspec, fp.ndust, fp.nburn, fp.nsteps) == (150, 400, 2, 0, 301, 15001)
    assert fp.fname == '2298' and fp.dirname == 'koi2298' and fp.spmin == 0.645 and fp.spmax == 0.84


def test_parse_param_text(tmp_path):
    f = tmp_path / 'param_x.txt'
    f.write_text(PARAM_TEXT)
    _check_koi2298(paramfile.parse_param_file(str(f)))


@pytest.mark.skipif(not os.path.exists(REF), reason='reference checkout not present (GPU box)')
def test_parse_reference_param_file():
    _check_koi2298(paramfile.parse_param_file(os.path.join(REF, 'param_koi2298.txt')))


def test_kic_to_sdss_conversion():
    pmag = [13.815, 13.505, 13.355, 12.323, 11.826, 11.735]
    pf = ['sdss,r', 'sdss,i', 'sdss,z', 'j', 'h', 'k']
    out = paramfile.convert_photometry(pmag, pf, 'param_koi2298.txt')
    assert out[0] == pytest.approx(-0.0383 + 0.0548 * (13.815 - 13.505) + 13.815)      # r uses r-i
    assert out[1] == pytest.approx(-0.0583 + 0.0696 * (13.815 - 13.505) + 13.505)      # i uses r-i
    assert out[2] == pytest.approx(-0.0597 + 0.1587 * (13.505 - 13.355) + 13.355)      # z uses i-z
    assert list(out[3:]) == pmag[3:]
    assert list(paramfile.convert_photometry(pmag, pf, 'param_synth.txt')) == pmag     # 'synth' files are left alone


def test_load_observed_spectrum_mask_cut_normalise(tmp_path):
    wl = np.linspace(0.60, 0.90, 3001)
    flux = 2.0e-15 * (1 + 0.1 * np.sin(wl * 50)); err = flux / 100
    f = tmp_path / 's.txt'
    np.savetxt(f, np.column_stack([wl, flux, err]))
    w, d, e = paramfile.load_observed_spectrum(str(f), 0.645, 0.84, True)
    assert w.min() > 0.645 and w.max() < 0.84
    for lo, hi in ((0.6860, 0.6880), (0.7600, 0.7660), (0.8210, 0.8240)):
        assert not np.any((w > lo) & (w < hi))
    assert np.median(d) == pytest.approx(1.0, rel=1e-15) and np.allclose(e / d, 0.01, rtol=1e-12)
    w2, _, _ = paramfile.load_observed_spectrum(str(f), 0.645, 0.84, False)
    assert w2.size > w.size


def _write_bintable(path, cols):
    names = list(cols)
    n = len(cols[names[0]])
    def card(k, v, quote=False):                                                   # noqa: E306
        v = f"'{v:<8}'" if quote else f"{v:>20}"
        return f"{k:<8}= {v}".ljust(80)
    prim = ''.join([card('SIMPLE', 'T'), card('BITPIX', 16), card('NAXIS', 0), card('EXTEND', 'T'), 'END'.ljust(80)])
    prim = prim.ljust(2880)
    hdr = [card('XTENSION', 'BINTABLE', True), card('BITPIX', 8), card('NAXIS', 2), card('NAXIS1', 4 * len(names)),
           card('NAXIS2', n), card('PCOUNT', 0), card('GCOUNT', 1), card('TFIELDS', len(names))]
    for k, nm in enumerate(names, 1):
        hdr += [card(f'TTYPE{k}', nm, True), card(f'TFORM{k}', 'E', True)]
    hdr = (''.join(hdr) + 'END'.ljust(80))
    hdr = hdr.ljust(-(-len(hdr) // 2880) * 2880)
    body = b''.join(struct.pack('>' + 'f' * len(names), *[float(cols[nm][i]) for nm in names]) for i in range(n))
    body = body.ljust(-(-len(body) // 2880) * 2880, b'\0')
    with open(path, 'wb') as fh:
        fh.write(prim.encode() + hdr.encode() + body)


def test_fits_reader_and_float32_mean(tmp_path):
    wl = np.linspace(10600, 14400, 107).astype(np.float32)
    thr = np.hanning(107).astype(np.float32)
    os.makedirs(tmp_path / 'bps')
    _write_bintable(tmp_path / 'bps' / '2mass_j_001_syn.fits', {'WAVELENGTH': wl, 'THROUGHPUT': thr, 'ERROR': thr * 0})
    w, t, nre, cwl = passbands.get_transmission('j', 1700, str(tmp_path / 'bps'))
    assert w.dtype == np.float32 and np.array_equal(w, wl) and np.array_equal(t, thr)
    assert isinstance(cwl, np.float32) and cwl == np.mean(wl)                        # float32 mean, as astropy would give


def test_name_resolution_quirks(tmp_path):
    d = tmp_path / 'bps'; os.makedirs(d)
    np.savetxt(d / 'keck_kp.txt', np.column_stack([np.linspace(1.9, 2.4, 20), np.hanning(20)]))
    np.savetxt(d / 'lp600.csv', np.column_stack([np.linspace(600, 1000, 30), np.hanning(30)]), delimiter=',')
    np.savetxt(d / '2MASS_2MASS.Ks.dat', np.column_stack([np.linspace(19000, 24000, 25), 0.5 * np.hanning(25)]))
    np.savetxt(d / 'DSSI_880nm.dat', np.column_stack([np.linspace(8500, 9100, 15), np.hanning(15)]))
    np.savetxt(d / 'Kepler_Kepler.K.dat', np.column_stack([np.linspace(4000, 9000, 15), np.hanning(15)]))
    w, t, _, c = passbands.get_transmission('Kp', 1700, str(d))          # bare 'Kp' is Keck K', microns -> A
    assert w[0] == pytest.approx(19000.0) and w[-1] == pytest.approx(24000.0)
    w, t, _, _ = passbands.get_transmission('lp600', 1700, str(d))       # nm -> A
    assert w[0] == pytest.approx(6000.0)
    w, t, _, _ = passbands.get_transmission('k', 1700, str(d))           # 2MASS Ks normalised to max
    assert t.max() == pytest.approx(1.0)
    assert passbands.get_transmission('880', 1700, str(d))[0][0] == 8500.0
    assert passbands.get_transmission('kepler', 1700, str(d))[0][0] == 4000.0


@pytest.mark.skipif(not os.path.exists(REF), reason='reference checkout not present (GPU box)')
def test_reference_bps_tables_match_committed_fixture():
    import golden_io
    fx = np.load(os.path.join(os.path.dirname(__file__), 'golden', 'c2_koi2298.npz'))
    ctm, ptm, tmi, tma = passbands.assemble_filter_tables(['lp600', 'Kp'], ['sdss,r', 'sdss,i', 'sdss,z', 'j', 'h', 'k'],
                                                          1700, os.path.join(REF, 'bps'))
    cw, ct = golden_io.unpack_curves(fx, 'c')
    assert all(np.array_equal(a, b) for a, b in zip(cw, ctm[0])) and float(tmi) == float(fx['tmi'])
    assert [float(c) for c in ptm[3]] == list(fx['p_cwl'])
    w, d, e = paramfile.load_observed_spectrum(os.path.join(REF, 'koi2298_HET.txt'), 0.645, 0.84, True)
    assert w.size == 1945 and np.array_equal(d, fx['data_flux'])
