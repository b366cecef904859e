"""Parameter-file interface of mft6.py, kept as is.

Reference: /root/reference/mft6.py:3458-3466 (``key value`` lines, ``#`` comments, value ends at
the first blank or tab), :3469-3572 (typed reads), :3644, :3679-3681; example file
/root/reference/param_koi2298.txt.  Arrays are bracketed, comma separated, without spaces.
``lgmin`` / ``lgmax`` are accepted but ignored, as in the reference (log g range is the fixed
[4, 5.5] of mft6.py:3512).
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np


def read_pardict(path):
    """Raw key -> string dictionary (mft6.py:3458-3466)."""
    pardict = {}
    with open(path) as fi:
        for line in fi:
            if not line.startswith('#') and not line.strip() == '':
                key, val = line.split(' ')[0:2]
                val = val.split('\t')[0]
                pardict[str(key)] = val.strip()
    return pardict


def _truthy(s):
    return 't' in s.lower()


def _float_list(s):
    out = []
    for p in s.strip('[]\n').split(','):
        out.append(np.nan if p.strip() == 'np.nan' else float(p))
    return out


def _name_list(s):
    """Quoted filter names, e.g. ``['sdss,r','j']`` (mft6.py:3526-3527, :3541-3542)."""
    parts = [p.strip('\\') for p in s.strip('[] ').split('\'')]
    return [p for p in parts if len(p) >= 1 and not p == ',']


@dataclass
class FitParams:
    """Typed view of the parameter file; attribute names are the file's keys."""
    models: str = 'btsettl'
    res: int = 1700
    mask: bool = False
    rad_prior: bool = False
    filename: str = ''
    spmin: float = 0.0
    spmax: float = 0.0
    tmin: int = 3000
    tmax: int = 5500
    specmin: int = 3000
    specmax: int = 30000
    plx: float = 0.0
    plx_err: float = 0.0
    dist_fit: bool = True
    cmag: list = field(default_factory=list)
    cerr: list = field(default_factory=list)
    cfilt: list = field(default_factory=list)
    pmag: list = field(default_factory=list)
    perr: list = field(default_factory=list)
    pfilt: list = field(default_factory=list)
    av: float = 0.0
    av_err: float = 0.0
    ra: float = 0.0
    dec: float = 0.0
    nwalk: int = 0
    nstep: int = 0
    nspec: int = 2
    ndust: int = 0
    dirname: str = ''
    fname: str = ''
    nsteps: int = 0
    nburn: int = 0
    real_values: list = field(default_factory=list)
    raw: dict = field(default_factory=dict)


def parse_param_file(path) -> FitParams:
    d = read_pardict(path)
    fp = FitParams(raw=d)
    fp.models = d['models']
    fp.res = int(d['res'])
    fp.mask = _truthy(d.get('mask', 'f'))
    fp.rad_prior = _truthy(d.get('rad_prior', 'f'))
    fp.filename = d['filename']
    fp.spmin, fp.spmax = float(d['spmin']), float(d['spmax'])
    fp.tmin, fp.tmax = int(d['tmin']), int(d['tmax'])
    fp.specmin, fp.specmax = int(d['specmin']), int(d['specmax'])
    fp.plx, fp.plx_err = float(d['plx']), float(d['plx_err'])
    fp.dist_fit = _truthy(d['dist_fit'])
    fp.cmag, fp.cerr = _float_list(d['cmag']), _float_list(d['cerr'])
    fp.cfilt = _name_list(d['cfilt'])
    fp.pmag, fp.perr = _float_list(d['pmag']), _float_list(d['perr'])
    fp.pfilt = _name_list(d['pfilt'])
    fp.av, fp.av_err = float(d['av']), float(d['av_err'])
    fp.ra, fp.dec = float(d['ra']), float(d['dec'])
    fp.nwalk, fp.nstep = int(d['nwalk']), int(d['nstep'])
    fp.nspec, fp.ndust = int(d['nspec']), int(d['ndust'])
    fp.dirname, fp.fname = d['dirname'], d['fname']
    fp.nsteps, fp.nburn = int(d.get('nsteps', 0)), int(d.get('nburn', 0))
    if 'real_values' in d:
        fp.real_values = _float_list(d['real_values'])
    return fp


# KIC -> SDSS photometric conversion of mft6.py:3548-3562 (applied unless 'synth' is in the
# parameter-file name).
_KIC_SLOPE = {'g': 0.0921, 'r': 0.0548, 'i': 0.0696, 'z': 0.1587}
_KIC_INT = {'g': -0.0985, 'r': -0.0383, 'i': -0.0583, 'z': -0.0597}
_KIC_COLOR = {'g': 'g-r', 'r': 'r-i', 'i': 'r-i', 'z': 'i-z'}


def convert_photometry(pmag, pfilt, parfile_name=''):
    """mft6.py:3553-3562: KIC magnitudes of the SDSS bands moved onto the SDSS system."""
    oldphot = list(pmag)
    pfilt = np.array(pfilt)
    if 'synth' in parfile_name:
        return np.array(oldphot)
    phot = np.zeros(len(oldphot))
    for n, p in enumerate(pfilt):
        if 'sdss' in p.lower():
            band = p.split(',')[1]
            c1, c2 = _KIC_COLOR[band].split('-')
            color = (oldphot[np.where('sdss,' + c1 == pfilt)[0][0]]
                     - oldphot[np.where('sdss,' + c2 == pfilt)[0][0]])
            phot[n] = _KIC_INT[band] + _KIC_SLOPE[band] * color + oldphot[n]
        else:
            phot[n] = oldphot[n]
    return phot


def load_observed_spectrum(path, spmin, spmax, mask):
    """mft6.py:3492-3509: read (wl [um], flux, err), telluric mask, window cut, median normalise."""
    data_wl, dsp, de = np.genfromtxt(path, unpack=True)
    if mask:
        keep = ((data_wl <= 0.6860) | ((data_wl >= 0.6880) & (data_wl <= 0.7600))
                | ((data_wl >= 0.7660) & (data_wl <= 0.8210)) | (data_wl > 0.8240))
        data_wl, dsp, de = data_wl[keep], dsp[keep], de[keep]
    cut = np.where((data_wl > spmin) & (data_wl < spmax))
    data_wl, dsp, de = data_wl[cut], dsp[cut], de[cut]
    de = de / np.median(dsp)
    dsp = dsp / np.median(dsp)
    return data_wl, dsp, de
