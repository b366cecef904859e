"""The C-ABI library builds, loads and exports every symbol include/mcmcspec.h declares (no GPU)."""
import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_functions():
    src = open(os.path.join(ROOT, 'include', 'mcmcspec.h')).read()
    src = re.sub(r'/\*.*?\*/', '', src, flags=re.S)
    names = re.findall(r'\b(ms_[a-z_]+)\s*\(', src)
    return sorted(set(names))


def test_header_declares_the_boundary():
    names = declared_functions()
    for must in ('ms_create', 'ms_destroy', 'ms_lnprob', 'ms_lnprob_parts', 'ms_model_spectrum',
                 'ms_last_error', 'ms_broaden_rows', 'ms_stretch_propose', 'ms_stretch_accept'):
        assert must in names


def test_library_exports_every_declared_symbol(native_lib):
    from mcmc_spec_b200 import native
    for name in declared_functions():
        assert hasattr(native_lib, name), name
    assert sorted(native.EXPORTS) == declared_functions()


def test_struct_mirror_matches(native_lib):
    from mcmc_spec_b200 import native
    assert native_lib.ms_config_size() == ctypes.sizeof(native.MsConfig)
    assert native_lib.ms_abi_version() == native.MS_ABI_VERSION


def test_header_is_plain_c(tmp_path):
    # the boundary must be consumable from C (no C++ / torch types in the signatures)
    src = tmp_path / 't.c'
    src.write_text('#include "mcmcspec.h"\nint main(void){ return (int)sizeof(ms_config) == 0; }\n')
    rc = os.system(f'gcc -std=c99 -Wall -Werror -I {ROOT}/include -c {src} -o {tmp_path}/t.o')
    assert rc == 0


def test_create_fails_loudly_without_gpu_or_on_bad_config(native_lib, small_problem):
    import torch
    from mcmc_spec_b200 import native
    from mcmc_spec_b200.tables import build_tables
    t = build_tables(small_problem)
    cfg, keep = native.make_config(t)
    cfg.abi_version = 99
    h = ctypes.c_void_p()
    assert native_lib.ms_create(ctypes.byref(cfg), ctypes.byref(h)) == 1          # MS_ERR_INVALID
    assert b'ABI' in native_lib.ms_last_error(None)
    if not torch.cuda.is_available():
        with pytest.raises(native.NativeError):
            native.Engine(t)                                                        # no CPU fallback


def test_missing_library_is_an_error(tmp_path):
    from mcmc_spec_b200 import native
    with pytest.raises(native.NativeError):
        native.load_library(str(tmp_path / 'nope.so'))
