"""CPU: the C-ABI library loads and exports exactly what include/lsi_b200.h declares, and the ctypes table in
lsi/_b200.py mirrors it (no compute calls here -- there is no GPU in the build container)."""
import ctypes
import os
import re
import subprocess

import pytest

from _util import ROOT

HEADER = os.path.join(ROOT, 'include', 'lsi_b200.h')


def _declared():
    src = open(HEADER).read()
    return sorted(set(re.findall(r'LSI_B200_API\s+[\w\s\*]+?\b(lsi_b200_\w+)\s*\(', src)))


def test_header_declares_entry_points():
    names = _declared()
    assert len(names) >= 24
    for must in ('lsi_b200_forward_splat', 'lsi_b200_forward_splat_backward', 'lsi_b200_forward_splat_host',
                 'lsi_b200_splat', 'lsi_b200_bilinear', 'lsi_b200_zbuf_composition_loss', 'lsi_b200_photo_loss'):
        assert must in names


def test_library_exports_every_declared_symbol():
    from lsi import _b200
    assert os.path.exists(_b200.LIB_PATH), 'build first: python -c "import __graft_entry__ as g; g.build()"'
    lib = ctypes.CDLL(_b200.LIB_PATH)
    for name in _declared():
        assert hasattr(lib, name), name
    out = subprocess.run(['nm', '-D', '--defined-only', _b200.LIB_PATH], capture_output=True, text=True).stdout
    exported = sorted(set(re.findall(r' T (lsi_b200_\w+)', out)))
    assert exported == _declared()


def test_ctypes_table_matches_header():
    from lsi import _b200
    assert sorted(_b200.SIGNATURES) == _declared()
    src = open(HEADER).read()
    for name, (_, argtypes) in _b200.SIGNATURES.items():
        m = re.search(r'LSI_B200_API[^;(]*\b%s\s*\(([^;]*?)\)\s*;' % name, src, re.S)
        assert m, name
        args = m.group(1).strip()
        n = 0 if args in ('', 'void') else len(args.split(','))
        assert n == len(argtypes), (name, n, len(argtypes))


def test_version_and_error_string_without_gpu():
    from lsi import _b200
    lib = _b200.lib()
    assert lib.lsi_b200_version() >= 100
    desc = _b200.SplatDesc(0, 1, 8, 8, 8, 8, 1.0, 0.0, 1.0, 10.0, 1, 0, 3, 1, 1, 0)     # n_layers = 0: rejected
    assert lib.lsi_b200_forward_splat_workspace_bytes(desc) == 0
    assert b'n_layers' in lib.lsi_b200_last_error()
    desc = _b200.SplatDesc(4, 64, 256, 832, 256, 832, 1.0, 1e-3, 0.4, 50.0, 1, 0, 3, 1, 1, 0)
    assert lib.lsi_b200_forward_splat_workspace_bytes(desc) > 64 * 16 * 4
    assert int(lib.lsi_b200_loss_partials_count()) >= 256


def test_missing_library_fails_loudly(monkeypatch):
    from lsi import _b200
    monkeypatch.setattr(_b200, '_lib', None)
    monkeypatch.setattr(_b200, 'LIB_PATH', '/nonexistent/liblsi_b200.so')
    with pytest.raises(RuntimeError, match='no CPU/torch fallback'):
        _b200.lib()


def test_cpu_tensors_are_rejected():
    import torch
    from lsi.geometry import ldi
    from lsi.loss import loss
    with pytest.raises(RuntimeError, match='CUDA only'):
        ldi.forward_splat((torch.zeros(1, 1, 4, 4, 3), None, torch.zeros(1, 1, 4, 4, 1)), None, None, None, None, None)
    with pytest.raises(RuntimeError, match='CUDA only'):
        loss.decreasing_disp_loss(torch.zeros(2, 1, 4, 4, 1))
