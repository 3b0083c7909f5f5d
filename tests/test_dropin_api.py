"""CPU: the product `lsi` package is a drop-in for the reference's modules at the level the two scripts use them
(SURVEY.md section 8b): every public function of lsi/geometry/{ldi,sampling,projection}.py, lsi/nnutils/{helpers,nets}.py and
lsi/loss/loss.py exists under the same module path with the same argument names, order and literal defaults (product-only
arguments are `_`-prefixed and come last), and every call ldi_enc_dec.py / ldi_pred_eval.py make into those modules binds against
the product's signature.  The table is extracted from the reference tree by oracle/gen_api_signatures.py; in the build container (where that tree is
mounted) the committed table is also checked to be current."""
import importlib
import inspect
import json
import os

import pytest

from _util import GOLD

TABLE = json.load(open(os.path.join(GOLD, 'ref_api_signatures.json')))
# the data-generator / TF-runtime helpers whose role is taken over by other code, with the reason
EXEMPT = {}


def _sig(mod, fn):
    return inspect.signature(getattr(importlib.import_module(mod), fn))


@pytest.mark.parametrize('mod', sorted(TABLE['functions']))
def test_public_functions_have_the_reference_signature(mod):
    m = importlib.import_module(mod)
    for fn, ref_args in sorted(TABLE['functions'][mod].items()):
        if (mod, fn) in EXEMPT:
            continue
        assert hasattr(m, fn), '%s.%s is missing' % (mod, fn)
        params = list(_sig(mod, fn).parameters.values())
        public = [p for p in params if not p.name.startswith('_')]
        assert [p.name for p in public] == [a['name'] for a in ref_args], (mod, fn, [p.name for p in public])
        assert all(p.name.startswith('_') for p in params[len(public):]), (mod, fn)       # product-only knobs come last
        for p, a in zip(public, ref_args):
            d = a['default']
            if d is None:
                assert p.default is inspect.Parameter.empty, (mod, fn, p.name)
            elif 'value' in d:
                assert p.default == d['value'] and type(p.default) == type(d['value']), (mod, fn, p.name, p.default, d['value'])
        for p in params[len(public):]:
            assert p.default is not inspect.Parameter.empty, (mod, fn, p.name)             # ... and are optional


@pytest.mark.parametrize('script', sorted(TABLE['call_sites']))
def test_script_call_sites_bind(script):
    """The model / geometry / loss call sites of the scripts (ldi_enc_dec.py:175-228,265-410; ldi_pred_eval.py:117-224,297-548)."""
    sites = TABLE['call_sites'][script]
    assert len(sites) >= 10
    for s in sites:
        sig = _sig(s['module'], s['function'])
        sig.bind(*([None] * s['n_positional']), **{k: None for k in s['keywords']})       # raises TypeError on a mismatch


def test_committed_table_is_current(tmp_path, monkeypatch):
    from oracle import gen_api_signatures as G
    if not os.path.isdir(G.REF):
        pytest.skip('reference tree not mounted')
    monkeypatch.setattr(G, 'OUT', str(tmp_path / 'sig.json'))
    G.main()
    assert json.load(open(str(tmp_path / 'sig.json'))) == TABLE
