"""TEST INFRASTRUCTURE -- extracts, from the reference tree (read-only, /root/reference), the public API surface that the
drop-in `lsi` package must reproduce: for every function of lsi/geometry/{ldi,sampling,projection}.py, lsi/nnutils/{helpers,nets}.py
and lsi/loss/loss.py its argument names (in order) and literal defaults; and for the two scripts (ldi_enc_dec.py,
ldi_pred_eval.py) every call they make into those modules with the keyword names they pass.  Only names / literals are
recorded (no source text).  Output: tests/golden/ref_api_signatures.json, consumed by tests/test_dropin_api.py.

    python oracle/gen_api_signatures.py
"""
import ast
import json
import os

REF = '/root/reference'
HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(os.path.dirname(HERE), 'tests', 'golden', 'ref_api_signatures.json')
MODULES = {'lsi.geometry.ldi': 'lsi/geometry/ldi.py', 'lsi.geometry.sampling': 'lsi/geometry/sampling.py',
           'lsi.geometry.projection': 'lsi/geometry/projection.py', 'lsi.nnutils.helpers': 'lsi/nnutils/helpers.py',
           'lsi.nnutils.nets': 'lsi/nnutils/nets.py', 'lsi.loss.loss': 'lsi/loss/loss.py',
           'lsi.geometry.homography': 'lsi/geometry/homography.py', 'lsi.geometry.layers': 'lsi/geometry/layers.py'}
SCRIPTS = ['ldi_enc_dec.py', 'ldi_pred_eval.py']


def _literal(node):
    try:
        return {'value': ast.literal_eval(node)}
    except Exception:
        return {'expr': True}


def functions(path):
    tree = ast.parse(open(path).read().replace('\t', '    '))
    out = {}
    for node in tree.body:
        if isinstance(node, ast.FunctionDef) and not node.name.startswith('_'):
            args = [a.arg for a in node.args.args]
            defaults = [None] * (len(args) - len(node.args.defaults)) + [_literal(d) for d in node.args.defaults]
            out[node.name] = [{'name': a, 'default': d} for a, d in zip(args, defaults)]
    return out


def call_sites(path, aliases):
    """calls of the form <alias>.<fn>(...) where <alias> is the local name of one of MODULES."""
    tree = ast.parse(open(path).read())
    sites = []
    for node in ast.walk(tree):
        if isinstance(node, ast.Call) and isinstance(node.func, ast.Attribute) and isinstance(node.func.value, ast.Name) \
                and node.func.value.id in aliases:
            sites.append({'module': aliases[node.func.value.id], 'function': node.func.attr, 'n_positional': len(node.args),
                          'keywords': sorted(k.arg for k in node.keywords if k.arg), 'line': node.lineno})
    return sites


def import_aliases(path):
    tree = ast.parse(open(path).read())
    al = {}
    for node in ast.walk(tree):
        if isinstance(node, ast.ImportFrom) and node.module:
            for n in node.names:
                full = node.module + '.' + n.name
                if full in MODULES:
                    al[n.asname or n.name] = full
    return al


def main():
    api = {m: functions(os.path.join(REF, p)) for m, p in MODULES.items()}
    calls = {}
    for s in SCRIPTS:
        p = os.path.join(REF, s)
        calls[s] = call_sites(p, import_aliases(p))
    with open(OUT, 'w') as f:
        json.dump({'functions': api, 'call_sites': calls}, f, indent=1, sort_keys=True)
    print('wrote %s: %d functions, %d call sites' % (OUT, sum(len(v) for v in api.values()), sum(len(v) for v in calls.values())))


if __name__ == '__main__':
    main()
