"""TEST INFRASTRUCTURE -- recipe that places an UNMODIFIED copy of the reference package under baseline/_ref/
(git-ignored, but shipped to the GPU box by gpurun) so that the reference itself -- not a port -- can be timed and
run there:  bench.py's cfg-1 / reference legs and tests/test_reference_alias.py import it from that directory on top
of oracle/fdm_shim.py (the reference's `fdm` dependency is absent everywhere).

The reference has no setup.py / pyproject.toml, so `pip install --target baseline/_ref /root/reference` cannot work
(DESIGN.md section 8); this copy is what such an install would have produced.  Nothing is copied into tracked paths.

    python -m oracle.make_ref          # needs /root/reference (build container only)
"""
import hashlib
import os
import shutil
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = '/root/reference'
DST = os.path.join(REPO, 'baseline', '_ref')


def available():
    return os.path.isfile(os.path.join(DST, 'fractulus', 'equation.py'))


def make(force=False):
    if not os.path.isdir(os.path.join(SRC, 'fractulus')):
        return available()
    if available() and not force:
        same = open(os.path.join(SRC, 'fractulus', 'equation.py'), 'rb').read() == \
            open(os.path.join(DST, 'fractulus', 'equation.py'), 'rb').read()
        if same:
            return True
    if os.path.isdir(DST):
        shutil.rmtree(DST)
    shutil.copytree(os.path.join(SRC, 'fractulus'), os.path.join(DST, 'fractulus'),
                    ignore=shutil.ignore_patterns('__pycache__', '*.pyc'))
    for extra in ('LICENSE', 'README.md'):
        if os.path.isfile(os.path.join(SRC, extra)):
            shutil.copy(os.path.join(SRC, extra), os.path.join(DST, extra))
    lines = []
    for root, _, files in sorted(os.walk(DST)):
        for name in sorted(files):
            path = os.path.join(root, name)
            lines.append('%s  %s' % (hashlib.sha256(open(path, 'rb').read()).hexdigest(), os.path.relpath(path, DST)))
    with open(os.path.join(DST, 'PROVENANCE.txt'), 'w') as f:
        f.write('unmodified copy of %s made by oracle/make_ref.py (sha256 of every file below)\n' % SRC)
        f.write('\n'.join(lines) + '\n')
    return True


def load(shim_as_fdm=True):
    """Import the unmodified reference from baseline/_ref on top of oracle/fdm_shim.py -> (fractulus, fdm)."""
    if not available():
        raise ImportError('baseline/_ref/fractulus is missing: run `python -m oracle.make_ref` where /root/reference exists')
    from oracle import fdm_shim
    fdm = fdm_shim.install_as_fdm() if shim_as_fdm else sys.modules['fdm']
    if DST not in sys.path:
        sys.path.insert(0, DST)
    import fractulus
    assert os.path.realpath(fractulus.__file__).startswith(os.path.realpath(DST)), fractulus.__file__
    return fractulus, fdm


if __name__ == '__main__':
    ok = make(force='--force' in sys.argv)
    print('baseline/_ref:', 'ready' if ok else 'NOT available (no /root/reference here)')
