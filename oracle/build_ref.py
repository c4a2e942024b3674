"""
ORACLE (test infrastructure) -- builds ``oracle/_ref``: the reference's OWN hot-path modules,
made importable under Python 3.

The reference (``/root/reference``) is Python 2 and imports an un-vendored ``utils`` package and
``shapely``, so it cannot be imported as it lies.  This script reads the few
source files of the hot path *where they lie*, applies a token-level Python 2 -> 3 conversion
(``lib2to3`` is not part of this image's interpreter, so the handful of fixers the files need are
restated here), and writes the result only into ``oracle/_ref/`` (git-ignored: derived reference
code never enters the history; not gpurun-ignored, so the built tree travels to the GPU box like a
compiled ``.so`` would).  Next to the converted modules it writes *stub* modules for what the
reference imports but the hot path never executes:

  utils.misc.display_progress            -> identity              (video/io/base.py:16, analysis/video.py:11)
  utils.math.get_number_range            -> np.iinfo / finfo range (video/filters.py:27,117)
  utils.math.safe_typecast, contiguous_true_regions -> minimal     (io/computed.py:12, io/composer.py:13)
  utils.data_structures.cache.cached_property -> functools-based  (analysis/image.py:19)
  shapely.geometry / shapely.geos        -> empty namespaces       (analysis/regions.py:15; only if shapely is absent)
  external.simplify_polygon_visvalingam  -> empty namespace        (analysis/regions.py:18)
  video.debug, video.analysis.curves     -> empty namespaces       (analysis/regions.py:17,20)

Converted verbatim (apart from the fixers): video/filters.py, video/io/base.py, video/io/memory.py,
video/analysis/regions.py, video/analysis/video.py, video/analysis/image.py.

Fixers (each is one token-level rule; nothing else in the files is touched):
  * ``xrange`` -> ``range``; ``itertools.izip`` -> ``zip``; ``np.int`` -> ``int`` (removed in NumPy 2)
  * ``import curves`` (implicit relative import) -> ``from . import curves``
  * classes that define ``next`` get ``__next__ = next`` (appended module footer)
  * ``ndimage.measurements.label`` is left alone: SciPy 1.18 still resolves it

Usage:  python oracle/build_ref.py [--src /root/reference] [--force]
"""

import io
import os
import sys
import tokenize

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, '_ref')
DEFAULT_SRC = '/root/reference'

CONVERT = ['video/filters.py', 'video/io/base.py', 'video/io/memory.py',
           'video/analysis/regions.py', 'video/analysis/video.py', 'video/analysis/image.py']

FOOTER = '''

# --- appended by oracle/build_ref.py: Python 3 iterator protocol for classes defining next() ---
for _c in list(globals().values()):
    if isinstance(_c, type) and 'next' in vars(_c) and '__next__' not in vars(_c):
        _c.__next__ = _c.next
'''

STUBS = {
    'video/__init__.py': '# package marker written by oracle/build_ref.py\n',
    'video/io/__init__.py': 'from .base import VideoFork\nfrom .memory import VideoMemory\n',
    'video/analysis/__init__.py': '',
    'video/debug.py': '# stub: matplotlib debugging helpers are never executed on the hot path\n',
    'video/analysis/curves.py': '# stub: polyline helpers are never executed on the hot path\n',
    'external/__init__.py': '',
    'external/simplify_polygon_visvalingam.py': '# stub\n',
    'utils/__init__.py': '',
    'utils/misc.py': (
        'def display_progress(iterator, total=None):\n'
        '    """ stand-in for the un-vendored utils.misc.display_progress: no progress bar """\n'
        '    return iterator\n'),
    'utils/math.py': (
        'import numpy as np\n\n'
        'def get_number_range(dtype):\n'
        '    """ (min, max) of a numpy dtype """\n'
        '    info = np.iinfo(dtype) if np.issubdtype(dtype, np.integer) else np.finfo(dtype)\n'
        '    return info.min, info.max\n\n'
        'def safe_typecast(data, dtype):\n'
        '    lo, hi = get_number_range(dtype)\n'
        '    return np.clip(data, lo, hi).astype(dtype)\n\n'
        'def contiguous_true_regions(condition):\n'
        '    d = np.diff(np.r_[0, np.asarray(condition, np.int8), 0])\n'
        '    return np.c_[np.nonzero(d == 1)[0], np.nonzero(d == -1)[0]]\n'),
    'utils/data_structures/__init__.py': '',
    'utils/data_structures/cache.py': (
        'import functools\n\n'
        'class cached_property(object):\n'
        '    """ stand-in for utils.data_structures.cache.cached_property (used as @cached_property()) """\n'
        '    def __init__(self, *args, **kwargs):\n'
        '        pass\n'
        '    def __call__(self, method):\n'
        '        name = "_cache_" + method.__name__\n'
        '        @functools.wraps(method)\n'
        '        def getter(obj):\n'
        '            if not hasattr(obj, name):\n'
        '                setattr(obj, name, method(obj))\n'
        '            return getattr(obj, name)\n'
        '        return property(getter)\n'),
}

SHAPELY_STUBS = {
    'shapely/__init__.py': '# stub written by oracle/build_ref.py (shapely is not installed)\n',
    'shapely/geometry.py': '',
    'shapely/geos.py': '',
}


def convert(source):
    """ token-level Python 2 -> 3 conversion of one module's text """
    toks = list(tokenize.generate_tokens(io.StringIO(source).readline))
    out = []
    i = 0
    while i < len(toks):
        t = toks[i]
        if t.type == tokenize.NAME and t.string == 'xrange':
            t = t._replace(string='range')
        elif (t.type == tokenize.NAME and t.string in ('np', 'itertools') and i + 2 < len(toks)
              and toks[i + 1].string == '.' and toks[i + 2].type == tokenize.NAME):
            nxt = toks[i + 2].string
            rep = None
            if t.string == 'np' and nxt == 'int':
                rep = 'int'
            elif t.string == 'itertools' and nxt == 'izip':
                rep = 'zip'
            if rep is not None:
                # keep the column span so untokenize's spacing stays valid
                out.append(toks[i + 2]._replace(string=rep, start=t.start))
                i += 3
                continue
        out.append(t)
        i += 1
    text = tokenize.untokenize(out)
    # implicit relative imports (whole-line statements only)
    lines = []
    for line in text.split('\n'):
        s = line.strip()
        if s in ('import curves', 'import shapes', 'import image'):
            line = line.replace('import ', 'from . import ')
        lines.append(line)
    text = '\n'.join(lines)
    if 'def next(self)' in text:
        text += FOOTER
    return text


def build(src=DEFAULT_SRC, force=False):
    """ writes oracle/_ref; returns its path, or None when the reference sources are absent """
    if not os.path.isdir(os.path.join(src, 'video')):
        return OUT if os.path.isdir(os.path.join(OUT, 'video')) else None
    stamp = os.path.join(OUT, '.built')
    newest = max(os.path.getmtime(os.path.join(src, f)) for f in CONVERT)
    newest = max(newest, os.path.getmtime(os.path.abspath(__file__)))
    if not force and os.path.exists(stamp) and os.path.getmtime(stamp) >= newest:
        return OUT
    files = dict(STUBS)
    try:
        import shapely  # noqa: F401
        if os.path.abspath(os.path.dirname(shapely.__file__)).startswith(OUT):
            raise ImportError
    except ImportError:
        files.update(SHAPELY_STUBS)
    for rel in CONVERT:
        with open(os.path.join(src, rel), 'r') as fh:
            files[rel] = convert(fh.read())
    for rel, text in files.items():
        path = os.path.join(OUT, rel)
        os.makedirs(os.path.dirname(path), exist_ok=True)
        with open(path, 'w') as fh:
            fh.write(text)
    with open(stamp, 'w') as fh:
        fh.write('converted from %s\n' % src)
    return OUT


def import_ref(src=DEFAULT_SRC):
    """ builds (when the sources are here) and imports the converted reference.
    Returns a namespace with the modules, or None when neither sources nor a built tree exist. """
    path = build(src)
    if path is None:
        return None
    import importlib
    import types
    import warnings
    if path not in sys.path:
        sys.path.insert(0, path)
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        ns = types.SimpleNamespace(
            path=path,
            filters=importlib.import_module('video.filters'),
            base=importlib.import_module('video.io.base'),
            memory=importlib.import_module('video.io.memory'),
            regions=importlib.import_module('video.analysis.regions'),
            video=importlib.import_module('video.analysis.video'),
            image=importlib.import_module('video.analysis.image'))
    return ns


if __name__ == '__main__':
    src = DEFAULT_SRC
    if '--src' in sys.argv:
        src = sys.argv[sys.argv.index('--src') + 1]
    p = build(src, force='--force' in sys.argv)
    print('oracle/_ref:', p)
    if p is not None:
        ref = import_ref(src)
        print('imported:', ', '.join(sorted(k for k in vars(ref) if k != 'path')))
