"""blueice_b200: the likelihood-evaluation hot path of blueice, rebuilt B200-native.

Same top-level names as the reference package (blueice/__init__.py:1-11).
"""
from .exceptions import *          # noqa: F401,F403
from .likelihood import *          # noqa: F401,F403
from .model import *               # noqa: F401,F403
from .source import *              # noqa: F401,F403

__version__ = '0.1.0'
__reference_version__ = '1.2.1'


_ALIASED = ('exceptions', 'likelihood', 'model', 'source', 'pdf_morphers', 'inference', 'utils', 'data_reading',
            'test_helpers')


def install_as_blueice(force=False):
    """Make `import blueice` (and `blueice.likelihood`, `.source`, `.model`, `.inference`, ...) resolve to this package,
    so that analysis code written against the reference (blueice/__init__.py:1-11 and its submodules) runs on the B200
    path without editing its imports.  Call it before the first `import blueice`.

    Raises ImportError if a different `blueice` is already imported (pass force=True to replace it).  The reference's
    `blueice.parallel` (process-pool model construction, out of scope) is not aliased and raises ImportError on import.
    """
    import importlib
    import sys
    me = sys.modules[__name__]
    have = sys.modules.get('blueice')
    if have is not None and have is not me and not force:
        raise ImportError("a different 'blueice' package is already imported (%r); call install_as_blueice() first "
                          "or pass force=True" % getattr(have, '__file__', have))
    if force:
        for name in [n for n in sys.modules if n == 'blueice' or n.startswith('blueice.')]:
            del sys.modules[name]
    sys.modules['blueice'] = me
    for sub in _ALIASED:
        sys.modules['blueice.' + sub] = importlib.import_module(__name__ + '.' + sub)
    return me
