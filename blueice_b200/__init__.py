"""blueice_b200: the likelihood-evaluation hot path of blueice, rebuilt B200-native.

Same top-level names as the reference package (blueice/__init__.py:1-11).
"""
from .exceptions import *          # noqa: F401,F403
from .likelihood import *          # noqa: F401,F403
from .model import *               # noqa: F401,F403
from .source import *              # noqa: F401,F403

__version__ = '0.1.0'
__reference_version__ = '1.2.1'
