"""TEST-ONLY stand-in for `atomicwrites.atomic_write` (not installed here).
Used only so the unmodified reference can be imported by tests/golden/make_golden.py."""
import contextlib
import os
import tempfile


@contextlib.contextmanager
def atomic_write(path, mode="w", overwrite=False, **kwargs):
    d = os.path.dirname(os.path.abspath(path))
    fd, tmp = tempfile.mkstemp(dir=d)
    os.close(fd)
    try:
        with open(tmp, mode) as f:
            yield f
        if overwrite:
            os.replace(tmp, path)
        else:
            os.link(tmp, path)
            os.unlink(tmp)
    except BaseException:
        if os.path.exists(tmp):
            os.unlink(tmp)
        raise
