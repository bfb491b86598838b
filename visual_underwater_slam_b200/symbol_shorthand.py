"""gtsam.symbol_shorthand as a module, so that batch.py:26 (`from gtsam.symbol_shorthand import B, V, X, L`) works with the
package swapped in: any single letter is a key constructor, as in gtsam (`X(j)` = Symbol('x', j).key())."""
import string as _string
from .symbol import symbol as _symbol


def _make(ch):
    def f(j):
        return _symbol(ch, j)
    f.__name__ = ch.upper()
    return f


for _c in _string.ascii_uppercase:
    globals()[_c] = _make(_c.lower())
__all__ = list(_string.ascii_uppercase)
