"""gtsam.Symbol / gtsam.symbol_shorthand key codec (batch.py:26).

key = (uint8 chr << 56) | index  -- bit-exact with gtsam/inference/Symbol.h, so Values
iterate b < l < v < x exactly like gtsam's std::map<Key, Value>.
"""
import numpy as np

_INDEX_BITS = 56
_INDEX_MASK = (1 << _INDEX_BITS) - 1


def symbol(c, j):
    c = ord(c) if isinstance(c, str) else int(c)
    j = int(j)
    if not (0 <= j <= _INDEX_MASK):
        raise ValueError("Symbol index out of range")
    return (c << _INDEX_BITS) | j


def symbolChr(key):
    return chr((int(key) >> _INDEX_BITS) & 0xFF)


def symbolIndex(key):
    return int(key) & _INDEX_MASK


def symbols(c, idx):
    """Vectorised: uint64 keys for an integer array (bulk-construction API)."""
    c = ord(c) if isinstance(c, str) else int(c)
    return (np.uint64(c) << np.uint64(_INDEX_BITS)) | np.asarray(idx, dtype=np.uint64)


class Symbol:
    def __init__(self, c, j=None):
        if j is None:
            self._key = int(c)
        else:
            self._key = symbol(c, j)

    def key(self):
        return self._key

    def chr(self):
        return ord(symbolChr(self._key))

    def index(self):
        return symbolIndex(self._key)

    def string(self):
        return f"{symbolChr(self._key)}{symbolIndex(self._key)}"

    __repr__ = string


class _Shorthand:
    """from gtsam.symbol_shorthand import B, V, X, L"""

    def __getattr__(self, name):
        if len(name) == 1 and name.isalpha():
            ch = name.lower()
            return lambda j, _c=ch: symbol(_c, j)
        raise AttributeError(name)


symbol_shorthand = _Shorthand()
B = symbol_shorthand.B
V = symbol_shorthand.V
X = symbol_shorthand.X
L = symbol_shorthand.L
