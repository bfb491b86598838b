"""gtsam.Values facade (batch.py:81, :271, :274, :283-288, :297-298; read-out :57-68).

Internally four typed tables (Pose3 as 12 doubles R|t, velocity 3, bias 6, landmark 3) that the
packer sorts by uint64 key -- the same ascending-key order gtsam's std::map iterates in
(b < l < v < x) -- and transposes into the structure-of-arrays device buffers.
"""
import numpy as np
from .geometry import Pose3
from .navigation import ConstantBias
from .symbol import symbolChr

KINDS = ("pose", "vel", "bias", "lm")
_DIM = {"pose": 12, "vel": 3, "bias": 6, "lm": 3}


class Values:
    def __init__(self, other=None):
        self._rows = {k: [] for k in KINDS}      # list of np rows / blocks
        self._keys = {k: [] for k in KINDS}      # list of python ints / arrays
        self._index = {}                          # key -> kind
        self._cache = None
        if other is not None:
            for k in KINDS:
                keys, data = other.table(k)
                self.insert_bulk(k, keys, data)

    # ------------------------------------------------------------------ gtsam API
    def insert(self, key, value=None):
        if isinstance(key, Values):                           # Values.insert(Values)
            for k in KINDS:
                keys, data = key.table(k)
                self.insert_bulk(k, keys, data)
            return
        key = int(key)
        if key in self._index:
            raise RuntimeError(f"Attempting to add a key-value pair with key \"{symbolChr(key)}"
                               f"{key & ((1 << 56) - 1)}\", key already exists.")
        if isinstance(value, Pose3):
            kind, row = "pose", value.as_row()
        elif isinstance(value, ConstantBias):
            kind, row = "bias", value.vector()
        else:
            row = np.asarray(value, dtype=np.float64).reshape(-1)
            if row.size != 3:
                raise NotImplementedError("Values.insert: only Pose3, ConstantBias and 3-vectors are supported")
            kind = "lm" if symbolChr(key) == 'l' else "vel"
        self._rows[kind].append(row[None, :])
        self._keys[kind].append(np.array([key], dtype=np.uint64))
        self._index[key] = kind
        self._cache = None

    def insert_bulk(self, kind, keys, data):
        """Bulk-construction API: keys uint64[n], data [n, 12|3|6|3]."""
        keys = np.asarray(keys, dtype=np.uint64).reshape(-1)
        data = np.asarray(data, dtype=np.float64).reshape(len(keys), _DIM[kind])
        if len(keys) == 0:
            return
        kl = keys.tolist()
        if len(set(kl)) != len(kl) or any(k in self._index for k in kl):
            raise RuntimeError("Values.insert_bulk: key already exists")
        self._rows[kind].append(data.copy())
        self._keys[kind].append(keys.copy())
        self._index.update(dict.fromkeys(kl, kind))
        self._cache = None

    def clear(self):
        self.__init__()

    def exists(self, key):
        return int(key) in self._index

    def size(self):
        return len(self._index)

    def keys(self):
        return sorted(self._index.keys())

    def _lookup(self, key, kind, what):
        key = int(key)
        k = self._index.get(key)
        if k is None:
            raise RuntimeError(f"Attempting to retrieve the key \"{symbolChr(key)}{key & ((1 << 56) - 1)}\", "
                               "which does not exist in the Values.")
        if k not in kind:
            raise RuntimeError(f"Values.{what}: key holds a different type")
        keys, data = self.table(k)
        pos = int(np.searchsorted(keys, np.uint64(key)))
        return k, data[pos]

    def atPose3(self, key):
        return Pose3.from_row(self._lookup(key, ("pose",), "atPose3")[1])

    def atVector(self, key):
        return self._lookup(key, ("vel", "lm", "bias"), "atVector")[1].copy()

    def atPoint3(self, key):
        return self._lookup(key, ("lm", "vel"), "atPoint3")[1].copy()

    def atConstantBias(self, key):
        row = self._lookup(key, ("bias",), "atConstantBias")[1]
        return ConstantBias(row[:3], row[3:])

    # ------------------------------------------------------------------ packer side
    def table(self, kind):
        """(keys uint64[n] ascending, data [n,d]) for one kind."""
        if self._cache is None:
            self._cache = {}
        if kind not in self._cache:
            if self._keys[kind]:
                keys = np.concatenate(self._keys[kind])
                data = np.concatenate(self._rows[kind], axis=0)
                order = np.argsort(keys, kind="stable")
                self._cache[kind] = (keys[order], data[order])
            else:
                self._cache[kind] = (np.zeros(0, dtype=np.uint64), np.zeros((0, _DIM[kind])))
        return self._cache[kind]

    @staticmethod
    def from_tables(tables):
        v = Values()
        for kind in KINDS:
            keys, data = tables[kind]
            v.insert_bulk(kind, keys, data)
        return v
