"""Per-step diagnostics table -- same public surface as reference chsimpy/timedata.py
(`insert`, `data()`, column properties, `energy_falls`), but backed by a growable buffer
(the reference re-allocates the whole table with np.append on every step)."""
import numpy as np

COLUMNS = ("it_range", "E", "E2", "SA", "domtime", "Ra", "L2", "PS", "delt")   # timedata.py:8-9 order


class TimeData:
    def __init__(self, capacity=1024):
        self._buf = np.empty((max(int(capacity), 1), 9))
        self._n = 0

    # -- writers -------------------------------------------------------------------------
    def _reserve(self, extra):
        need = self._n + extra
        if need > self._buf.shape[0]:
            grown = np.empty((max(need, 2 * self._buf.shape[0]), 9))
            grown[:self._n] = self._buf[:self._n]
            self._buf = grown

    def insert(self, it, delt, E, E2, SA, domtime, Ra, L2, PS):
        """One row (keyword order of reference timedata.py:8); NaN is fatal as there (:10)."""
        self.extend(np.array([[it, E, E2, SA, domtime, Ra, L2, PS, delt]], dtype=np.float64))

    def extend(self, rows):
        """Appends device-produced rows (n x 9).  Raises AssertionError at the first row that
        holds a NaN, after appending it -- the reference asserts right after its np.append."""
        rows = np.asarray(rows, dtype=np.float64).reshape(-1, 9)
        bad = np.flatnonzero(np.isnan(rows).any(axis=1))
        if bad.size:
            rows = rows[:bad[0] + 1]
        self._reserve(rows.shape[0])
        self._buf[self._n:self._n + rows.shape[0]] = rows
        self._n += rows.shape[0]
        assert not bad.size, "NaN in TimeData row %d" % (self._n - 1)

    # -- readers -------------------------------------------------------------------------
    @property
    def _data(self):
        return self._buf[:self._n]

    def data(self):
        return self._buf[:self._n]

    def __len__(self):
        return self._n

    def energy_falls(self, it=None):
        """E2[it-1] > E2[it] > E2[0]  (reference timedata.py:63)."""
        e2 = self.E2
        return bool(e2[it - 1] > e2[it] > e2[0])


def _col(i):
    return property(lambda self: self._buf[:self._n, i])


for _i, _name in enumerate(COLUMNS):
    setattr(TimeData, _name, _col(_i))
