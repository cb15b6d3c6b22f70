"""Stand-in for the `intervaltree` package (absent from this image), used ONLY by
tests/golden/make_*.py and bench.py (python_reference baseline) to import the unmodified reference.

Implements the two calls the reference makes: `IntervalTree.addi(b, e, data)`
(utils.py:100) and `IntervalTree.overlap(b, e)` (read_operations.py:63) with
the package's half-open semantics `iv.begin < e and iv.end > b`.  The real
package returns a *set* (unspecified order); this stand-in returns hits sorted
by (begin, end), which is the canonical order the oracle and the CUDA path use.
"""
from collections import namedtuple

Interval = namedtuple("Interval", "begin end data")


class IntervalTree:
    def __init__(self):
        self._ivs = []

    def addi(self, begin, end, data=None):
        self._ivs.append(Interval(begin, end, data))

    def overlap(self, begin, end):
        return sorted((iv for iv in self._ivs if iv.begin < end and iv.end > begin),
                      key=lambda iv: (iv.begin, iv.end))

    def __len__(self):
        return len(self._ivs)

    def __iter__(self):
        return iter(self._ivs)
