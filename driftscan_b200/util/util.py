"""File-name patterns and a one-slot memoiser (mirrors drift/util/util.py:6-32).

The zero-padded ``m`` directory names are part of the on-disk product layout
(``<bt>/beam_m/<m>/beam.hdf5``), so the patterns must agree digit for digit.
"""

import functools
import math


def _ndigits(n):
    return int(math.ceil(math.log10(n + 1)))


def intpattern(n):
    """printf pattern for a signed integer of magnitude up to ``n``."""
    return "%+0" + repr(_ndigits(n) + 1) + "d"


def natpattern(n):
    """printf pattern for a natural number up to ``n`` (zero padded)."""
    return "%0" + repr(_ndigits(n)) + "d"


def cache_last(func):
    """Remember the result of the most recent call (same object is returned for a
    repeated identical call; callers must not mutate it)."""
    last = {}

    @functools.wraps(func)
    def wrapper(*args, **kwargs):
        key = (args, tuple(sorted(kwargs.items())))
        if last.get("key", None) != key or "val" not in last:
            last["val"] = func(*args, **kwargs)
            last["key"] = key
        return last["val"]

    return wrapper
