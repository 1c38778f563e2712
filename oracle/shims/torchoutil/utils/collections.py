from typing import Iterable


def all_eq(it: Iterable) -> bool:
    it = list(it)
    return all(x == it[0] for x in it[1:])
