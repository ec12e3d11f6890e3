"""TEST-ONLY: importable names only (nn/conv.py:18-19, nn/pool.py:7)."""


def dense_diff_pool(*a, **k):
    raise NotImplementedError


def global_sort_pool(*a, **k):
    raise NotImplementedError
