"""TEST-ONLY shim: torch-sparse is used by the reference only for an isinstance check (nn/conv.py:152)."""


class SparseTensor:  # pragma: no cover
    pass
