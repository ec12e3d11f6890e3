class SparseTensor:  # isinstance check only (nn/conv.py:152)
    pass
