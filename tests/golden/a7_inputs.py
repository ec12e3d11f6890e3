"""Seeded inputs of the A7 pin, shared by the golden generator (make_golden_a7.py) and the tests."""
import torch

# (name, level, Fin, Fout, K, bias)
CASES = [("l1_6_16", 1, 6, 16, 6, True), ("l2_16_16", 2, 16, 16, 6, True), ("l3_16_32", 3, 16, 32, 6, True),
         ("l3_16_16_K3", 3, 16, 16, 3, False)]


def case_inputs(ci, n, fin, fout, K, nb=2):
    g = torch.Generator().manual_seed(700 + ci)
    x = torch.randn(nb, n, fin, generator=g)
    w = torch.randn(K, fin, fout, generator=g) * 0.1
    b = torch.randn(fout, generator=g) * 0.1
    dy = torch.randn(nb, n, fout, generator=g)
    return x, w, b, dy
