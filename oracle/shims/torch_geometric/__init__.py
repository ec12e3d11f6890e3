"""TEST-ONLY shim of the torch-geometric==2.0.4 leaves the reference touches."""
import random
import numpy as np
import torch


def seed_everything(seed):
    random.seed(seed)
    np.random.seed(seed)
    torch.manual_seed(seed)
