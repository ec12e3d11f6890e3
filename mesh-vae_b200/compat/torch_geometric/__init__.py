import random
import numpy as np
import torch


def seed_everything(seed):          # main.py:210
    random.seed(seed)
    np.random.seed(seed)
    torch.manual_seed(seed)
    torch.cuda.manual_seed_all(seed)
