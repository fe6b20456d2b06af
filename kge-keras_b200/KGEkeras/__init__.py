# -*- coding: utf-8 -*-
"""B200-native drop-in for the triple-scoring hot path of KGE-Keras: the same class / function names and signatures as
the reference's KGEkeras/models.py and KGEkeras/utils.py; the arithmetic runs in libkgeb200's sm_100a kernels
(kge-keras_b200/csrc, C ABI in include/kgeb200.h).  Importing the package needs neither CUDA nor TensorFlow; using a
model does need a GPU (there is no CPU fallback)."""
__version__ = '0.1.0'
from .models import *  # noqa: F401,F403
from .utils import *   # noqa: F401,F403
from .data import DataGenerator, build_mappings, load_dataset, map_triples  # noqa: F401
