# -*- coding: utf-8 -*-
"""B200-native drop-in for the triple-scoring hot path of KGE-Keras (same class / function names and signatures as
KGEkeras/models.py and KGEkeras/utils.py of the reference; the arithmetic runs in libkgeb200's sm_100a kernels)."""
__version__ = '0.1.0'
