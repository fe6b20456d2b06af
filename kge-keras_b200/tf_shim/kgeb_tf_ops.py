"""tf.load_op_library loader + registered gradients for the shim (SOURCE ONLY: never imported in this repository's tests;
TensorFlow is not installable in the build image).  See INTEGRATION.md section 2.

    scores = kgeb_score(E, R, idx, score_kind=..., gamma=...)           # differentiable: the gradient below is registered
    loss   = kgeb_train_step(E, R, mE, vE, mR, vR, pos, step, ...)      # one fused step in place on the variables' buffers
    counts = kgeb_rank1n(E, R, triples, offsets, ids, side=0|1, ...)    # rank = 1 + counts
"""
import os

import tensorflow as tf  # noqa: F401  (raises ImportError here -- by design there is no fallback)

_ops = tf.load_op_library(os.path.join(os.path.dirname(__file__), "libkgeb_tf_ops.so"))
kgeb_score = _ops.kgeb_score
kgeb_train_step = _ops.kgeb_train_step
kgeb_rank1n = _ops.kgeb_rank1n


@tf.RegisterGradient("KgebScore")
def _kgeb_score_grad(op, dscore):
    """Entity gradient as tf.IndexedSlices (what a tf.keras Embedding hands its optimizer, models.py:89-112: only the touched rows
    reach the optimizer), relation gradient dense (a few hundred rows)."""
    ids, rows, n, g_r = _ops.kgeb_score_grad_slices(op.inputs[0], op.inputs[1], op.inputs[2], dscore,
                                                   score_kind=op.get_attr("score_kind"), gamma=op.get_attr("gamma"))
    n = tf.cast(n, tf.int32)
    g_e = tf.IndexedSlices(values=rows[:n], indices=ids[:n], dense_shape=tf.shape(op.inputs[0]))
    return g_e, g_r, None
