"""tf.load_op_library loader + registered gradient for the shim (SOURCE ONLY: never imported in this repository's tests;
TensorFlow is not installable in the build image).  See INTEGRATION.md section 2."""
import os

import tensorflow as tf  # noqa: F401  (raises ImportError here -- by design there is no fallback)

_ops = tf.load_op_library(os.path.join(os.path.dirname(__file__), "libkgeb_tf_ops.so"))
kgeb_score = _ops.kgeb_score


@tf.RegisterGradient("KgebScore")
def _kgeb_score_grad(op, dscore):
    g_e, g_r = _ops.kgeb_score_grad(op.inputs[0], op.inputs[1], op.inputs[2], dscore,
                                    score_kind=op.get_attr("score_kind"), gamma=op.get_attr("gamma"))
    return g_e, g_r, None
