"""oracle/ -- TEST INFRASTRUCTURE ONLY.

CPU restatement (NumPy, fp32 and fp64) of the KGE-Keras triple-scoring hot path:
score functions, losses, negative sampling, optimizer step and filtered ranking.

Nothing under this directory is product code.  Only `tests/`, `__graft_entry__.smoke()`
and the `cpu_baseline` / `--impl reference` legs of `bench.py` may import it, and only as the
checker (or as the timed CPU baseline), never as the thing shipped.  The product path
(`kge-keras_b200/`) never imports `oracle` and fails loudly without its CUDA library.

Parity pinning status: the reference (`/root/reference`, TensorFlow/Keras) cannot be imported
as it stands in this image (no tensorflow, keras, rdflib, spacy), and its own `tests.py` asserts
nothing.  Two mechanisms run the reference's OWN code here and freeze its outputs as fixtures:
  * `oracle/ref_ast.py` executes the NumPy/SciPy function bodies of utils.py (AST-extracted,
    unmodified): `generate_negative`, `oversample_data`, `gen_tail_data`, `gen_head_data`,
    `validate`, `mrr`, `hits`, `pad` -> `tests/golden/make_golden.py`, `make_golden_full.py`;
  * `oracle/ref_tf_exec.py` loads models.py and utils.py unmodified over a torch-backed
    `tensorflow` stand-in: score functions, losses, regulariser, literal branch and their autodiff
    gradients, in fp32 and fp64 -> `tests/golden/make_golden_fp.py` -> `tests/golden/fp_ref.npz`.
The restatement in `kge_oracle.py` is held to both fixture sets (tests/test_cpu_*goldens*.py).
Not pinned by reference code, because the reference only calls them: the Keras optimizers,
`binary_crossentropy`, `GlorotUniform`, `Dense` (restated from their published definitions) and
the rounding of TensorFlow's own kernels (the bars are stated against fp64 runs of the source).
"""
