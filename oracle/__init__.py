"""oracle/ -- TEST INFRASTRUCTURE ONLY.

CPU restatement (NumPy, fp32 and fp64) of the KGE-Keras triple-scoring hot path:
score functions, losses, negative sampling, optimizer step and filtered ranking.

Nothing under this directory is product code.  Only `tests/`, `__graft_entry__.smoke()`
and the `cpu_baseline` / `--impl reference` legs of `bench.py` may import it, and only as the
checker (or as the timed CPU baseline), never as the thing shipped.  The product path
(`kge-keras_b200/`) never imports `oracle` and fails loudly without its CUDA library.

Parity pinning status: the reference (`/root/reference`, TensorFlow/Keras) cannot be imported
in this image (no tensorflow, keras, rdflib, spacy), and its own `tests.py` asserts nothing, so
the TF-dependent part (score functions, losses, optimizer) is **parity unpinned** by reference
golden vectors: it is a line-by-line restatement, cross-checked fp32-vs-fp64 and against torch
autograd.  The NumPy/SciPy part (`generate_negative`, `oversample_data`, `gen_tail_data`,
`gen_head_data`, `validate`) IS pinned: `oracle/ref_ast.py` executes the reference's own function
bodies (AST-extracted, unmodified) in this container and `tests/golden/make_golden.py` froze
their outputs into `tests/golden/*.npz|json`, which the restatement is checked against.
"""
