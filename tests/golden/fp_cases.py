"""Case list shared by tests/golden/make_golden_fp.py (which runs the reference's own models.py / utils.py source through
oracle/ref_tf_exec.py and freezes the outputs in tests/golden/fp_ref.npz) and by the tests that read the fixture.

A case names the reference CLASS whose `func` is executed (models.py line range in the comment), the constructor
keywords, the oracle / C-ABI score kind the result pins, and the stored table widths as multiples of d."""
import os

FIXTURE = os.path.join(os.path.dirname(os.path.abspath(__file__)), "fp_ref.npz")

N_ENT, N_REL, N_TRIPLES = 60, 7, 48

# name, reference class, ctor kwargs, kind name (oracle.kge_oracle constant), entity width / d, relation width / d, d
SCORE_CASES = [
    ("DistMult_d16", "DistMult", {}, "DISTMULT", 1, 1, 16),                          # models.py:171-172
    ("DistMult_d100", "DistMult", {}, "DISTMULT", 1, 1, 100),
    ("TransE_L1_d16", "TransE", {"norm": 1}, "TRANSE_L1", 1, 1, 16),                 # models.py:186-190
    ("TransE_L1_d200", "TransE", {"norm": 1}, "TRANSE_L1", 1, 1, 200),
    ("TransE_L2_d16", "TransE", {"norm": 2}, "TRANSE_L2", 1, 1, 16),
    ("TransE_L2_d100", "TransE", {"norm": 2}, "TRANSE_L2", 1, 1, 100),
    ("TransE_L1_gamma0_d16", "TransE", {"norm": 1, "gamma": 0}, "TRANSE_L1", 1, 1, 16),     # the gamma <= 0 branch, 189-190
    ("TransE_L2_gammaNeg_d16", "TransE", {"norm": 2, "gamma": -1}, "TRANSE_L2", 1, 1, 16),
    ("ComplEx_d16", "ComplEx", {}, "COMPLEX", 2, 2, 16),                             # models.py:212-222
    ("ComplEx_d100", "ComplEx", {}, "COMPLEX", 2, 2, 100),
    ("HolE_d16", "HolE", {}, "HOLE", 1, 1, 16),                                      # models.py:231-237
    ("HolE_d64", "HolE", {}, "HOLE", 1, 1, 64),
    ("RotatE_d16", "RotatE", {}, "ROTATE", 2, 1, 16),                                # models.py:439-459
    ("RotatE_gamma0_d16", "RotatE", {"gamma": 0}, "ROTATE", 2, 1, 16),
    ("pRotatE_d16", "pRotatE", {}, "PROTATE", 2, 2, 16),                             # models.py:479-489
    ("ModE_L1_d16", "ModE", {"norm": 1}, "MODE_L1", 1, 1, 16),                       # models.py:417-418 (func only, see generator)
    ("ModE_L2_d16", "ModE", {"norm": 2}, "MODE_L2", 1, 1, 16),
    ("HAKE_d16", "HAKE", {}, "HAKE", 2, 3, 16),                                      # models.py:383-400
]

LOSS_NAMES = ["pointwize_hinge", "pointwize_logistic", "pointwize_square_loss", "pointwize_cross_entropy",
              "pairwize_hinge", "pairwize_logistic", "pairwize_square_loss"]            # utils.py:288-330
LOSS_B, LOSS_K = 24, 5
LOSS_MARGINS = [1.0, 0.5]

# name, score case it reuses the class / widths of, loss name, pair mode ("PAIRED": true (B,1), negative_samples=K;
# "ALLPAIRS": the harness call loss(score, fscore, margin, 1) on two equally long lists, tests.py:74-83), extras
STEP_CASES = [
    ("step_DistMult_pairwize_hinge_allpairs", "DistMult_d16", "pairwize_hinge", "ALLPAIRS", {}),
    ("step_TransE_L1_pairwize_hinge_paired", "TransE_L1_d16", "pairwize_hinge", "PAIRED", {}),
    ("step_TransE_L2_pairwize_logistic_paired", "TransE_L2_d16", "pairwize_logistic", "PAIRED", {}),
    ("step_ComplEx_pointwize_logistic", "ComplEx_d16", "pointwize_logistic", "PAIRED", {}),
    ("step_HolE_pairwize_logistic_paired", "HolE_d16", "pairwize_logistic", "PAIRED", {}),
    ("step_HAKE_pairwize_logistic_paired", "HAKE_d16", "pairwize_logistic", "PAIRED", {}),
    ("step_DistMult_l3reg", "DistMult_d16", "pairwize_hinge", "PAIRED", {"regularization": 0.01}),   # models.py:10-11,80-92
    ("step_literal_DistMult_tanh", "DistMult_d16", "pairwize_hinge", "ALLPAIRS", {"literals": 5, "literal_activation": "tanh"}),
]
STEP_B, STEP_K = 16, 3
