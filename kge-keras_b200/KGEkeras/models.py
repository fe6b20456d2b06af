"""Host-side mirror of KGEkeras/models.py (reference): same class names, constructor keywords, `call` / `func` /
`fit` / `evaluate` / `predict` entry points -- but no TensorFlow underneath.  Every score, loss, gradient and update
is computed by libkgeb200's sm_100a kernels through KGEkeras._engine; there is no CPU fallback.

Reference lines mirrored: EmbeddingModel.__init__ models.py:14-129, call 134-162, DistMult 164-172, TransE 175-190,
ComplEx 203-222, HolE 224-237; siblings ModE 403-418, RotatE 421-459, pRotatE 461-489 (score_fwd/score_bwd + ranking
kernels; they train through the unfused kernel sequence), HAKE 362-400 (scoring and training as written, no ranking).
CosinE keeps its NameError.
"""
from __future__ import annotations

import warnings

import numpy as np

from . import _abi as abi


def l3_reg(weight_matrix, w=0.01):
    """models.py:10-11: w * ||W||_3^3 = w * sum |W|^3 of a weight matrix (an EmbeddingTable, a CUDA tensor or an array), computed
    by kgeb_l3_reg -- the kernel the training step uses when `regularization != 0` (models.py:79-82).  Returns a Python float."""
    import torch
    from ._engine import Engine, _ptr, _stream
    if isinstance(weight_matrix, EmbeddingTable):
        weight_matrix = weight_matrix.embeddings
    W = torch.as_tensor(np.asarray(weight_matrix, dtype=np.float32)) if not isinstance(weight_matrix, torch.Tensor) else weight_matrix.detach().to(torch.float32)
    if W.dim() != 2:
        raise ValueError(f"l3_reg expects a 2-D weight matrix, got shape {tuple(W.shape)}")
    rows, width = W.shape
    eng = Engine(abi.DISTMULT, rows, 1, width, width, E=W, R=np.zeros((1, width), np.float32))     # pads rows to the 128-bit granule with zeros
    G = torch.zeros_like(eng.E)
    flags = torch.zeros(rows, dtype=torch.uint8, device=eng.device)
    out = torch.zeros(1, dtype=torch.float32, device=eng.device)
    eng._check(eng.lib.kgeb_l3_reg(eng.ctx, _ptr(eng.E), rows, eng.we_k, eng.we_k, float(w), _ptr(G), _ptr(flags), _ptr(out), _stream()))
    return float(out.item())


class Dropout:
    """Stands in for keras.layers.Dropout(dp) (models.py:122): only the rate is kept."""

    def __init__(self, rate):
        self.rate = float(rate)


class EmbeddingTable:
    """Stands in for keras.layers.Embedding (models.py:89-112): owns nothing, views a table of the model's Engine."""

    def __init__(self, model, which, input_dim, output_dim, name):
        self._model, self._which = model, which
        self.input_dim, self.output_dim, self.name = input_dim, output_dim, name
        self.trainable = True

    def get_weights(self):
        E, R = self._model._engine_().get_tables()
        return [E if self._which == "E" else R]

    def set_weights(self, weights):
        (w,) = weights
        if self._model._head is not None:
            self._model._head.version += 1   # the cached inference table of a literal model depends on both tables
        if self._which == "E":
            self._model._engine_().set_tables(E=w)
        else:
            self._model._engine_().set_tables(R=w)

    @property
    def embeddings(self):
        eng = self._model._engine_()
        return eng.E[:, :eng.we] if self._which == "E" else eng.Rel[:, :eng.wr]

    def __call__(self, ids):
        import torch
        ids = torch.as_tensor(np.asarray(ids)).to(self.embeddings.device).long()
        return self.embeddings[ids]


class History:
    def __init__(self):
        self.history = {}
        self.epoch = []


class _OptimizerSpec:
    """Accepts what Keras' compile accepts for this path: a name or an object with Keras-style attributes."""

    def __init__(self, opt):
        name = opt if isinstance(opt, str) else type(opt).__name__
        name = name.lower()
        cfg = dict(learning_rate=None, beta_1=0.9, beta_2=0.999, epsilon=1e-7)
        if not isinstance(opt, str):
            for k in list(cfg):
                if hasattr(opt, k):
                    cfg[k] = float(getattr(opt, k))
            if cfg["learning_rate"] is None and hasattr(opt, "lr"):
                cfg["learning_rate"] = float(opt.lr)
        if name not in abi.OPT_KINDS:
            raise ValueError(f"optimizer {name!r} is not served by the B200 hot path (have {sorted(abi.OPT_KINDS)})")
        self.kind = abi.OPT_KINDS[name]
        # Keras defaults: Adam 1e-3, Adagrad 1e-3 (TF2), SGD 1e-2
        self.lr = cfg["learning_rate"] if cfg["learning_rate"] is not None else (1e-2 if name == "sgd" else 1e-3)
        self.beta1, self.beta2, self.eps = cfg["beta_1"], cfg["beta_2"], cfg["epsilon"]


class EmbeddingModel:
    """Base class (models.py:13-162).  Constructor signature identical to the reference's; `negative_samples`,
    `batch_size` and `loss_function`, which the reference accepts and drops, drive `fit` here."""

    _score_kind = None  # set by the four hot-path subclasses
    _e_mult, _r_mult = 1, 1

    def __init__(self,
                 e_dim,
                 r_dim,
                 num_entities,
                 num_relations,
                 negative_samples=2,
                 batch_size=16,
                 loss_function='pointwize_hinge',
                 dp=0.2,
                 margin=1,
                 loss_weight=1,
                 regularization=0.0,
                 literal_activation=None,
                 literals=None,
                 use_batch_norm=False,
                 entity_embedding_args=None,
                 relational_embedding_args=None,
                 init_entities=None,
                 init_relations=None,
                 name='embedding_model',
                 **kwargs):
        self.regularization = regularization
        self.num_entities = num_entities
        self.num_relations = num_relations
        self.name = name
        self.entity_embedding = EmbeddingTable(self, "E", num_entities, e_dim, name + '_entity_embedding')
        self.relational_embedding = EmbeddingTable(self, "R", num_relations, r_dim, name + '_relational_embedding')
        # literal branch (models.py:114-120): non-trainable literal table + one shared Dense(e_dim, activation)
        self._literals = literals if literal_activation else None
        self._literal_activation = literal_activation if literal_activation else None
        self._head = None
        self.dp = Dropout(dp)
        self.e_dim = e_dim
        self.r_dim = r_dim
        self.margin = margin
        self.loss_weight = loss_weight
        self.negative_samples = negative_samples
        self.batch_size = batch_size
        self.loss_function = loss_function
        self._init_E, self._init_R = init_entities, init_relations
        self._eng = None
        self._opt = _OptimizerSpec('adam')
        self._seed, self._step_counter = int(np.random.randint(0, 2 ** 31 - 1)), 0
        self.pair_mode = 'paired'          # 'allpairs' reproduces the (m,m) broadcast of the reference harness (tests.py:81)
        self.corrupt = 'both'              # reference behaviour (utils.py:117-120); 'head' | 'tail' | 'either' honour the flags
        self.stop_training = False
        self.__dict__.update(kwargs)       # models.py:129: unknown keywords become attributes

    # ---- engine ------------------------------------------------------------------------------------------------
    def _gamma_(self):
        return float(getattr(self, "gamma", 12.0))

    def _kind_(self):
        if self._score_kind is None:
            raise NotImplementedError(f"{type(self).__name__}: no score kernel (served: DistMult, TransE, ComplEx, HolE, HAKE, "
                                      f"ModE, RotatE, pRotatE)")
        return self._score_kind

    def _param_(self):
        return 1.0

    def _engine_(self):
        if self._eng is None:
            from ._engine import Engine
            self._eng = Engine(self._kind_(), self.num_entities, self.num_relations, self.e_dim, self.r_dim,
                               gamma=self._gamma_(), E=self._init_E, R=self._init_R, param=self._param_())
            self._init_E = self._init_R = None
        return self._eng

    def _head_(self):
        """The literal Dense head (None for models without literals)."""
        if self._literal_activation is None:
            return None
        if self._head is None:
            from ._literal import LiteralHead
            act = self._literal_activation if isinstance(self._literal_activation, str) else getattr(self._literal_activation, "__name__", None)
            self._head = LiteralHead(self._engine_(), self._literals, act)
            self._literals = None
        return self._head

    def _eval_engine_(self):
        """The engine inference runs on: the raw tables, or E' = act([E | L] W + b) for models with literals."""
        head = self._head_()
        return self._engine_() if head is None else head.effective_engine()

    def get_config(self):
        return self.__dict__

    # ---- forward -----------------------------------------------------------------------------------------------
    def call(self, inputs, training=False):
        """inputs: (batch, 3) ids (int or the harness' float32) -> (batch,) scores as a CUDA torch tensor.
        training=True applies Dropout(dp) to the gathered rows (models.py:144-150) with a fresh mask per call."""
        head = self._head_()
        if training and self.dp.rate > 0:
            self._call_counter = getattr(self, "_call_counter", 0) + 1
            if head is not None:
                return head.score(inputs, dropout=(self.dp.rate, self._seed ^ 0x5EED, 2 ** 41 + self._call_counter))
            return self._engine_().score(inputs, dropout=(self.dp.rate, self._seed ^ 0x5EED, 2 ** 41 + self._call_counter, 0))
        return self._eval_engine_().score(inputs)

    __call__ = call

    def func(self, s, p, o, training=False):
        """Score of already-gathered rows (the per-model hook of the reference).  Runs the same kernel on a scratch
        engine whose tables are the given rows."""
        import torch
        from ._engine import Engine
        s, p, o = (torch.as_tensor(np.asarray(x.detach().cpu() if hasattr(x, "detach") else x), dtype=torch.float32) for x in (s, p, o))
        n = s.shape[0]
        eng = Engine(self._kind_(), 2 * n, n, s.shape[1], p.shape[1], gamma=self._gamma_(), E=torch.cat([s, o]).numpy(), R=p.numpy(),
                     param=self._param_())
        idx = np.stack([np.arange(n), np.arange(n), n + np.arange(n)], axis=1).astype(np.int32)
        return eng.score(idx)

    def predict(self, x, batch_size=32, verbose=0, **kwargs):
        """Keras Model.predict: numpy (n,) scores.  The whole input is scored in one launch (batch_size is accepted)."""
        return self._eval_engine_().score(x).cpu().numpy()

    # ---- training ----------------------------------------------------------------------------------------------
    def compile(self, optimizer='adam', loss=None, **kwargs):
        self._opt = _OptimizerSpec(optimizer)
        if isinstance(loss, str):
            self.loss_function = loss

    def _loss_kind(self):
        if self.loss_function not in abi.LOSS_KINDS:
            raise KeyError(self.loss_function)  # same exception as loss_function_lookup (utils.py:321-330)
        return abi.LOSS_KINDS[self.loss_function]

    def _step_kwargs(self):
        return dict(loss_kind=self._loss_kind(), margin=float(self.margin if self.margin is not None else 1), loss_weight=float(self.loss_weight),
                    pair_mode={'paired': abi.PAIRED, 'allpairs': abi.ALLPAIRS}[self.pair_mode],
                    corrupt_mode={'both': abi.CORRUPT_BOTH, 'head': abi.CORRUPT_HEAD, 'tail': abi.CORRUPT_TAIL,
                                  'either': abi.CORRUPT_EITHER}[self.corrupt])

    @staticmethod
    def _split_xy(x):
        """([pos, neg], y) | [pos, neg] | pos  ->  (pos, neg or None)"""
        if isinstance(x, tuple) and len(x) == 2 and (x[1] is None or np.ndim(x[1]) <= 1) and isinstance(x[0], (list, tuple)):
            x = x[0]
        if isinstance(x, (list, tuple)) and len(x) == 2 and np.ndim(x[0]) >= 2:
            return x[0], x[1]
        return x, None

    def train_on_batch(self, x, y=None, **kwargs):
        pos, neg = self._split_xy(x)
        o = self._opt
        self._step_counter += 1
        stepper = self._head_() or self._engine_()
        res = stepper.train_step(pos, neg, opt_kind=o.kind, lr=o.lr, beta1=o.beta1, beta2=o.beta2, eps=o.eps,
                                 K=int(self.negative_samples), seed=self._seed, offset=self._step_counter,
                                 l3=float(self.regularization), dropout=float(self.dp.rate),
                                 **self._step_kwargs())
        return res["loss"].clone()   # the engine reuses one device scalar for every step's loss

    def test_on_batch(self, x, y=None, **kwargs):
        pos, neg = self._split_xy(x)
        eng = self._eval_engine_()
        kw = self._step_kwargs()
        if neg is None:
            neg = eng.negatives(pos, int(self.negative_samples), self._seed, 2 ** 40 + self._step_counter, kw["corrupt_mode"])
        t, f = eng.score(pos), eng.score(neg)
        K = max(1, f.numel() // max(1, t.numel()))
        loss, _, _ = eng.loss(t, f, K, kw["loss_kind"], kw["pair_mode"], kw["margin"], kw["loss_weight"], want_grads=False)
        return loss

    def fit(self, x=None, y=None, batch_size=None, epochs=1, verbose=1, callbacks=None, validation_data=None, shuffle=True,
            steps_per_epoch=None, **kwargs):
        """Keras-style fit.  x is a Sequence like the harness' DataGenerator (tests.py:29-61, yielding ([pos, neg], None)),
        a [pos, neg] pair of (m,3) arrays, or a (n,3) array of positives (negatives then come from the in-kernel Philox
        sampler, `negative_samples` per positive)."""
        hist = History()
        callbacks = list(callbacks or [])
        for cb in callbacks:
            if hasattr(cb, "set_model"):
                cb.set_model(self)
            else:
                cb.model = self
            if hasattr(cb, "on_train_begin"):
                cb.on_train_begin({})
        bs = int(batch_size or self.batch_size)
        is_seq = hasattr(x, "__getitem__") and hasattr(x, "__len__") and not isinstance(x, (list, tuple, np.ndarray)) and not hasattr(x, "shape")
        logs = {}
        for epoch in range(epochs):
            losses = []
            if is_seq:
                nb = len(x) if steps_per_epoch is None else min(len(x), steps_per_epoch)
                for b in range(nb):
                    losses.append(self.train_on_batch(x[b]))
                if hasattr(x, "on_epoch_end"):
                    x.on_epoch_end()
            else:
                # array input: the triples go to the device once, every batch is a device-side row gather (the shuffle
                # order comes from the global NumPy stream, like Keras' shuffle=True)
                import torch
                from ._engine import as_idx
                pos, neg = self._split_xy(x)
                dev = self._engine_().device
                if epoch == 0 or getattr(self, "_fit_cache_src", None) is not x:
                    self._fit_cache_src = x
                    self._fit_pos = as_idx(pos, dev)
                    self._fit_neg = None if neg is None else as_idx(neg, dev)
                pos_d, neg_d = self._fit_pos, self._fit_neg
                n = pos_d.shape[0]
                order = torch.from_numpy(np.random.permutation(n) if shuffle else np.arange(n)).to(dev)
                nb = (n + bs - 1) // bs if steps_per_epoch is None else steps_per_epoch
                per = None if neg_d is None else neg_d.shape[0] // n
                ar = None if per is None else torch.arange(per, device=dev)
                for b in range(nb):
                    sel = order[b * bs:(b + 1) * bs]
                    if sel.numel() == 0:
                        break
                    if neg_d is None:
                        losses.append(self.train_on_batch(pos_d[sel]))
                    else:
                        nsel = (sel[:, None] * per + ar[None, :]).reshape(-1)
                        losses.append(self.train_on_batch([pos_d[sel], neg_d[nsel]]))
            if losses:   # one read-back per epoch
                import torch
                logs = {"loss": float(torch.stack([l.reshape(()) for l in losses]).double().mean().item())}
            else:
                logs = {"loss": float("nan")}
            if validation_data is not None:
                logs["val_loss"] = float(self.evaluate(validation_data, verbose=0))
            for cb in callbacks:
                if hasattr(cb, "on_epoch_end"):
                    cb.on_epoch_end(epoch, logs)
            hist.epoch.append(epoch)
            for k, v in logs.items():
                hist.history.setdefault(k, []).append(v)
            if verbose:
                print(f"Epoch {epoch + 1}/{epochs} - " + " - ".join(f"{k}: {v:.4f}" for k, v in logs.items()))
            if self.stop_training:
                break
        for cb in callbacks:
            if hasattr(cb, "on_train_end"):
                cb.on_train_end(logs)
        return hist

    def evaluate(self, x=None, y=None, batch_size=None, verbose=1, **kwargs):
        """Mean loss over x (same input forms as fit), no parameter update."""
        is_seq = hasattr(x, "__getitem__") and hasattr(x, "__len__") and not isinstance(x, (list, tuple, np.ndarray)) and not hasattr(x, "shape")
        if is_seq:
            vals = [float(self.test_on_batch(x[b]).item()) for b in range(len(x))]
            return float(np.mean(vals)) if vals else float("nan")
        return float(self.test_on_batch(x).item())

    # ---- weights -----------------------------------------------------------------------------------------------
    def get_weights(self):
        """[entity table, relation table] (+ [Dense kernel, Dense bias] for models with literals)."""
        E, R = self._engine_().get_tables()
        head = self._head_()
        return [E, R] if head is None else [E, R] + head.get_weights()

    def set_weights(self, weights):
        self._engine_().set_tables(weights[0], weights[1])
        head = self._head_()
        if head is not None:
            if len(weights) >= 4:
                head.set_weights(weights[2], weights[3])
            else:
                head.version += 1


class DistMult(EmbeddingModel):
    _score_kind = abi.DISTMULT

    def __init__(self, name='DistMult', **kwargs):
        """DistMult (models.py:164-172): sum(s*p*o)."""
        super(DistMult, self).__init__(name=name, **kwargs)


class TransE(EmbeddingModel):
    def __init__(self, name='TransE', norm=1, gamma=12, **kwargs):
        """TransE (models.py:175-190): gamma - ||s+p-o||_norm (bare norm if gamma <= 0)."""
        super(TransE, self).__init__(name=name, **kwargs)
        self.gamma = gamma
        self.norm = norm

    def _kind_(self):
        if self.norm == 1:
            return abi.TRANSE_L1
        if self.norm == 2:
            return abi.TRANSE_L2
        raise NotImplementedError(f"TransE norm={self.norm!r}: the kernels serve ord 1 and 2")


class ComplEx(EmbeddingModel):
    _score_kind = abi.COMPLEX

    def __init__(self, name='ComplEx', **kwargs):
        """ComplEx (models.py:203-222): rows are [re | im]; like the reference, e_dim / r_dim must be passed by keyword
        and are doubled."""
        kwargs['e_dim'] = 2 * kwargs['e_dim']
        kwargs['r_dim'] = 2 * kwargs['r_dim']
        super(ComplEx, self).__init__(name=name, **kwargs)


class HolE(EmbeddingModel):
    _score_kind = abi.HOLE

    def __init__(self, name='HolE', **kwargs):
        """HolE (models.py:224-237): sum_k p_k (s (*) o)_k, circular correlation.  `fft=True` (default) scores and differentiates
        through the Fourier domain like the reference (models.py:233-234) when the width is a power of two in 32..1024;
        `fft=False` keeps the direct O(d^2) kernels with their sequential fp32 contract (other widths always use them)."""
        self.fft = bool(kwargs.pop("fft", True))
        super(HolE, self).__init__(name=name, **kwargs)

    def _engine_(self):
        fresh = self._eng is None
        eng = super(HolE, self)._engine_()
        if fresh:
            eng.set_hole_mode(self.fft)
        return eng


# ---- sibling classes: constructors kept (same keywords, same width multipliers), score kernels are SURVEY 8f "next" ----
class CosinE(EmbeddingModel):
    def __init__(self, name='CosinE', **kwargs):
        super(CosinE, self).__init__(name=name, **kwargs)

    def func(self, s, p, o, training=False):
        # the reference calls an undefined name here (models.py:200); the defect is kept, not papered over
        raise NameError("name 'cosine_similarity' is not defined")


class ModE(EmbeddingModel):
    def __init__(self, gamma=12, norm=2, name='ModE', **kwargs):
        kwargs['e_dim'] = 2 * kwargs['e_dim']
        kwargs['r_dim'] = 3 * kwargs['r_dim']
        super(ModE, self).__init__(name=name, **kwargs)
        self.norm = norm
        self.gamma = gamma  # the reference reads self.gamma without assigning it (models.py:415); fixed here

    def _kind_(self):
        """gamma - ||s*p - o||_norm (models.py:417-418).  The product needs equal stored widths, i.e. 2*e_dim == 3*r_dim
        as passed by the caller (the reference would fail in TF's broadcast otherwise; here the C ABI reports it)."""
        if self.norm == 1:
            return abi.MODE_L1
        if self.norm == 2:
            return abi.MODE_L2
        raise NotImplementedError(f"ModE norm={self.norm!r}: the kernels serve ord 1 and 2")


class HAKE(EmbeddingModel):
    def __init__(self, epsilon=2, gamma=12, phase_weight=0.5, mod_weight=1, name='HAKE', **kwargs):
        kwargs['e_dim'] = 2 * kwargs['e_dim']
        kwargs['r_dim'] = 3 * kwargs['r_dim']
        super(HAKE, self).__init__(name=name, **kwargs)
        self.gamma, self.epsilon, self.phase_weight, self.mod_weight = gamma, epsilon, phase_weight, mod_weight
        self.pi = np.pi
        self.embedding_range = (self.gamma + self.epsilon) / self.e_dim / 2

    _score_kind = abi.HAKE   # models.py:383-400 as written: the modulus norm spans the whole call (kgeb_set_hake_params)

    def _param_(self):
        return (float(self.embedding_range / self.pi), float(self.phase_weight), float(self.mod_weight))   # the divisor of models.py:390-392

    def predict(self, x, batch_size=32, verbose=0, **kwargs):
        """Keras Model.predict scores its input in consecutive batches of `batch_size` rows (default 32) and every batch is its
        own call of func(): the axis-less norm of models.py:397 therefore couples the rows of ONE batch, and HAKE's predictions
        -- hence the ranks the reference's validate derives from them (utils.py:211) -- depend on the batching.  Reproduced
        here by kgeb_set_hake_call_rows: one launch, one norm per group of batch_size rows (batch_size=None: one group)."""
        eng = self._eval_engine_()
        eng._check(eng.lib.kgeb_set_hake_call_rows(eng.ctx, int(batch_size) if batch_size else 0))
        try:
            return eng.score(x).cpu().numpy()
        finally:
            eng._check(eng.lib.kgeb_set_hake_call_rows(eng.ctx, 0))


class RotatE(EmbeddingModel):
    _score_kind = abi.ROTATE   # models.py:439-459; needs e_dim == r_dim as passed (entity rows 2d, d relation phases)

    def __init__(self, gamma=12, norm=2, epsilon=2, name='RotatE', **kwargs):
        kwargs['e_dim'] = 2 * kwargs['e_dim']
        kwargs['r_dim'] = kwargs['r_dim']
        super(RotatE, self).__init__(name=name, **kwargs)
        self.gamma, self.epsilon, self.norm = gamma, epsilon, norm
        self.pi = np.pi
        self.embedding_range = (self.gamma + self.epsilon) / self.e_dim / 2


class pRotatE(EmbeddingModel):
    _score_kind = abi.PROTATE  # models.py:479-489

    def _param_(self):
        return float(self.modulus)

    def __init__(self, gamma=12, epsilon=2, modulus=0.5, name='pRotatE', **kwargs):
        kwargs['e_dim'] = 2 * kwargs['e_dim']
        kwargs['r_dim'] = 2 * kwargs['r_dim']
        super(pRotatE, self).__init__(name=name, **kwargs)
        self.gamma, self.epsilon = gamma, epsilon
        self.pi = np.pi
        self.embedding_range = (self.gamma + self.epsilon) / self.e_dim / 2
        self.modulus = modulus * self.embedding_range


class _ConvOutOfScope(EmbeddingModel):
    def __init__(self, *args, **kwargs):
        raise NotImplementedError(f"{type(self).__name__} is a dense conv net (models.py:239-360), outside the "
                                  f"gather/score/scatter hot path this build accelerates (SURVEY 2 #4)")


class ConvE(_ConvOutOfScope):
    pass


class ConvR(_ConvOutOfScope):
    pass


class ConvKB(_ConvOutOfScope):
    pass
