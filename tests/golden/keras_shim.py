"""TEST INFRASTRUCTURE (never imported by the product): a torch-backed stand-in for the slice of the
TensorFlow / Keras / TensorFlow-Recommenders API that the reference's model-building code calls, so that the
reference's OWN functions can be EXECUTED in this container (TensorFlow cannot be installed here):

    src/models/NeuMFModel.py:53-100    NeuMFModel.compileModel           (functional Keras graph)
    src/models/BPRModel.py:49-74,124-144  BPRModel.compileModel, bprTripletLoss, identityLoss
    trainers/twoTower.py:19-111        TwoTowerModel.__init__/call/setCandidates/computeEmb/computeLoss*/train_step

What executing them pins: the WIRING -- which tensor feeds which layer, concat order, BatchNorm after the
activation, Dot axes, loss choice, optimizer and learning rate, the gradient-tape / apply_gradients sequence,
StringLookup offsets, the arguments handed to tfrs.tasks.Retrieval.  What it does not pin: the arithmetic inside
each Keras / TFRS layer, which is restated here from the upstream documentation exactly
as oracle/ restates it -- that residue is "upstream numerics, unpinned" (DESIGN.md section 2).

Layer semantics restated (Keras 2.3/2.4 defaults): Embedding = row lookup after casting ids to int; Dense = x W + b
with glorot-uniform W [in, out]; BatchNormalization(axis=-1, momentum .99, eps 1e-3), training: biased batch
variance, moving = moving * .99 + batch * .01; Dropout: inverted, mask injected by the test (TF's stream is not
reproducible); Dot(axes) = sum of products over `axes`, keepdims; Concatenate(axis=-1); Flatten; Lambda.
Optimizers: Adam (eps 1e-7 outside the root, bias correction folded into the step size; sparse gradients are
applied dense-equivalently), Adagrad (initial accumulator 0.1, eps 1e-7).  TFRS: Retrieval task (in-batch softmax,
SUM reduction, accidental-hit removal with finfo(float32).min / 100), BruteForce / Streaming top-k.
"""
import sys
import types

import numpy as np
import torch

DT = torch.float64          # the shim runs in float64: it is the reference side of a parity check
_CREATED = []               # every layer in creation order (the golden scripts assign weights by this order)
DROPOUT_MASKS = {}          # Dropout layer index (creation order among Dropout layers) -> mask tensor or None


def reset():
    _CREATED.clear()
    DROPOUT_MASKS.clear()


def _t(x):
    if isinstance(x, torch.Tensor):
        return x
    a = np.asarray(x)
    if a.dtype.kind in "US":
        return a
    return torch.as_tensor(a, dtype=DT if a.dtype.kind == "f" else torch.int64)


# ---- symbolic functional API ---------------------------------------------------------------------------------------
class Sym:
    def __init__(self, layer=None, inputs=None, name=None):
        self.layer, self.inputs, self.name = layer, inputs, name


def _has_sym(x):
    if isinstance(x, Sym):
        return True
    if isinstance(x, (list, tuple)):
        return any(_has_sym(v) for v in x)
    return False


def Input(shape=None, name=None, **kw):
    return Sym(None, None, name)


class Layer:
    def __init__(self, name=None, **kw):
        self.name = name
        self.built = False
        _CREATED.append(self)

    def variables(self):
        return []

    def __call__(self, x, training=False, **kw):
        if _has_sym(x):
            return Sym(self, x)
        return self.forward(x, training)


class Embedding(Layer):
    def __init__(self, input_dim, output_dim, name=None, input_length=None, **kw):
        super().__init__(name)
        self.embeddings = torch.empty(input_dim, output_dim, dtype=DT).uniform_(-0.05, 0.05).requires_grad_()

    def variables(self):
        return [self.embeddings]

    def forward(self, ids, training):
        ids = _t(ids)
        return self.embeddings[ids.to(torch.int64)]      # Keras casts float ids to int32


class Dense(Layer):
    def __init__(self, units, activation=None, name=None, **kw):
        super().__init__(name)
        self.units, self.activation = units, activation
        self.kernel = self.bias = None

    def variables(self):
        return [self.kernel, self.bias] if self.kernel is not None else []

    def build(self, fan_in):
        lim = float(np.sqrt(6.0 / (fan_in + self.units)))
        self.kernel = torch.empty(fan_in, self.units, dtype=DT).uniform_(-lim, lim).requires_grad_()
        self.bias = torch.zeros(self.units, dtype=DT).requires_grad_()

    def forward(self, x, training):
        if self.kernel is None:
            self.build(x.shape[-1])
        y = x @ self.kernel + self.bias
        if self.activation in (None, "linear"):
            return y
        if self.activation == "relu":
            return torch.relu(y)
        if self.activation == "sigmoid":
            return torch.sigmoid(y)
        raise NotImplementedError(self.activation)


class BatchNormalization(Layer):
    def __init__(self, name=None, momentum=0.99, epsilon=1e-3, **kw):
        super().__init__(name)
        self.momentum, self.epsilon = momentum, epsilon
        self.gamma = self.beta = self.moving_mean = self.moving_variance = None

    def variables(self):
        return [self.gamma, self.beta] if self.gamma is not None else []

    def build(self, n):
        self.gamma = torch.ones(n, dtype=DT).requires_grad_()
        self.beta = torch.zeros(n, dtype=DT).requires_grad_()
        self.moving_mean = torch.zeros(n, dtype=DT)
        self.moving_variance = torch.ones(n, dtype=DT)

    def forward(self, x, training):
        if self.gamma is None:
            self.build(x.shape[-1])
        if training:
            mu, var = x.mean(0), x.var(0, unbiased=False)
            with torch.no_grad():
                self.moving_mean = self.moving_mean * self.momentum + mu * (1 - self.momentum)
                self.moving_variance = self.moving_variance * self.momentum + var * (1 - self.momentum)
        else:
            mu, var = self.moving_mean, self.moving_variance
        return self.gamma * (x - mu) / torch.sqrt(var + self.epsilon) + self.beta


class Dropout(Layer):
    def __init__(self, rate, **kw):
        super().__init__(None)
        self.rate = rate
        self.index = sum(isinstance(l, Dropout) for l in _CREATED) - 1

    def forward(self, x, training):
        m = DROPOUT_MASKS.get(self.index) if training else None
        return x if m is None else x * m


class Concatenate(Layer):
    def __init__(self, axis=-1, **kw):
        super().__init__(None)
        self.axis = axis

    def forward(self, xs, training):
        return torch.cat(list(xs), dim=self.axis)


class Flatten(Layer):
    def forward(self, x, training):
        return x.reshape(x.shape[0], -1)


class Dot(Layer):
    def __init__(self, axes, **kw):
        super().__init__(None)
        self.axes = axes

    def forward(self, xs, training):
        a, b = xs
        return (a * b).sum(dim=self.axes, keepdim=True)


class Lambda(Layer):
    def __init__(self, function, output_shape=None, **kw):
        super().__init__(None)
        self.function = function

    def forward(self, xs, training):
        return self.function(xs)


class StringLookup(Layer):
    """TF 2.3 / 2.4 experimental StringLookup: index 0 = mask token, 1 = out-of-vocabulary, vocabulary from 2."""

    def __init__(self, vocabulary=None, **kw):
        super().__init__(None)
        self.table = {str(v): j + 2 for j, v in enumerate(vocabulary)}

    def forward(self, x, training):
        a = np.asarray(x).reshape(-1)
        return torch.as_tensor([self.table.get(str(v), 1) for v in a], dtype=torch.int64)


class Sequential(Layer):
    def __init__(self, layers=None, **kw):
        self.layers = list(layers or [])
        self.name = None

    def add(self, l):
        self.layers.append(l)

    def variables(self):
        return [v for l in self.layers for v in l.variables()]

    def __call__(self, x, training=False, **kw):
        for l in self.layers:
            x = l(x, training=training)
        return x


def _evaluate(node, feed, training, cache):
    if isinstance(node, (list, tuple)):
        return [_evaluate(n, feed, training, cache) for n in node]
    if id(node) in cache:
        return cache[id(node)]
    if node.layer is None:
        v = _t(feed[node.name])
    else:
        v = node.layer.forward(_evaluate(node.inputs, feed, training, cache), training)
    cache[id(node)] = v
    return v


class Model:
    """tf.keras.Model: functional (inputs, outputs) or subclassed."""

    def __init__(self, *args, inputs=None, outputs=None, name=None, **kw):
        if len(args) >= 2 and (_has_sym(args[0]) or _has_sym(args[1])):
            inputs, outputs = args[0], args[1]
        object.__setattr__(self, "_tracked", [])
        self.inputs, self.outputs, self.name = inputs, outputs, name
        self.optimizer = self.loss = self.compiled_loss = None
        self.compiled_metrics_names = None

    def __setattr__(self, k, v):
        if isinstance(v, (Layer, Sequential)) and hasattr(self, "_tracked"):
            self._tracked.append(v)
        object.__setattr__(self, k, v)

    # -- functional --
    def _graph_layers(self):
        seen, order = set(), []

        def walk(n):
            if isinstance(n, (list, tuple)):
                for m in n:
                    walk(m)
                return
            if n.layer is not None:
                walk(n.inputs)
                if id(n.layer) not in seen:
                    seen.add(id(n.layer)); order.append(n.layer)
        walk(self.outputs)
        return sorted(order, key=_CREATED.index)

    @property
    def layers(self):
        if self.outputs is not None:
            return self._graph_layers()
        out = []
        for l in self._tracked:
            out.extend(l.layers if isinstance(l, Sequential) else [l])
        return out

    @property
    def trainable_variables(self):
        seen, out = set(), []
        for l in self.layers:
            for v in l.variables():
                if id(v) not in seen:
                    seen.add(id(v)); out.append(v)
        return out

    @property
    def metrics(self):
        return []

    def compile(self, optimizer=None, loss=None, metrics=None, **kw):
        self.optimizer, self.loss, self.compiled_metrics_names = optimizer, loss, metrics
        self.compiled_loss = _resolve_loss(loss)

    def __call__(self, feed, training=False):
        if self.outputs is None:
            return self.call(feed)
        if not isinstance(feed, dict):
            feed = {s.name: v for s, v in zip(self.inputs, feed)}
        return _evaluate(self.outputs, feed, training, {})


# ---- losses / optimizers ---------------------------------------------------------------------------------------------
def _mse(y_true, y_pred):
    return ((y_pred - _t(y_true).reshape(y_pred.shape)) ** 2).mean()


class BinaryCrossentropy:
    """Keras BinaryCrossentropy on probabilities: clip to [eps, 1 - eps], mean over everything."""

    def __call__(self, y_true, y_pred):
        eps = 1e-7
        p = y_pred.clamp(eps, 1 - eps)
        y = _t(y_true).reshape(p.shape).to(p.dtype)
        return -(y * torch.log(p) + (1 - y) * torch.log(1 - p)).mean()


def _resolve_loss(loss):
    if loss in ("mean_squared_error", "mse"):
        return _mse
    return loss


class Adam:
    def __init__(self, learning_rate=1e-3, lr=None, beta_1=0.9, beta_2=0.999, epsilon=1e-7, **kw):
        self.lr = lr if lr is not None else learning_rate
        self.b1, self.b2, self.eps, self.t, self.state = beta_1, beta_2, epsilon, 0, {}

    def apply_gradients(self, grads_and_vars):
        self.t += 1
        alpha = self.lr * np.sqrt(1 - self.b2 ** self.t) / (1 - self.b1 ** self.t)
        with torch.no_grad():
            for g, v in grads_and_vars:
                g = torch.zeros_like(v) if g is None else g
                m, s = self.state.setdefault(id(v), (torch.zeros_like(v), torch.zeros_like(v)))
                m.mul_(self.b1).add_(g, alpha=1 - self.b1)
                s.mul_(self.b2).addcmul_(g, g, value=1 - self.b2)
                v.sub_(alpha * m / (s.sqrt() + self.eps))


class Adagrad:
    def __init__(self, learning_rate=1e-3, initial_accumulator_value=0.1, epsilon=1e-7, **kw):
        self.lr, self.init, self.eps, self.state = learning_rate, initial_accumulator_value, epsilon, {}

    def apply_gradients(self, grads_and_vars):
        with torch.no_grad():
            for g, v in grads_and_vars:
                g = torch.zeros_like(v) if g is None else g
                acc = self.state.setdefault(id(v), torch.full_like(v, self.init))
                acc.add_(g * g)
                v.sub_(self.lr * g / (acc.sqrt() + self.eps))


class GradientTape:
    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False

    def gradient(self, loss, variables):
        return list(torch.autograd.grad(loss, list(variables), allow_unused=True))


# ---- tf.data (just enough for setCandidates) -----------------------------------------------------------------------------
class Dataset:
    def __init__(self, items):
        self.items = list(items)

    @staticmethod
    def from_tensor_slices(x):
        return Dataset(list(x))

    def batch(self, n):
        return Dataset([self.items[k:k + n] for k in range(0, len(self.items), n)])

    def map(self, fn):
        return Dataset([fn(x) for x in self.items])

    def __iter__(self):
        return iter(self.items)

    def __len__(self):
        return len(self.items)


# ---- TensorFlow Recommenders ---------------------------------------------------------------------------------------------
MIN_FLOAT = float(np.finfo(np.float32).min) / 100.0


class Retrieval:
    """tfrs.tasks.Retrieval: scores = Q C^T, labels = identity; with candidate_ids, off-diagonal entries whose candidate
    id equals the row's positive id get MIN_FLOAT added (RemoveAccidentalHits); categorical cross-entropy from logits,
    reduction SUM (the task's default loss)."""

    def __init__(self, loss=None, **kw):
        self.loss = loss
        self.calls = []

    def __call__(self, query_embeddings, candidate_embeddings, sample_weight=None, candidate_sampling_probability=None,
                 candidate_ids=None, compute_metrics=True, training=False, **kw):
        self.calls.append(dict(compute_metrics=compute_metrics, training=training, has_ids=candidate_ids is not None))
        scores = query_embeddings @ candidate_embeddings.T
        B = scores.shape[0]
        if candidate_ids is not None:
            ids = np.asarray(candidate_ids).reshape(-1)
            same = torch.as_tensor(ids[:, None] == ids[None, :]) & ~torch.eye(B, dtype=torch.bool)
            scores = scores + same.to(scores.dtype) * MIN_FLOAT
        if self.loss is not None:
            return self.loss(torch.eye(B, dtype=scores.dtype), scores)
        return torch.nn.functional.cross_entropy(scores, torch.arange(B), reduction="sum")


class BruteForce(Layer):
    def __init__(self, query_model=None, k=10, **kw):
        self.k = k
        self.C = self.ids = None

    def index(self, candidates, identifiers=None):
        self.C = torch.cat([c for c in candidates], dim=0)
        self.ids = np.asarray(list(identifiers)) if identifiers is not None else np.arange(self.C.shape[0])
        return self

    def __call__(self, queries, k=None, **kw):
        k = k or self.k
        scores = queries @ self.C.T
        order = torch.argsort(-scores, dim=1, stable=True)[:, :k]          # tf.math.top_k: ties -> lower index first
        return torch.gather(scores, 1, order), self.ids[order.numpy()]


class Streaming(BruteForce):
    pass


# ---- module tree ---------------------------------------------------------------------------------------------------------
def install():
    """Registers the shim as `tensorflow`, `tensorflow.keras...`, `keras...` and `tensorflow_recommenders`."""
    def mod(name, **attrs):
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        sys.modules[name] = m
        return m

    layer_names = dict(Input=Input, Embedding=Embedding, Dense=Dense, BatchNormalization=BatchNormalization,
                       Dropout=Dropout, Concatenate=Concatenate, Flatten=Flatten, Dot=Dot, Lambda=Lambda, Layer=Layer)
    preprocessing = mod("tensorflow.keras.layers.experimental.preprocessing", StringLookup=StringLookup)
    experimental = mod("tensorflow.keras.layers.experimental", preprocessing=preprocessing)
    layers = mod("tensorflow.keras.layers", experimental=experimental, **layer_names)
    models = mod("tensorflow.keras.models", Model=Model, Sequential=Sequential)
    optimizers = mod("tensorflow.keras.optimizers", Adam=Adam, Adagrad=Adagrad)
    losses = mod("tensorflow.keras.losses", BinaryCrossentropy=BinaryCrossentropy)
    activations = mod("tensorflow.keras.activations", sigmoid=torch.sigmoid)
    utils = mod("tensorflow.keras.utils", model_to_dot=lambda *a, **k: None)
    keras = mod("tensorflow.keras", layers=layers, models=models, optimizers=optimizers, losses=losses,
                activations=activations, utils=utils, Model=Model, Sequential=Sequential, Input=Input)
    math = mod("tensorflow.math",
               reduce_sum=lambda x, axis=None, keepdims=False: x.sum(dim=axis, keepdim=keepdims) if axis is not None else x.sum(),
               reduce_mean=lambda x, axis=None, keepdims=False: x.mean(dim=axis, keepdim=keepdims) if axis is not None else x.mean(),
               multiply=lambda a, b: a * b, subtract=lambda a, b: a - b, top_k=None)
    data = mod("tensorflow.data", Dataset=Dataset)
    tf = mod("tensorflow", keras=keras, math=math, data=data, function=lambda f: f, GradientTape=GradientTape,
             constant=lambda v, **k: torch.as_tensor(v, dtype=DT), sigmoid=torch.sigmoid,
             convert_to_tensor=_t, ones=lambda n: torch.ones(n, dtype=DT), float32="float32")
    # import-only names the reference files pull in
    dso = mod("tensorflow.python.data.ops.dataset_ops", DatasetV2=Dataset)
    mod("tensorflow.python.data.ops", dataset_ops=dso)
    mod("tensorflow.python.data", ops=sys.modules["tensorflow.python.data.ops"])
    dl = mod("tensorflow.python.distribute.distribute_lib", Strategy=object)
    mod("tensorflow.python.distribute", distribute_lib=dl)
    pk = mod("tensorflow.python.keras.models", Model=Model)
    mod("tensorflow.python.keras", models=pk)
    mod("tensorflow.python", data=sys.modules["tensorflow.python.data"], distribute=sys.modules["tensorflow.python.distribute"],
        keras=sys.modules["tensorflow.python.keras"])
    tf.python = sys.modules["tensorflow.python"]
    ftk = mod("tensorflow_recommenders.layers.factorized_top_k", BruteForce=BruteForce, Streaming=Streaming)
    tl = mod("tensorflow_recommenders.layers", factorized_top_k=ftk)
    tt = mod("tensorflow_recommenders.tasks", Retrieval=Retrieval)
    mod("tensorflow_recommenders", layers=tl, tasks=tt)
    return tf
