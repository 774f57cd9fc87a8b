"""ORACLE (test infrastructure, never the product path): embedding gather, sparse-gradient
scatter-add and the Keras optimizers, restated on the CPU in NumPy.

UPSTREAM NUMERICS UNPINNED: the reference's tests hold no numbers for these ops (SURVEY.md section
0.3) and TensorFlow cannot be installed here, so these functions restate the upstream Keras/TensorFlow
(>=2.3.1, /root/reference/requirements.txt:1) semantics that the reference's call sites rely on.
tests/test_oracle_models.py pins them against hand-computed fp64 cases; how the reference WIRES them
(which table feeds which lookup, which optimizer with which learning rate) is pinned by executed
reference code (tests/test_oracle_wiring.py).

Call sites restated:
  Embedding(...)(ids)           NeuMFModel.py:58-63, BPRModel.py:55-61, bpr.py:178-184, twoTower.py:34,36
  Adam(1e-3) / Adam(lr=0.005)   NeuMFModel.py:89, BPRModel.py:70, bpr.py:201, NFC_plain.py:153
  "Adagrad", 0.1                twoTower.py:278-279 (through the missing trainers.model_utils.getOptimizer)
"""
import numpy as np


def gather_rows(table, ids):
    """out[b,:] = table[ids[b],:]  (tf.gather / ResourceGather under keras Embedding)."""
    return np.ascontiguousarray(table[np.asarray(ids, dtype=np.int64)])


def scatter_add_rows(num_rows, ids, values, dtype=None):
    """Dense [num_rows,d] sum of IndexedSlices(values, ids): duplicates are summed
    (tf UnsortedSegmentSum in the optimizer's de-duplication)."""
    values = np.asarray(values)
    out = np.zeros((num_rows, values.shape[1]), dtype=dtype or values.dtype)
    np.add.at(out, np.asarray(ids, dtype=np.int64), values)
    return out


def keras_adam_alpha(lr, beta1, beta2, t):
    """Keras Adam folds the bias corrections into the step size:
    alpha_t = lr * sqrt(1 - beta2^t) / (1 - beta1^t)   (t = 1 for the first step)."""
    return lr * np.sqrt(1.0 - beta2 ** t) / (1.0 - beta1 ** t)


def adam_dense_keras(w, m, v, g, t, lr=1e-3, beta1=0.9, beta2=0.999, eps=1e-7):
    """One Keras Adam step on a whole variable, in place.  For a sparse (IndexedSlices)
    gradient Keras is dense-equivalent: m and v of EVERY row decay and every row moves
    (SURVEY.md section 0.4), so g is the dense scatter-added gradient (zero for untouched rows).
    eps is added OUTSIDE the square root (Keras default epsilon 1e-7)."""
    dt = w.dtype
    a = dt.type(keras_adam_alpha(lr, beta1, beta2, t))
    m *= dt.type(beta1)
    m += dt.type(1.0 - beta1) * g
    v *= dt.type(beta2)
    v += dt.type(1.0 - beta2) * g * g
    w -= a * m / (np.sqrt(v) + dt.type(eps))


def adam_rows_lazy(w, m, v, g_dense, rows, t, lr=1e-3, beta1=0.9, beta2=0.999, eps=1e-7):
    """Lazy (row-sparse) Adam: only `rows` (unique row ids hit by the batch) are updated.
    Deviates from Keras for rows with stale momentum; used for tables too large for a dense
    pass (BASELINE.json configs[3]) and stated as such in DESIGN.md."""
    rows = np.asarray(rows, dtype=np.int64)
    dt = w.dtype
    a = dt.type(keras_adam_alpha(lr, beta1, beta2, t))
    g = g_dense[rows]
    m[rows] = dt.type(beta1) * m[rows] + dt.type(1.0 - beta1) * g
    v[rows] = dt.type(beta2) * v[rows] + dt.type(1.0 - beta2) * g * g
    w[rows] -= a * m[rows] / (np.sqrt(v[rows]) + dt.type(eps))


def adagrad_rows(w, acc, g_dense, rows, lr=0.1, eps=1e-7):
    """Keras Adagrad sparse apply (ResourceSparseApplyAdagradV2): duplicates summed first,
    acc += g^2, w -= lr * g / (sqrt(acc) + eps); accumulator starts at 0.1; only touched rows move."""
    rows = np.asarray(rows, dtype=np.int64)
    dt = w.dtype
    g = g_dense[rows]
    acc[rows] += g * g
    w[rows] -= dt.type(lr) * g / (np.sqrt(acc[rows]) + dt.type(eps))


def adagrad_dense(w, acc, g, lr=0.1, eps=1e-7):
    """Keras Adagrad on a dense variable (ResourceApplyAdagradV2)."""
    dt = w.dtype
    acc += g * g
    w -= dt.type(lr) * g / (np.sqrt(acc) + dt.type(eps))


def keras_embedding_init(rng, rows, dim, dtype=np.float32):
    """Keras Embedding default initializer 'uniform' = U(-0.05, 0.05)."""
    return rng.uniform(-0.05, 0.05, size=(rows, dim)).astype(dtype)


def glorot_uniform(rng, fan_in, fan_out, dtype=np.float32):
    """Keras Dense default kernel initializer; kernel layout [fan_in, fan_out]."""
    lim = np.sqrt(6.0 / (fan_in + fan_out))
    return rng.uniform(-lim, lim, size=(fan_in, fan_out)).astype(dtype)
