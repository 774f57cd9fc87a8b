"""Biased-SVD matrix factorisation -- drop-in mirror of the reference's src/origin_models/svd/SVD.py
(SURVEY.md section 8 row f4).

The reference is a module of constants and functions driven by pandas chunks: digest (:105-124) builds the id
dictionaries and the global mean, fit_model (:187-221) walks the ratings one `iterrows()` row at a time,
mean_square_error / mean_absolute_error (:223-253) walk them again, do_topk (:418-442) pushes the float32
matrices through TFRS BruteForce and trainers/topKmetrics.  The names, argument order and return values are kept;
what changes is underneath:

  * the chunk iterator becomes a `Ratings` frame: dense int32 user / item ids, float64 ratings and the
    dependency tickets of the file order, resident in HBM;
  * fit_model is one cooperative kernel launch per epoch that applies the ratings with EXACTLY the sequential
    semantics (csrc/svd.cu) -- not a Hogwild approximation;
  * matrices and bias vectors are float64 device tensors, updated in place like the NumPy arrays.

No CPU fallback: every function raises BrkError without the CUDA library / a device.
"""
import ctypes as C
import os

import numpy as np
import torch

from . import _native as N
from . import distributed as D
from . import hotpath as H
from . import pipeline as PL
from . import topKmetrics as topk

# module constants of the reference (SVD.py:14-62)
EPOCHS = 1
LEARNING_RATE = 0.01
EMBEDDING_REGULARIZATION = 0
BIAS_REGULARIZATION = 0.01
NUMBER_OF_EMBEDDINGS = 50
TOPK_BATCH_SIZE = 5000
EPOCH_ERROR_CALCULATION_FREQUENCY = 1
EVALUATE = True
VERBOSE = False
NUMBER_OF_FILES = 5
NUMBER_OF_CHUNKS_TO_EAT = 5
USER_ID_COLUMN = "CUSTOMER_ID"
ITEM_ID_COLUMN = "PRODUCT_ID"
RATING_COLUMN = "RATING_TYPE"                 # None: ratings from the quintiles of the two columns below
TRANSACTION_COUNT_COLUMN = "TRANSACTION_COUNT"
QUANTITY_SUM_COLUMN = "QUANTITY_SUM"
FILE_PATH = None
TRANSACTION_COUNT_SCALE = 0.5
QUANTITY_SUM_SCALE = 0.5
TRANSACTION_COUNT_QUINTILES = (1, 2, 4)
QUANTITY_SUM_QUINTILES = (1, 1, 2)


def _dev(device=None):
    if not torch.cuda.is_available():
        raise N.BrkError("no CUDA device: the brk_b200 hot path is sm_100a CUDA only (no CPU fallback)")
    return torch.device(device) if device is not None else torch.device(f"cuda:{torch.cuda.current_device()}")


def _f64(t, name, dev=None):
    if not isinstance(t, torch.Tensor):
        t = torch.from_numpy(np.ascontiguousarray(t, dtype=np.float64)).to(_dev(dev))
    if t.dtype != torch.float64 or not t.is_cuda or not t.is_contiguous():
        raise TypeError(f"{name} must be a contiguous float64 CUDA tensor")
    return t


def _i32(t, name, dev=None):
    if not isinstance(t, torch.Tensor):
        t = torch.from_numpy(np.ascontiguousarray(t, dtype=np.int32)).to(_dev(dev))
    if t.dtype != torch.int32 or not t.is_cuda or not t.is_contiguous():
        raise TypeError(f"{name} must be a contiguous int32 CUDA tensor")
    return t


def _key_kind(column):
    return "i" if isinstance(column, torch.Tensor) or np.asarray(column).dtype.kind in "iu" else "S"


def _reduce_ws(dev):
    return torch.empty(max(int(N.lib().brk_svd_reduce_workspace_bytes(N.ctx(dev))), 256), dtype=torch.uint8, device=dev)


class Ratings:
    """The reference's `dataset` (an iterable of DataFrame chunks) as device columns in file order:
    users / items int32 dense ids, ratings float64, and `sched` int32 [n, 4] = (user, item, user ticket,
    item ticket) -- how many earlier ratings touch the same user / item row (brk_svd_schedule)."""

    def __init__(self, users, items, ratings, num_users=None, num_items=None, device=None):
        dev = _dev(device)
        self.users = _i32(users, "users", dev)
        self.items = _i32(items, "items", dev)
        self.ratings = _f64(ratings, "ratings", dev)
        n = self.users.numel()
        if self.items.numel() != n or self.ratings.numel() != n:
            raise ValueError("users, items and ratings must have one entry per rating")
        self.num_users = int(num_users) if num_users is not None else (int(self.users.max().item()) + 1 if n else 1)
        self.num_items = int(num_items) if num_items is not None else (int(self.items.max().item()) + 1 if n else 1)
        self.device = self.users.device
        self.sched = torch.empty((max(n, 1), 4), dtype=torch.int32, device=self.device)
        lib = N.lib()
        ws_bytes = int(lib.brk_svd_schedule_workspace_bytes(n, self.num_users, self.num_items))
        if ws_bytes < 0:
            raise ValueError(f"unsupported sizes: n={n} users={self.num_users} items={self.num_items}")
        ws = torch.empty(max(ws_bytes, 256), dtype=torch.uint8, device=self.device)
        bad = torch.zeros(1, dtype=torch.int32, device=self.device)
        N.check(lib.brk_svd_schedule(N.ctx(self.device), N.ptr(self.users), N.ptr(self.items), n, self.num_users,
                                     self.num_items, N.ptr(self.sched), N.ptr(bad), N.ptr(ws), ws.numel(),
                                     N.stream_ptr()), "brk_svd_schedule")
        if int(bad.item()):
            raise IndexError(f"rating ids outside [0, {self.num_users}) x [0, {self.num_items})")
        self._fit_ws = None           # device scratch of fit_model, sized on first use (depends on d)
        # rating count of the busiest row = a lower bound of the critical path; n / that = the parallelism on offer
        self.max_row_count = int(self.sched[:n, 2:4].max().item()) + 1 if n else 0

    def __len__(self):
        return self.users.numel()

    def critical_path(self):
        """Length of the longest chain of ratings that must run one after the other (max dependency level) -- what
        bounds an epoch on the device.  Host recurrence over the file; diagnostics only."""
        u = self.users.cpu().numpy(); i = self.items.cpu().numpy()
        lu = np.zeros(self.num_users, dtype=np.int64); li = np.zeros(self.num_items, dtype=np.int64)
        for a, b in zip(u.tolist(), i.tolist()):
            l = (lu[a] if lu[a] > li[b] else li[b]) + 1
            lu[a] = l; li[b] = l
        return int(max(lu.max(initial=0), li.max(initial=0)))


class RatingChunks:
    """The reference's chunked datasets -- movielens_cross_validation (SVD.py:301-347: one file, shuffled, split into
    five) and grundfos_network_drive_files (:349-409: five files) -- over in-memory columns.  Same protocol:
    iteration yields every chunk except the held-out one; use_no_test_set(), next_cross_validation_distribution(),
    get_test_set().  A chunk is a (raw users, raw items, ratings) column triple; after digest() the chunks also hold
    dense ids and `train_frame()` / `test_frame()` give the device frames fit_model and the error functions take."""

    def __init__(self, chunks, test_positive_only=False):
        self.chunks = [(np.asarray(u), np.asarray(i), np.asarray(r, dtype=np.float64)) for u, i, r in chunks]
        if not self.chunks:
            raise ValueError("no chunks")
        self.number_of_chunks = len(self.chunks)
        self.test_set_index = self.number_of_chunks - 1           # start with the last chunk as the test set
        # grundfos_network_drive_files.get_test_set keeps only rows whose rating column is 1 (:389-392)
        self.test_positive_only = bool(test_positive_only)
        self.dense = None                                         # per chunk (users int32, items int32) after digest
        self.num_users = self.num_items = None
        self.device = None
        self._frames = {}

    @classmethod
    def split(cls, raw_users, raw_items, ratings, number_of_chunks=5, shuffle_seed=None, **kw):
        """movielens_cross_validation: (optionally shuffled, the reference's `.sample(frac=1)` is unseeded) rows cut
        into number_of_chunks pieces with np.array_split (:308-310)."""
        u, i, r = np.asarray(raw_users), np.asarray(raw_items), np.asarray(ratings, dtype=np.float64)
        if shuffle_seed is not None:
            perm = np.random.default_rng(shuffle_seed).permutation(len(u))
            u, i, r = u[perm], i[perm], r[perm]
        return cls(list(zip(np.array_split(u, number_of_chunks), np.array_split(i, number_of_chunks),
                            np.array_split(r, number_of_chunks))), **kw)

    def use_no_test_set(self):
        self.test_set_index = self.number_of_chunks

    def next_cross_validation_distribution(self):
        self.test_set_index -= 1
        return self.test_set_index >= 0

    def _test_rows(self):
        if self.test_set_index == self.number_of_chunks:
            raise Exception("There is no test set because use_no_test_set was called. Call "
                            "next_cross_validation_distribution to set the first chunk to be the test set again.")
        u, i, r = self.chunks[self.test_set_index]
        keep = (r == 1) if self.test_positive_only else np.ones(len(r), dtype=bool)
        return keep

    def get_test_set(self):
        keep = self._test_rows()
        u, i, r = self.chunks[self.test_set_index]
        return u[keep], i[keep], r[keep]

    def training_chunk_indices(self):
        return [c for c in range(self.number_of_chunks) if c != self.test_set_index]

    def __iter__(self):
        """The dataset is its own iterator, as in the reference (:332-347): a new `for` restarts it."""
        self.next_chunk = 0
        return self

    def __next__(self):
        while self.next_chunk == self.test_set_index:
            self.next_chunk += 1                                  # the held-out chunk is skipped
        if self.next_chunk >= self.number_of_chunks:
            raise StopIteration
        chunk = self.chunks[self.next_chunk]
        self.next_chunk += 1
        return chunk

    # ---- device side (after digest) ------------------------------------------------------------------------
    def _need_digest(self):
        if self.dense is None:
            raise N.BrkError("RatingChunks: call digest(dataset) first (dense ids are assigned there)")

    def train_frame(self):
        """Ratings frame of the training chunks in iteration order (cached per held-out chunk)."""
        self._need_digest()
        key = ("train", self.test_set_index)
        if key not in self._frames:
            idx = self.training_chunk_indices()
            u = np.concatenate([self.dense[c][0] for c in idx]) if idx else np.zeros(0, np.int32)
            i = np.concatenate([self.dense[c][1] for c in idx]) if idx else np.zeros(0, np.int32)
            r = np.concatenate([self.chunks[c][2] for c in idx]) if idx else np.zeros(0)
            self._frames[key] = Ratings(u, i, r, num_users=self.num_users, num_items=self.num_items, device=self.device)
        return self._frames[key]

    def test_frame(self):
        self._need_digest()
        key = ("test", self.test_set_index)
        if key not in self._frames:
            keep = self._test_rows()
            du, di = self.dense[self.test_set_index]
            self._frames[key] = Ratings(du[keep], di[keep], self.chunks[self.test_set_index][2][keep],
                                        num_users=self.num_users, num_items=self.num_items, device=self.device)
        return self._frames[key]


def _read_chunk(path_or_file, columns):
    """One rating file -> (raw users, raw items, ratings): `columns` is what the reference hands to read_csv as
    usecols (SVD.py:503-506): [user, item, rating] or [user, item, transaction count, quantity sum]."""
    import pandas as pd
    df = pd.read_csv(path_or_file, usecols=list(columns))
    u, i = df[columns[0]].to_numpy(), df[columns[1]].to_numpy()
    if len(columns) == 3:
        r = df[columns[2]].to_numpy(dtype=np.float64)
    else:
        r = get_rating(df[columns[2]].to_numpy(dtype=np.float64), df[columns[3]].to_numpy(dtype=np.float64)).cpu().numpy()
    return u, i, r


class movielens_cross_validation(RatingChunks):
    """SVD.py:301-347 by name: one CSV, rows shuffled (`.sample(frac=1)` is unseeded there; shuffle_seed here, None =
    file order), cut into five chunks; number_of_chunks_to_eat is the number of chunks the iterator walks."""

    def __init__(self, file_path, number_of_chunks_to_eat, columns, shuffle_seed=0):
        u, i, r = _read_chunk(file_path, columns)
        split = RatingChunks.split(u, i, r, 5, shuffle_seed=shuffle_seed)
        super().__init__(split.chunks)
        self.file_path, self.number_of_chunks_to_eat, self.columns = file_path, number_of_chunks_to_eat, columns
        self.test_set_index = self.number_of_chunks_to_eat - 1


class grundfos_network_drive_files(RatingChunks):
    """SVD.py:349-409 by name: number_of_files CSVs `file_path.format(1..n)` (read from the local file system; the
    SMB credentials are accepted and ignored), one chunk per file; the test set keeps only rating == 1 rows when the
    files carry a rating column (:389-392)."""

    def __init__(self, file_path, number_of_files, credentials, columns):
        chunks = [_read_chunk(file_path.format(n), columns) for n in range(1, number_of_files + 1)]
        super().__init__(chunks, test_positive_only=(len(columns) == 3))
        self.file_path, self.number_of_files, self.columns = file_path, number_of_files, columns
        self.username, self.password = credentials if credentials else (None, None)
        self.files = self.chunks


def get_data(file_path=None, grundfos=True, credentials=(None, None)):
    """SVD.py:499-514 without the prompts: the dataset object for FILE_PATH (or file_path) with the module's column
    constants -- grundfos_network_drive_files over NUMBER_OF_FILES files, or movielens_cross_validation."""
    path = file_path if file_path is not None else FILE_PATH
    if path is None:
        raise ValueError("get_data: set SVD.FILE_PATH or pass file_path (the reference hard-codes an SMB share)")
    if RATING_COLUMN is not None:
        columns = [USER_ID_COLUMN, ITEM_ID_COLUMN, RATING_COLUMN]
    else:
        columns = [USER_ID_COLUMN, ITEM_ID_COLUMN, TRANSACTION_COUNT_COLUMN, QUANTITY_SUM_COLUMN]
    if grundfos:
        return grundfos_network_drive_files(path, NUMBER_OF_FILES, credentials, columns)
    return movielens_cross_validation(path, NUMBER_OF_CHUNKS_TO_EAT, columns)


def get_config():
    """SVD.py:79-103: the hyper-parameters of the run (the reference adds the git commit of its checkout)."""
    import subprocess
    try:
        sha = subprocess.run(["git", "rev-parse", "HEAD"], capture_output=True, text=True, timeout=5,
                             cwd=os.path.dirname(os.path.abspath(__file__))).stdout.strip() or None
    except Exception:
        sha = None
    return {"epochs": EPOCHS, "learning_rate": LEARNING_RATE, "regularization": EMBEDDING_REGULARIZATION,
            "number_of_embeddings": NUMBER_OF_EMBEDDINGS, "file_path": FILE_PATH, "git_commit_sha": sha,
            "rating_column": RATING_COLUMN, "transaction_count_column": TRANSACTION_COUNT_COLUMN,
            "transaction_count_scale": TRANSACTION_COUNT_SCALE, "transaction_count_quintiles": TRANSACTION_COUNT_QUINTILES,
            "quantity_sum_column": QUANTITY_SUM_COLUMN, "quantity_sum_scale": QUANTITY_SUM_SCALE,
            "quantity_sum_quintiles": QUANTITY_SUM_QUINTILES}


_verbose_print_count = 0


def print_verbose(message):
    """SVD.py:64-71: progress line with a spinner, only when VERBOSE."""
    global _verbose_print_count
    if VERBOSE:
        print("-\\|/-\\|/"[_verbose_print_count % 8] + "  " + message, end="\r")
        _verbose_print_count += 1


def clear_verbose_print():
    """SVD.py:73-77."""
    global _verbose_print_count
    if VERBOSE:
        _verbose_print_count = 0
        print("\r" + " " * 80, end="\r")


def _frame_of(dataset):
    """fit_model / the error functions take what the reference passes: the chunk iterator (its training chunks), a
    list holding the test frame (`[test_dataframe]`, SVD.py:462) or a Ratings frame."""
    if isinstance(dataset, Ratings):
        return dataset
    if isinstance(dataset, RatingChunks):
        return dataset.train_frame()
    if isinstance(dataset, (list, tuple)) and len(dataset) == 1 and isinstance(dataset[0], Ratings):
        return dataset[0]
    raise TypeError("expected a Ratings frame, a RatingChunks dataset or [Ratings]")


def get_rating(transaction_count, quantity_sum, tc_scale=TRANSACTION_COUNT_SCALE, qs_scale=QUANTITY_SUM_SCALE,
               tc_quintiles=TRANSACTION_COUNT_QUINTILES, qs_quintiles=QUANTITY_SUM_QUINTILES, device=None):
    """get_rating with RATING_COLUMN = None (SVD.py:255-262) over whole columns: float64 ratings on the device."""
    tc = _f64(transaction_count, "transaction_count", device)
    qs = _f64(quantity_sum, "quantity_sum", device)
    if tc.numel() != qs.numel():
        raise ValueError("transaction_count and quantity_sum differ in length")
    out = torch.empty_like(tc)
    a = (C.c_double * 3)(*[float(x) for x in tc_quintiles]); b = (C.c_double * 3)(*[float(x) for x in qs_quintiles])
    N.check(N.lib().brk_svd_quintile_ratings(N.ctx(tc.device), N.ptr(tc), N.ptr(qs), tc.numel(), float(tc_scale),
                                             float(qs_scale), a, b, N.ptr(out), N.stream_ptr()),
            "brk_svd_quintile_ratings")
    return out


def place_in_quintile(value, quintiles):
    """SVD.py:264-270 for one value (host scalar helper, kept for callers of the reference's name)."""
    q1, median, q3 = quintiles
    return 4 if value > q3 else 3 if value > median else 2 if value > q1 else 1


def digest(raw_users, raw_items=None, ratings=None, device=None):
    """digest (SVD.py:105-124).  Two call forms:
      digest(dataset)  with a RatingChunks -- the reference's call: ids and the global mean over the chunks the
        iterator currently yields (all of them after use_no_test_set()); returns the reference's 5-tuple
        (user_ids, item_ids, uid_max, iid_max, global_bias) and leaves the dense ids in the dataset;
      digest(raw_users, raw_items, ratings)  over whole columns: returns the 5-tuple plus the Ratings frame.
    user_ids / item_ids are the reference's dicts raw id -> dense id (order of first appearance)."""
    if isinstance(raw_users, RatingChunks):
        ds = raw_users
        idx = ds.training_chunk_indices()
        u = np.concatenate([ds.chunks[c][0] for c in idx]); i = np.concatenate([ds.chunks[c][1] for c in idx])
        r = np.concatenate([ds.chunks[c][2] for c in idx])
        user_ids, item_ids, uid_max, iid_max, mu, frame = digest(u, i, r, device=device)
        ds.device, ds.num_users, ds.num_items = frame.device, uid_max + 1, iid_max + 1
        du, di = frame.users.cpu().numpy(), frame.items.cpu().numpy()
        cuts = np.cumsum([0] + [len(ds.chunks[c][0]) for c in idx])
        ds.dense = [None] * ds.number_of_chunks
        for n, c in enumerate(idx):
            ds.dense[c] = (du[cuts[n]:cuts[n + 1]], di[cuts[n]:cuts[n + 1]])
        for c in range(ds.number_of_chunks):                      # a chunk held out during digest: ids must be known
            if ds.dense[c] is None:
                try:
                    ds.dense[c] = (np.array([user_ids[x] for x in ds.chunks[c][0].tolist()], np.int32),
                                   np.array([item_ids[x] for x in ds.chunks[c][1].tolist()], np.int32))
                except KeyError as e:                             # the reference fails the same way in fit_model
                    raise KeyError(f"id {e} of the held-out chunk was not seen by digest") from None
        ds._frames = {}
        return user_ids, item_ids, uid_max, iid_max, mu
    dev = _dev(device)
    uv, iv = PL.Vocabulary(dev), PL.Vocabulary(dev)
    u = uv.build(raw_users)
    i = iv.build(raw_items)
    r = _f64(ratings, "ratings", dev)
    if r.numel() == 0:
        raise ZeroDivisionError("digest of an empty rating file")          # the reference divides by the row count
    out = torch.empty(2, dtype=torch.float64, device=dev)
    N.check(N.lib().brk_svd_mean(N.ctx(dev), N.ptr(r), r.numel(), N.ptr(out), N.ptr(_reduce_ws(dev)), N.stream_ptr()),
            "brk_svd_mean")
    kind, kind_i = _key_kind(raw_users), _key_kind(raw_items)
    uk, ik = uv.host_keys(kind), iv.host_keys(kind_i)
    user_ids = {(int(k) if kind == "i" else k): j for j, k in enumerate(uk)}
    item_ids = {(int(k) if kind_i == "i" else k): j for j, k in enumerate(ik)}
    frame = Ratings(u, i, r, num_users=uv.size, num_items=iv.size, device=dev)
    return user_ids, item_ids, uv.size - 1, iv.size - 1, float(out[0].item()), frame


def convert_ids(chunk, chunk_number, user_ids, item_ids, next_user_id, next_item_id):
    """SVD.py:126-137 for one chunk (a (raw users, raw items, ...) column tuple): raw ids not seen before get the next
    dense id, in row order; the dicts are updated in place.  Host dict logic (digest does the same on the device for
    whole files through pipeline.Vocabulary)."""
    for raw in np.asarray(chunk[0]).tolist():
        if raw not in user_ids:
            user_ids[raw] = next_user_id
            next_user_id += 1
    for raw in np.asarray(chunk[1]).tolist():
        if raw not in item_ids:
            item_ids[raw] = next_item_id
            next_item_id += 1
    return next_user_id, next_item_id


def calculate_average(chunk, chunk_number, total_so_far, average_so_far):
    """SVD.py:139-161: the running mean rating after one more chunk (chunk[2] = its ratings)."""
    r = np.asarray(chunk[2], dtype=np.float64)
    chunk_total = len(r)
    chunk_average = float(r.sum()) / chunk_total                  # ZeroDivisionError on an empty chunk, like the reference
    total_so_far += chunk_total
    average_so_far = ((chunk_total / total_so_far) * chunk_average) + \
        (((total_so_far - chunk_total) / total_so_far) * average_so_far)
    return total_so_far, average_so_far


def get_idset(chunks):
    """SVD.py:410-416: the set of (raw user, raw item) pairs of the given chunks."""
    result = set()
    for chunk in chunks:
        result.update(zip(np.asarray(chunk[0]).tolist(), np.asarray(chunk[1]).tolist()))
    return result


def init_parameters(number_of_users, number_of_items, number_of_embeddings=NUMBER_OF_EMBEDDINGS, seed=None, device=None):
    """(user_matrix, item_matrix, user_bias_vector, item_bias_vector) as train_and_evaluate creates them
    (SVD.py:446-449): U(0, 1) / d matrices, zero biases, float64."""
    dev = _dev(device)
    rng = np.random.default_rng(seed)
    P = torch.from_numpy(rng.random((number_of_users, number_of_embeddings)) * (1 / number_of_embeddings)).to(dev)
    Q = torch.from_numpy(rng.random((number_of_items, number_of_embeddings)) * (1 / number_of_embeddings)).to(dev)
    return P, Q, torch.zeros(number_of_users, dtype=torch.float64, device=dev), \
        torch.zeros(number_of_items, dtype=torch.float64, device=dev)


def _check_params(frame, user_matrix, item_matrix, user_bias_vector, item_bias_vector):
    P, Q = _f64(user_matrix, "user_matrix"), _f64(item_matrix, "item_matrix")
    bu, bi = _f64(user_bias_vector, "user_bias_vector"), _f64(item_bias_vector, "item_bias_vector")
    if P.dim() != 2 or Q.dim() != 2 or P.shape[1] != Q.shape[1]:
        raise ValueError("user_matrix / item_matrix must be [rows, d] with one d")
    if P.shape[0] < frame.num_users or Q.shape[0] < frame.num_items or bu.numel() < frame.num_users or \
            bi.numel() < frame.num_items:
        raise IndexError("parameter tables are smaller than the id range of the ratings")
    return P, Q, bu, bi


def fit_model(dataset, user_matrix, item_matrix, user_bias_vector, item_bias_vector, global_bias, user_ids=None,
              item_ids=None, learning_rate=None, embedding_regularization=None, bias_regularization=None,
              warps_per_sm=None):
    """One pass over `dataset` (a Ratings frame) in file order, parameters updated in place (SVD.py:187-221).
    Hyper-parameters default to the module constants, like the reference's globals.  warps_per_sm: resident warps
    per SM (0 = as many as fit); by default 8 when the file offers little parallelism (fewer than 512 ratings per
    link of its longest row chain: more waiting warps only add polling traffic -- measured 5.8 vs 6.4 ms on the
    ML-1M-shaped file), otherwise 0.  The result does not depend on it."""
    dataset = _frame_of(dataset)
    P, Q, bu, bi = _check_params(dataset, user_matrix, item_matrix, user_bias_vector, item_bias_vector)
    lr = LEARNING_RATE if learning_rate is None else learning_rate
    ereg = EMBEDDING_REGULARIZATION if embedding_regularization is None else embedding_regularization
    breg = BIAS_REGULARIZATION if bias_regularization is None else bias_regularization
    if warps_per_sm is None:
        warps_per_sm = 8 if len(dataset) < 512 * max(dataset.max_row_count, 1) else 0
    need = int(N.lib().brk_svd_fit_workspace_bytes(dataset.num_users, dataset.num_items, P.shape[1]))
    if need < 0:
        raise ValueError(f"unsupported row width d={P.shape[1]} (1..512)")
    if dataset._fit_ws is None or dataset._fit_ws.numel() < need:
        dataset._fit_ws = torch.empty(need, dtype=torch.uint8, device=dataset.device)
    N.check(N.lib().brk_svd_fit_epoch(N.ctx(dataset.device), N.ptr(dataset.sched), N.ptr(dataset.ratings), len(dataset),
                                      N.ptr(P), N.ptr(Q), N.ptr(bu), N.ptr(bi), dataset.num_users, dataset.num_items,
                                      P.shape[1], float(global_bias), float(lr), float(ereg), float(breg),
                                      N.ptr(dataset._fit_ws), dataset._fit_ws.numel(), int(warps_per_sm),
                                      N.stream_ptr()), "brk_svd_fit_epoch")


def check_fit(dataset):
    """Raises if the last fit_model over `dataset` gave up on a wait (synchronises)."""
    dataset = _frame_of(dataset)
    if dataset._fit_ws is not None and int(dataset._fit_ws[:4].view(torch.int32).item()) != 0:
        raise N.BrkError("brk_svd_fit_epoch aborted: the schedule does not belong to these ratings")


def predict(user, item, user_matrix, item_matrix, user_bias_vector, item_bias_vector, global_bias):
    """predict (SVD.py:179-185) for one pair (Python ints -> float) or for id arrays (-> float64 device tensor)."""
    P, Q = _f64(user_matrix, "user_matrix"), _f64(item_matrix, "item_matrix")
    bu, bi = _f64(user_bias_vector, "user_bias_vector"), _f64(item_bias_vector, "item_bias_vector")
    scalar = np.isscalar(user) and np.isscalar(item)
    u = _i32(np.atleast_1d(user) if not isinstance(user, torch.Tensor) else user, "user", P.device)
    i = _i32(np.atleast_1d(item) if not isinstance(item, torch.Tensor) else item, "item", P.device)
    if u.numel() != i.numel():
        raise ValueError("user and item differ in length")
    if u.numel() and (int(u.min()) < 0 or int(u.max()) >= P.shape[0] or int(i.min()) < 0 or int(i.max()) >= Q.shape[0]):
        raise IndexError("id outside the parameter tables")
    out = torch.empty(u.numel(), dtype=torch.float64, device=P.device)
    N.check(N.lib().brk_svd_predict(N.ctx(P.device), N.ptr(u), N.ptr(i), u.numel(), N.ptr(P), N.ptr(Q), N.ptr(bu),
                                    N.ptr(bi), P.shape[1], float(global_bias), N.ptr(out), N.stream_ptr()),
            "brk_svd_predict")
    return float(out.item()) if scalar else out


def _errors(dataset, user_matrix, item_matrix, user_bias_vector, item_bias_vector, global_bias):
    dataset = _frame_of(dataset)
    P, Q, bu, bi = _check_params(dataset, user_matrix, item_matrix, user_bias_vector, item_bias_vector)
    if len(dataset) == 0:
        raise ZeroDivisionError("mean error over an empty rating set")      # accumulator / count, SVD.py:247
    out = torch.empty(2, dtype=torch.float64, device=dataset.device)
    N.check(N.lib().brk_svd_errors(N.ctx(dataset.device), N.ptr(dataset.users), N.ptr(dataset.items),
                                   N.ptr(dataset.ratings), len(dataset), N.ptr(P), N.ptr(Q), N.ptr(bu), N.ptr(bi),
                                   P.shape[1], float(global_bias), N.ptr(out), N.ptr(_reduce_ws(dataset.device)),
                                   N.stream_ptr()), "brk_svd_errors")
    return out


def mean_generic_error(generic, dataset, user_matrix, item_matrix, user_bias_vector, item_bias_vector, global_bias,
                       user_ids=None, item_ids=None):
    """SVD.py:223-247: mean of generic(rating - prediction).  abs and squaring are recognised (by probing the callable)
    and reduced on the device; any other callable is applied on the host to the device-computed errors."""
    probe = (generic(-2.0), generic(0.5))
    if probe == (2.0, 0.5):
        return mean_absolute_error(dataset, user_matrix, item_matrix, user_bias_vector, item_bias_vector, global_bias)
    if probe == (4.0, 0.25):
        return mean_square_error(dataset, user_matrix, item_matrix, user_bias_vector, item_bias_vector, global_bias)
    frame = _frame_of(dataset)
    if len(frame) == 0:
        raise ZeroDivisionError("mean error over an empty rating set")
    pred = predict(frame.users, frame.items, user_matrix, item_matrix, user_bias_vector, item_bias_vector, global_bias)
    err = (frame.ratings - pred).cpu().numpy()
    return float(sum(generic(float(e)) for e in err) / len(err))


def mean_square_error(dataset, user_matrix, item_matrix, user_bias_vector, item_bias_vector, global_bias,
                      user_ids=None, item_ids=None):
    """SVD.py:252-253."""
    return float(_errors(dataset, user_matrix, item_matrix, user_bias_vector, item_bias_vector, global_bias)[0].item())


def mean_absolute_error(dataset, user_matrix, item_matrix, user_bias_vector, item_bias_vector, global_bias,
                        user_ids=None, item_ids=None):
    """SVD.py:249-250."""
    return float(_errors(dataset, user_matrix, item_matrix, user_bias_vector, item_bias_vector, global_bias)[1].item())


def recommend_users(users, user_matrix, item_matrix, k):
    """(scores [n, k] float64, item rows [n, k] int32) for a list of dense user ids: plain dot products over the
    whole item matrix, best first, ties -> lower item row (recommend, SVD.py:286-299, batched)."""
    P, Q = _f64(user_matrix, "user_matrix"), _f64(item_matrix, "item_matrix")
    u = _i32(np.atleast_1d(users) if not isinstance(users, torch.Tensor) else users, "users", P.device)
    if u.numel() and (int(u.min()) < 0 or int(u.max()) >= P.shape[0]):
        raise IndexError("user id outside user_matrix")
    if int(k) < 1:
        raise ValueError("k must be >= 1")
    n, I = u.numel(), Q.shape[0]
    vals = torch.empty((n, int(k)), dtype=torch.float64, device=P.device)
    ids = torch.empty((n, int(k)), dtype=torch.int32, device=P.device)
    chunk = max(1, min(n, (1 << 28) // max(I, 1)))                         # <= 2 GiB of float64 scores per launch
    scores = torch.empty((min(chunk, max(n, 1)), I), dtype=torch.float64, device=P.device)
    for a in range(0, n, chunk):
        b = min(n, a + chunk)
        N.check(N.lib().brk_svd_recommend(N.ctx(P.device), N.ptr(P), N.ptr(u[a:b]), b - a, N.ptr(Q), I, P.shape[1],
                                          int(k), N.ptr(scores), N.ptr(vals[a:b]), N.ptr(ids[a:b]), N.stream_ptr()),
                "brk_svd_recommend")
    return vals, ids


class rating_prediction:
    """SVD.py:272-283: what recommend returns -- .index, .prediction, printed as "(prediction @ index)".  Iterating
    yields (index, prediction), so `for i, p in recommend(...)` works as well."""

    def __init__(self, index, prediction):
        self.index = index
        self.prediction = prediction

    def __iter__(self):
        return iter((self.index, self.prediction))

    def __str__(self):
        return f"({self.prediction} @ {self.index})"

    __repr__ = __str__


class svd_prediction_doer:
    """SVD.py:163-177: the adaptor that gives the factor matrices the `predict` interface of the NFC model, queried
    with raw ids: predict([[raw user], [raw item]]) -> predict(user_ids[raw user], item_ids[raw item], ...)."""

    def __init__(self, user_ids, item_ids, user_matrix, item_matrix, user_bias_vector, item_bias_vector, global_bias):
        self.user_matrix = user_matrix
        self.item_matrix = item_matrix
        self.user_bias_vector = user_bias_vector
        self.item_bias_vector = item_bias_vector
        self.global_bias = global_bias
        self.user_ids = user_ids
        self.item_ids = item_ids

    def predict(self, weird_array):
        user = self.user_ids[weird_array[0][0]]
        item = self.item_ids[weird_array[1][0]]
        return predict(user, item, self.user_matrix, self.item_matrix, self.user_bias_vector, self.item_bias_vector,
                       self.global_bias)


def recommend(user_vector, item_matrix, k):
    """recommend (SVD.py:286-299): the k best (index, prediction) of one user vector, best first.  The reference
    returns its rating_prediction objects in replacement order; here the same objects come sorted by descending
    prediction (ties: lower index); unfilled slots (k > number of items) keep index None / prediction -inf as there."""
    Q = _f64(item_matrix, "item_matrix")
    p = _f64(user_vector, "user_vector", Q.device).reshape(1, -1)
    vals, ids = recommend_users(torch.zeros(1, dtype=torch.int32, device=Q.device), p, Q, k)
    vals, ids = vals[0].cpu().numpy(), ids[0].cpu().numpy()
    return [rating_prediction(int(i) if i >= 0 else None, float(v)) for v, i in zip(vals, ids)]


def do_topk(user_matrix, item_matrix, test_idset, train_idset, user_ids, item_ids, k=10):
    """do_topk (SVD.py:418-442): float32 BruteForce top-10 of every user over the item matrix (the fused scoring +
    top-K kernel), then topKMetrics against the test pairs and against all pairs.  user_ids / item_ids: the
    digest dicts raw id -> dense row."""
    P, Q = _f64(user_matrix, "user_matrix"), _f64(item_matrix, "item_matrix")
    index = H.BruteForceIndex(k).index(Q.to(torch.float32))                # tf.convert_to_tensor(..., float32), :422
    q32 = P.to(torch.float32)
    vals, ids = [], []
    for a in range(0, P.shape[0], TOPK_BATCH_SIZE):
        v, i = index(q32[a:a + TOPK_BATCH_SIZE])
        vals.append(v); ids.append(i)
    vals = torch.cat(vals).cpu().numpy(); ids = torch.cat(ids).cpu().numpy()
    # the reference labels predictions with dense rows (user_id counter, item row) but probes sets of RAW ids
    # (get_idset, :410-416) -- kept: predictions carry dense ids exactly as there
    predictions = [(u, [(vals[u, j], int(ids[u, j])) for j in range(ids.shape[1])]) for u in range(ids.shape[0])]
    actual_user_ids, actual_item_ids = list(user_ids.keys()), list(item_ids.keys())
    return {"all_data": topk.topKMetrics(predictions, test_idset, actual_user_ids, actual_item_ids),
            "train_set": topk.topKMetrics(predictions, train_idset, actual_user_ids, actual_item_ids)}


def train_and_evaluate(dataset, user_ids, item_ids, uid_max, iid_max, global_bias, bmThread=None, epochs=None,
                       seed=None, evaluate=None, verbose=False):
    """train_and_evaluate (SVD.py:437-497) for the current held-out chunk of a digested RatingChunks dataset: random
    U(0,1)/d matrices, zero biases, EPOCHS passes of fit_model over the training chunks with the training / test MSE
    after each, then the test MSE and do_topk.  Returns the reference's dict {'all_data', 'train_set', 'mse'} plus
    'epoch_mse', 'epoch_test_mse' and 'parameters'.  bmThread (the reference's resource logger) is ignored."""
    train, test = dataset.train_frame(), dataset.test_frame()
    P, Q, bu, bi = init_parameters(uid_max + 1, iid_max + 1, NUMBER_OF_EMBEDDINGS, seed, train.device)
    epoch_mse, epoch_test_mse = [], []
    for e in range(1, (EPOCHS if epochs is None else epochs) + 1):
        fit_model(dataset, P, Q, bu, bi, global_bias, user_ids, item_ids)
        if e % EPOCH_ERROR_CALCULATION_FREQUENCY == 0:
            epoch_mse.append(mean_square_error(dataset, P, Q, bu, bi, global_bias, user_ids, item_ids))
            epoch_test_mse.append(mean_square_error([test], P, Q, bu, bi, global_bias, user_ids, item_ids))
            if verbose:
                print(f"::::EPOCH {e:=3}::::    MSE: {epoch_mse[-1]}", flush=True)
                print(f"             And on the test set MSE: {epoch_test_mse[-1]}")
    check_fit(train)
    result = {"epoch_mse": epoch_mse, "epoch_test_mse": epoch_test_mse, "parameters": (P, Q, bu, bi)}
    if EVALUATE if evaluate is None else evaluate:
        result["mse"] = mean_square_error([test], P, Q, bu, bi, global_bias, user_ids, item_ids)

        test_idset = get_idset([dataset.get_test_set()])            # sets of RAW (user, item) pairs (:478-479)
        train_idset = get_idset(list(dataset) + [dataset.get_test_set()])
        result.update(do_topk(P, Q, test_idset, train_idset, user_ids, item_ids))
    return result


def cross_validate(dataset, epochs=None, seed=None, verbose=False):
    """The reference's main program (SVD.py:519-566): digest every chunk, then hold each chunk out in turn (last one
    first), train_and_evaluate, and average the metric dicts with getAverage.  Returns {'all_data', 'train_set',
    'mse': [per fold], 'folds': [held-out chunk per fold]}.
    The folds are independent, which is the only parallelism this model offers across GPUs ("replicas only"): under
    torch.distributed fold f runs on rank f % world_size and the per-fold results are all-gathered."""
    dataset.use_no_test_set()
    user_ids, item_ids, uid_max, iid_max, global_bias = digest(dataset)
    folds = []
    while dataset.next_cross_validation_distribution():
        folds.append(dataset.test_set_index)
    mine = fold_assignment(len(folds), D.rank(), D.world_size())
    local = {}
    for f in mine:
        dataset.test_set_index = folds[f]
        if verbose:
            print("=" * 16 + f"\nCross validation {f + 1} of {len(folds)}\n" + "=" * 16)
        res = train_and_evaluate(dataset, user_ids, item_ids, uid_max, iid_max, global_bias, None, epochs=epochs,
                                 seed=None if seed is None else seed + f, verbose=verbose)
        local[f] = {"all_data": res["all_data"], "train_set": res["train_set"], "mse": res["mse"]}
    merged = merge_fold_results(local, len(folds))
    return {"all_data": topk.getAverage([merged[f]["all_data"] for f in range(len(folds))]),
            "train_set": topk.getAverage([merged[f]["train_set"] for f in range(len(folds))]),
            "mse": [merged[f]["mse"] for f in range(len(folds))], "folds": folds}


def fold_assignment(n_folds, rank, world):
    """Folds of this rank: f with f % world == rank (host logic, no device)."""
    return [f for f in range(n_folds) if f % world == rank]


def merge_fold_results(local, n_folds):
    """All ranks' {fold: result} dicts -> one dict holding every fold (all_gather_object under torch.distributed)."""
    if D.world_size() == 1:
        merged = dict(local)
    else:
        import torch.distributed as dist
        parts = [None] * D.world_size()
        dist.all_gather_object(parts, local)
        merged = {}
        for part in parts:
            merged.update(part)
    missing = [f for f in range(n_folds) if f not in merged]
    if missing:
        raise N.BrkError(f"cross_validate: folds {missing} were not run by any rank")
    return merged
