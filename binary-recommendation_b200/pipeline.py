"""Device input pipeline: the caller's side of the training step.

  * epoch_permutation / epoch_permutation_host -- the seeded stand-in for the reference's unseeded shuffles
    (src/models/NeuMFModel.py:109,117-121; trainers/twoTower.py:197): "brk perm v1", a keyed bijection.
  * neumf_epoch_build -- bootstrapDataset's frame (NeuMFModel.py:102-109) written by one kernel.
  * Vocabulary -- `pd.unique` order ids / StringLookup (trainers/loadBinaryMovieLens.py:16-19,58-61,
    trainers/twoTower.py:33-36) by a device hash table.

Everything here forwards to libbrk_b200.so (csrc/pipeline.cu); there is no host fallback for the device calls.
"""
import numpy as np
import torch

from . import _native as N

SALT_ROWS = 0
SALT_BATCHES = 1
U32 = 0xFFFFFFFF
RESERVED_KEY = 0xFFFFFFFFFFFFFFFF


def epoch_permutation(n, seed, epoch, salt=SALT_ROWS, first=0, count=None, device=None, out=None):
    """int64 device tensor: perm(first .. first+count) of the keyed bijection of [0, n)."""
    count = n - first if count is None else count
    dev = torch.device(device) if device is not None else torch.device(f"cuda:{torch.cuda.current_device()}")
    if out is None:
        out = torch.empty(count, dtype=torch.int64, device=dev)
    N.check(N.lib().brk_epoch_permutation(N.ctx(dev), n, first, count, seed & U32, epoch & U32, salt & U32,
                                          N.ptr(out), N.stream_ptr()), "brk_epoch_permutation")
    return out


def epoch_permutation_host(n, seed, epoch, salt=SALT_BATCHES, first=0, count=None):
    """The same permutation evaluated by the library on the host (batch orders of host-driven loops)."""
    count = n - first if count is None else count
    out = np.empty(count, dtype=np.int64)
    N.check(N.lib().brk_epoch_permutation_host(n, first, count, seed & U32, epoch & U32, salt & U32,
                                               out.ctypes.data), "brk_epoch_permutation_host")
    return out


def neumf_epoch_build(pos_users, pos_items, n_neg, seed, epoch, reject=False, csr_indptr=None, csr_items=None,
                      first=0, count=None, out=None):
    """(users int32, items int32, labels float32) -- rows [first, first+count) of the shuffled frame of
    positives (label 1) and n_neg Philox negatives (label 0).  reject=True re-draws negatives that are known
    positives (needs the per-user sorted positive lists, synth.build_csr)."""
    if pos_users.dtype != torch.int32 or pos_items.dtype != torch.int32:
        raise TypeError("pos_users / pos_items must be int32")
    pos_users = pos_users.contiguous(); pos_items = pos_items.contiguous()
    P = pos_users.numel()
    if pos_items.numel() != P or P == 0:
        raise ValueError("need equally long, non-empty positive columns")
    n = P + int(n_neg)
    count = n - first if count is None else count
    dev = pos_users.device
    if reject:
        if csr_indptr is None or csr_items is None:
            raise ValueError("reject=True needs csr_indptr / csr_items")
        if csr_indptr.dtype != torch.int64 or csr_items.dtype != torch.int32:
            raise TypeError("csr_indptr must be int64 and csr_items int32")
    if out is None:
        out = (torch.empty(count, dtype=torch.int32, device=dev), torch.empty(count, dtype=torch.int32, device=dev),
               torch.empty(count, dtype=torch.float32, device=dev))
    u, i, y = out
    N.check(N.lib().brk_neumf_epoch_build(N.ctx(dev), N.ptr(pos_users), N.ptr(pos_items), P, int(n_neg), first, count,
                                          seed & U32, epoch & U32, 1 if reject else 0,
                                          N.ptr(csr_indptr if reject else None), N.ptr(csr_items if reject else None),
                                          N.ptr(u), N.ptr(i), N.ptr(y), N.stream_ptr()), "brk_neumf_epoch_build")
    return u, i, y


def pack_keys(values):
    """Exact 64-bit keys (NumPy uint64) for an id column: integers as they are (two's complement), byte / unicode
    strings of at most 8 bytes packed big-endian.  Longer strings cannot be keyed exactly and raise."""
    a = np.asarray(values)
    if a.dtype.kind in "iu":
        keys = a.astype(np.int64).view(np.uint64) if a.dtype.kind == "i" else a.astype(np.uint64)
    elif a.dtype.kind in "USO":
        b = a.astype("S") if a.dtype.kind != "S" else a
        if b.dtype.itemsize > 8:
            raise ValueError(f"string ids longer than 8 bytes ({b.dtype.itemsize}) have no exact 64-bit key")
        w = np.zeros((len(b), 8), dtype=np.uint8)
        raw = np.frombuffer(b.tobytes(), dtype=np.uint8).reshape(len(b), b.dtype.itemsize) if len(b) else \
            np.zeros((0, b.dtype.itemsize), dtype=np.uint8)
        # numpy pads 'S' on the right with NULs: right-align the used bytes (big-endian integer of the string)
        lens = (raw != 0).sum(axis=1)
        if len(b) and not ((raw != 0) == (np.arange(b.dtype.itemsize)[None, :] < lens[:, None])).all():
            raise ValueError("string ids must not contain NUL bytes")
        for L in np.unique(lens):
            rows = np.nonzero(lens == L)[0]
            w[rows, 8 - L:] = raw[rows, :L]
        keys = w.view(">u8").reshape(-1).astype(np.uint64)
    else:
        raise TypeError(f"unsupported id dtype {a.dtype}")
    if len(keys) and (keys == np.uint64(RESERVED_KEY)).any():
        raise ValueError("key 0xFFFFFFFFFFFFFFFF (int64 -1) is reserved")
    return np.ascontiguousarray(keys)


def unpack_keys(keys, kind):
    """Inverse of pack_keys for vocabularies: kind 'i' -> int64 array, 'S' -> list of str."""
    keys = np.asarray(keys, dtype=np.uint64)
    if kind == "i":
        return keys.view(np.int64)
    return [int(k).to_bytes(8, "big").lstrip(b"\0").decode() for k in keys]


class Vocabulary:
    """Device-resident key -> id table.  build() assigns ids in order of first appearance (pd.unique);
    lookup() maps further keys, unknown ones to `oov` (StringLookup: offset 2, oov 1)."""

    def __init__(self, device=None):
        self.device = torch.device(device) if device is not None else torch.device(f"cuda:{torch.cuda.current_device()}")
        self.table_keys = self.table_vals = None
        self.capacity = 0
        self.offset = 0
        self.size = 0
        self.keys = None          # uint64 bit patterns as int64 device tensor [size], first-occurrence order

    @staticmethod
    def _dev_keys(keys, dev):
        if isinstance(keys, torch.Tensor):
            if keys.dtype != torch.int64:
                raise TypeError("device keys must be int64 (the 64-bit key pattern)")
            if keys.numel() and bool((keys == -1).any().item()):      # the empty-slot marker of the hash table
                raise ValueError("key 0xFFFFFFFFFFFFFFFF (int64 -1) is reserved")
            return keys.contiguous()
        return torch.from_numpy(pack_keys(keys).view(np.int64)).to(dev)

    def build(self, keys, offset=0):
        """keys: int64 device tensor (bit patterns) or a host id column.  Returns int32 ids (device)."""
        k = self._dev_keys(keys, self.device)
        n = k.numel()
        lib = N.lib()
        self.capacity = int(lib.brk_vocab_capacity(n))
        self.offset = int(offset)
        self.table_keys = torch.empty(self.capacity, dtype=torch.int64, device=self.device)
        self.table_vals = torch.empty(self.capacity, dtype=torch.int32, device=self.device)
        ids = torch.empty(n, dtype=torch.int32, device=self.device)
        vocab = torch.empty(max(n, 1), dtype=torch.int64, device=self.device)
        n_unique = torch.zeros(1, dtype=torch.int64, device=self.device)
        ws = torch.empty(max(int(lib.brk_vocab_workspace_bytes(n)), 256), dtype=torch.uint8, device=self.device)
        N.check(lib.brk_vocab_build_u64(N.ctx(self.device), N.ptr(k), n, self.offset, N.ptr(self.table_keys),
                                        N.ptr(self.table_vals), self.capacity, N.ptr(ids), N.ptr(vocab),
                                        N.ptr(n_unique), N.ptr(ws), N.stream_ptr()), "brk_vocab_build_u64")
        self.size = int(n_unique.item())
        self.keys = vocab[:self.size].clone()
        return ids

    def lookup(self, keys, oov=1):
        if self.table_keys is None:
            raise N.BrkError("Vocabulary.lookup before build")
        k = self._dev_keys(keys, self.device)
        ids = torch.empty(k.numel(), dtype=torch.int32, device=self.device)
        N.check(N.lib().brk_vocab_lookup_u64(N.ctx(self.device), N.ptr(k), k.numel(), N.ptr(self.table_keys),
                                             N.ptr(self.table_vals), self.capacity, int(oov), N.ptr(ids),
                                             N.stream_ptr()), "brk_vocab_lookup_u64")
        return ids

    def host_keys(self, kind="i"):
        return unpack_keys(self.keys.cpu().numpy().view(np.uint64), kind)
