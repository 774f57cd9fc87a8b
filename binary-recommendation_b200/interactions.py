"""Interaction files and the binary columnar cache: the data formats on the
caller's side of the hot path.

The reference re-parses a CSV with pandas on every run and rebuilds its id vocabularies with `pd.unique`:
  * "neumf"    CUSTOMER_ID,PRODUCT_ID,MATERIAL,QUANTITY, integer ids, table sizes = max id + 1
               (src/models/NeuMFModel.py:21-27; src/generator/NegativeDataSetGenerator.py:56-57);
  * "ncf"      customer_id,normalized_customer_id,material,product_id,rating_type, header row replaced
               (trainers/NFC_plain.py:72-73; src/models/NCFModel.py:61-64);
  * "twotower" CUSTOMER_ID,MATERIAL or CUSTOMER_ID,NORMALIZED_CUSTOMER_ID,MATERIAL,PRODUCT_ID,RATING_TYPE with string
               ids, first row dropped (trainers/loadBinaryMovieLens.py:41-62);
  * "ml-100k"  tab-separated user_id, movie_id, rating, unix_timestamp; ids become strings
               (trainers/loadBinaryMovieLens.py:8-21).
Here a file is parsed once (pandas' C parser on the host: text parsing is not the hot path), its id columns are
factorised on the device in first-occurrence order (pipeline.Vocabulary, bit-equal to `pd.unique`), and the
result is written as a **binary columnar cache** that later runs map straight into pinned memory.

Cache layout (".brkc", little-endian):
  bytes 0..7    magic  b"BRKCOL1\\0"
  u32 version (1) | u32 n_columns | u64 n_rows | u64 meta_bytes
  meta_bytes of UTF-8 JSON: {"columns": [{"name", "dtype", "offset", "nbytes", "role"}...],
                             "vocab":   [{"name", "kind" ("i"|"S"), "dtype", "offset", "nbytes", "size"}...],
                             "attrs": {...}}
  column / vocabulary blobs, each starting on a 4096-byte boundary (offsets are absolute file offsets)
Columns are raw arrays (int32 dense ids, float32 values); a vocabulary blob holds the 64-bit keys of an id column
in id order (ids index it directly; StringLookup adds its offset of 2 at lookup time).
"""
import json
import os
import struct

import numpy as np

MAGIC = b"BRKCOL1\0"
VERSION = 1
ALIGN = 4096
HEADER = struct.Struct("<8sIIQQ")

SCHEMAS = {
    # name: (read_csv kwargs, user column, item column, value column or None, ids are strings?)
    "neumf": (dict(), "CUSTOMER_ID", "PRODUCT_ID", None, False),
    "ncf": (dict(header=0, names=["customer_id", "normalized_customer_id", "material", "product_id", "rating_type"]),
            "normalized_customer_id", "product_id", "rating_type", False),
    "twotower": (dict(header=0, names=["CUSTOMER_ID", "MATERIAL"], dtype={"MATERIAL": str, "CUSTOMER_ID": str}),
                 "CUSTOMER_ID", "MATERIAL", None, True),
    "twotower-rdzero": (dict(header=0, names=["CUSTOMER_ID", "NORMALIZED_CUSTOMER_ID", "MATERIAL", "PRODUCT_ID",
                                              "RATING_TYPE"], dtype={"MATERIAL": str, "CUSTOMER_ID": str}),
                        "CUSTOMER_ID", "MATERIAL", "RATING_TYPE", True),
    "ml-100k": (dict(sep="\t", names=["user_id", "movie_id", "rating", "unix_timestamp"], encoding="latin-1",
                     dtype={"user_id": str, "movie_id": str}), "user_id", "movie_id", "rating", True),
}


class InteractionError(ValueError):
    pass


def read_csv_columns(path_or_file, schema="neumf", rowLimit=None):
    """Host parse of one of the reference's CSV schemas -> {"user": column, "item": column[, "value": float32]}
    (raw ids, not yet dense)."""
    if schema not in SCHEMAS:
        raise InteractionError(f"unknown schema {schema!r}; one of {sorted(SCHEMAS)}")
    import pandas as pd
    kw, ucol, icol, vcol, _ = SCHEMAS[schema]
    df = pd.read_csv(path_or_file, nrows=rowLimit, **kw)
    for c in (ucol, icol) + ((vcol,) if vcol else ()):
        if c not in df.columns:
            raise InteractionError(f"column {c!r} missing (schema {schema!r} has {list(df.columns)})")
    out = {"user": df[ucol].to_numpy(), "item": df[icol].to_numpy()}
    if vcol:
        out["value"] = df[vcol].to_numpy(dtype=np.float32)
    return out


def _pad_to(f, align=ALIGN):
    pos = f.tell()
    pad = (-pos) % align
    if pad:
        f.write(b"\0" * pad)
    return pos + pad


def write_cache(path, columns, vocab=None, attrs=None, roles=None):
    """columns: {name: 1-D ndarray (int32 / int64 / float32 / float64)}; vocab: {name: (keys uint64 ndarray, kind)}.
    Written to a temporary file and renamed, so a reader never sees a partial cache."""
    names = list(columns)
    if not names:
        raise InteractionError("a cache needs at least one column")
    n_rows = len(columns[names[0]])
    arrays = {}
    for k in names:
        a = np.ascontiguousarray(columns[k])
        if a.ndim != 1 or len(a) != n_rows:
            raise InteractionError(f"column {k!r}: need a 1-D array of {n_rows} rows")
        if a.dtype.kind not in "iuf":
            raise InteractionError(f"column {k!r}: dtype {a.dtype} is not numeric (factorise id columns first)")
        arrays[k] = a.astype(a.dtype.newbyteorder("<"), copy=False)
    vocab = vocab or {}
    # two passes: the meta block's size decides the blob offsets, which the meta block records
    def layout(meta_bytes):
        off = HEADER.size + meta_bytes
        cols, vocs = [], []
        for k in names:
            off += (-off) % ALIGN
            cols.append({"name": k, "dtype": arrays[k].dtype.str, "offset": off, "nbytes": arrays[k].nbytes,
                         "role": (roles or {}).get(k, "id" if arrays[k].dtype.kind in "iu" else "value")})
            off += arrays[k].nbytes
        for k, (keys, kind) in vocab.items():
            keys = np.ascontiguousarray(keys, dtype="<u8")
            off += (-off) % ALIGN
            vocs.append({"name": k, "kind": kind, "dtype": "<u8", "offset": off, "nbytes": keys.nbytes, "size": len(keys)})
            off += keys.nbytes
        return json.dumps({"columns": cols, "vocab": vocs, "attrs": attrs or {}}).encode()
    meta = layout(0)
    meta_bytes = len(meta) + 64                 # slack: offsets can gain digits between the passes
    meta = layout(meta_bytes)
    if len(meta) > meta_bytes:
        raise InteractionError("cache meta block overflow")
    meta = meta + b" " * (meta_bytes - len(meta))
    tmp = f"{path}.tmp{os.getpid()}"
    with open(tmp, "wb") as f:
        f.write(HEADER.pack(MAGIC, VERSION, len(names), n_rows, meta_bytes))
        f.write(meta)
        for k in names:
            _pad_to(f)
            f.write(arrays[k].tobytes())
        for k, (keys, kind) in vocab.items():
            _pad_to(f)
            f.write(np.ascontiguousarray(keys, dtype="<u8").tobytes())
    os.replace(tmp, path)
    return path


class InteractionCache:
    """A cache opened read-only: columns are np.memmap views (no copy until they are staged)."""

    def __init__(self, path):
        self.path = path
        size = os.path.getsize(path)
        with open(path, "rb") as f:
            head = f.read(HEADER.size)
            if len(head) < HEADER.size:
                raise InteractionError(f"{path}: truncated header")
            magic, version, n_cols, n_rows, meta_bytes = HEADER.unpack(head)
            if magic != MAGIC:
                raise InteractionError(f"{path}: not a brk columnar cache (bad magic)")
            if version != VERSION:
                raise InteractionError(f"{path}: cache version {version}, this reader handles {VERSION}")
            try:
                meta = json.loads(f.read(meta_bytes).decode())
            except Exception as e:
                raise InteractionError(f"{path}: unreadable meta block ({e})")
        if len(meta["columns"]) != n_cols:
            raise InteractionError(f"{path}: header says {n_cols} columns, meta lists {len(meta['columns'])}")
        self.n_rows = int(n_rows)
        self.attrs = meta.get("attrs", {})
        self.columns, self.roles, self.vocab = {}, {}, {}
        for c in meta["columns"]:
            dt = np.dtype(c["dtype"])
            if c["offset"] + c["nbytes"] > size or c["nbytes"] != self.n_rows * dt.itemsize:
                raise InteractionError(f"{path}: column {c['name']!r} is truncated or mis-sized")
            self.columns[c["name"]] = np.memmap(path, dtype=dt, mode="r", offset=c["offset"], shape=(self.n_rows,)) \
                if self.n_rows else np.zeros(0, dtype=dt)
            self.roles[c["name"]] = c.get("role", "id")
        for v in meta.get("vocab", []):
            if v["offset"] + v["nbytes"] > size:
                raise InteractionError(f"{path}: vocabulary {v['name']!r} is truncated")
            keys = np.memmap(path, dtype="<u8", mode="r", offset=v["offset"], shape=(v["size"],)) if v["size"] else \
                np.zeros(0, dtype="<u8")
            self.vocab[v["name"]] = (keys, v["kind"])

    def __len__(self):
        return self.n_rows

    def vocabulary(self, name):
        """The raw ids of column `name` in id order (what `pd.unique` returned): int64 array or list of str."""
        from . import pipeline as PL
        keys, kind = self.vocab[name]
        return PL.unpack_keys(np.asarray(keys), kind)

    def to_device(self, names=None, device=None, rowLimit=None):
        """Stages columns through pinned memory to the device (one async copy per column)."""
        import torch
        dev = torch.device(device) if device is not None else torch.device(f"cuda:{torch.cuda.current_device()}")
        out = {}
        for k in (names or list(self.columns)):
            a = self.columns[k][:rowLimit]
            t = torch.from_numpy(np.array(a))                     # copy out of the read-only memory map
            out[k] = t.pin_memory().to(dev, non_blocking=True) if t.numel() else t.to(dev)
        return out


def factorize_on_device(values, device=None, offset=0):
    """(ids int32 device tensor, keys uint64 ndarray in id order, kind): dense ids in first-occurrence order --
    `pd.unique` + positional index (loadBinaryMovieLens.py:16-19,58-61) computed by the device hash table."""
    from . import pipeline as PL
    a = np.asarray(values)
    kind = "i" if a.dtype.kind in "iu" else "S"
    v = PL.Vocabulary(device)
    ids = v.build(a, offset=offset)
    return ids, v.keys.cpu().numpy().view(np.uint64), kind, v


def build_cache(path, columns, device=None, attrs=None, factorize=("user", "item")):
    """Factorises the id columns named in `factorize` on the device and writes the cache.  Returns the opened cache."""
    cols, vocab = {}, {}
    for k, a in columns.items():
        if k in factorize:
            ids, keys, kind, _ = factorize_on_device(a, device)
            cols[k] = ids.cpu().numpy()
            vocab[k] = (keys, kind)
        else:
            cols[k] = np.asarray(a)
    at = dict(attrs or {})
    for k in factorize:
        if k in vocab:
            at[f"num_{k}"] = int(len(vocab[k][0]))
    write_cache(path, cols, vocab, at)
    return InteractionCache(path)


def csv_to_cache(csv_path, cache_path, schema="neumf", rowLimit=None, device=None, dense_ids=None):
    """CSV in one of the reference's schemas -> cache.  dense_ids=False keeps integer ids as they are (the
    "neumf" rule: table sizes are max id + 1, NeuMFModel.py:25-26); True factorises them (the two-tower rule)."""
    cols = read_csv_columns(csv_path, schema, rowLimit)
    strings = SCHEMAS[schema][4]
    dense = strings if dense_ids is None else dense_ids
    attrs = {"schema": schema, "source": os.path.basename(str(csv_path))}
    if dense:
        return build_cache(cache_path, cols, device=device, attrs=attrs)
    for k in ("user", "item"):
        a = cols[k]
        if a.dtype.kind not in "iu" or (len(a) and (a.min() < 0 or a.max() >= 2 ** 31)):
            raise InteractionError(f"column {k!r}: raw ids must be non-negative int32 values (or pass dense_ids=True)")
        cols[k] = a.astype(np.int32)
        attrs[f"num_{k}"] = int(a.max()) + 1 if len(a) else 0
    write_cache(cache_path, cols, None, attrs)
    return InteractionCache(cache_path)
