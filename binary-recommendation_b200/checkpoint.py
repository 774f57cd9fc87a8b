"""Checkpoint / serve handoff.

The reference saves one Keras SavedModel at `checkpointPath` after training (src/models/RModel.py:139,
getModelSaveLocation :175-179: only the chief's copy is kept, workers write to a temp dir that is deleted) and
reloads it with `tf.keras.models.load_model` for serving (`restoreFromLatestCheckPoint`, RModel.py:172-173, called by
the REST endpoint, src/restful/RecommendationEndpoint.py:19-23).  With mirrored variables every worker holds the
whole model, so "the chief saves" is enough there.  Here tables can be row-sharded over GPUs (sharded.py: row r
lives on rank r % G at local row r // G), so a checkpoint is a directory:

  manifest.json                      written last by rank 0 (atomic rename): format version, world size, entries.
                                     Re-saving into an existing directory first REMOVES the old manifest (and the shard
                                     files of other world sizes), so a crash in the middle of a save leaves a directory
                                     without manifest -- "incomplete", never a manifest naming a mix of old and new tensors
  <name>.bin                         replicated tensors (dense block, BatchNorm statistics, optimizer step), rank 0
  <name>.shard<r>-of-<G>.bin         row-sharded tensors (weights and Adam / Adagrad slots), one file per rank,
                                     raw little-endian arrays [local_rows, d]

and loading takes the *reader's* (rank, world): shards are re-interleaved through memory maps, so a run saved on
8 GPUs can be restored on 1 (serving) or on 4 without materialising the whole table on any host.
"""
import json
import os

import numpy as np

FORMAT = "brk-checkpoint"
VERSION = 1


class CheckpointError(RuntimeError):
    pass


def shard_rows(rows, world):
    return (rows + world - 1) // world


def _np(t):
    if isinstance(t, np.ndarray):
        return t
    return t.detach().cpu().numpy()


def _write(path, a):
    tmp = f"{path}.tmp{os.getpid()}"
    a = np.ascontiguousarray(a)
    a.astype(a.dtype.newbyteorder("<"), copy=False).tofile(tmp)
    os.replace(tmp, path)


def shard_file(name, rank, world):
    return f"{name}.shard{rank:03d}-of-{world:03d}.bin"


def save_checkpoint(dirpath, replicated=None, sharded=None, rank=0, world=1, meta=None, barrier=None):
    """replicated: {name: tensor} (identical on every rank; rank 0 writes them);
    sharded: {name: (local shard tensor [local_rows, d], global_rows)} with the r % G / r // G row layout.
    barrier: callable run between the shard writes and the manifest (torch.distributed.barrier under N > 1), so the
    manifest only ever names complete shard sets."""
    replicated, sharded = replicated or {}, sharded or {}
    for name in list(replicated) + list(sharded):
        if not name or name != os.path.basename(name) or name in (".", "..") or "\\" in name or name.startswith("manifest"):
            raise CheckpointError(f"tensor name {name!r} is not a plain file name")
    for name, (t, rows) in sharded.items():                 # validate everything before touching what is on disk
        shp = tuple(_np(t).shape) if not hasattr(t, "shape") else tuple(t.shape)
        if len(shp) != 2 or shp[0] != shard_rows(int(rows), world):
            raise CheckpointError(f"{name}: local shard has shape {shp}, expected [{shard_rows(int(rows), world)}, d]")
    os.makedirs(dirpath, exist_ok=True)
    if rank == 0:
        # invalidate what is there before the first byte of the new save is written
        old = os.path.join(dirpath, "manifest.json")
        if os.path.exists(old):
            os.remove(old)
        keep = {shard_file(n, q, world) for n in sharded for q in range(world)} | {n + ".bin" for n in replicated}
        for f in os.listdir(dirpath):
            if f.endswith(".bin") and f not in keep:
                os.remove(os.path.join(dirpath, f))
    if barrier is not None:
        barrier()
    entries = {}
    for name, (t, rows) in sharded.items():
        a = _np(t)
        if a.ndim != 2 or a.shape[0] != shard_rows(int(rows), world):
            raise CheckpointError(f"{name}: local shard has shape {a.shape}, expected [{shard_rows(int(rows), world)}, d]")
        _write(os.path.join(dirpath, shard_file(name, rank, world)), a)
        entries[name] = {"sharding": "row_mod", "dtype": a.dtype.str, "rows": int(rows), "d": int(a.shape[1])}
    if rank == 0:
        for name, t in replicated.items():
            a = _np(t)
            _write(os.path.join(dirpath, name + ".bin"), a)
            entries[name] = {"sharding": "replicated", "dtype": a.dtype.str, "shape": list(a.shape)}
    if barrier is not None:
        barrier()
    if rank == 0:
        manifest = {"format": FORMAT, "version": VERSION, "world": int(world), "entries": entries, "meta": meta or {}}
        tmp = os.path.join(dirpath, f"manifest.json.tmp{os.getpid()}")
        with open(tmp, "w") as f:
            json.dump(manifest, f, indent=1, sort_keys=True)
        os.replace(tmp, os.path.join(dirpath, "manifest.json"))
    if barrier is not None:
        barrier()
    return dirpath


def read_manifest(dirpath):
    p = os.path.join(dirpath, "manifest.json")
    if not os.path.exists(p):
        raise CheckpointError(f"{dirpath}: no manifest.json (incomplete or missing checkpoint)")
    with open(p) as f:
        m = json.load(f)
    if m.get("format") != FORMAT:
        raise CheckpointError(f"{dirpath}: not a {FORMAT} directory")
    if m.get("version") != VERSION:
        raise CheckpointError(f"{dirpath}: checkpoint version {m.get('version')}, this reader handles {VERSION}")
    return m


def load_rows(dirpath, name, entry, saved_world, global_rows):
    """Rows `global_rows` (int64 array of global row ids) of a sharded tensor, read through memory maps."""
    rows, d, dt = entry["rows"], entry["d"], np.dtype(entry["dtype"])
    g = np.asarray(global_rows, dtype=np.int64)
    if len(g) and (g.min() < 0 or g.max() >= rows):
        raise CheckpointError(f"{name}: row ids out of range [0, {rows})")
    out = np.zeros((len(g), d), dtype=dt)
    lr = shard_rows(rows, saved_world)
    for q in range(saved_world):
        sel = np.nonzero(g % saved_world == q)[0]
        if not len(sel):
            continue
        p = os.path.join(dirpath, shard_file(name, q, saved_world))
        if not os.path.exists(p) or os.path.getsize(p) != lr * d * dt.itemsize:
            raise CheckpointError(f"{p}: shard file missing or of the wrong size")
        mm = np.memmap(p, dtype=dt, mode="r", shape=(lr, d))
        out[sel] = mm[g[sel] // saved_world]
    return out


def load_checkpoint(dirpath, rank=0, world=1, names=None):
    """-> (replicated {name: ndarray}, sharded {name: local shard ndarray [shard_rows(rows, world), d] for THIS
    (rank, world), zero-padded past the last row}, meta).  The saved world size may differ from `world`."""
    m = read_manifest(dirpath)
    rep, shd = {}, {}
    for name, e in m["entries"].items():
        if names is not None and name not in names:
            continue
        if name != os.path.basename(name) or name in (".", ".."):
            raise CheckpointError(f"{dirpath}: manifest entry {name!r} is not a plain file name")
        if e["sharding"] == "replicated":
            p = os.path.join(dirpath, name + ".bin")
            dt = np.dtype(e["dtype"])
            n = int(np.prod(e["shape"])) if e["shape"] else 1
            if not os.path.exists(p) or os.path.getsize(p) != n * dt.itemsize:
                raise CheckpointError(f"{p}: missing or of the wrong size")
            rep[name] = np.fromfile(p, dtype=dt).reshape(e["shape"])
        elif e["sharding"] == "row_mod":
            lr = shard_rows(e["rows"], world)
            mine = np.arange(rank, e["rows"], world, dtype=np.int64)         # global rows this reader owns
            local = np.zeros((lr, e["d"]), dtype=np.dtype(e["dtype"]))
            local[:len(mine)] = load_rows(dirpath, name, e, m["world"], mine)
            shd[name] = local
        else:
            raise CheckpointError(f"{name}: unknown sharding {e['sharding']!r}")
    return rep, shd, m.get("meta", {})


# ---- model-level helpers ------------------------------------------------------------------------------------
def save_state_dict(dirpath, sd, meta=None):
    """An unsharded model's state_dict (RModel.saveCheckPoint): every tensor replicated."""
    return save_checkpoint(dirpath, replicated=sd, meta=meta)


def load_state_dict(dirpath):
    """Counterpart of save_state_dict; sharded entries (a checkpoint written by a sharded run) come back as whole
    tables -- the 8-GPU-train / 1-GPU-serve handoff."""
    import torch
    m = read_manifest(dirpath)
    rep, shd, _ = load_checkpoint(dirpath, 0, 1)
    sd = {k: torch.from_numpy(np.ascontiguousarray(v)) for k, v in rep.items()}
    for k, v in shd.items():
        sd[k] = torch.from_numpy(np.ascontiguousarray(v[:m["entries"][k]["rows"]]))
    return sd
