"""Checkpoint directories (binary-recommendation_b200/checkpoint.py): the row layout
(row r on rank r % G at local row r // G) must survive any change of world size bit for bit.  CPU only: NumPy shards,
plus a world-size-2 gloo run in which each rank writes its own shard files."""
import json
import os

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from binrec_b200 import checkpoint as CK


def _shards(full, G):
    lr = CK.shard_rows(len(full), G)
    out = []
    for r in range(G):
        part = np.zeros((lr, full.shape[1]), dtype=full.dtype)
        part[:len(full[r::G])] = full[r::G]
        out.append(part)
    return out


@pytest.mark.parametrize("rows", [1, 7, 64, 1001])
@pytest.mark.parametrize("G_save,G_load", [(1, 1), (1, 3), (2, 1), (3, 2), (8, 1), (8, 5), (4, 8)])
def test_reshard_round_trip(tmp_path, rows, G_save, G_load):
    rng = np.random.default_rng(rows * 31 + G_save)
    full = rng.standard_normal((rows, 6)).astype(np.float32)
    slot = rng.standard_normal((rows, 6)).astype(np.float32)
    dense = rng.standard_normal(37).astype(np.float32)
    state = np.array([5, 123, 456], dtype=np.int64)
    d = str(tmp_path / "cp")
    for r in range(G_save - 1, -1, -1):
        CK.save_checkpoint(d, replicated={"dense": dense, "opt_state": state} if r == 0 else None,
                           sharded={"user": (_shards(full, G_save)[r], rows), "user_m": (_shards(slot, G_save)[r], rows)},
                           rank=r, world=G_save, meta={"step": 5})
    want_w, want_m = _shards(full, G_load), _shards(slot, G_load)
    for r in range(G_load):
        rep, shd, meta = CK.load_checkpoint(d, r, G_load)
        assert np.array_equal(shd["user"], want_w[r]) and np.array_equal(shd["user_m"], want_m[r])
        assert np.array_equal(rep["dense"], dense) and np.array_equal(rep["opt_state"], state) and rep["opt_state"].dtype == np.int64
        assert meta == {"step": 5}
    sd = CK.load_state_dict(d)
    assert np.array_equal(sd["user"].numpy(), full) and np.array_equal(sd["user_m"].numpy(), slot)


def test_load_rows_reads_only_what_is_asked(tmp_path):
    full = np.arange(40, dtype=np.float32).reshape(10, 4)
    d = str(tmp_path / "cp")
    for r in (2, 1, 0):
        CK.save_checkpoint(d, sharded={"item": (_shards(full, 3)[r], 10)}, rank=r, world=3)
    m = CK.read_manifest(d)
    got = CK.load_rows(d, "item", m["entries"]["item"], m["world"], [9, 0, 4, 4])
    assert np.array_equal(got, full[[9, 0, 4, 4]])
    with pytest.raises(CK.CheckpointError):
        CK.load_rows(d, "item", m["entries"]["item"], m["world"], [10])


def test_damaged_and_incomplete_checkpoints_fail_loudly(tmp_path):
    d = str(tmp_path / "cp")
    full = np.ones((9, 2), dtype=np.float32)
    with pytest.raises(CK.CheckpointError, match="manifest"):
        CK.load_checkpoint(d)
    CK.save_checkpoint(d, sharded={"user": (_shards(full, 2)[1], 9)}, rank=1, world=2)      # rank 0 never finished
    with pytest.raises(CK.CheckpointError, match="manifest"):
        CK.load_checkpoint(d)
    CK.save_checkpoint(d, replicated={"x": np.zeros(3)}, sharded={"user": (_shards(full, 2)[0], 9)}, rank=0, world=2)
    CK.load_checkpoint(d)
    os.remove(os.path.join(d, CK.shard_file("user", 1, 2)))
    with pytest.raises(CK.CheckpointError, match="shard file"):
        CK.load_checkpoint(d)
    with pytest.raises(CK.CheckpointError, match="local shard"):
        CK.save_checkpoint(d, sharded={"user": (np.ones((3, 2), np.float32), 9)}, rank=0, world=2)
    m = json.load(open(os.path.join(d, "manifest.json"))); m["version"] = 99
    json.dump(m, open(os.path.join(d, "manifest.json"), "w"))
    with pytest.raises(CK.CheckpointError, match="version"):
        CK.load_checkpoint(d)


def _worker(rank, world, port, d):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    full = np.arange(26 * 3, dtype=np.float32).reshape(26, 3)
    CK.save_checkpoint(d, replicated={"dense": np.arange(5, dtype=np.float32)},
                       sharded={"user": (torch.from_numpy(_shards(full, world)[rank]), 26)},
                       rank=rank, world=world, barrier=dist.barrier)
    rep, shd, _ = CK.load_checkpoint(d, rank, world)           # every rank reads back its own shard after the barrier
    assert np.array_equal(shd["user"], _shards(full, world)[rank])
    dist.destroy_process_group()


def test_two_ranks_write_one_checkpoint_gloo(tmp_path):
    d = str(tmp_path / "cp")
    port = 29500 + os.getpid() % 2000
    mp.spawn(_worker, args=(2, port, d), nprocs=2, join=True)
    full = np.arange(26 * 3, dtype=np.float32).reshape(26, 3)
    assert np.array_equal(CK.load_state_dict(d)["user"].numpy(), full)
    for r in range(3):
        assert np.array_equal(CK.load_checkpoint(d, r, 3)[1]["user"], _shards(full, 3)[r])


def test_chief_and_worker_save_locations_follow_the_reference(tmp_path):
    """getModelSaveLocation / isMaster / getSlaveTempDir / clearSlaveTempDir (src/models/RModel.py:175-196)."""
    import os
    import types
    from binrec_b200.RModel import RModel
    m = RModel("NeuMFModel", workDir=str(tmp_path))
    strat = lambda t, i: types.SimpleNamespace(cluster_resolver=types.SimpleNamespace(task_type=t, task_id=i))
    assert m.isMaster(None, 0) and m.isMaster("chief", 3) and m.isMaster("worker", 0) and not m.isMaster("worker", 1)
    assert not m.isMaster("ps", 0)
    assert m.getModelSaveLocation(None) == m.checkpointPath
    assert m.getModelSaveLocation(strat("chief", 0)) == m.checkpointPath
    assert m.getModelSaveLocation(strat("worker", 0)) == m.checkpointPath
    loc = m.getModelSaveLocation(strat("worker", 2))
    assert loc == os.path.join(m.checkpointPath, "workertemp_2") and os.path.isdir(loc)
    m.clearSlaveTempDir(strat("worker", 0))                                 # the master clears nothing
    assert os.path.isdir(loc)
    m.clearSlaveTempDir(strat("worker", 2))
    assert not os.path.exists(loc) and os.path.isdir(m.checkpointPath)
    assert m.getNumberOfWorkers({"cluster": {"worker": ["a:1", "b:2"]}}) == 2 and m.getNumberOfWorkers(None) == 1


def test_rmodel_plot_writes_the_history_series(tmp_path):
    """RModel.plot (RModel.py:100-113): loss / val_loss / the requested metrics per epoch, written beside the checkpoint."""
    import json
    from binrec_b200.RModel import RModel
    m = RModel("NeuMFModel", workDir=str(tmp_path))
    hist = {"loss": [0.5, 0.4], "val_loss": [0.6, 0.5], "mse": [0.3, 0.2], "unused": [1, 2]}
    out = m.plot(hist, [("mse", "MSE")])
    assert json.load(open(out)) == {"loss": [0.5, 0.4], "val_loss": [0.6, 0.5], "mse": [0.3, 0.2]}
    assert m.getPredictDataFrame(3) is None and m.predictForUser(3) is None and m.getPredictableUsers() == []


def test_resave_invalidates_the_old_checkpoint_first_and_rejects_path_names(tmp_path, monkeypatch):
    """A crash in the middle of a re-save must leave an INCOMPLETE directory, never a manifest naming a mix of old and
    new tensors; shards of an earlier world size are removed; tensor names are plain file names."""
    d = str(tmp_path / "cp")
    full = np.arange(18, dtype=np.float32).reshape(9, 2)
    for r in (1, 0):
        CK.save_checkpoint(d, replicated={"x": np.zeros(3)}, sharded={"user": (_shards(full, 2)[r], 9)}, rank=r, world=2)
    CK.load_checkpoint(d)
    calls = {"n": 0}
    real = CK._write

    def dying_write(path, a):
        calls["n"] += 1
        if calls["n"] == 2:
            raise OSError("disk full")
        real(path, a)

    monkeypatch.setattr(CK, "_write", dying_write)
    with pytest.raises(OSError):
        CK.save_checkpoint(d, replicated={"x": np.ones(3), "y": np.ones(2)}, sharded={"user": (full + 1, 9)}, rank=0, world=1)
    monkeypatch.setattr(CK, "_write", real)
    with pytest.raises(CK.CheckpointError, match="manifest"):
        CK.load_checkpoint(d)                                   # incomplete, not half old / half new
    CK.save_checkpoint(d, replicated={"x": np.ones(3)}, sharded={"user": (full + 1, 9)}, rank=0, world=1)
    rep, shd, _ = CK.load_checkpoint(d)
    assert np.array_equal(shd["user"], full + 1) and np.array_equal(rep["x"], np.ones(3))
    assert not any("-of-002" in f for f in os.listdir(d))       # the 2-rank shards are gone
    for bad in ("../evil", "a/b", "", "..", "manifest.json"):
        with pytest.raises(CK.CheckpointError, match="plain file name"):
            CK.save_checkpoint(d, replicated={bad: np.zeros(1)})
    CK.load_checkpoint(d)                                       # the rejected saves touched nothing
