"""Multi-GPU plumbing: one process per GPU over torch.distributed (NCCL on the GPU box, gloo in the
CPU tests).  Replaces the reference's tf.distribute.experimental.MultiWorkerMirroredStrategy
(/root/reference/src/models/RModel.py:119-121; cluster from TF_CONFIG,
test/NeuMFModelWorker01.py:9): synchronous data parallelism.

Three pieces:
  * mirrored training (what the reference does): every rank holds the full tables, trains its slice
    of the global batch, the flat gradient arena is summed with ONE all-reduce per step and every rank
    applies the identical optimizer step.  The fused kernels scale gradients by 1/global_batch so
    the sum is the gradient of the global mean loss.  Right for tables that fit one GPU many times
    over (ML-1M shape: 2.5 MB).
  * row-sharded tables (owner = id mod G, local row = id div G) with an all-to-all of ids, rows and
    row gradients -- for tables too large to mirror (BASELINE.json configs[3]); the routing helpers
    here are pure index arithmetic and run on CPU or GPU tensors alike.
  * top-K over item-range shards: every rank scores all queries against its item range, the [U, k]
    lists are all-gathered and merged (score desc, id asc) -- identical to an unsharded scan.
"""
import os

import torch
import torch.distributed as dist


def is_dist():
    return dist.is_available() and dist.is_initialized()


def world_size():
    return dist.get_world_size() if is_dist() else 1


def rank():
    return dist.get_rank() if is_dist() else 0


def init_from_env(backend=None):
    """RANK / WORLD_SIZE / MASTER_ADDR / MASTER_PORT as torchrun sets them (the reference reads
    TF_CONFIG instead)."""
    if is_dist() or int(os.environ.get("WORLD_SIZE", "1")) == 1:
        return
    backend = backend or ("nccl" if torch.cuda.is_available() else "gloo")
    if backend == "nccl":
        local = int(os.environ.get("LOCAL_RANK", "0"))
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
    else:
        dist.init_process_group(backend)


def all_reduce_sum_(flat):
    """In-place sum over ranks of the flat gradient arena."""
    if is_dist() and world_size() > 1:
        dist.all_reduce(flat, op=dist.ReduceOp.SUM)
    return flat


def broadcast_(t, src=0):
    if is_dist() and world_size() > 1:
        dist.broadcast(t, src)
    return t


# ---- batch slicing ------------------------------------------------------------------------------------
def local_slice(n, r=None, w=None):
    """Contiguous slice [lo, hi) of n items owned by rank r of w (remainder to the low ranks)."""
    r = rank() if r is None else r
    w = world_size() if w is None else w
    base, rem = divmod(n, w)
    lo = r * base + min(r, rem)
    return lo, lo + base + (1 if r < rem else 0)


# ---- row-shard routing (owner = id mod G) -----------------------------------------------------------------
def owner_of(ids, G):
    return ids % G


def local_row(ids, G):
    return torch.div(ids, G, rounding_mode="floor")


def shard_rows(num_rows, r, G):
    """Number of rows rank r owns under owner = id mod G."""
    return (num_rows - r + G - 1) // G


def bucket_by_owner(ids, G):
    """Stable partition of `ids` by owner.  Returns (perm, counts): ids[perm] is grouped by owner
    0..G-1 (original order kept inside a bucket), counts[g] = bucket size.  inverse: out[perm] = x."""
    own = owner_of(ids.long(), G)
    perm = torch.sort(own, stable=True).indices
    counts = torch.bincount(own, minlength=G)
    return perm, counts


def exchange_counts(counts):
    """all-to-all of the per-peer counts: returns how many ids each peer will send to this rank."""
    out = torch.empty_like(counts)
    dist.all_to_all_single(out, counts)
    return out


def exchange_rows(send, send_counts, recv_counts):
    """Variable-size all-to-all of rows ([n, ...] tensors); counts are host lists."""
    recv = send.new_empty((int(sum(recv_counts)),) + tuple(send.shape[1:]))
    dist.all_to_all_single(recv, send, output_split_sizes=[int(c) for c in recv_counts],
                           input_split_sizes=[int(c) for c in send_counts])
    return recv


class ShardedLookup:
    """Routing of one batch of ids to the owners of a row-sharded table and back.
    forward : ids -> (all-to-all ids) -> owners gather rows -> (all-to-all rows) -> rows in batch order
    backward: row gradients in batch order -> (all-to-all) -> owners scatter-add into their shard.
    `gather` and `scatter_add` are callables so that the CPU tests can run the same routing with
    NumPy-style indexing while the GPU path passes the C-ABI kernels."""

    def __init__(self, ids, G):
        self.G = G
        self.perm, counts = bucket_by_owner(ids, G)
        self.send_counts = counts.tolist()
        self.recv_counts = exchange_counts(counts).tolist()
        sorted_ids = ids[self.perm]
        self.owner_local_ids = local_row(exchange_rows(sorted_ids.contiguous(), self.send_counts, self.recv_counts), G)

    def forward(self, gather):
        rows_for_peers = gather(self.owner_local_ids)                       # [n_recv, d] from my shard
        got = exchange_rows(rows_for_peers.contiguous(), self.recv_counts, self.send_counts)
        out = torch.empty_like(got)
        out[self.perm] = got                                                 # back to batch order
        return out

    def backward(self, grads, scatter_add):
        send = grads[self.perm].contiguous()
        got = exchange_rows(send, self.send_counts, self.recv_counts)
        scatter_add(self.owner_local_ids, got)


# ---- top-K over item-range shards ----------------------------------------------------------------------------
def gather_topk_parts(vals, ids):
    """all-gather of the per-shard [U, k] lists -> ([G, U, k], [G, U, k])."""
    if not is_dist() or world_size() == 1:
        return vals.unsqueeze(0), ids.unsqueeze(0)
    G = world_size()
    pv = [torch.empty_like(vals) for _ in range(G)]
    pi = [torch.empty_like(ids) for _ in range(G)]
    dist.all_gather(pv, vals.contiguous())
    dist.all_gather(pi, ids.contiguous())
    return torch.stack(pv), torch.stack(pi)


def sharded_topk(queries, item_vectors_local, item_lo, k, merge=None, local_topk=None):
    """Top-k of `queries` against an item catalog range-sharded over the ranks: this rank holds rows
    [item_lo, item_lo + len(item_vectors_local)).  Result identical on every rank and identical to
    the unsharded scan (same tie rule)."""
    from . import hotpath as H
    local_topk = local_topk or (lambda q, c, off: H.BruteForceIndex(k).index(c, id_offset=off)(q))
    merge = merge or H.topk_merge
    v, i = local_topk(queries, item_vectors_local, item_lo)
    pv, pi = gather_topk_parts(v, i)
    return merge(pv, pi)
