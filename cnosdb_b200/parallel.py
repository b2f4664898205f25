"""Multi-GPU plumbing: one process per GPU, series ids sharded across ranks, one exchange step — the
element-wise all-reduce of the per-bucket partials. torch.distributed only
moves the (tiny) partial sections; decode / filter / reduce never leave the rank.

The reference shards series the same way (hash(SeriesKey) % n_shards, common/models/src/meta_data.rs:81-85)
and merges per-partition partial aggregates in DataFusion's final AggregateExec.
"""
import numpy as np
import torch
import torch.distributed as dist

KEY_FIRST_IDENTITY = 0x7FFFFFFFFFFFFFFF
KEY_LAST_IDENTITY = -0x8000000000000000


class _DevArray:
    """Zero-copy view of library-owned device memory for torch (CUDA array interface v3)."""

    def __init__(self, ptr, n, typestr):
        self.__cuda_array_interface__ = {"shape": (int(n),), "typestr": typestr, "data": (int(ptr), False),
                                         "version": 3, "strides": None}


def device_tensor(ptr, n, dtype, device):
    if n == 0:
        return torch.empty(0, dtype=dtype, device=device)
    typestr = {torch.int64: "<i8", torch.float64: "<f8", torch.uint8: "|u1"}[dtype]
    return torch.as_tensor(_DevArray(ptr, n, typestr), device=device)


def shard_range(n_series, rank, world_size):
    """Contiguous series-id range [lo, hi) of a rank, like the reference's per-task series chunks
    (tskv/src/reader/iterator.rs:232-235). (BASELINE's `id % N` would put all i64 series of the mixed C4
    workload on the even ranks and all f64 series on the odd ones.)"""
    per = (n_series + world_size - 1) // world_size
    return min(n_series, rank * per), min(n_series, (rank + 1) * per)


def shard_of(series_id, n_series, world_size):
    per = (n_series + world_size - 1) // world_size
    return series_id // per


def select_tag_subset(n_series, keep_one_in=10):
    """Deterministic 'tag predicate': ids with hash(id) % keep_one_in == 0, as a sorted u32 list
    (what get_series_id_by_filter would hand to the scan, tskv/src/kvcore.rs:249-279)."""
    ids = np.arange(n_series, dtype=np.uint64)
    h = (ids * np.uint64(0x9E3779B97F4A7C15)) >> np.uint64(32)
    return ids[h % np.uint64(keep_one_in) == 0].astype(np.uint32)


def allreduce_sections(sections, group=None):
    """Element-wise combine of the partial sections of every rank, in place.

    sections: dict with int64 tensors 'sum_i64', 'min_i64', 'max_i64', 'sel_val', float64 'sum_f64' and
    ints 'first_len', 'last_len' (FIRST keys are the tail of 'min_i64', LAST keys the tail of 'max_i64',
    'sel_val' holds FIRST then LAST values in the same order).
      counts / integer sums : SUM (wrapping int64 add == the reference's wrapping u64/i64 sum)
      f64 sums              : SUM
      min / max keys        : MIN / MAX on order-preserving int64 keys
      first / last          : keys are unique per (timestamp, series slot); after the key MIN/MAX only the
                              owning rank keeps its value, the others contribute 0 to a SUM.
    """
    nf, nl = sections["first_len"], sections["last_len"]
    mn, mx, sv = sections["min_i64"], sections["max_i64"], sections["sel_val"]
    local_first = mn[mn.numel() - nf:].clone() if nf else None
    local_last = mx[mx.numel() - nl:].clone() if nl else None
    if sections["sum_i64"].numel():
        dist.all_reduce(sections["sum_i64"], op=dist.ReduceOp.SUM, group=group)
    if sections["sum_f64"].numel():
        dist.all_reduce(sections["sum_f64"], op=dist.ReduceOp.SUM, group=group)
    if mn.numel():
        dist.all_reduce(mn, op=dist.ReduceOp.MIN, group=group)
    if mx.numel():
        dist.all_reduce(mx, op=dist.ReduceOp.MAX, group=group)
    if nf:
        won = mn[mn.numel() - nf:]
        sv[:nf].masked_fill_((local_first != won) | (won == KEY_FIRST_IDENTITY), 0)
    if nl:
        won = mx[mx.numel() - nl:]
        sv[nf:nf + nl].masked_fill_((local_last != won) | (won == KEY_LAST_IDENTITY), 0)
    if nf + nl:
        dist.all_reduce(sv, op=dist.ReduceOp.SUM, group=group)
    return sections


def scan_sections(scan, device):
    """Wrap the device partials of a PreparedScan as torch tensors (no copies)."""
    v = scan.partials()
    return {
        "sum_i64": device_tensor(v.sum_i64_ptr, v.sum_i64_len, torch.int64, device),
        "sum_f64": device_tensor(v.sum_f64_ptr, v.sum_f64_len, torch.float64, device),
        "min_i64": device_tensor(v.min_i64_ptr, v.min_i64_len, torch.int64, device),
        "max_i64": device_tensor(v.max_i64_ptr, v.max_i64_len, torch.int64, device),
        "sel_val": device_tensor(v.sel_val_ptr, v.sel_val_len, torch.int64, device),
        "first_len": int(v.sel_first_len), "last_len": int(v.sel_last_len),
    }


def allreduce_scan(scan, engine, sections=None, group=None):
    """All-reduce the partials of an enqueued scan, ordered on the engine's stream."""
    device = torch.device("cuda", engine.device)
    sections = sections or scan_sections(scan, device)
    with torch.cuda.stream(torch.cuda.ExternalStream(engine.stream(), device=device)):
        allreduce_sections(sections, group=group)
    return sections


def init_engine_comm(engine, rank, world_size, group=None):
    """Creates the library's NCCL communicator for this rank's engine: rank 0 draws the unique id, torch.distributed
    (any backend) only carries those 128 bytes to the other ranks. Returns False if the library cannot load NCCL."""
    try:
        ident = [engine.comm_unique_id() if rank == 0 else None]
    except Exception:
        ident = [None]
    dist.broadcast_object_list(ident, src=0, group=group)
    if ident[0] is None:
        return False
    engine.comm_init(ident[0], rank, world_size)
    return True


class GatherExchange:
    """The one collective of a multi-GPU scan: all-gather every rank's exchange region, merge locally. Default: inside
    the library (tskvgpu_scan_exchange: ncclAllGather + merge kernel on the engine stream; the engine needs
    init_engine_comm first). use_torch=True keeps the round-1 path - torch.distributed's all_gather into a torch buffer +
    tskvgpu_scan_merge_gathered - for hosts that drive their own collectives."""

    def __init__(self, scan, engine, world_size, group=None, use_torch=False):
        self.scan, self.engine, self.world, self.group = scan, engine, world_size, group
        self.use_torch = use_torch
        if use_torch:
            self.device = torch.device("cuda", engine.device)
            ptr, words = scan.exchange_view()
            self.local = device_tensor(ptr, words, torch.int64, self.device)
            self.gathered = torch.empty(world_size * words, dtype=torch.int64, device=self.device)
            self.stream = torch.cuda.ExternalStream(engine.stream(), device=self.device)

    def run(self):
        if not self.use_torch:
            self.scan.exchange()
            return
        with torch.cuda.stream(self.stream):
            dist.all_gather_into_tensor(self.gathered, self.local, group=self.group)
        self.scan.merge_gathered(self.gathered.data_ptr(), self.world)
