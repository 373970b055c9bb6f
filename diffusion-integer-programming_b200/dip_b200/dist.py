"""Multi-GPU layer: the path shards by independent (instance, sample) units -- no collective inside
the step loop; ONE all_gather of fixed-size per-instance records at the end.

One process per GPU (torchrun); ``torch.distributed`` with NCCL over NVLink on the GPU box, gloo in
the CPU tests.  Noise is a pure function of (instance_id, sample_id), so any world size produces
the same samples.
"""
import numpy as np
import torch
import torch.distributed as dist

INT64_MAX = np.iinfo(np.int64).max


def partition(n_instances, samples_per_instance, rank, world):
    """Work items of `rank`: list of (instance_id, sample_begin, sample_end).

    #instances >= world: whole instances round-robin (keeps the CSR / conditions rank-local);
    otherwise the sample range of every instance is split into `world` contiguous chunks."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError("bad rank/world")
    if n_instances >= world:
        return [(i, 0, samples_per_instance) for i in range(rank, n_instances, world)]
    items = []
    for i in range(n_instances):
        lo = (samples_per_instance * rank) // world
        hi = (samples_per_instance * (rank + 1)) // world
        if hi > lo:
            items.append((i, lo, hi))
    return items


def record_width(n_vars):
    """bytes per instance record: best_obj int64 | n_feasible int32 | n_samples int32 | bit-packed best x."""
    return 16 + (n_vars + 7) // 8


def make_records(n_instances, n_vars):
    rec = np.zeros((n_instances, record_width(n_vars)), dtype=np.uint8)
    rec[:, :8] = np.frombuffer(np.int64(INT64_MAX).tobytes(), dtype=np.uint8)
    return rec


def update_record(rec, instance_id, xbits, viol_int, obj_int):
    """Fold one wave's results (numpy: xbits (B,n) uint8, viol_int (B,), obj_int (B,)) into the record
    of `instance_id`: count feasible samples, keep the feasible assignment of minimum objective."""
    row = rec[instance_id]
    best = int(np.frombuffer(row[:8].tobytes(), dtype=np.int64)[0])
    counts = np.frombuffer(row[8:16].tobytes(), dtype=np.int32).copy()
    feas = viol_int == 0
    counts[0] += int(feas.sum())
    counts[1] += int(len(viol_int))
    if feas.any():
        objs = np.where(feas, obj_int, INT64_MAX)
        j = int(np.argmin(objs))
        if int(objs[j]) < best:
            best = int(objs[j])
            row[16:] = np.packbits(xbits[j].astype(np.uint8))
    row[:8] = np.frombuffer(np.int64(best).tobytes(), dtype=np.uint8)
    row[8:16] = np.frombuffer(counts.tobytes(), dtype=np.uint8)


def merge_records(stacked):
    """(world, n_instances, width) -> (n_instances, width): min objective wins (lowest rank on ties), counts add."""
    world, n_inst, width = stacked.shape
    out = stacked[0].copy()
    for i in range(n_inst):
        best = INT64_MAX
        feas = tot = 0
        for r in range(world):
            row = stacked[r, i]
            o = int(np.frombuffer(row[:8].tobytes(), dtype=np.int64)[0])
            c = np.frombuffer(row[8:16].tobytes(), dtype=np.int32)
            feas += int(c[0]); tot += int(c[1])
            if o < best:
                best = o
                out[i, 16:] = row[16:]
        out[i, :8] = np.frombuffer(np.int64(best).tobytes(), dtype=np.uint8)
        out[i, 8:16] = np.frombuffer(np.array([feas, tot], dtype=np.int32).tobytes(), dtype=np.uint8)
    return out


def decode_record(row, n_vars):
    best = int(np.frombuffer(row[:8].tobytes(), dtype=np.int64)[0])
    c = np.frombuffer(row[8:16].tobytes(), dtype=np.int32)
    x = np.unpackbits(row[16:])[:n_vars]
    return {"best_obj": None if best == INT64_MAX else best, "n_feasible": int(c[0]), "n_samples": int(c[1]), "best_x": x}


def all_gather_records(rec, device=None):
    """The single collective of the path: every rank ends with the merged per-instance bests."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return rec
    world = dist.get_world_size()
    backend = dist.get_backend()
    t = torch.from_numpy(rec)
    if backend == "nccl":
        t = t.to(device or torch.device("cuda", torch.cuda.current_device()))
    out = torch.empty((world * t.shape[0], t.shape[1]), dtype=t.dtype, device=t.device)   # concatenated layout (gloo and nccl)
    dist.all_gather_into_tensor(out, t.contiguous())
    return merge_records(out.cpu().numpy().reshape(world, t.shape[0], t.shape[1]))
