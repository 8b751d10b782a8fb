"""Multi-GPU partitioning of the path: the units are independent clips (or frame pairs);
clip c goes to rank c mod G; there is NO data-path collective -- only a final host gather of the per-clip
results (tracks, VP rows) to rank 0, which writes the CSVs.  Works with any torch.distributed backend
(NCCL on the GPU box, gloo in the CPU tests)."""
import os

import torch.distributed as dist


def assign(n_units, world_size, rank):
    """Indices of the units (clips / frame pairs) owned by `rank`: round-robin, c mod G."""
    return list(range(rank, n_units, world_size))


def gather_to_root(local_results, root=0, group=None):
    """local_results: dict unit_index -> picklable result.  Returns the merged dict on `root`, None
    elsewhere.  A single-process run (no process group) returns the input."""
    if not (dist.is_available() and dist.is_initialized()):
        return dict(local_results)
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    bucket = [None] * world if rank == root else None
    dist.gather_object(dict(local_results), bucket, dst=root, group=group)
    if rank != root:
        return None
    merged = {}
    for part in bucket:
        for k, v in part.items():
            if k in merged:
                raise RuntimeError(f"unit {k} was processed by two ranks")
            merged[k] = v
    return merged


def write_vp_csvs(rows_by_clip, names, out_dir):
    """One ./vps/vps_<name>.csv per clip with the reference's schema (save_csv, VP.py:388-392).  A value may be the rows
    of a clip or its CSV text already rendered by the rank that tracked it (render_csvs)."""
    from .vp import csv_text
    os.makedirs(out_dir, exist_ok=True)
    paths = []
    for c in sorted(rows_by_clip):
        path = os.path.join(out_dir, f"vps_{names[c]}.csv")
        v = rows_by_clip[c]
        with open(path, "w", newline="") as f:
            f.write(v if isinstance(v, str) else csv_text(v))
        paths.append(path)
    return paths


def render_csvs(rows_by_clip):
    """{clip: rows} -> {clip: CSV text}: formatting the rows is host work that shards with the clips, so every rank
    renders its own and the root only writes files."""
    from .vp import csv_text
    return {c: csv_text(r) for c, r in rows_by_clip.items()}
