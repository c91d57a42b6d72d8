"""Width-bucket sharding of text lines across data-parallel ranks (host logic; SURVEY.md §8e, reference
src/ocr_dataset.py:46-93 width buckets + src/datautils.py:4-51 GroupedSampler).

Lines are independent, so the path shards with NO data-path collective at inference; in training the only exchange is
the gradient all-reduce.  What matters is balance: a step costs ~ (padded width x batch), so every rank should draw
its batch from the SAME width bucket at each step.
"""
import numpy as np

# normalised-height-30 width limits used by the reference (ocr_dataset.py:60)
BUCKET_LIMITS = (150, 200, 300, 350, 450, 600, float("inf"))


def bucket_of(width, line_height, limits=BUCKET_LIMITS):
    w30 = width * 30.0 / line_height
    for i, lim in enumerate(limits):
        if w30 <= lim:
            return i
    return len(limits) - 1


def shard_batches(widths, line_height, batch_size, world_size, rank, limits=BUCKET_LIMITS, drop_last=True):
    """Deterministic plan: returns this rank's list of batches (each a list of line indices sorted by width
    descending, the SortByWidthCollater contract).  Batches are formed inside width buckets; consecutive groups of
    `world_size` batches from the same bucket are dealt one per rank, so at every step all ranks hold batches of
    similar padded width.  Buckets whose batch count is not a multiple of world_size give their tail batches to the
    ranks round-robin only if drop_last is False."""
    widths = np.asarray(widths)
    by_bucket = {}
    for i, w in enumerate(widths):
        by_bucket.setdefault(bucket_of(int(w), line_height, limits), []).append(i)
    mine = []
    for b in sorted(by_bucket):
        idx = sorted(by_bucket[b], key=lambda i: (-int(widths[i]), i))
        batches = [idx[k:k + batch_size] for k in range(0, len(idx), batch_size)]
        if drop_last and batches and len(batches[-1]) < batch_size:
            batches.pop()
        full_groups = len(batches) // world_size
        for g in range(full_groups):
            mine.append(batches[g * world_size + rank])
        if not drop_last:
            tail = batches[full_groups * world_size:]
            if rank < len(tail):
                mine.append(tail[rank])
    return mine
